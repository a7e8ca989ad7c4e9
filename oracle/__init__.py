"""CPU oracle: test infrastructure only (see graph_oracle.py / model_oracle.py headers).

The product package (gnn-lm_b200/, importable as gnnlm_b200) must never import this."""
