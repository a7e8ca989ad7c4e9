"""`--user-dir` entry point (fairseq/utils.py:315-330 imports this directory as a top-level module):

    fairseq-eval-lm DATA --user-dir /path/to/repo/gnn-lm_b200/fairseq_plugin --path ckpt.pt --graph --use-precompute-feat ...

Registers the B200 graph LM with fairseq's own registries (gnnlm_b200.registry.register_with_fairseq): model `hgt_lm` and its
architectures, task `graph_language_modeling`, and -- unless GNNLM_PLUGIN_OVERRIDE=0 -- re-points `transformer_lm` /
`language_modeling` so that existing checkpoints and scripts resolve here when `--graph_layer > 0` / `--graph` are set (anything else
is handed back to stock fairseq), and swaps the scorer `fairseq_cli.eval_lm` constructs.  fairseq itself is not importable in the
build image (numpy 2.x, no dgl / faiss), so the registration logic is tested against a stand-in registry
(tests/test_multi_rank_cpu.py::test_fairseq_registration_*)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from gnnlm_b200 import registry as _registry            # noqa: E402
from gnnlm_b200.sequence_scorer import SequenceScorer    # noqa: E402


def _install():
    import fairseq.models as fm
    import fairseq.tasks as ft
    override = os.environ.get("GNNLM_PLUGIN_OVERRIDE", "1") != "0"
    names = _registry.register_with_fairseq(fm, ft, override=override)
    if override:
        import fairseq.sequence_scorer as fs
        fs.SequenceScorer = SequenceScorer
        cli = sys.modules.get("fairseq_cli.eval_lm")           # it bound the name at import time (eval_lm.py:22)
        if cli is not None:
            cli.SequenceScorer = SequenceScorer
    return names


ARCHITECTURES = _install()
