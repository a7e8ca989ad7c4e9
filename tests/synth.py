"""Test-side glue: run the same synthetic problem through the CUDA path and through the CPU oracle."""
import numpy as np
import torch

from gnnlm_b200 import synth
from oracle import model_oracle as mo

_CACHE = {}


def make_problem(name: str, n_d: int = None, seed: int = 0):
    key = (name, n_d, seed)
    if key not in _CACHE:
        cfg = dict(synth.CONFIGS[name])
        if n_d:
            cfg["n_d"] = n_d
        model = synth.make_model(cfg)
        data = synth.make_data(cfg, seed=seed, device="cpu")
        _CACHE[key] = (cfg, model, data)
    return _CACHE[key]


def oracle_model(cfg, model) -> dict:
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    hgt = {k[len("decoder.hgt_decoder."):]: v for k, v in sd.items() if k.startswith("decoder.hgt_decoder.")}
    q = model.decoder.tgt_quantizer
    if cfg["cutoff"] is not None:
        soft = mo.adaptive_weights({k[len("decoder.adaptive_softmax."):]: v for k, v in sd.items()
                                    if k.startswith("decoder.adaptive_softmax.")})
        cutoff = list(cfg["cutoff"]) + [cfg["V"]]
    else:
        soft, cutoff = {"plain": sd["decoder.embed_out"]}, None
    return {"sd": hgt, "n_heads": cfg["H"], "n_layers": cfg["NL"], "centroids": q.centroids_torch.numpy(),
            "A": q.A.numpy(), "b": q.b.numpy(), "softmax": soft, "cutoff": cutoff}


def run_oracle(prob, dtype=torch.float32, device="cpu", dense_tt=None) -> dict:
    """The oracle over one synthetic problem.  device="cpu": everything on the host (data may live anywhere).  A CUDA device:
    the oracle's torch statements run there (fp64 checker for the full-size configurations); the datastore tables stay where
    they are (the code rows of the graph's nodes are gathered by torch indexing)."""
    cfg, model, data = prob
    B, L = cfg["B"], cfg["L"]
    cpu = lambda t: t.cpu()
    codes = data["codes"].numpy() if str(device) == "cpu" and data["codes"].device.type == "cpu" else data["codes"]
    if str(device) == "cpu" and torch.is_tensor(codes):
        codes = codes.cpu()
    batch = {"nbr": cpu(data["nbr"]).numpy(), "offsets": cpu(data["positions"]).numpy(), "tgt_feats": data["feats"].float(),
             "target": data["target"], "codes": codes, "cl": cfg["c"], "cr": cfg["c"], "n_d": data["n_d"]}
    knn = {"dists": data["knn_dists"], "ids": data["knn_ids"], "vals": data["vals"].long(), "lmbda": cfg["lmbda"],
           "temperature": cfg["temp"]}
    if str(device) == "cpu":
        batch = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in batch.items()}
        knn = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in knn.items()}
    out = mo.eval_batch(oracle_model(cfg, model), batch, knn, dtype=dtype, device=device, dense_tt=dense_tt)
    nll2, ppl = mo.perplexity(out["score_sum"], out["count"])
    out["ppl"], out["nll"] = ppl, -out["score_sum"] / out["count"]
    return out


def run_gpu(prob, dev, math="fp32") -> dict:
    cfg, model, data = prob
    import copy
    return synth.run_gpu(cfg, copy.deepcopy(model), data, dev, math)


def oracle_train(prob, dtype=torch.float64, deprecated=False):
    """Loss of fairseq/criterions/adaptive_loss.py:31-83 through the oracle's HGT statements and its gradients w.r.t. every
    decoder.hgt_decoder.* tensor by torch.autograd -- which is how the reference itself obtains them (autograd over hgt.py)."""
    from oracle import graph_oracle as go
    cfg, model, data = prob
    om = oracle_model(cfg, model)
    B, L = cfg["B"], cfg["L"]
    if deprecated:
        g = go.build_batch(data["nbr"].numpy(), data["positions"].numpy(), data["n_d"], cfg["c"], cfg["c"], deprecated=True)
    else:
        g = go.build_batch_vectorised(data["nbr"].numpy(), data["positions"].numpy(), data["n_d"], cfg["c"], cfg["c"])
    codes = data["codes"].numpy()[g["ntgt_offsets"]]
    h_ntgt = torch.from_numpy(mo.pq_decode(codes, om["centroids"], om.get("A"), om.get("b"), np.float64)).to(dtype)
    sd = {k: v.detach().to(dtype).requires_grad_(True) for k, v in om["sd"].items()}
    h = mo.hgt_forward_csr(sd, data["feats"].to(dtype), h_ntgt, g, (B, L), om["n_heads"], om["n_layers"])
    soft = {k: ([t.to(dtype) for t in v] if isinstance(v, list) else v.to(dtype)) for k, v in om["softmax"].items()}
    loss = mo.adaptive_loss(soft, om["cutoff"], h["tgt"], data["target"])
    names = list(sd)
    grads = torch.autograd.grad(loss, [sd[n] for n in names], allow_unused=True)
    return float(loss.detach()), {n: g_ for n, g_ in zip(names, grads)}
