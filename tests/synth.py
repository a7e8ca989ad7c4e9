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


def run_oracle(prob, dtype=torch.float32) -> dict:
    cfg, model, data = prob
    B, L = cfg["B"], cfg["L"]
    batch = {"nbr": data["nbr"].numpy(), "offsets": data["positions"].numpy(), "tgt_feats": data["feats"].float(),
             "target": data["target"], "codes": data["codes"].numpy(), "cl": cfg["c"], "cr": cfg["c"], "n_d": data["n_d"]}
    knn = {"dists": data["knn_dists"], "ids": data["knn_ids"], "vals": data["vals"].long(), "lmbda": cfg["lmbda"],
           "temperature": cfg["temp"]}
    out = mo.eval_batch(oracle_model(cfg, model), batch, knn, dtype=dtype)
    nll2, ppl = mo.perplexity(out["score_sum"], out["count"])
    out["ppl"], out["nll"] = ppl, -out["score_sum"] / out["count"]
    return out


def run_gpu(prob, dev, math="fp32") -> dict:
    cfg, model, data = prob
    import copy
    return synth.run_gpu(cfg, copy.deepcopy(model), data, dev, math)
