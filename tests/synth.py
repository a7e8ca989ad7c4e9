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


# ---- replay of the library's dropout masks (include/gnnlm_sm100.h: gnnlm_dropout_f32 / attention p_drop, seed) through the oracle
def _dm_mix(seed: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def dropout_multiplier(seed: int, idx: np.ndarray, p: float) -> np.ndarray:
    thresh = np.uint64(int(np.float32(p) * np.float32(16777216.0)))
    keep = (_dm_mix(seed, idx) >> np.uint64(40)) >= thresh
    return np.where(keep, np.float32(1.0) / (np.float32(1.0) - np.float32(p)), np.float32(0.0)).astype(np.float64)


def oracle_dropout_hooks(cfg, seed: int, p_feat: float, p_att: float, p_soft: float):
    """(hooks(l) for hgt_forward_csr, drop_in, drop_tail) replaying train.py's masks: site seeds of train.site_seed, element index
    row * cols + col, edge index (dst << 38) ^ (src << 6) ^ head with inter sources in compact centre numbering."""
    from gnnlm_b200.train import site_seed
    H = cfg["H"]

    def elem(x, site, p):
        if p <= 0:
            return x
        r, c = x.shape
        m = dropout_multiplier(site_seed(seed, site), np.arange(r * c, dtype=np.uint64), p).reshape(r, c)
        return x * torch.from_numpy(m).to(x.dtype)

    def hooks(l):
        def attn(etype, att, src, dst):
            if p_att <= 0:
                return att
            site = 16 * l + {("tgt", "intra", "tgt"): 0, ("ntgt", "inter", "tgt"): 1, ("ntgt", "intra", "ntgt"): 2}[etype]
            s_np, d_np = src.numpy().astype(np.uint64), dst.numpy().astype(np.uint64)
            if etype[1] == "inter":                     # compact centre numbering = rank of the centre node id
                s_np = np.searchsorted(np.unique(s_np), s_np).astype(np.uint64)
            idx = ((d_np[:, None] << np.uint64(38)) ^ (s_np[:, None] << np.uint64(6)) ^ np.arange(H, dtype=np.uint64)[None, :])
            return att * torch.from_numpy(dropout_multiplier(site_seed(seed, site), idx, p_att)).to(att.dtype)
        return {"attn": attn, "feat": lambda t, out: elem(out, 16 * l + (3 if t == "tgt" else 4), p_feat)}
    return hooks, (lambda x: elem(x, 1000, p_soft)), (lambda i, hid, idx: elem(hid, 1001 + i, p_soft))


def oracle_train(prob, dtype=torch.float64, deprecated=False, dropout=None):
    """`dropout` = (seed, p_feat, p_att, p_soft): replay the library's masks (new builder only)."""
    # Loss of fairseq/criterions/adaptive_loss.py:31-83 through the oracle's HGT statements and its gradients w.r.t. every
    # decoder.hgt_decoder.* tensor by torch.autograd -- which is how the reference itself obtains them (autograd over hgt.py).
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
    hooks = drop_in = drop_tail = None
    if dropout is not None:
        hooks, drop_in, drop_tail = oracle_dropout_hooks(cfg, *dropout)
    h = mo.hgt_forward_csr(sd, data["feats"].to(dtype), h_ntgt, g, (B, L), om["n_heads"], om["n_layers"], dense_tt=False, hooks=hooks)
    soft = {k: ([t.to(dtype) for t in v] if isinstance(v, list) else v.to(dtype)) for k, v in om["softmax"].items()}
    loss = mo.adaptive_loss(soft, om["cutoff"], h["tgt"], data["target"], drop_in=drop_in, drop_tail=drop_tail)
    names = list(sd)
    grads = torch.autograd.grad(loss, [sd[n] for n in names], allow_unused=True)
    return float(loss.detach()), {n: g_ for n, g_ in zip(names, grads)}
