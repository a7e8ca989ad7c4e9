"""Seeded synthetic problems of the shapes BASELINE.json names: random-init weights
initialised like the reference modules, uniform tokens / codes / neighbour ids, N(0,1) features.
Used by bench.py, __graft_entry__.smoke() and the tests.  No oracle imports here."""
from argparse import Namespace
from types import SimpleNamespace

import numpy as np
import torch

from . import ops
from .dataset import DeviceDatastore
from .graph import build_token_graph
from .knn_model import KNNModel
from .model import TransformerLanguageModel, default_args
from .pq_codec import TorchPQCodec
from .sequence_scorer import SequenceScorer

CONFIGS = {
    # BASELINE.json configs[0]: tiny, CPU-runnable
    "c1": dict(d=512, H=8, V=10000, cutoff=[2000, 6000], tied=False, B=2, L=256, k=8, c=1, M=64, NL=1, n_d=1 << 20,
               k_nn=64, lmbda=0.25, temp=1.0),
    # configs[1]: enwik8 shape (char vocab, plain softmax)
    "c2": dict(d=512, H=8, V=204, cutoff=None, tied=False, B=1, L=512, k=32, c=1, M=64, NL=3, n_d=1 << 24,
               k_nn=1024, lmbda=0.25, temp=1.0),
    # configs[2]: wiki103 shape (the headline metric's config)
    "c3": dict(d=1024, H=8, V=267744, cutoff=[20000, 60000], tied=True, B=1, L=3072, k=32, c=1, M=128, NL=3,
               n_d=103227021, k_nn=1024, lmbda=0.25, temp=1.0),
    # the Wiki103 shape at the setting of the reference's own evaluation script (gnnlm_scripts/wiki103/hgt_lm_wiki103_reproduce.sh:
    # 127-147: one 256-token sample per batch, --gcn-k 128, --neighbor-context 2, --k 1024, --lmbda 0.1, --temperature 0.01)
    "c3e": dict(d=1024, H=8, V=267744, cutoff=[20000, 60000], tied=True, B=1, L=256, k=128, c=2, M=128, NL=3,
                n_d=103227021, k_nn=1024, lmbda=0.1, temp=0.01),
    # configs[3]: One-Billion-Word shape (vocab 793,471, SURVEY.md 8 C4: d=1024 inferred from the +0.02B parameters; adaptive
    # cutoffs are the checkpoint's -- 60000/160000 assumed; ~0.8 G datastore tokens x 128 B = 98 GB of codes in HBM).  The
    # reference batches this corpus by sentence (--sample-break-mode eos); blocks here are packed to 3072 tokens.
    "c4": dict(d=1024, H=8, V=793471, cutoff=[60000, 160000], tied=True, B=1, L=3072, k=32, c=1, M=128, NL=3,
               n_d=768 * 1000 * 1000, k_nn=1024, lmbda=0.25, temp=1.0),
    # a small wiki103-like problem for tests (adaptive, tied, 3 layers)
    "c3mini": dict(d=1024, H=8, V=30000, cutoff=[4000, 12000], tied=True, B=1, L=192, k=8, c=1, M=128, NL=3,
                   n_d=1 << 18, k_nn=128, lmbda=0.25, temp=1.0),
}


class Dictionary:
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def pad(self):
        return 1

    def eos(self):
        return 2


def make_model(cfg: dict, seed: int = 1) -> TransformerLanguageModel:
    """Random-init model on the CPU, `torch.manual_seed(seed)` like the scripts' --seed 1."""
    torch.manual_seed(seed)
    c = SimpleNamespace(**cfg)
    rng = np.random.RandomState(seed)
    dsub = c.d // c.M
    cen = rng.randn(c.M, 256, dsub).astype(np.float32)
    A = np.linalg.qr(rng.randn(c.d, c.d))[0].astype(np.float32)
    codec = TorchPQCodec(centroids=cen, A=A, b=np.zeros(0, np.float32))
    args = default_args(decoder_embed_dim=c.d, decoder_attention_heads=c.H, graph_layer=c.NL, decoder_gcn_dim=cfg.get("gcn_dim"),
                        adaptive_softmax_cutoff=c.cutoff, tie_adaptive_weights=c.tied)
    model = TransformerLanguageModel.build_model(args, dictionary=Dictionary(c.V), quantizer=codec)
    with torch.no_grad():   # make biases / pri / LN affine non-trivial so that parity exercises them
        for n_, p in model.named_parameters():
            if p.dim() == 1 or n_.endswith("relation_pri"):
                p.add_(0.05 * torch.randn_like(p))
    return model.eval()


def make_tables(cfg: dict, seed: int = 0, n_d: int = None, device="cpu") -> dict:
    """The datastore tables (quantized-keys.npy / vals.npy stand-ins), generated directly on `device`."""
    c = SimpleNamespace(**cfg)
    n_d = n_d or c.n_d
    g = torch.Generator(device=device).manual_seed(seed)
    codes = torch.randint(0, 256, (n_d, c.M), generator=g, device=device, dtype=torch.uint8)
    vals = torch.randint(4, c.V, (n_d,), generator=g, device=device, dtype=torch.int32)
    return dict(codes=codes, vals=vals, n_d=n_d)


def local_neighbours(B: int, L_: int, k: int, n_d: int, c: int, p_continue: float, n_hot: int, generator, device):
    """Neighbour ids with the locality of real kNN graphs instead of uniform draws: with probability `p_continue` neighbour j of
    token t continues neighbour j of token t-1 (row o + 1: adjacent tokens retrieve adjacent datastore rows, so their clusters
    overlap), otherwise it is a fresh retrieval from `n_hot` popular rows with Zipfian (1 / rank) popularity, scattered over
    the datastore -- recurring contexts retrieve the SAME row, which is what makes centres repeat inside a block."""
    u = torch.rand((B, L_, k), generator=generator, device=device, dtype=torch.float64)
    rank = torch.exp(u * float(np.log(n_hot))).long().clamp_(1, n_hot)                  # P(rank = r) ~ 1 / r
    fresh = (rank * 2654435761 + 12345) % (n_d - 2 * c - L_) + c
    cont = torch.rand((B, L_, k), generator=generator, device=device) < p_continue
    cont[:, 0] = False
    t = torch.arange(L_, device=device).view(1, L_, 1).expand(B, L_, k)
    start = torch.cummax(torch.where(cont, torch.zeros_like(t), t), dim=1).values       # first token of the run each entry is in
    return torch.gather(fresh, 1, start) + (t - start)


def make_batch(cfg: dict, tables: dict, seed: int = 0, device="cpu", stress: bool = True, locality=None) -> dict:
    """One batch of per-token inputs: neighbour ids, fp16 features, targets, kNN-LM search results.
    `locality` = (p_continue, n_hot): neighbour ids from local_neighbours instead of uniform draws."""
    c = SimpleNamespace(**cfg)
    n_d, vals = tables["n_d"], tables["vals"]
    g = torch.Generator(device=device).manual_seed(seed + 12345)
    T = c.B * c.L
    nbr = torch.randint(c.c, n_d - c.c, (c.B, c.L, c.k), generator=g, device=device, dtype=torch.int64)
    if locality is not None:
        nbr = local_neighbours(c.B, c.L, c.k, n_d, c.c, float(locality[0]), int(locality[1]), g, device)
    if stress:   # 1% missing, 0.1% within c of either boundary
        r = torch.rand((c.B, c.L, c.k), generator=g, device=device)
        nbr[r < 0.01] = -1
        edge = (r >= 0.01) & (r < 0.011)
        nbr[edge] = torch.where(torch.rand(int(edge.sum()), generator=g, device=device) < 0.5, 0, n_d - 1)
    feats = torch.randn((T, c.d), generator=g, device=device).half()
    target = torch.randint(4, c.V, (c.B, c.L), generator=g, device=device, dtype=torch.int64)
    dists = torch.randn((T, c.k_nn), generator=g, device=device)
    ids = torch.randint(0, n_d, (T, c.k_nn), generator=g, device=device, dtype=torch.int64)
    ids[torch.rand((T, c.k_nn), generator=g, device=device) < 0.002] = -1
    # make a share of the retrieved values hit the target, as a real datastore would
    hit = torch.rand((T,), generator=g, device=device) < 0.5
    j = torch.randint(0, c.k_nn, (T,), generator=g, device=device)
    rows = ids[torch.arange(T, device=device), j]
    ok = hit & (rows >= 0)
    flat_t = target.reshape(-1).clone()
    flat_t[ok] = vals[rows[ok].to(vals.device)].long().to(device)
    target = flat_t.view(c.B, c.L)
    return dict(nbr=nbr, feats=feats, target=target, knn_dists=dists, knn_ids=ids,
                positions=torch.arange(T, device=device).view(c.B, c.L))


def make_data(cfg: dict, seed: int = 0, n_d: int = None, device="cpu", stress: bool = True, locality=None) -> dict:
    tables = make_tables(cfg, seed, n_d, device)
    out = dict(tables)
    out.update(make_batch(cfg, tables, seed, device, stress, locality))
    return out


def to_device(data: dict, device) -> dict:
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in data.items()}


def write_data_dir(root: str, cfg: dict, n_blocks: int, tail_tokens: int = 0, seed: int = 0, split: str = "valid",
                   device="cpu") -> dict:
    """A synthetic data directory in the reference's layout (knn/path_utils.py:13-41, language_modeling.py:266-298):
    dict.txt, {split}.bin/.idx, {split}_dstore/{info.json, keys.npy, neighbors.mmap.{k}, neighbors.mmap.{k_nn}, dists.{k_nn}},
    train_dstore/{info.json, vals.npy, quantized-keys.npy} -- n_blocks blocks of cfg['L'] tokens (+ a ragged tail).  The kNN-LM
    neighbour distances are not a reference file (find_knn.py:65-66 drops them): they go to `dists.{k_nn}` for
    eval_lm's --knn-dists-file.  Arrays are generated on `device` and written from the host."""
    import json
    import os
    from .formats import write_mmap_indexed
    c = SimpleNamespace(**cfg)
    n_tok = n_blocks * c.L + tail_tokens
    g = torch.Generator(device=device).manual_seed(seed + 777)
    os.makedirs(os.path.join(root, f"{split}_dstore"), exist_ok=True)
    os.makedirs(os.path.join(root, "train_dstore"), exist_ok=True)
    tables = make_tables(cfg, seed, device=device)
    host = lambda t: t.cpu().numpy()
    tokens = host(torch.randint(4, c.V, (n_tok,), generator=g, device=device, dtype=torch.int64))
    tokens[c.L - 1::c.L] = 2                                             # an eos per block keeps sentences = blocks
    tokens[-1] = 2
    bounds = list(range(0, n_tok, c.L)) + [n_tok]
    tdt = np.uint16 if c.V < 65500 else np.int32                       # indexed_dataset.py:29-33 (__best_fitting_dtype)
    write_mmap_indexed(os.path.join(root, split), [tokens[a:b] for a, b in zip(bounds[:-1], bounds[1:])], tdt)
    with open(os.path.join(root, "dict.txt"), "w") as f:
        f.write("".join(f"w{i} {c.V - i}\n" for i in range(4, c.V)))
    nbr = torch.randint(c.c, c.n_d - c.c, (n_tok, c.k), generator=g, device=device, dtype=torch.int64)
    nbr[torch.rand((n_tok, c.k), generator=g, device=device) < 0.01] = -1
    host(nbr).tofile(os.path.join(root, f"{split}_dstore", f"neighbors.mmap.{c.k}"))
    host(torch.randn((n_tok, c.d), generator=g, device=device).half()).tofile(os.path.join(root, f"{split}_dstore", "keys.npy"))
    ids = torch.randint(0, c.n_d, (n_tok, c.k_nn), generator=g, device=device, dtype=torch.int64)
    host(ids).tofile(os.path.join(root, f"{split}_dstore", f"neighbors.mmap.{c.k_nn}"))
    host(torch.randn((n_tok, c.k_nn), generator=g, device=device)).tofile(os.path.join(root, f"{split}_dstore", f"dists.{c.k_nn}"))
    info = {"hidden_size": c.d, "vocab_size": c.V, "dstore_fp16": True, "val_size": 1}
    json.dump(dict(info, dstore_size=n_tok), open(os.path.join(root, f"{split}_dstore", "info.json"), "w"))
    json.dump(dict(info, dstore_size=c.n_d), open(os.path.join(root, "train_dstore", "info.json"), "w"))
    vdt = np.int16 if c.V < 2 ** 15 else np.int32                       # language_modeling.py:272 (dstore_fp16)
    host(tables["vals"]).astype(vdt).reshape(-1, 1).tofile(os.path.join(root, "train_dstore", "vals.npy"))
    np.save(os.path.join(root, "train_dstore", "quantized-keys.npy"), host(tables["codes"]))
    return {"n_tokens": n_tok, "n_blocks": len(bounds) - 1, "tables": tables,
            "dists_file": os.path.join(root, f"{split}_dstore", f"dists.{c.k_nn}")}


class Runner:
    """The call a user makes per batch, on device-resident tables: graph assembly -> PQ decode -> HGT ->
    log-probs -> kNN mix -> NLL accumulation.  `step(host_batch)` is the e2e variant (pinned host
    inputs, H2D inside)."""

    def __init__(self, cfg: dict, model: TransformerLanguageModel, data: dict, device, math: str = "fp32",
                 prune_unreachable: bool = True):
        self.cfg, self.c, self.device, self.prune_unreachable = cfg, SimpleNamespace(**cfg), device, prune_unreachable
        self.model = model.to(device).set_math(math)
        self.dstore = DeviceDatastore(data["codes"].to(device), data["vals"].to(device))
        self.knn = KNNModel(self.dstore.vals, vocab_size=self.c.V, metric_type="do_not_recomp_ip", k=self.c.k_nn)
        self.args = Namespace(lmbda=self.c.lmbda, knn_keytype=None)
        self.scorer = SequenceScorer(Dictionary(self.c.V), args=self.args)
        self.acc = torch.zeros(2, dtype=torch.float64, device=device)
        from .eval_lm import GraphedScorer
        self._graphed = GraphedScorer(self._score, device=device)
        self._stager = None

    def sample_from(self, nbr, feats, target, dists, ids):
        c = self.c
        g = build_token_graph(nbr, self.dstore.size, c.c, c.c, reach=c.NL - 1 if self.prune_unreachable else None,
                              dedup=bool(self.cfg.get("deprecated", False)))
        g.codes_table = self.dstore.codes
        g.nodes["tgt"].data["h"] = feats.view(-1, feats.shape[-1])
        self.knn.set_search_results(dists, ids)
        return {"net_input": {"src_tokens": target, "graph": g}, "target": target, "ntokens": target.numel()}

    KEYS = ("nbr", "feats", "target", "knn_dists", "knn_ids")

    def _score(self, inp: dict, dry: bool = False, want_knn: bool = False):
        s = self.sample_from(inp["nbr"], inp["feats"], inp["target"], inp["knn_dists"], inp["knn_ids"])
        return self.scorer.score_tokens(self.model, s, self.knn, self.c.temp, nll_acc=None if dry else self.acc,
                                        want_knn=want_knn)

    def step_resident(self, dev_batch: dict, want_knn=False, cuda_graph=False):
        """cuda_graph=True: replay the step as one captured CUDA graph (eval_lm.GraphedScorer)."""
        inp = {k: dev_batch[k] for k in self.KEYS}
        if cuda_graph and not want_knn:
            return self._graphed(inp)
        return self._score(inp, want_knn=want_knn)

    def run_host(self, host_batches, steps: int, cuda_graph=False):
        """`steps` e2e steps over pinned host batches (rotated), inputs double-buffered: step i is launched, then the H2D
        copy of step i+1 is issued on a copy stream (it runs under the kernels of step i), then the step's result (the
        16 B accumulator) is read back.  Returns the last read-back."""
        from .eval_lm import Stager
        if self._stager is None:
            self._stager = Stager(self.device)
        st, n, out = self._stager, len(host_batches), None
        pick = lambda i: {k: host_batches[i % n][k] for k in self.KEYS}
        staged = st.stage(pick(0))
        for i in range(steps):
            inp = st.wait(staged)
            if cuda_graph:
                self._graphed(inp)
            else:
                self._score(inp)
            st.done(staged)
            if i + 1 < steps:
                staged = st.stage(pick(i + 1))
            out = self.acc.cpu()      # D2H read of the step's result, synchronises
        return out

    def step_host(self, host_batch: dict, cuda_graph=False):
        """host_batch: pinned CPU tensors.  Copies in, scores, reads the scalar result back."""
        if cuda_graph:       # pinned host -> the graph's static device buffers, then one replay
            self._graphed({k: host_batch[k] for k in self.KEYS})
        else:
            self._score({k: host_batch[k].to(self.device, non_blocking=True) for k in self.KEYS})
        return self.acc.cpu()     # D2H read of the step's result (16 B), synchronises


def run_gpu(cfg, model, data, device, math="fp32") -> dict:
    r = Runner(cfg, model, data, device, math)
    d = to_device({k: data[k] for k in ("nbr", "feats", "target", "knn_dists", "knn_ids")}, device)
    lp, p, rec, dec_out = r.step_resident(d, want_knn=True)
    s, n = r.acc.tolist()
    nll2 = -s / n / np.log(2)
    return {"logprob": lp.reshape(-1).cpu().numpy(), "knn_prob": p.cpu().numpy(), "recall": rec.cpu().numpy(),
            "score_sum": s, "count": int(n), "ppl": float(2 ** nll2), "nll": -s / n,
            "gcn_feat": dec_out[0].reshape(-1, dec_out[0].shape[-1]).float().cpu().numpy()}
