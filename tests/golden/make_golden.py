#!/usr/bin/env python
"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

Nothing here is imported by the product.  The reference package cannot be imported as a
package in this image (np.float at fairseq/data/indexed_dataset.py:89 under numpy 2.x; dgl,
faiss and pyarrow.plasma are not installed), so each function is lifted out of its file by
`ast` and executed unmodified in a namespace that stubs only the missing imports:

  graph_*.npz      GraphTokenBlockDataset.{new_build_graph, build_ntgt_edges,
                   auto_regressive_edges} (fairseq/data/token_block_dataset.py:338-412,545-594)
                   under a recording `dgl.heterograph` stub.  `(c,0)` cases run the unmodified
                   source; `(c,c)` cases apply the one-token fix of SURVEY.md Q1
                   (`len(self.neighbor_offsets.shape[0])` -> `len(self.neighbor_tokens)`).
  adaptive_*.npz   fairseq/modules/adaptive_softmax.py + adaptive_input.py loaded by file path
                   (pure torch): get_log_prob in target mode and full mode, tied and untied.
  pq_*.npz         NumpyPQCodec.decode and TorchPQCodec.decode bodies (knn/pq_wrapper.py:70-84,
                   169-203) run against a fake `self`.
  knn_*.npz        KNNModel.get_knn_prob (knn/knn_model.py:103-217) with get_knns replaced by
                   precomputed (dists, ids) -- the faiss search is out of scope.
  scorer_*.npz     SequenceScorer.generate (fairseq/sequence_scorer.py:27-194) with a scripted
                   model + the reference AdaptiveSoftmax + the reference get_knn_prob.
  hgt_*.npz        HGT / HGTLayer (fairseq/models/hgt.py:22-79,299-420,459-513) executed under
                   a ~100-line DGL stub that implements DGL's *documented* semantics for the
                   five calls the layer makes.  This pins the layer code (projection order,
                   einsum, scaling, residual, LayerNorm); the DGL semantics themselves are
                   restated, not executed (DGL is absent) -> "parity unpinned" at that boundary.
  hgt_hetero4 / hgt_two_stream / hgt_infer_*   the same module on a GENERAL heterograph (its own self-test topology, hgt.py:516-552),
                   with two_stream=True (after the one-statement fix of quirk Q11: the unmodified source raises KeyError('k_tilde'),
                   asserted here) and through HGTLayer.infer / reorder_incremental_state step by step (hgt.py:81-297,422-438); the
                   stub gains apply_edges(edges=subset), edges(), update_all(etype=), a scoping local_scope, and the reference's own
                   fairseq/incremental_decoding_utils.py mixin loaded by file path.  (`--hetero-only` regenerates just these.)
  adaptive_input_* fairseq/modules/adaptive_input.py AdaptiveInput.forward (the `--reinit-nfeat` ntgt features).
  registry.json    flag defaults of the model / task / eval-lm parsers and the `transformer_lm*` architecture presets, from the
                   reference's own add_args and architecture functions.
  fmt_*            {split}.bin/.idx + dict.txt written AND read back by the reference's own
                   MMapIndexedDatasetBuilder / MMapIndexedDataset / Dictionary (fmt.npz = what its readers return).
"""
import ast
import importlib.util
import math
import os
import sys
import textwrap
import types
from functools import lru_cache
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _src(path):
    with open(os.path.join(REF, path)) as f:
        return f.read()


def _class_source(path, cls):
    src = _src(path)
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            return ast.get_source_segment(src, node)
    raise KeyError(cls)


def _method_source(path, cls, name):
    src = _src(path)
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == name:
                    lines = src.splitlines()[item.lineno - 1 - len(item.decorator_list):item.end_lineno]
                    return textwrap.dedent("\n".join(lines))
    raise KeyError((cls, name))


def _load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------
# graph builder
# ------------------------------------------------------------------------------------------
class _RecGraph:
    def __init__(self, edges):
        self.edges = edges
        self._nodes = {"tgt": types.SimpleNamespace(data={}), "ntgt": types.SimpleNamespace(data={})}

    @property
    def nodes(self):
        return self._nodes


def _graph_builder(fix_q1):
    ns = {"torch": torch, "np": np, "lru_cache": lru_cache, "Dict": Dict, "Tuple": Tuple, "List": List,
          "dgl": types.SimpleNamespace(heterograph=lambda e: _RecGraph(e))}
    body = []
    for m in ("new_build_graph", "deprecated_build_graph", "build_ntgt_edges", "auto_regressive_edges"):
        s = _method_source("fairseq/data/token_block_dataset.py", "GraphTokenBlockDataset", m)
        if fix_q1 and m == "new_build_graph":
            assert "len(self.neighbor_offsets.shape[0])" in s
            s = s.replace("len(self.neighbor_offsets.shape[0])", "len(self.neighbor_tokens)")
        body.append(textwrap.indent(s, "    "))
    exec("class G:\n" + "\n".join(body), ns)
    return ns["G"]


def make_graph_case(name, L, k, n_d, cl, cr, invalid_ctx, intra_ctx, M, seed, stress):
    rng = np.random.RandomState(seed)
    G = _graph_builder(fix_q1=cr > 0)
    g = G()
    start = int(rng.randint(0, 1000))
    offsets = np.arange(start, start + L, dtype=np.int64)
    nbr = rng.randint(0, n_d, size=(L, k)).astype(np.int64)
    if stress:
        nbr[rng.rand(L, k) < 0.1] = -1
        edge = rng.rand(L, k) < 0.15
        nbr[edge] = rng.choice([0, 1, 2, n_d - 1, n_d - 2, n_d - 3], size=int(edge.sum()))
        nbr[L // 2] = -1                 # a token with no neighbour at all
    codes = rng.randint(0, 256, size=(n_d, M)).astype(np.uint8)
    vals = rng.randint(4, 1000, size=(n_d, 1)).astype(np.int32)
    g.neighbor_offsets = nbr  # only .shape is read (buggy line) -- kept for fidelity
    g.neighbor_tokens = vals
    g.quant_neighbor_feats = codes
    g.left_neighbor_context, g.right_neighbor_context = cl, cr
    g.invalid_neighbor_context = invalid_ctx
    g.max_intra_context = intra_ctx
    src_tok = torch.zeros(L, dtype=torch.long)
    graph = g.new_build_graph(src_tok, offsets, nbr, src_tok)
    e = graph.edges
    out = dict(
        L=L, k=k, n_d=n_d, cl=cl, cr=cr, invalid_ctx=invalid_ctx, intra_ctx=intra_ctx,
        offsets=offsets, nbr=nbr, codes=codes, vals=vals,
        tt_src=e[("tgt", "intra", "tgt")][0].numpy(), tt_dst=e[("tgt", "intra", "tgt")][1].numpy(),
        inter_src=e[("ntgt", "inter", "tgt")][0].numpy(), inter_dst=e[("ntgt", "inter", "tgt")][1].numpy(),
        nn_src=e[("ntgt", "intra", "ntgt")][0].numpy(), nn_dst=e[("ntgt", "intra", "ntgt")][1].numpy(),
        ntgt_codes=graph.nodes["ntgt"].data["h"].numpy(),
        ntgt_labels=graph.nodes["ntgt"].data["labels"].numpy(),
    )
    np.savez_compressed(os.path.join(OUT, f"graph_{name}.npz"), **out)
    print(f"graph_{name}: ntgt={len(out['ntgt_labels'])} E_nn={len(out['nn_src'])}")


def make_dedup_case(name, L, k, n_d, cl, cr, invalid_ctx, intra_ctx, M, seed):
    """`--deprecated` graph build (token_block_dataset.py:414-479): runs unmodified (it clips the right context with
    len(self.neighbor_tokens), :458).  Neighbour ids are drawn from a narrow range so that rows repeat and runs of
    consecutive rows form across neighbours and tokens."""
    rng = np.random.RandomState(seed)
    g = _graph_builder(False)()
    start = int(rng.randint(0, 1000))
    offsets = np.arange(start, start + L, dtype=np.int64)
    nbr = rng.randint(0, min(n_d, 6 * L), size=(L, k)).astype(np.int64)
    nbr[rng.rand(L, k) < 0.1] = -1
    edge = rng.rand(L, k) < 0.1
    nbr[edge] = rng.choice([0, 1, n_d - 1, n_d - 2], size=int(edge.sum()))
    nbr[L // 3] = -1
    nbr[1, :] = nbr[0, :]                     # a token repeating its predecessor's neighbours
    if k > 1:
        nbr[2, 1] = nbr[2, 0]                 # the same row twice for one token
    codes = rng.randint(0, 256, size=(n_d, M)).astype(np.uint8)
    vals = rng.randint(4, 1000, size=(n_d, 1)).astype(np.int32)
    g.neighbor_offsets, g.neighbor_tokens, g.quant_neighbor_feats = nbr, vals, codes
    g.left_neighbor_context, g.right_neighbor_context = cl, cr
    g.invalid_neighbor_context, g.max_intra_context = invalid_ctx, intra_ctx
    z = torch.zeros(L, dtype=torch.long)
    graph = g.deprecated_build_graph(z, offsets, nbr, z)
    e = graph.edges
    out = dict(L=L, k=k, n_d=n_d, cl=cl, cr=cr, invalid_ctx=invalid_ctx, intra_ctx=intra_ctx, offsets=offsets, nbr=nbr, codes=codes,
               tt_src=e[("tgt", "intra", "tgt")][0].numpy(), tt_dst=e[("tgt", "intra", "tgt")][1].numpy(),
               inter_src=e[("ntgt", "inter", "tgt")][0].numpy(), inter_dst=e[("ntgt", "inter", "tgt")][1].numpy(),
               nn_src=e[("ntgt", "intra", "ntgt")][0].numpy(), nn_dst=e[("ntgt", "intra", "ntgt")][1].numpy(),
               ntgt_codes=graph.nodes["ntgt"].data["h"].numpy())
    np.savez_compressed(os.path.join(OUT, f"dedup_{name}.npz"), **out)
    print(f"dedup_{name}: ntgt={len(out['ntgt_codes'])} (of {int((nbr >= 0).sum())} neighbours) E_nn={len(out['nn_src'])}")


def make_edges_doctest():
    G = _graph_builder(False)
    a = G.build_ntgt_edges({0: 0, 1: 1, 2: 2, 12: 3, 13: 4}, 3)
    b = G.build_ntgt_edges({0: 0, 1: 1, 2: 2, 12: 3, 13: 4}, 0)
    assert a == ([0, 0, 1, 0, 1, 2, 3, 3, 4], [0, 1, 1, 2, 2, 2, 3, 4, 4])   # token_block_dataset.py:552
    assert b == ([0, 1, 2, 3, 4], [0, 1, 2, 3, 4])                             # :554
    c = G.build_ntgt_edges({7: 3, 5: 1, 6: 0, 9: 2}, 1, bidirect=True)
    u, v = G.auto_regressive_edges(6, 0)
    u3, v3 = G.auto_regressive_edges(6, 3)
    np.savez_compressed(os.path.join(OUT, "edges_misc.npz"),
                        bidirect_src=np.array(c[0]), bidirect_dst=np.array(c[1]),
                        ar6_u=u.numpy(), ar6_v=v.numpy(), ar6c3_u=u3.numpy(), ar6c3_v=v3.numpy())


# ------------------------------------------------------------------------------------------
# adaptive softmax
# ------------------------------------------------------------------------------------------
def _ref_adaptive():
    asm = _load_by_path("ref_adaptive_softmax", "fairseq/modules/adaptive_softmax.py")
    ain = _load_by_path("ref_adaptive_input", "fairseq/modules/adaptive_input.py")
    return asm, ain


def make_adaptive_case(name, V, d, cutoff, tied, T, seed, tie_proj=None):
    asm, ain = _ref_adaptive()
    torch.manual_seed(seed)
    emb = None
    if tied:
        emb = ain.AdaptiveInput(V, 1, d, 4, d, list(cutoff))
    tie_proj = tied if tie_proj is None else tie_proj
    m = asm.AdaptiveSoftmax(V, d, list(cutoff), dropout=0.0, factor=4.0, adaptive_inputs=emb, tie_proj=tie_proj)
    m.eval()
    x = torch.randn(1, T, d)
    g = torch.Generator().manual_seed(seed)
    target = torch.randint(4, V, (1, T), generator=g)
    target[0, :4] = torch.tensor([cutoff[0] - 1, cutoff[0], cutoff[1] - 1, V - 1])[:4]
    with torch.no_grad():
        lp_t = m.get_log_prob(x, target)
        lp_f = m.get_log_prob(x, None)
    out = {"x": x.numpy(), "target": target.numpy(), "cutoff": np.array(m.cutoff),
           "lp_target_mode_at_target": lp_t.gather(2, target.unsqueeze(-1)).squeeze(-1).numpy(),
           "lp_full": lp_f.numpy(), "tied": np.array(int(tied)), "tie_proj": np.array(int(tie_proj))}
    for k_, v_ in m.state_dict().items():
        out["sd." + k_] = v_.numpy()
    if tied:
        for k_, v_ in emb.state_dict().items():
            out["emb." + k_] = v_.numpy()
    np.savez_compressed(os.path.join(OUT, f"adaptive_{name}.npz"), **out)
    print(f"adaptive_{name}: keys={[k for k in out if k.startswith('sd.')]}")


# ------------------------------------------------------------------------------------------
# PQ codec
# ------------------------------------------------------------------------------------------
def make_pq_case(name, n, M, dsub, with_b, seed):
    rng = np.random.RandomState(seed)
    d = M * dsub
    cen = rng.randn(M, 256, dsub).astype(np.float32)
    A = np.linalg.qr(rng.randn(d, d))[0].astype(np.float32)
    b = rng.randn(d).astype(np.float32) if with_b else np.zeros(0, np.float32)
    codes = rng.randint(0, 256, size=(n, M)).astype(np.uint8)
    ns = {"np": np, "torch": torch}
    exec(_method_source("knn/pq_wrapper.py", "NumpyPQCodec", "decode").replace("def decode", "def np_decode"), ns)
    exec(_method_source("knn/pq_wrapper.py", "TorchPQCodec", "decode").replace("def decode", "def th_decode"), ns)
    fake = types.SimpleNamespace(centroids=cen, pre=(A, b), centroids_torch=torch.from_numpy(cen),
                                 A=torch.from_numpy(A), b=torch.from_numpy(b))
    x_np = ns["np_decode"](fake, codes)
    x_th = ns["th_decode"](fake, torch.from_numpy(codes)).numpy()
    fake2 = types.SimpleNamespace(centroids=cen, pre=None, centroids_torch=torch.from_numpy(cen))
    x_raw = ns["th_decode"](fake2, torch.from_numpy(codes)).numpy()
    np.savez_compressed(os.path.join(OUT, f"pq_{name}.npz"), cen=cen, A=A, b=b, codes=codes,
                        x_numpy=x_np, x_torch=x_th, x_nopre=x_raw)
    print(f"pq_{name}: max|np-torch|={np.abs(x_np - x_th).max():.2e}")


# ------------------------------------------------------------------------------------------
# kNN-LM probability + scorer
# ------------------------------------------------------------------------------------------
def _ref_get_knn_prob():
    ns = {"np": np, "torch": torch, "F": torch.nn.functional, "Union": Union, "Tuple": Tuple}
    exec(_method_source("knn/knn_model.py", "KNNModel", "get_knn_prob"), ns)
    return ns["get_knn_prob"]


class _FakeKNN:
    """Stands in for KNNModel; only get_knns (the faiss search) is replaced."""

    def __init__(self, dists, knns, vals, vocab, metric):
        self._d, self._k = dists, knns
        self.vals = vals
        self.vocab_size = vocab
        self.metric_type = metric
        self.index_file = "faiss_store.ip"
        self.k = knns.shape[1]
        self.data_store = types.SimpleNamespace(val_size=1)
        self.keys = None

    def get_knns(self, queries, k=0):
        return self._d.copy(), self._k.copy()

    get_knn_prob = _ref_get_knn_prob()


def make_knn_case(name, T, knn, n_d, V, temp, metric, seed, with_missing):
    rng = np.random.RandomState(seed)
    dists = rng.randn(T, knn).astype(np.float32)
    ids = rng.randint(0, n_d, size=(T, knn)).astype(np.int64)
    if with_missing:
        ids[rng.rand(T, knn) < 0.05] = -1
    vals = rng.randint(4, V, size=(n_d,)).astype(np.int32)
    targets = torch.from_numpy(rng.randint(4, V, size=(T,)).astype(np.int64))
    # make some targets actually hit
    for t in range(0, T, 2):
        j = rng.randint(0, knn)
        if ids[t, j] >= 0:
            targets[t] = int(vals[ids[t, j]])
    m = _FakeKNN(dists, ids, vals, V, metric)
    q = torch.zeros(T, 4)
    p_t, recall = m.get_knn_prob(q, t=temp, targets=targets, return_recall=True)
    p_full = m.get_knn_prob(q, t=temp)
    np.savez_compressed(os.path.join(OUT, f"knn_{name}.npz"), dists=dists, ids=ids, vals=vals,
                        targets=targets.numpy(), temp=np.float64(temp), metric=np.array(metric), V=V,
                        p_target=p_t.numpy(), recall=recall.numpy(), p_full=p_full.numpy())
    print(f"knn_{name}: mean p={float(p_t.mean()):.4f}")


def make_knn_recompute_case(name, T, knn, n_d, d, V, temp, metric, index_file, seed, fp16_keys):
    """metric_type l2 / ip: similarities recomputed from keys[knns] (knn_model.py:159-177), cosine index included."""
    rng = np.random.RandomState(seed)
    keys = rng.randn(n_d, d).astype(np.float16 if fp16_keys else np.float32)
    queries = rng.randn(T, d).astype(np.float32)
    ids = rng.randint(0, n_d, size=(T, knn)).astype(np.int64)
    ids[rng.rand(T, knn) < 0.05] = -1
    vals = rng.randint(4, V, size=(n_d,)).astype(np.int32)
    targets = torch.from_numpy(rng.randint(4, V, size=(T,)).astype(np.int64))
    for t in range(0, T, 2):
        j = rng.randint(0, knn)
        if ids[t, j] >= 0:
            targets[t] = int(vals[ids[t, j]])
    m = _FakeKNN(np.zeros((T, knn), np.float32), ids, vals, V, metric)
    m.index_file, m.keys = index_file, keys
    q = torch.from_numpy(queries)
    p_t, recall = m.get_knn_prob(q, t=temp, targets=targets, return_recall=True)
    _, sims, _ = m.get_knn_prob(q, t=temp, return_knn=True)
    np.savez_compressed(os.path.join(OUT, f"knn_{name}.npz"), keys=keys, queries=queries, ids=ids, vals=vals,
                        targets=targets.numpy(), temp=np.float64(temp), metric=np.array(metric),
                        cosine=np.bool_("cosine" in index_file), V=V, p_target=p_t.numpy(), recall=recall.numpy(),
                        sims=sims.numpy())
    print(f"knn_{name}: mean p={float(p_t.mean()):.4f}")


def _ref_slice_indices():
    """fairseq/data/token_block_utils_fast.pyx:22-105 is Cython whose body is plain Python once the C declarations are
    stripped: drop `cdef` lines / decorators / casts, turn the two `cdef`/`cpdef` signatures into `def`s, and exec it."""
    import re
    from itertools import chain
    from math import ceil
    src = open(os.path.join(REF, "fairseq/data/token_block_utils_fast.pyx")).read()
    start = src.index("cdef np.ndarray[DTYPE_t, ndim=2] _get_slice_indices_none_mode")
    end = src.index("cpdef np.ndarray[DTYPE_t, ndim=2] _get_block_to_dataset_index_fast")
    out = []
    for line in src[start:end].splitlines():
        st = line.strip()
        if st.startswith("@cython"):
            continue
        m = re.match(r"^c?p?def np\.ndarray\[DTYPE_t, ndim=\d\] (\w+)\((.*)\):$", line)
        if m:
            args = ", ".join(a.strip().split(" ")[-1] for a in re.sub(r"\[[^\]]*\]", "", m.group(2)).split(","))
            out.append(f"def {m.group(1)}({args}):")
            continue
        if st.startswith("cdef "):
            d = re.match(r"^cdef\s+[\w\.]+(?:\[[^\]]*\])?\s+(\w+)\s*=\s*(.*)$", st)
            if d:              # typed declaration with an initialiser -> plain assignment
                out.append(line[:len(line) - len(line.lstrip())] + d.group(1) + " = " + d.group(2))
            continue
        out.append(line)
    code = "\n".join(out).replace("<DTYPE_t> ", "").replace("<double> ", "")
    ns = {"np": np, "DTYPE": np.int64, "chain": chain, "ceil": ceil}
    exec(code, ns)
    return ns["_get_slice_indices_fast"]


def make_slice_cases():
    """--sample-break-mode none / complete / complete_doc / eos block boundaries for a few sentence-length vectors."""
    f = _ref_slice_indices()
    rng = np.random.RandomState(0)
    out = {}
    cases = {"short": rng.randint(1, 12, size=40), "long": rng.randint(1, 60, size=25), "one": np.array([7]),
             "docs": np.array([5, 9, 1, 4, 4, 1, 1, 30, 2, 1, 8])}
    for name, sizes in cases.items():
        sizes = sizes.astype(np.int64)
        out[f"{name}.sizes"] = sizes
        for mode in ("none", "complete", "complete_doc", "eos"):
            for bs in (16, 50):
                out[f"{name}.{mode}.{bs}"] = np.asarray(f(sizes, mode, bs, 1), dtype=np.int64).reshape(-1, 2)
    np.savez_compressed(os.path.join(OUT, "slices.npz"), **out)
    print("slices:", len(out), "arrays")


def make_scorer_case(name, B, L, d, V, cutoff, knn, lmbda, temp, seed):
    asm, _ = _ref_adaptive()
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    soft = asm.AdaptiveSoftmax(V, d, list(cutoff), dropout=0.0, factor=4.0).eval()
    feats = torch.randn(B, L, d)
    target = torch.from_numpy(rng.randint(4, V, size=(B, L)).astype(np.int64))
    n_d = 5000
    dists = rng.randn(B * L, knn).astype(np.float32)
    ids = rng.randint(0, n_d, size=(B * L, knn)).astype(np.int64)
    vals = rng.randint(4, V, size=(n_d,)).astype(np.int32)
    knn_model = _FakeKNN(dists, ids, vals, V, "do_not_recomp_ip")

    class Model:
        def eval(self):
            return self

        def __call__(self, **kw):
            return feats, {"inner_states": [feats.transpose(0, 1)]}

        def get_normalized_probs(self, net_output, log_probs, sample):
            out = soft.get_log_prob(net_output[0], sample["target"])
            return out if log_probs else out.exp_()

    utils = types.SimpleNamespace(strip_pad=lambda t, pad: t[t.ne(pad)])
    ns = {"torch": torch, "sys": sys, "np": np, "utils": utils, "KNNModel": object}
    exec(_class_source("fairseq/sequence_scorer.py", "SequenceScorer"), ns)
    d_ = types.SimpleNamespace(pad=lambda: 1, eos=lambda: 2)
    args = types.SimpleNamespace(lmbda=lmbda, knn_keytype=None)
    sc = ns["SequenceScorer"](d_, softmax_batch=10 ** 9, args=args)
    start = torch.zeros(B, 1, dtype=torch.long)
    start[-1, 0] = 3   # gcn-context-window style loss_start_idx on the last sample
    sample = {"net_input": {}, "target": target, "start_indices": start}
    hy_lm = ns["SequenceScorer"](d_, softmax_batch=10 ** 9, args=types.SimpleNamespace(lmbda=0.0, knn_keytype=None)
                                 ).generate([Model()], dict(sample))
    out = {"feats": feats.numpy(), "target": target.numpy(), "dists": dists, "ids": ids, "vals": vals,
           "lmbda": np.float64(lmbda), "temp": np.float64(temp), "cutoff": np.array(soft.cutoff),
           "start_indices": start.numpy()}
    for k_, v_ in soft.state_dict().items():
        out["sd." + k_] = v_.numpy()
    for i, h in enumerate(hy_lm):
        out[f"lm_pos_{i}"] = h[0]["positional_scores"].numpy()
    if B == 1:   # the kNN branch of the reference is only index-consistent for B == 1
        hy = sc.generate([Model()], dict(sample), knn_dstore=knn_model, temperature=temp)
        for i, h in enumerate(hy):
            out[f"knn_pos_{i}"] = h[0]["positional_scores"].numpy()
            out[f"knn_recall_{i}"] = h[0]["knn_recall"].numpy()
            out[f"knn_score_{i}"] = h[0]["score"].numpy()
    np.savez_compressed(os.path.join(OUT, f"scorer_{name}.npz"), **out)
    print(f"scorer_{name}: done")


# ------------------------------------------------------------------------------------------
# HGT under a minimal DGL stub (DGL's documented semantics, restated)
# ------------------------------------------------------------------------------------------
class _StubFn:
    @staticmethod
    def v_dot_u(a, b, out):
        return ("v_dot_u", a, b, out)

    @staticmethod
    def u_mul_e(a, b, out):
        return ("u_mul_e", a, b, out)

    @staticmethod
    def sum(msg, out):
        return ("sum", msg, out)


class _SubGraph:
    def __init__(self, g, cet):
        self.g, self.cet = g, cet
        self.src, self.dst = g.edges_of[cet]
        self.srcdata, self.dstdata = {}, {}
        self.edata = g.edata_of[cet]
        self.n_dst = g.num_nodes(cet[2])

    def apply_edges(self, f, edges=None):
        kind, a, b, out = f
        assert kind == "v_dot_u"   # out[e] = <dst[a][v_e], src[b][u_e]> over the last dim, keepdim
        val = (self.dstdata[a][self.dst] * self.srcdata[b][self.src]).sum(-1, keepdim=True)
        if edges is None:
            self.edata[out] = val
        else:                      # a subset of the edges: a new edge field starts as zeros (DGL's default initialiser)
            if out not in self.edata:
                self.edata[out] = torch.zeros_like(val)
            self.edata[out][edges] = val[edges]
        self.g.srcdata_of[self.cet] = self.srcdata

    def edges(self, form="uv", order="eid", etype=None):
        assert form == "uv" and order == "eid" and (etype is None or tuple(etype) == tuple(self.cet))
        return self.src, self.dst


def _stub_edge_softmax(sub, score, norm_by="dst"):
    assert norm_by == "dst"
    H = score.shape[1:]
    mx = torch.full((sub.n_dst,) + H, -float("inf"), dtype=score.dtype)
    mx = mx.scatter_reduce(0, sub.dst.view(-1, *[1] * len(H)).expand_as(score), score, "amax", include_self=True)
    ex = torch.exp(score - mx[sub.dst])
    den = torch.zeros((sub.n_dst,) + H, dtype=score.dtype).index_add_(0, sub.dst, ex)
    return ex / den[sub.dst]


class _StubGraph:
    def __init__(self, edges, num_nodes):
        self.edges_of = {k: (torch.as_tensor(s), torch.as_tensor(d)) for k, (s, d) in edges.items()}
        self._n = num_nodes
        self.canonical_etypes = list(edges.keys())
        self.ntypes = sorted(num_nodes.keys())   # DGL sorts node types alphabetically
        self.nodes = {nt: types.SimpleNamespace(data={}) for nt in num_nodes}
        self.edata_of = {k: {} for k in edges}
        self.srcdata_of = {}
        self._subs = {}

    def num_nodes(self, nt):
        return self._n[nt]

    def local_scope(self):
        import contextlib

        @contextlib.contextmanager
        def scope():               # feature writes inside the scope do not outlive it (DGLGraph.local_scope)
            saved = {nt: dict(ns.data) for nt, ns in self.nodes.items()}
            try:
                yield
            finally:
                for nt, ns in self.nodes.items():
                    ns.data = saved[nt]
                self.edata_of = {k: {} for k in self.edges_of}
                self.srcdata_of, self._subs = {}, {}
        return scope()

    def __getitem__(self, cet):
        if cet not in self._subs:
            self._subs[cet] = _SubGraph(self, cet)
        return self._subs[cet]

    def multi_update_all(self, funcs, cross_reducer):
        assert cross_reducer == "mean"
        per_dst = {}
        for cet, (mf, rf) in funcs.items():
            sub = self[cet]
            _, uname, ename, _ = mf
            _, _, oname = rf
            m = sub.srcdata[uname][sub.src] * sub.edata[ename]
            agg = torch.zeros((sub.n_dst,) + m.shape[1:], dtype=m.dtype).index_add_(0, sub.dst, m)
            per_dst.setdefault((cet[2], oname), []).append(agg)   # zero rows for 0-in-degree dst
        for (nt, oname), lst in per_dst.items():
            self.nodes[nt].data[oname] = torch.stack(lst, 0).mean(0)

    def update_all(self, mf, rf, etype):
        sub = self[tuple(etype)]
        _, uname, ename, _ = mf
        _, _, oname = rf
        m = sub.srcdata[uname][sub.src] * sub.edata[ename]
        self.nodes[etype[2]].data[oname] = torch.zeros((sub.n_dst,) + m.shape[1:], dtype=m.dtype).index_add_(0, sub.dst, m)


def _ref_hgt(patch=None):
    dgl = types.ModuleType("dgl")
    dgl.DGLHeteroGraph = _StubGraph
    dgl.DGLGraph = _StubGraph
    dgl_fn = types.ModuleType("dgl.function")
    for n in ("v_dot_u", "u_mul_e", "sum"):
        setattr(dgl_fn, n, getattr(_StubFn, n))
    dgl_ops = types.ModuleType("dgl.ops")
    dgl_ops.edge_softmax = _stub_edge_softmax
    dgl.function, dgl.ops = dgl_fn, dgl_ops
    inc = _load_by_path("ref_incremental_decoding_utils", "fairseq/incremental_decoding_utils.py")   # pure torch: the real mixin
    fs = types.ModuleType("fairseq")
    saved = {k: sys.modules.get(k) for k in ("dgl", "dgl.function", "dgl.ops", "fairseq",
                                              "fairseq.incremental_decoding_utils")}
    sys.modules.update({"dgl": dgl, "dgl.function": dgl_fn, "dgl.ops": dgl_ops, "fairseq": fs,
                        "fairseq.incremental_decoding_utils": inc})
    try:
        src = _src("fairseq/models/hgt.py")
        src = src[:src.index("if __name__ == '__main__':")]
        if patch is not None:
            src = patch(src)
        mod = types.ModuleType("ref_hgt")
        exec(compile(src, "ref_hgt.py", "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def make_hgt_case(name, B, L, k, cl, cr, d, H, n_layers, seed, stress=True, hidden=None, out_dim=None):
    sys.path.insert(0, os.path.join(OUT, "..", ".."))
    from oracle import graph_oracle as go
    rng = np.random.RandomState(seed)
    n_d = 500
    nbr = rng.randint(0, n_d, size=(B, L, k)).astype(np.int64)
    if stress:
        nbr[rng.rand(B, L, k) < 0.15] = -1
        nbr[0, 1] = -1
        nbr[0, 2, 0] = 0
        nbr[0, 2, 1] = n_d - 1
    offsets = np.arange(B * L, dtype=np.int64).reshape(B, L)
    # edge lists straight from the REFERENCE builder (not the oracle), batched like dgl.batch
    G = _graph_builder(fix_q1=cr > 0)
    graphs = []
    for b in range(B):
        g = G()
        g.neighbor_offsets = nbr[b]
        g.neighbor_tokens = np.zeros((n_d, 1), np.int32)
        g.quant_neighbor_feats = None
        g.left_neighbor_context, g.right_neighbor_context = cl, cr
        g.invalid_neighbor_context, g.max_intra_context = 0, 0
        z = torch.zeros(L, dtype=torch.long)
        e = g.new_build_graph(z, offsets[b], nbr[b], z).edges
        n_ntgt = int(g.new_build_graph(z, offsets[b], nbr[b], z).nodes["ntgt"].data["labels"].shape[0])
        graphs.append({"n_tgt": L, "n_ntgt": n_ntgt,
                       "tt": tuple(t.numpy() for t in e[("tgt", "intra", "tgt")]),
                       "inter": tuple(t.numpy() for t in e[("ntgt", "inter", "tgt")]),
                       "nn": tuple(t.numpy() for t in e[("ntgt", "intra", "ntgt")])})
    bg = go.batch_graphs(graphs)
    hgt = _ref_hgt()
    torch.manual_seed(seed)
    model = hgt.HGT(ntype2idx={"tgt": 0, "ntgt": 1}, etype2idx={"intra": 0, "inter": 1}, in_dim=d,
                    hidden_dim=hidden or d, out_dim=out_dim or d, n_layers=n_layers, n_heads=H, dropout=0.0,
                    attn_drop=0.0).eval()
    with torch.no_grad():
        for p in model.parameters():          # make every parameter non-trivial (biases, pri, LN affine)
            if p.dim() <= 2 and p.shape[-1] != d or p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    etypes = [("tgt", "intra", "tgt"), ("ntgt", "inter", "tgt"), ("ntgt", "intra", "ntgt")]
    sg = _StubGraph({etypes[0]: bg["tt"], etypes[1]: bg["inter"], etypes[2]: bg["nn"]},
                    {"tgt": bg["n_tgt"], "ntgt": bg["n_ntgt"]})
    h_tgt = torch.randn(bg["n_tgt"], d)
    h_ntgt = torch.randn(bg["n_ntgt"], d)
    sg.nodes["ntgt"].data["h"] = h_ntgt
    with torch.no_grad():
        out = model(sg, features={"tgt": h_tgt}, etypes=etypes)
    res = {"nbr": nbr, "offsets": offsets, "n_d": n_d, "cl": cl, "cr": cr, "H": H, "n_layers": n_layers,
           "h_tgt": h_tgt.numpy(), "h_ntgt": h_ntgt.numpy(),
           "out_tgt": out["tgt"].numpy(), "out_ntgt": out["ntgt"].numpy()}
    for k_, v_ in model.state_dict().items():
        res["sd." + k_] = v_.numpy()
    np.savez_compressed(os.path.join(OUT, f"hgt_{name}.npz"), **res)
    print(f"hgt_{name}: n_tgt={bg['n_tgt']} n_ntgt={bg['n_ntgt']} |out|={float(out['tgt'].abs().mean()):.3f}")



# ------------------------------------------------------------------------------------------
# HGT on a general heterograph, two_stream, incremental infer() (hgt.py:81-297,324-330,360-394)
# ------------------------------------------------------------------------------------------
def _perturb(model, d):
    with torch.no_grad():
        for p in model.parameters():          # make every parameter non-trivial (biases, pri, LN affine)
            if p.dim() <= 2 and p.shape[-1] != d or p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))


def _save_hgt(name, model, res):
    for k_, v_ in model.state_dict().items():
        res["sd." + k_] = v_.numpy()
    np.savez_compressed(os.path.join(OUT, f"hgt_{name}.npz"), **res)


def _rand_edges(rng, n_src, n_dst, n_edges, skip_dst=()):
    """Random COO list (duplicates allowed, as DGL multigraphs allow them); destinations in `skip_dst` get no in-edge."""
    dst = rng.randint(0, n_dst, size=n_edges)
    dst = dst[~np.isin(dst, list(skip_dst))]
    return rng.randint(0, n_src, size=dst.shape[0]).astype(np.int64), dst.astype(np.int64)


def make_hetero_case(name, d, H, n_layers, seed):
    """The reference module's own self-test topology (hgt.py:516-552: four node types, ten canonical edge types), random sizes."""
    rng = np.random.RandomState(seed)
    n = {"src": 9, "nsrc": 14, "tgt": 7, "ntgt": 11}
    cets = [("src", "intra", "src"), ("src", "intra", "tgt"), ("tgt", "intra", "tgt"), ("nsrc", "inter", "src"),
            ("src", "inter", "nsrc"), ("ntgt", "inter", "tgt"), ("ntgt", "intra", "ntgt"), ("ntgt", "intra", "nsrc"),
            ("nsrc", "intra", "ntgt"), ("nsrc", "intra", "nsrc")]
    edges = {}
    for i, (s_, r_, t_) in enumerate(cets):
        if (s_, r_, t_) == ("tgt", "intra", "tgt"):
            u, v = np.triu_indices(n["tgt"])
            edges[(s_, r_, t_)] = (u.astype(np.int64), v.astype(np.int64))
        else:
            edges[(s_, r_, t_)] = _rand_edges(rng, n[s_], n[t_], 3 * n[t_], skip_dst=(1,) if i % 2 else ())
    hgt = _ref_hgt()
    torch.manual_seed(seed)
    ntype2idx = {"src": 0, "nsrc": 1, "tgt": 2, "ntgt": 3}
    model = hgt.HGT(ntype2idx=ntype2idx, etype2idx={"intra": 0, "inter": 1}, in_dim=d, hidden_dim=d, out_dim=d,
                    n_layers=n_layers, n_heads=H, dropout=0.0, attn_drop=0.0).eval()
    _perturb(model, d)
    sg = _StubGraph(edges, n)
    feats = {nt: torch.randn(n[nt], d) for nt in n}
    for nt in n:
        sg.nodes[nt].data["h"] = feats[nt]
    with torch.no_grad():
        out = model(sg, features={"tgt": feats["tgt"]})          # the other types come from G.nodes[.].data["h"] (hgt.py:501-503)
    res = {"H": H, "n_layers": n_layers, "ntypes": np.array(list(n.keys())), "num_nodes": np.array(list(n.values())),
           "cets": np.array(["|".join(c) for c in cets])}
    for c in cets:
        res["src." + "|".join(c)], res["dst." + "|".join(c)] = edges[c]
    for nt in n:
        res["h." + nt], res["out." + nt] = feats[nt].numpy(), out[nt].numpy()
    _save_hgt(name, model, res)
    print(f"hgt_{name}: |out tgt|={float(out['tgt'].abs().mean()):.3f}")


TWO_STREAM_ANCHOR = "                sub_graph.srcdata['k'] = k\n                sub_graph.dstdata['q'] = q\n                sub_graph.srcdata[f'v_{srctype}_{etype}_{dsttype}'] = v\n\n                sub_graph.apply_edges(fn.v_dot_u('q', 'k', 't'))\n                attn_score = sub_graph.edata.pop('t').sum(-1) * relation_pri / self.sqrt_dk\n                attn_score = self.attn_drop(edge_softmax(sub_graph, attn_score, norm_by='dst'))\n\n                sub_graph.edata['t'] = attn_score.unsqueeze(-1)\n\n                if self.two_stream and"


def _patch_two_stream(src):
    """SURVEY.md-style quirk: the query stream's self-loop score reads srcdata['k_tilde'] (hgt.py:376), which the
    reference never assigns -- as written, two_stream raises on any tgt-intra-tgt edge set.  The fixture applies the evident intent
    as a one-statement fix: k_tilde = the relation-transformed keys of the query stream (tgt_tilde_k, hgt.py:330)."""
    assert src.count(TWO_STREAM_ANCHOR) == 1
    fix = ("                if self.two_stream and (srctype, etype, dsttype) == ('tgt', 'intra', 'tgt'):\n"
           "                    sub_graph.srcdata['k_tilde'] = torch.einsum('bij,ijk->bik', G.nodes['tgt'].data['tgt_tilde_k'], relation_att)\n")
    return src.replace(TWO_STREAM_ANCHOR, fix + TWO_STREAM_ANCHOR)


def make_two_stream_case(name, d, H, n_layers, seed):
    rng = np.random.RandomState(seed)
    n = {"src": 10, "tgt": 8, "ntgt": 12}
    u, v = np.triu_indices(n["tgt"])
    su, sv = np.meshgrid(np.arange(n["src"]), np.arange(n["tgt"]), indexing="ij")
    edges = {("src", "intra", "src"): _rand_edges(rng, n["src"], n["src"], 30),
             ("src", "intra", "tgt"): (su.reshape(-1).astype(np.int64), sv.reshape(-1).astype(np.int64)),
             ("tgt", "intra", "tgt"): (u.astype(np.int64), v.astype(np.int64)),
             ("ntgt", "inter", "tgt"): _rand_edges(rng, n["ntgt"], n["tgt"], 20, skip_dst=(2,)),
             ("ntgt", "intra", "ntgt"): _rand_edges(rng, n["ntgt"], n["ntgt"], 30)}
    ntype2idx = {"src": 0, "tgt": 1, "ntgt": 2}
    feats = {nt: torch.randn(n[nt], d, generator=torch.Generator().manual_seed(seed + 7 + i)) for i, nt in enumerate(n)}

    def run(patch):
        hgt = _ref_hgt(patch)
        torch.manual_seed(seed)
        model = hgt.HGT(ntype2idx=ntype2idx, etype2idx={"intra": 0, "inter": 1}, in_dim=d, hidden_dim=d, out_dim=d,
                        n_layers=n_layers, n_heads=H, dropout=0.0, attn_drop=0.0, two_stream=True).eval()
        _perturb(model, d)
        sg = _StubGraph(edges, n)
        for nt in n:
            sg.nodes[nt].data["h"] = feats[nt]
        with torch.no_grad():
            return model, model(sg, features=dict(feats))

    try:
        run(None)
        raise SystemExit("the unmodified two_stream path was expected to raise (k_tilde never assigned)")
    except KeyError as e:
        print(f"hgt_{name}: unmodified reference raises KeyError({e}) as documented")
    model, out = run(_patch_two_stream)
    res = {"H": H, "n_layers": n_layers, "ntypes": np.array(list(n.keys())), "num_nodes": np.array(list(n.values())),
           "cets": np.array(["|".join(c) for c in edges])}
    for c in edges:
        res["src." + "|".join(c)], res["dst." + "|".join(c)] = edges[c]
    for nt in n:
        res["h." + nt] = feats[nt].numpy()
    for nt in out:
        res["out." + nt] = out[nt].numpy()
    _save_hgt(name, model, res)
    print(f"hgt_{name}: keys {sorted(out)} |out tgt_tilde|={float(out['tgt_tilde'].abs().mean()):.3f}")


def make_infer_case(name, bsz, steps, n_per, d, H, n_layers, seed, reorder_at=None):
    """HGTLayer.infer (hgt.py:81-297) driven through HGT.forward(incremental_state=...): `steps` decoding steps over a batched
    graph of bsz blocks with max_len = 512 tgt nodes each (the constant at :93), causal tgt-intra-tgt edges, ntgt -> tgt inter edges
    towards the first `steps` positions and random ntgt-intra-ntgt edges.  h['tgt'] is the current step's [bsz, d] feature."""
    max_len = 512
    rng = np.random.RandomState(seed)
    n = {"tgt": bsz * max_len, "ntgt": bsz * n_per}
    u, v = np.triu_indices(max_len)
    tt = (np.concatenate([u + b * max_len for b in range(bsz)]).astype(np.int64),
          np.concatenate([v + b * max_len for b in range(bsz)]).astype(np.int64))
    i_src = np.arange(n["ntgt"], dtype=np.int64)
    i_dst = (i_src // n_per) * max_len + rng.randint(0, steps, size=n["ntgt"])
    i_dst[i_dst % max_len == 1] += 1                                      # position 1 of every block: no inter in-edge
    nn_s, nn_d = [], []
    for b in range(bsz):
        s_, d_ = _rand_edges(rng, n_per, n_per, 3 * n_per)
        nn_s += [s_ + b * n_per, np.arange(n_per) + b * n_per]
        nn_d += [d_ + b * n_per, np.arange(n_per) + b * n_per]
    edges = {("tgt", "intra", "tgt"): tt, ("ntgt", "inter", "tgt"): (i_src, i_dst.astype(np.int64)),
             ("ntgt", "intra", "ntgt"): (np.concatenate(nn_s).astype(np.int64), np.concatenate(nn_d).astype(np.int64))}
    hgt = _ref_hgt()
    torch.manual_seed(seed)
    model = hgt.HGT(ntype2idx={"tgt": 0, "ntgt": 1}, etype2idx={"intra": 0, "inter": 1}, in_dim=d, hidden_dim=d, out_dim=d,
                    n_layers=n_layers, n_heads=H, dropout=0.0, attn_drop=0.0).eval()
    _perturb(model, d)
    etypes = list(edges.keys())
    h_ntgt = torch.randn(n["ntgt"], d)
    h_steps = torch.randn(steps, bsz, d)
    sg = _StubGraph(edges, n)
    sg.nodes["ntgt"].data["h"] = h_ntgt
    inc: dict = {}
    outs, outs_n = [], []
    order = torch.arange(bsz)
    with torch.no_grad():
        for s_ in range(steps):
            if reorder_at is not None and s_ == reorder_at:
                order = torch.arange(bsz).flip(0)
                for layer in model.gcs:
                    layer.reorder_incremental_state(inc, order)
            x = h_steps[s_][order] if reorder_at is not None and s_ >= reorder_at else h_steps[s_]
            out = model(sg, features={"tgt": x}, etypes=etypes, incremental_state=inc)
            outs.append(out["tgt"].clone())
            outs_n.append(out["ntgt"].clone())
    res = {"H": H, "n_layers": n_layers, "bsz": bsz, "steps": steps, "max_len": max_len, "reorder_at": -1 if reorder_at is None else reorder_at,
           "ntypes": np.array(list(n.keys())), "num_nodes": np.array(list(n.values())), "cets": np.array(["|".join(c) for c in edges]),
           "h.ntgt": h_ntgt.numpy(), "h_steps": h_steps.numpy(), "out_steps": torch.stack(outs).numpy(),
           "out_ntgt_first": outs_n[0].numpy(), "out_ntgt_last": outs_n[-1].numpy()}
    for c in edges:
        if c != ("tgt", "intra", "tgt"):          # the causal list is regenerated by the test (triu per block)
            res["src." + "|".join(c)], res["dst." + "|".join(c)] = edges[c]
    _save_hgt(name, model, res)
    print(f"hgt_{name}: steps={steps} |out|={float(outs[-1].abs().mean()):.3f}")


def make_adaptive_input_case(name, V, d, cutoff, T, seed):
    """adaptive_input_*.npz: the reference's AdaptiveInput (fairseq/modules/adaptive_input.py, loaded by file path) on random
    tokens that hit every band and both edges of every cutoff -- what `--reinit-nfeat` feeds the ntgt nodes."""
    _, ain = _ref_adaptive()
    torch.manual_seed(seed)
    m = ain.AdaptiveInput(V, 1, d, 4, d, list(cutoff))
    m.eval()
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(0, V, (T,), generator=g)
    edge = [0, 1, cutoff[0] - 1, cutoff[0], cutoff[1] - 1, cutoff[1], V - 1]
    tok[:len(edge)] = torch.tensor(edge)
    with torch.no_grad():
        out = m(tok)
    rec = {"tokens": tok.numpy(), "out": out.numpy(), "cutoff": np.array(m.cutoff), "V": np.array(V), "d": np.array(d)}
    for k_, v_ in m.state_dict().items():
        rec["sd." + k_] = v_.numpy()
    np.savez_compressed(os.path.join(OUT, f"adaptive_input_{name}.npz"), **rec)
    print(f"adaptive_input_{name}: keys={[k for k in rec if k.startswith('sd.')]}")


# ---------------------------------------------------------------------------------------- flags and architecture presets
def make_registry_fixture():
    """registry.json: (1) the defaults argparse produces for the reference's own add_args of the transformer_lm model
    (fairseq/models/transformer_lm.py:52-139), the language_modeling task (fairseq/tasks/language_modeling.py:68-153) and the
    eval-lm option group (fairseq/options.py:456-501); (2) what each `transformer_lm*` architecture function
    (transformer_lm.py:190-333) makes of an empty Namespace.  The functions are lifted by ast and executed unmodified; only
    `utils.get_available_activation_fns` and the registry decorators are stubbed."""
    import argparse
    import json as _json
    out = {}
    # --- model / task / eval flags
    src = _method_source("fairseq/models/transformer_lm.py", "TransformerLanguageModel", "add_args")
    ns = {"utils": types.SimpleNamespace(get_available_activation_fns=lambda: ["relu", "gelu", "gelu_fast", "gelu_accurate", "tanh", "linear"])}
    exec(textwrap.dedent(src).replace("@staticmethod\n", ""), ns)
    def options(parser):      # dest -> spelling, kind and type of every option, so that the mirror keeps the command line
        return {a.dest: {"flags": list(a.option_strings), "action": type(a).__name__,
                         "type": getattr(a.type, "__name__", None), "choices": list(a.choices) if a.choices else None,
                         "nargs": a.nargs, "const": a.const}
                for a in parser._actions if a.dest != "help"}
    pm = argparse.ArgumentParser()
    ns["add_args"](pm)
    out["model_flags"] = vars(pm.parse_args([]))
    out["model_options"] = options(pm)
    src = _method_source("fairseq/tasks/language_modeling.py", "LanguageModelingTask", "add_args")
    ns = {}
    exec(textwrap.dedent(src).replace("@staticmethod\n", ""), ns)
    pt = argparse.ArgumentParser()
    ns["add_args"](pt)
    out["task_flags"] = vars(pt.parse_args(["DATA"]))
    out["task_options"] = options(pt)
    osrc = _src("fairseq/options.py")
    tree = ast.parse(osrc)
    ns = {"sys": sys}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("add_common_eval_args", "add_eval_lm_args"):
            exec(ast.get_source_segment(osrc, node), ns)
    pe = argparse.ArgumentParser()
    ns["add_eval_lm_args"](pe)
    out["eval_flags"] = vars(pe.parse_args([]))
    out["eval_options"] = options(pe)
    # --- architecture presets
    msrc = _src("fairseq/models/transformer_lm.py")
    tree = ast.parse(msrc)
    ns, archs = {}, {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            exec(ast.get_source_segment(msrc, node), ns)      # the function body, unmodified
            for dec in node.decorator_list:                   # @register_model_architecture('transformer_lm', '<arch>')
                if isinstance(dec, ast.Call) and getattr(dec.func, "id", "") == "register_model_architecture":
                    assert dec.args[0].value == "transformer_lm"
                    archs[dec.args[1].value] = ns[node.name]
    ns["_archs"] = archs
    out["archs"] = {}
    for arch, fn in ns["_archs"].items():
        a = argparse.Namespace()
        fn(a)
        out["archs"][arch] = vars(a)
    # an "old checkpoint" namespace exercises the backward-compatibility branch of base_lm_architecture (:193-200)
    a = argparse.Namespace(no_tie_adaptive_proj=False, decoder_final_norm=False)
    ns["_archs"]["transformer_lm"](a)
    out["archs_old_checkpoint"] = vars(a)
    with open(os.path.join(OUT, "registry.json"), "w") as f:
        _json.dump(out, f, indent=1, sort_keys=True)
    print("registry.json:", len(out["model_flags"]), "model flags,", len(out["task_flags"]), "task flags,", len(out["eval_flags"]),
          "eval flags,", len(out["archs"]), "architectures")


# ---------------------------------------------------------------------------------------- on-disk formats
def make_format_fixtures():
    """fmt_uint16.{bin,idx}, fmt_int32.{bin,idx}, fmt_dict.txt written by the reference's OWN writers
    (MMapIndexedDatasetBuilder, fairseq/data/indexed_dataset.py:496-523; Dictionary.save, fairseq/data/dictionary.py:230-252)
    and fmt.npz = what the reference's own readers return for them (MMapIndexedDataset.__getitem__ / .sizes,
    Dictionary.load numbering).  indexed_dataset.py is executed whole with only its package-relative import stubbed and
    `np.float` (removed in numpy 2) aliased to float; Dictionary is lifted by ast with its fairseq imports stubbed."""
    src = _src("fairseq/data/indexed_dataset.py").replace("from . import FairseqDataset", "FairseqDataset = object")
    had = hasattr(np, "float")
    if not had:
        np.float = float
    ns = {"__name__": "ref_indexed_dataset"}
    exec(compile(src, "indexed_dataset.py", "exec"), ns)
    dsrc = _class_source("fairseq/data/dictionary.py", "Dictionary")
    dns = {"torch": torch, "os": os, "Counter": __import__("collections").Counter,
           "PathManager": types.SimpleNamespace(open=open, mkdirs=lambda p: os.makedirs(p, exist_ok=True)),
           "data_utils": None, "safe_readline": None, "tokenize_line": lambda l: l.split(), "Pool": None}
    exec(compile(dsrc, "dictionary.py", "exec"), dns)
    RefDict = dns["Dictionary"]

    rng = np.random.RandomState(7)
    d = RefDict()
    words = [f"w{i}" for i in range(40)] + ["with space", "ünï", "madeupword0000", "madeupword0001"]
    for i, w in enumerate(words):
        d.add_symbol(w, n=1000 - i)
    dict_path = os.path.join(OUT, "fmt_dict.txt")
    d.save(dict_path)
    d2 = RefDict.load(dict_path)
    out = {"dict_len": len(d2), "dict_pad": d2.pad(), "dict_eos": d2.eos(), "dict_unk": d2.unk(), "dict_bos": d2.bos(),
           "dict_symbols": np.array(d2.symbols), "dict_index_w7": d2.index("w7"), "dict_index_missing": d2.index("nope")}
    for tag, dt in (("uint16", np.uint16), ("int32", np.int32)):
        prefix = os.path.join(OUT, f"fmt_{tag}")
        b = ns["MMapIndexedDatasetBuilder"](prefix + ".bin", dtype=dt)
        sents = []
        for n in (5, 1, 9, 3, 17, 2):
            sent = np.concatenate([rng.randint(4, len(d2), size=n - 1), [d2.eos()]]).astype(np.int64)
            sents.append(sent)
            b.add_item(torch.from_numpy(sent))
        b.finalize(prefix + ".idx")
        ref = ns["MMapIndexedDataset"](prefix)
        assert len(ref) == len(sents)
        out[f"{tag}_sizes"] = np.array(ref.sizes)
        out[f"{tag}_flat"] = np.concatenate([ref[i].numpy() for i in range(len(ref))])
        for i in range(len(ref)):
            out[f"{tag}_sent{i}"] = ref[i].numpy()
        del ref
    np.savez(os.path.join(OUT, "fmt.npz"), **out)
    if not had:
        del np.float
    print("fmt fixtures written")


if __name__ == "__main__":
    if "--formats-only" in sys.argv:
        make_format_fixtures()
        sys.exit(0)
    if "--registry-only" in sys.argv:
        make_registry_fixture()
        sys.exit(0)
    if "--adaptive-noproj-only" in sys.argv:
        # --tie-adaptive-weights WITHOUT --tie-adaptive-proj: tail projections are nn.Linear(d, dim_i) (adaptive_softmax.py:96-101)
        make_adaptive_case("tied_noproj", V=300, d=64, cutoff=[40, 120], tied=True, T=40, seed=2, tie_proj=False)
        sys.exit(0)
    if "--hetero-only" in sys.argv:
        make_hetero_case("hetero4", d=32, H=4, n_layers=2, seed=3)
        make_two_stream_case("two_stream", d=32, H=2, n_layers=2, seed=4)
        make_infer_case("infer_b2", bsz=2, steps=5, n_per=9, d=32, H=4, n_layers=2, seed=5)
        make_infer_case("infer_b3_reorder", bsz=3, steps=4, n_per=6, d=32, H=2, n_layers=1, seed=6, reorder_at=2)
        sys.exit(0)
    if "--adaptive-input-only" in sys.argv:
        make_adaptive_input_case("v300", V=300, d=64, cutoff=[40, 120], T=64, seed=0)
        sys.exit(0)
    make_slice_cases()
    make_edges_doctest()
    make_graph_case("c1_c1", L=24, k=4, n_d=400, cl=1, cr=1, invalid_ctx=0, intra_ctx=0, M=8, seed=0, stress=True)
    make_graph_case("c2_c0", L=16, k=3, n_d=300, cl=2, cr=0, invalid_ctx=0, intra_ctx=0, M=8, seed=1, stress=True)
    make_graph_case("c3_c3_intra5", L=20, k=5, n_d=600, cl=3, cr=3, invalid_ctx=0, intra_ctx=5, M=16, seed=2, stress=True)
    make_graph_case("c0_c0", L=12, k=6, n_d=200, cl=0, cr=0, invalid_ctx=0, intra_ctx=0, M=8, seed=3, stress=False)
    make_graph_case("c1_c2_invalid", L=32, k=4, n_d=1200, cl=1, cr=2, invalid_ctx=300, intra_ctx=0, M=8, seed=4, stress=True)
    make_dedup_case("c1", L=24, k=4, n_d=400, cl=1, cr=1, invalid_ctx=0, intra_ctx=0, M=8, seed=0)
    make_dedup_case("c2_c0", L=16, k=3, n_d=300, cl=2, cr=0, invalid_ctx=0, intra_ctx=0, M=8, seed=1)
    make_dedup_case("c0", L=12, k=5, n_d=60, cl=0, cr=0, invalid_ctx=0, intra_ctx=4, M=8, seed=2)
    make_dedup_case("c1_c3_invalid", L=32, k=4, n_d=1200, cl=1, cr=3, invalid_ctx=300, intra_ctx=0, M=16, seed=3)
    make_adaptive_case("untied", V=300, d=64, cutoff=[40, 120], tied=False, T=40, seed=0)
    make_adaptive_case("tied", V=300, d=64, cutoff=[40, 120], tied=True, T=40, seed=1)
    make_adaptive_case("tied_noproj", V=300, d=64, cutoff=[40, 120], tied=True, T=40, seed=2, tie_proj=False)
    make_pq_case("m8", n=50, M=8, dsub=4, with_b=False, seed=0)
    make_pq_case("m16b", n=30, M=16, dsub=8, with_b=True, seed=1)
    make_knn_case("ip_t1", T=24, knn=16, n_d=2000, V=300, temp=1.0, metric="do_not_recomp_ip", seed=0, with_missing=True)
    make_knn_case("l2_t001", T=24, knn=32, n_d=2000, V=300, temp=0.01, metric="do_not_recomp_l2", seed=1, with_missing=False)
    make_knn_recompute_case("recomp_l2", T=12, knn=16, n_d=500, d=64, V=300, temp=10.0, metric="l2",
                            index_file="faiss_store.l2", seed=2, fp16_keys=True)
    make_knn_recompute_case("recomp_ip", T=12, knn=16, n_d=500, d=64, V=300, temp=10.0, metric="ip",
                            index_file="faiss_store.ip", seed=3, fp16_keys=False)
    make_knn_recompute_case("recomp_cos", T=12, knn=16, n_d=500, d=96, V=300, temp=0.1, metric="ip",
                            index_file="faiss_store.cosine", seed=4, fp16_keys=True)
    make_scorer_case("b1_knn", B=1, L=20, d=64, V=300, cutoff=[40, 120], knn=16, lmbda=0.25, temp=1.0, seed=0)
    make_scorer_case("b2_lm", B=2, L=12, d=64, V=300, cutoff=[40, 120], knn=8, lmbda=0.25, temp=1.0, seed=1)
    make_hgt_case("l2_c1", B=2, L=6, k=3, cl=1, cr=1, d=32, H=4, n_layers=2, seed=0)
    make_hgt_case("l3_c2", B=1, L=8, k=2, cl=2, cr=2, d=32, H=2, n_layers=3, seed=1)
    make_hgt_case("l2_adapt", B=1, L=7, k=3, cl=1, cr=1, d=24, H=4, n_layers=2, seed=2, hidden=32, out_dim=24)
    make_hetero_case("hetero4", d=32, H=4, n_layers=2, seed=3)
    make_two_stream_case("two_stream", d=32, H=2, n_layers=2, seed=4)
    make_infer_case("infer_b2", bsz=2, steps=5, n_per=9, d=32, H=4, n_layers=2, seed=5)
    make_infer_case("infer_b3_reorder", bsz=3, steps=4, n_per=6, d=32, H=2, n_layers=1, seed=6, reorder_at=2)
    make_format_fixtures()
    make_adaptive_input_case("v300", V=300, d=64, cutoff=[40, 120], T=64, seed=0)
    make_registry_fixture()
