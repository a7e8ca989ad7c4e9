"""CPU ORACLE (test infrastructure, NOT product code) -- graph assembly.

Restates, line-faithfully, the reference's per-block kNN token-graph builder.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product path (gnn-lm_b200/) never does.

Reference lines followed (all under /root/reference):
  * fairseq/data/token_block_dataset.py:338-412  new_build_graph
  * fairseq/data/token_block_dataset.py:545-584  build_ntgt_edges
  * fairseq/data/token_block_dataset.py:586-594  auto_regressive_edges
  * fairseq/data/monolingual_dataset.py:237-262  collater -> dgl.batch (node-id offsets)
  * fairseq/data/token_block_utils_fast.pyx:22-35 block slicing in 'none' mode

Parity pin: `tests/golden/graph_*.npz` were produced by executing the reference's
*own* functions (extracted with `ast`, run under a stub `dgl`) -- see
tests/golden/make_golden.py.  `tests/test_oracle_graph.py` checks this restatement
against them and against the doctest vectors at token_block_dataset.py:549-554.

Deviation: token_block_dataset.py:384 calls
`len(self.neighbor_offsets.shape[0])` (len of an int -> TypeError) whenever the
right context is > 0.  The restatement uses `len(self.neighbor_tokens)` there,
which is what deprecated_build_graph uses (token_block_dataset.py:458).
Q3: an all-invalid block yields an empty ntgt set instead of raising at :410.
"""
from typing import Dict, List, Tuple

import numpy as np


def build_ntgt_edges(offsets2id: Dict[int, int], context: int = 0, bidirect: bool = False
                     ) -> Tuple[List[int], List[int]]:
    """token_block_dataset.py:545-584 -- sliding-window edges between ntgt nodes."""
    if not offsets2id:
        return [], []
    nodes = sorted(((nid, off) for off, nid in offsets2id.items()), key=lambda x: x[1])
    src, tgt = [], []
    start, end, length = 0, -1, len(nodes)
    while start < length:
        while end + 1 < length and nodes[end + 1][1] <= nodes[start][1] + context:
            end += 1
            for s in range(start, end + 1):
                src.append(nodes[s][0])
                tgt.append(nodes[end][0])
        start += 1
    if bidirect:
        for idx in range(len(src)):
            s, t = src[idx], tgt[idx]
            if s != t:
                src.append(t)
                tgt.append(s)
    return src, tgt


def auto_regressive_edges(length: int, max_context: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """token_block_dataset.py:586-594 -- (u, v) with u <= v (and v - u < max_context)."""
    mask = np.triu(np.ones((length, length), dtype=bool))
    if max_context:
        mask &= ~np.triu(np.ones((length, length), dtype=bool), k=max_context)
    us, vs = np.nonzero(mask)  # row-major, same order as torch.where
    return us.astype(np.int64), vs.astype(np.int64)


def new_build_graph(offsets: np.ndarray, neighbor_idxs: np.ndarray, n_datastore: int,
                    left_ctx: int, right_ctx: int, invalid_ctx: int = 0, max_intra_context: int = 0,
                    quant_feats: np.ndarray = None, neighbor_tokens: np.ndarray = None) -> dict:
    """token_block_dataset.py:338-412, one block.

    Returns COO edge lists in the reference's insertion order, the datastore row of
    every ntgt node (`ntgt_offsets`), gathered code rows / labels when tables are given.
    """
    L = len(offsets)
    ntgt_id = 0
    tgt2ntgt = [[], []]
    ntgt2ntgt = [[], []]
    ntgt_offsets = []
    for tgt_idx in range(L):
        for offset in neighbor_idxs[tgt_idx]:
            offset = int(offset)
            if offset == -1:
                continue
            if abs(int(offsets[tgt_idx]) - offset) < invalid_ctx:
                continue
            cur_ids, cur_offs = [ntgt_id], [offset]
            ntgt_offsets.append(offset)
            tgt2ntgt[0].append(tgt_idx)
            tgt2ntgt[1].append(ntgt_id)
            ntgt_id += 1
            ctx = []
            if left_ctx:
                ctx.extend(range(max(0, offset - left_ctx), offset))
            if right_ctx:
                ctx.extend(range(offset + 1, min(n_datastore, offset + 1 + right_ctx)))  # Q1 fix
            for o in ctx:
                cur_ids.append(ntgt_id)
                cur_offs.append(o)
                ntgt_offsets.append(o)
                ntgt_id += 1
            s, t = build_ntgt_edges({o: i for i, o in zip(cur_ids, cur_offs)}, context=1, bidirect=True)
            ntgt2ntgt[0].extend(s)
            ntgt2ntgt[1].extend(t)
    us, vs = auto_regressive_edges(L, max_intra_context)
    ntgt_offsets = np.asarray(ntgt_offsets, dtype=np.int64)
    out = {
        "n_tgt": L,
        "n_ntgt": ntgt_id,
        "tt": (us, vs),                                                     # ('tgt','intra','tgt')
        "inter": (np.asarray(tgt2ntgt[1], np.int64), np.asarray(tgt2ntgt[0], np.int64)),  # ntgt -> tgt
        "nn": (np.asarray(ntgt2ntgt[0], np.int64), np.asarray(ntgt2ntgt[1], np.int64)),   # ntgt -> ntgt
        "ntgt_offsets": ntgt_offsets,
    }
    if quant_feats is not None:
        out["ntgt_codes"] = quant_feats[ntgt_offsets] if ntgt_id else np.zeros((0, quant_feats.shape[1]), quant_feats.dtype)
    if neighbor_tokens is not None:
        out["ntgt_labels"] = neighbor_tokens.reshape(-1)[ntgt_offsets].astype(np.int64)
    return out


def deprecated_build_graph(offsets: np.ndarray, neighbor_idxs: np.ndarray, n_datastore: int, left_ctx: int,
                           right_ctx: int, invalid_ctx: int = 0, max_intra_context: int = 0,
                           quant_feats: np.ndarray = None) -> dict:
    """token_block_dataset.py:414-479 (`--deprecated`), one block: datastore rows are DE-DUPLICATED -- one ntgt node per
    distinct row of the block, numbered by first appearance (centre, then its left context ascending, then its right
    context ascending, neighbour by neighbour); ntgt-ntgt edges by build_ntgt_edges over ALL nodes of the block
    (context 1, bidirectional): rows at distance <= 1 are connected wherever they came from."""
    L = len(offsets)
    off2id: Dict[int, int] = {}
    tgt2ntgt = [[], []]
    for tgt_idx in range(L):
        for offset in neighbor_idxs[tgt_idx]:
            offset = int(offset)
            if offset == -1:                                                  # :432
                continue
            if abs(int(offsets[tgt_idx]) - offset) < invalid_ctx:             # :435
                continue
            if offset not in off2id:                                          # :440-448
                off2id[offset] = len(off2id)
            tgt2ntgt[0].append(tgt_idx)
            tgt2ntgt[1].append(off2id[offset])
            ctx = []
            if left_ctx:                                                      # :452-455
                ctx.extend(range(max(0, offset - left_ctx), offset))
            if right_ctx:                                                     # :456-460
                ctx.extend(range(offset + 1, min(n_datastore, offset + 1 + right_ctx)))
            for o in ctx:                                                     # :461-466
                if o not in off2id:
                    off2id[o] = len(off2id)
    s, t = build_ntgt_edges(off2id, 1, bidirect=True)                         # :469
    us, vs = auto_regressive_edges(L, max_intra_context)
    ntgt_offsets = np.asarray(list(off2id.keys()), dtype=np.int64)            # insertion order == node id order
    out = {"n_tgt": L, "n_ntgt": len(off2id), "tt": (us, vs),
           "inter": (np.asarray(tgt2ntgt[1], np.int64), np.asarray(tgt2ntgt[0], np.int64)),
           "nn": (np.asarray(s, np.int64), np.asarray(t, np.int64)), "ntgt_offsets": ntgt_offsets}
    if quant_feats is not None:
        out["ntgt_codes"] = quant_feats[ntgt_offsets] if len(off2id) else np.zeros((0, quant_feats.shape[1]), quant_feats.dtype)
    return out


def batch_graphs(graphs: List[dict]) -> dict:
    """dgl.batch semantics (monolingual_dataset.py:261): per-type node ids are offset by the
    cumulative node counts of the preceding graphs; edge lists are concatenated."""
    tgt_base = ntgt_base = 0
    acc = {"tt": [[], []], "inter": [[], []], "nn": [[], []]}
    extras = {k: [] for k in ("ntgt_offsets", "ntgt_codes", "ntgt_labels") if k in graphs[0]}
    for g in graphs:
        acc["tt"][0].append(g["tt"][0] + tgt_base)
        acc["tt"][1].append(g["tt"][1] + tgt_base)
        acc["inter"][0].append(g["inter"][0] + ntgt_base)
        acc["inter"][1].append(g["inter"][1] + tgt_base)
        acc["nn"][0].append(g["nn"][0] + ntgt_base)
        acc["nn"][1].append(g["nn"][1] + ntgt_base)
        for k in extras:
            extras[k].append(g[k])
        tgt_base += g["n_tgt"]
        ntgt_base += g["n_ntgt"]
    out = {"n_tgt": tgt_base, "n_ntgt": ntgt_base}
    for k, (s, d) in acc.items():
        out[k] = (np.concatenate(s), np.concatenate(d))
    for k, v in extras.items():
        out[k] = np.concatenate(v)
    return out


def canonical_csr(src: np.ndarray, dst: np.ndarray, n_dst: int) -> Tuple[np.ndarray, np.ndarray]:
    """stable-sort the insertion-ordered COO by dst ->
    indptr[n_dst+1] (int32), indices[E] = src (int32)."""
    order = np.argsort(dst, kind="stable")
    indices = src[order].astype(np.int32)
    counts = np.bincount(dst, minlength=n_dst)
    indptr = np.zeros(n_dst + 1, dtype=np.int32)
    np.cumsum(counts, out=indptr[1:])
    return indptr, indices


def build_batch(neighbor_idxs: np.ndarray, offsets: np.ndarray, n_datastore: int, left_ctx: int,
                right_ctx: int, invalid_ctx: int = 0, max_intra_context: int = 0, deprecated: bool = False) -> dict:
    """[B, L, k] neighbour ids + [B, L] stream positions -> batched graph + canonical CSRs."""
    one = deprecated_build_graph if deprecated else new_build_graph
    graphs = [one(offsets[b], neighbor_idxs[b], n_datastore, left_ctx, right_ctx,
                  invalid_ctx, max_intra_context) for b in range(neighbor_idxs.shape[0])]
    g = batch_graphs(graphs)
    g["tt_csr"] = canonical_csr(*g["tt"], g["n_tgt"])
    g["inter_csr"] = canonical_csr(*g["inter"], g["n_tgt"])
    g["nn_csr"] = canonical_csr(*g["nn"], g["n_ntgt"])
    return g


# ---------------------------------------------------------------------------------------------
# Vectorised form (numpy), used only to time the CPU baseline at sizes where the Python loop
# above would take minutes; tests pin it to new_build_graph on small inputs.
# ---------------------------------------------------------------------------------------------
def build_batch_vectorised(neighbor_idxs: np.ndarray, offsets: np.ndarray, n_datastore: int,
                           left_ctx: int, right_ctx: int, invalid_ctx: int = 0) -> dict:
    B, L, k = neighbor_idxs.shape
    o = neighbor_idxs.reshape(-1).astype(np.int64)
    pos = np.repeat(offsets.reshape(-1).astype(np.int64), k)
    valid = (o != -1) & ~(np.abs(pos - o) < invalid_ctx)
    nl = np.where(valid, np.minimum(left_ctx, o), 0)
    nr = np.where(valid, np.clip(np.minimum(n_datastore, o + 1 + right_ctx) - (o + 1), 0, None), 0)
    size = np.where(valid, 1 + nl + nr, 0)
    base = np.concatenate([[0], np.cumsum(size)])
    n_ntgt = int(base[-1])
    cl = np.repeat(np.arange(o.size), size)               # cluster of each node
    i = np.arange(n_ntgt) - base[cl]                       # index inside the cluster (creation order)
    nlc, oc = nl[cl], o[cl]
    # creation order: centre, left ascending, right ascending
    node_off = np.where(i == 0, oc, np.where(i <= nlc, oc - nlc + i - 1, oc + (i - nlc)))
    # sorted position inside the cluster
    p = np.where(i == 0, nlc, np.where(i <= nlc, i - 1, i))
    w = size[cl]
    pos2id = lambda q: base[cl] + np.where(q == nlc, 0, np.where(q < nlc, q + 1, q))
    deg = 1 + (p > 0) + (p < w - 1)
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    indices = np.empty(int(indptr[-1]), np.int32)
    cur = indptr[:-1].copy()
    m = p > 0
    indices[cur[m]] = pos2id(p - 1)[m]
    cur[m] += 1
    indices[cur] = np.arange(n_ntgt)
    cur += 1
    m = p < w - 1
    indices[cur[m]] = pos2id(p + 1)[m]
    vcnt = np.concatenate([[0], np.cumsum(valid)])
    inter_indptr = vcnt[np.arange(B * L + 1) * k].astype(np.int32)
    inter_indices = base[:-1][valid].astype(np.int32)
    return {"n_tgt": B * L, "n_ntgt": n_ntgt, "ntgt_offsets": node_off,
            "nn_csr": (indptr, indices), "inter_csr": (inter_indptr, inter_indices)}


# --------------------------------------------------------------------------- block boundaries
def slice_indices(sizes, break_mode, block_size: int, document_sep_len: int = 1) -> np.ndarray:
    """Block boundaries [n_blocks, 2] over the flat token stream for --sample-break-mode none / complete /
    complete_doc / eos (fairseq/data/token_block_utils_fast.pyx:22-105).  Plain loops on purpose."""
    sizes = [int(x) for x in sizes]
    total = sum(sizes)
    out = []
    if break_mode is None or break_mode == "none":                      # :22-35
        n = -(-total // block_size)
        out = [(i * block_size, min(i * block_size + block_size, total)) for i in range(n)]
    elif break_mode == "eos":                                           # :96-100: one sentence per block
        pos = 0
        for sz in sizes:
            out.append((pos, pos + sz))
            pos += sz
    elif break_mode in ("complete", "complete_doc"):                    # :63-95: whole sentences up to block_size
        doc = break_mode == "complete_doc"
        keep = 1 if doc else 0                                          # complete_doc drops blocks of <= 1 token
        tok, cur, i = 0, 0, 0
        while i < len(sizes):
            fits = cur + sizes[i] <= block_size or cur == 0
            if fits and not (doc and sizes[i] == document_sep_len):
                cur += sizes[i]
                i += 1
            else:
                if cur > keep:
                    out.append((tok, tok + cur))
                tok += cur
                cur = 0
                if doc and sizes[i] == document_sep_len:               # an empty sentence ends the document
                    tok += sizes[i]
                    i += 1
        if cur > keep:
            out.append((tok, tok + cur))
    else:
        raise ValueError("Invalid break_mode: " + str(break_mode))
    return np.asarray(out, dtype=np.int64).reshape(-1, 2)
