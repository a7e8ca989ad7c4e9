"""HGT on a GENERAL heterograph, its `two_stream` query stream and its incremental `infer()` -- the parts of
fairseq/models/hgt.py the LM evaluation never reaches: HGTLayer.forward for arbitrary node / edge types
(:299-420; the module's own self-test builds four node types and ten canonical edge types, :516-552), `two_stream` (:324-330,
360-394, 407-416) and HGTLayer.infer (:81-297) with reorder_incremental_state (:422-438).

Everything runs on libgnnlm_sm100.so kernels: per node type ONE projection GEMM emits Q | K'_r | V'_r for every relation r the
type is a source of (relation_att / relation_msg / relation_pri / sqrt(d_k) folded into the weights in fp64, as in hgt.py of this
package), gnnlm_hgt_edge_attn does score + edge softmax + aggregation + the cross-type mean per canonical edge type over a CSR by
destination, gnnlm_layernorm adds the residual and normalises.  torch is used for the container only (COO -> CSR of a graph the
caller hands over as edge lists, like `dgl.heterograph`).

Incremental decoding keeps, per layer, ONE cache row block per node type -- [bsz, nodes / bsz, Q | K'_r | V'_r ...] -- that the
step's projection writes in place (the GEMM's output rows are the cache rows of the current position), and asks the attention
kernel for the bsz current destinations only (`dst_ids`), instead of re-transforming every cached key / value and recomputing
every tgt node per step as the reference does (hgt.py:236-264).
"""
import contextlib
import uuid
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib as L
from . import ops

CET = Tuple[str, str, str]
MAX_LEN = 512          # hgt.py:93 (`max_len = 512  # todo`): tgt nodes per block of the incremental graph


class _NodeView:
    def __init__(self):
        self.data = {}


class HeteroGraph:
    """Stand-in for `dgl.heterograph(data_dict, num_nodes_dict)`: {(srctype, etype, dsttype): (src ids, dst ids)} kept as CSR by
    destination per canonical edge type (stable in the edge order, SURVEY.md §8c "canonical CSR").  Node types are sorted, as DGL
    sorts them (`G.ntypes`)."""

    def __init__(self, edges: Dict[CET, Tuple[torch.Tensor, torch.Tensor]], num_nodes: Optional[Dict[str, int]] = None,
                 device="cuda"):
        dev = torch.device(device)
        counts: Dict[str, int] = dict(num_nodes or {})
        self._coo = {}
        for (s, r, t), (u, v) in edges.items():
            u = torch.as_tensor(u, dtype=torch.int64).to(dev)
            v = torch.as_tensor(v, dtype=torch.int64).to(dev)
            if u.shape != v.shape or u.dim() != 1:
                raise ValueError(f"edge type {(s, r, t)}: src / dst must be 1-d arrays of equal length")
            if num_nodes is None:
                counts[s] = max(counts.get(s, 0), int(u.max()) + 1 if u.numel() else 0)
                counts[t] = max(counts.get(t, 0), int(v.max()) + 1 if v.numel() else 0)
            elif u.numel() and (int(u.max()) >= counts[s] or int(v.max()) >= counts[t] or int(u.min()) < 0 or int(v.min()) < 0):
                raise ValueError(f"edge type {(s, r, t)}: node id out of range")          # DGLError in the reference
            self._coo[(s, r, t)] = (u, v)
        self._n = counts
        self.ntypes = sorted(counts)
        self.canonical_etypes = list(edges.keys())
        self.nodes = {t: _NodeView() for t in counts}
        self.device = dev
        self._csr: Dict[Tuple[CET, bool], Tuple[torch.Tensor, torch.Tensor]] = {}
        self._shift: Dict[CET, torch.Tensor] = {}

    def num_nodes(self, ntype: str) -> int:
        return self._n[ntype]

    def num_edges(self, cet: CET) -> int:
        return int(self._coo[cet][0].numel())

    def local_scope(self):
        @contextlib.contextmanager
        def scope():
            saved = {t: dict(v.data) for t, v in self.nodes.items()}
            try:
                yield
            finally:
                for t, v in self.nodes.items():
                    v.data = saved[t]
        return scope()

    def to(self, device):
        dev = torch.device(device)
        if dev != self.device:
            self._coo = {c: (u.to(dev), v.to(dev)) for c, (u, v) in self._coo.items()}
            self._csr, self._shift, self.device = {}, {}, dev
        for nv in self.nodes.values():
            nv.data = {k: x.to(dev, non_blocking=True) for k, x in nv.data.items()}
        return self

    def csr(self, cet: CET, self_loops_shifted: bool = False):
        """(indptr [n_dst + 1], indices [E]) int32.  self_loops_shifted: the source id of every self loop u == v is moved to
        u + n_src -- the second half of a doubled source table (two_stream: the query stream's self loop sees its own key)."""
        key = (cet, self_loops_shifted)
        if key not in self._csr:
            u, v = self._coo[cet]
            n_dst = self._n[cet[2]]
            order = torch.sort(v, stable=True).indices
            src = u[order]
            if self_loops_shifted:
                src = torch.where(src == v[order], src + self._n[cet[0]], src)
            indptr = torch.zeros(n_dst + 1, dtype=torch.int64, device=self.device)
            indptr[1:] = torch.cumsum(torch.bincount(v, minlength=n_dst), 0)
            self._csr[key] = (indptr.to(torch.int32), src.to(torch.int32).contiguous())
        return self._csr[key]


# --------------------------------------------------------------------------------------------------------------------------
# weights: per node type [Q | K'_r V'_r ...] with the relation transforms folded
# --------------------------------------------------------------------------------------------------------------------------
def _prepare(layer, math_mode: int, ntypes: List[str], etypes: List[CET], two_stream: bool):
    from .hgt import _Weight, _fold
    key = ("hetero", math_mode, tuple(ntypes), tuple(etypes), two_stream, layer.relation_pri.device,
           tuple(int(p._version) for p in layer.parameters()))
    if getattr(layer, "_hprep_key", None) == key:
        return layer._hprep
    if not layer.use_norm:
        raise NotImplementedError("use_norm=False is never used by the reference decoder")
    d = layer.out_dim
    pri = layer.relation_pri.detach()
    P = {"w": {}, "cols": {}, "a": {}, "ln": {}, "d": d, "math": math_mode}

    def kv(tau, rel):
        Wk, bk = _fold(layer.k_linears[tau].weight, layer.k_linears[tau].bias, layer.relation_att[rel], pri[rel] / layer.sqrt_dk)
        Wv, bv = _fold(layer.v_linears[tau].weight, layer.v_linears[tau].bias, layer.relation_msg[rel], None)
        return Wk, bk, Wv, bv

    for nt in ntypes:
        tau = layer.ntype2idx[nt]
        Ws, bs, cols = [layer.q_linears[tau].weight.detach().float()], [layer.q_linears[tau].bias.detach().float()], {"q": 0}
        for r in sorted({r for (s, r, _) in etypes if s == nt}):
            Wk, bk, Wv, bv = kv(tau, layer.etype2idx[r])
            cols[("k", r)], cols[("v", r)] = len(Ws) * d, (len(Ws) + 1) * d
            Ws += [Wk, Wv]
            bs += [bk, bv]
        P["w"][nt] = _Weight(torch.cat(Ws, 0), torch.cat(bs, 0), math_mode)
        P["cols"][nt] = cols
        P["a"][nt] = _Weight(layer.a_linears[tau].weight.detach(), layer.a_linears[tau].bias.detach(), math_mode)
        P["ln"][nt] = (layer.norms[tau].weight.detach().float().contiguous(), layer.norms[tau].bias.detach().float().contiguous(),
                       layer.norms[tau].eps)
    if two_stream:
        # hgt.py:324-330: tgt_tilde_q = q_linears[tgt](h~); tgt_tilde_k = k_linear(h~) with k_linear the LAST node type's of the
        # projection loop (G.ntypes is sorted; 'tgt' sorts last in every graph the reference builds).  K~' = tgt_tilde_k through
        # relation_att['intra'] -- quirk Q11 (oracle/hetero_oracle.py): the reference reads it as srcdata['k_tilde'] (:376).
        t, last = layer.ntype2idx["tgt"], layer.ntype2idx[ntypes[-1]]
        rel = layer.etype2idx["intra"]
        Wk, bk = _fold(layer.k_linears[last].weight, layer.k_linears[last].bias, layer.relation_att[rel], pri[rel] / layer.sqrt_dk)
        P["tilde"] = _Weight(torch.cat([layer.q_linears[t].weight.detach().float(), Wk], 0),
                             torch.cat([layer.q_linears[t].bias.detach().float(), bk], 0), math_mode)
    layer._hprep, layer._hprep_key = P, key
    return P


def _attn_dtype(math_mode: int):
    return torch.bfloat16 if math_mode == L.MATH_BF16 else torch.float32


def _project(P, nt: str, x, out=None):
    from .hgt import _lin, as_act
    return _lin(as_act(x, P["math"]), P["w"][nt], P["math"], out=out, out_dtype=None if out is not None else _attn_dtype(P["math"]))


def _out(P, nt: str, t_agg: torch.Tensor, h_res):
    """LayerNorm(A-linear(t) + h) (hgt.py:399-405); the residual add is fused into the LayerNorm kernel."""
    from .hgt import _lin, act_dtype, as_act
    o = _lin(as_act(t_agg, P["math"]), P["a"][nt], P["math"])
    g, b, eps = P["ln"][nt]
    return ops.layernorm(o, g, b, eps, out_dtype=act_dtype(P["math"]), residual=h_res)


def _dst_groups(etypes: List[CET]) -> Dict[str, List[CET]]:
    by_dst: Dict[str, List[CET]] = {}
    for c in etypes:
        by_dst.setdefault(c[2], []).append(c)
    return by_dst


# --------------------------------------------------------------------------------------------------------------------------
# HGTLayer.forward on a general heterograph (+ two_stream)
# --------------------------------------------------------------------------------------------------------------------------
def layer_forward(layer, G: HeteroGraph, h: Dict[str, object], etypes: Optional[List[CET]], math_mode: int) -> Dict[str, object]:
    from .hgt import as_act
    etypes = list(etypes or G.canonical_etypes)
    ntypes = list(G.ntypes)
    P = _prepare(layer, math_mode, ntypes, etypes, layer.two_stream)
    d, H = P["d"], layer.n_heads
    dev = layer.relation_pri.device
    hx = {nt: as_act(h[nt], math_mode) for nt in ntypes}
    proj = {nt: _project(P, nt, hx[nt]) for nt in ntypes}             # hgt.py:315-322 + :347-348, one GEMM per node type
    col = lambda nt, name: proj[nt][:, P["cols"][nt][name]:P["cols"][nt][name] + d]
    by_dst = _dst_groups(etypes)
    new_h: Dict[str, object] = {}
    for nt in ntypes:
        if nt not in by_dst:
            raise KeyError("t")                                       # G.nodes[nt].data['t'] does not exist in the reference (:399)
        t_agg = torch.empty((G.num_nodes(nt), d), device=dev, dtype=torch.float32)
        group = by_dst[nt]
        for i, (s, r, t) in enumerate(group):                         # :350-358, :383-386 (cross_reducer='mean')
            indptr, indices = G.csr((s, r, t))
            ops.edge_attn(col(t, "q"), col(s, ("k", r)), col(s, ("v", r)), indptr, indices, H, t_agg, out_scale=1.0 / len(group),
                          accumulate=i > 0, tag=f"hetero:{s}-{r}-{t}")
        new_h[nt] = _out(P, nt, t_agg, hx[nt])
    if layer.two_stream:
        tt, st = ("tgt", "intra", "tgt"), ("src", "intra", "tgt")
        for c in (tt, st):
            if c not in etypes:
                raise KeyError(c)                                     # G.update_all(..., etype=c) on a missing edge type (:390-393)
        h_tilde = as_act(h["tgt_tilde"] if "tgt_tilde" in h else h["tgt"], math_mode)
        from .hgt import _lin
        qk = _lin(h_tilde, P["tilde"], math_mode, out_dtype=_attn_dtype(math_mode))          # Q~ | K~'
        n_t = G.num_nodes("tgt")
        t_agg = torch.empty((n_t, d), device=dev, dtype=torch.float32)
        # tgt-intra-tgt: self loops score against the query stream's own key (:371-376), values stay V' (:388 todo)
        k_ext = torch.cat([col("tgt", ("k", "intra")), qk[:, d:]], 0)
        v_ext = torch.cat([col("tgt", ("v", "intra"))] * 2, 0)
        indptr, indices = G.csr(tt, self_loops_shifted=True)
        ops.edge_attn(qk[:, :d], k_ext, v_ext, indptr, indices, H, t_agg, out_scale=0.5, tag="hetero:tilde-tt")
        indptr, indices = G.csr(st)                                   # src-intra-tgt: no leak possible (:369-370)
        ops.edge_attn(qk[:, :d], col("src", ("k", "intra")), col("src", ("v", "intra")), indptr, indices, H, t_agg, out_scale=0.5,
                      accumulate=True, tag="hetero:tilde-st")
        new_h["tgt_tilde"] = _out(P, "tgt", t_agg, h_tilde)           # :407-416
    return new_h


# --------------------------------------------------------------------------------------------------------------------------
# incremental state (fairseq/incremental_decoding_utils.py:12-46) and HGTLayer.infer
# --------------------------------------------------------------------------------------------------------------------------
class IncrementalState:
    """Mixin: per-module keys inside the caller's incremental_state dict (FairseqIncrementalState)."""

    def init_incremental_state(self):
        self._incremental_state_id = str(uuid.uuid4())

    def _get_full_incremental_state_key(self, key: str) -> str:
        return "{}.{}".format(self._incremental_state_id, key)

    def get_incremental_state(self, incremental_state, key: str):
        full_key = self._get_full_incremental_state_key(key)
        if incremental_state is None or full_key not in incremental_state:
            return None
        return incremental_state[full_key]

    def set_incremental_state(self, incremental_state, key: str, value):
        if incremental_state is not None:
            incremental_state[self._get_full_incremental_state_key(key)] = value
        return incremental_state


def layer_infer(layer, G: HeteroGraph, h: Dict[str, object], etypes: Optional[List[CET]], incremental_state: dict,
                math_mode: int) -> Dict[str, object]:
    """HGTLayer.infer (hgt.py:81-297): update the latest tgt node of every block only.  h['tgt'] = [bsz, d] features of the
    current position; the buffer under "prev_g" holds `step` [bsz] and, per node type, `{ntype}_qkv` [bsz, nodes / bsz, width]
    (Q | K'_r | V'_r ..., relation transforms applied) and `{ntype}_out_feat`; every entry is batch-major so that
    reorder_incremental_state is an index_select."""
    from .hgt import as_act, act_dtype, as_float
    assert not layer.two_stream, "not supported yet"                  # hgt.py:89
    etypes = list(etypes or G.canonical_etypes)
    ntypes = list(G.ntypes)
    P = _prepare(layer, math_mode, ntypes, etypes, False)
    d, H = P["d"], layer.n_heads
    dev = layer.relation_pri.device
    saved = layer.get_incremental_state(incremental_state, "prev_g") or {}
    x_t = as_act(h["tgt"], math_mode)
    bsz = x_t.shape[0]
    first = not saved
    for nt in ntypes:
        if G.num_nodes(nt) % bsz:
            raise ValueError(f"{G.num_nodes(nt)} '{nt}' nodes do not divide into {bsz} blocks")        # .view(bsz, n // bsz, -1), :123
    n_tgt = G.num_nodes("tgt")
    per = n_tgt // bsz
    if per != MAX_LEN:
        raise ValueError(f"the incremental graph must hold max_len = {MAX_LEN} tgt nodes per block (hgt.py:93), got {per}")
    step = 0 if first else int(saved["step_host"]) + 1
    if step >= per:
        raise IndexError(f"step {step} beyond max_len = {per}")
    buf = {"step": torch.full((bsz,), step, dtype=torch.long, device=dev), "step_host": step}
    idx = (torch.arange(bsz, dtype=torch.int32) * per + step).to(dev)                # tgt_idxs, :93-94 / :209-210
    adt = _attn_dtype(math_mode)
    width = {nt: P["w"][nt].W.shape[0] for nt in ntypes}
    if first:
        for nt in ntypes:
            n = G.num_nodes(nt)
            if nt == "tgt":                                           # :109-116: zero rows except the current position
                cache = torch.zeros((bsz, per, width[nt]), device=dev, dtype=adt)
            else:
                cache = torch.empty((bsz, n // bsz, width[nt]), device=dev, dtype=adt)
                _project(P, nt, h[nt], out=cache.view(n, width[nt]))
            buf[f"{nt}_qkv"] = cache
    else:
        for nt in ntypes:
            buf[f"{nt}_qkv"] = saved[f"{nt}_qkv"]
    cache_t = buf["tgt_qkv"]
    _project(P, "tgt", x_t, out=cache_t[:, step, :])                  # the GEMM writes the cache rows of this position in place
    flat = {nt: buf[f"{nt}_qkv"].view(-1, width[nt]) for nt in ntypes}
    col = lambda nt, name: flat[nt][:, P["cols"][nt][name]:P["cols"][nt][name] + d]
    by_dst = _dst_groups(etypes)
    new_h: Dict[str, object] = {}
    for nt in ntypes:
        if nt == "tgt":
            group = by_dst["tgt"]
            t_agg = torch.empty((bsz, d), device=dev, dtype=torch.float32)
            q_now = cache_t[:, step, :d]                              # compact rows of the bsz current destinations
            for i, (s, r, t) in enumerate(group):
                indptr, indices = G.csr((s, r, t))
                ops.edge_attn(q_now, col(s, ("k", r)), col(s, ("v", r)), indptr, indices, H, t_agg, dst_ids=idx, n_dst=bsz,
                              out_scale=1.0 / len(group), accumulate=i > 0, tag=f"infer:{s}-{r}-{t}")
            new_h[nt] = _out(P, nt, t_agg, x_t)
        elif first:                                                   # every other node type: full update on the first step only
            if nt not in by_dst:
                raise KeyError("t")
            n = G.num_nodes(nt)
            group = by_dst[nt]
            t_agg = torch.empty((n, d), device=dev, dtype=torch.float32)
            for i, (s, r, t) in enumerate(group):
                indptr, indices = G.csr((s, r, t))
                ops.edge_attn(col(t, "q"), col(s, ("k", r)), col(s, ("v", r)), indptr, indices, H, t_agg,
                              out_scale=1.0 / len(group), accumulate=i > 0, tag=f"infer:{s}-{r}-{t}")
            o = as_float(_out(P, nt, t_agg, as_act(h[nt], math_mode)))
            buf[f"{nt}_out_feat"] = o.view(bsz, n // bsz, d)
            new_h[nt] = o
        else:                                                         # :291-294: cached output features
            buf[f"{nt}_out_feat"] = saved[f"{nt}_out_feat"]
            new_h[nt] = buf[f"{nt}_out_feat"].view(-1, d)
    layer.set_incremental_state(incremental_state, "prev_g", buf)
    return new_h


def reorder_incremental_state(layer, incremental_state, new_order: torch.Tensor):
    """hgt.py:422-438: every tensor of the buffer is batch-major."""
    buf = layer.get_incremental_state(incremental_state, "prev_g")
    if buf is not None:
        for k, v in buf.items():
            if isinstance(v, torch.Tensor):
                buf[k] = v.index_select(0, new_order.to(v.device))
        layer.set_incremental_state(incremental_state, "prev_g", buf)
    return incremental_state
