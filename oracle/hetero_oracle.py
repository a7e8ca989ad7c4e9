"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain torch) of the reference HGT layer on a GENERAL heterograph, of its
`two_stream` query stream and of its incremental `infer()` (fairseq/models/hgt.py:81-297, 299-420, 494-513).  Only tests/ may
import this; the product path never does.

Pinned by tests/golden/hgt_{hetero4, two_stream, infer_*}.npz, produced by EXECUTING the reference module under the DGL stub of
tests/golden/make_golden.py (DGL's documented semantics restated: apply_edges / edge_softmax / multi_update_all(cross_reducer=
'mean') / update_all) -- DGL itself is absent here, so parity at that boundary is "unpinned", as for oracle/model_oracle.py.

Quirk Q11 (two_stream): hgt.py:376 reads srcdata['k_tilde'], which nothing assigns -- the unmodified reference raises on any
('tgt','intra','tgt') edge set (make_golden.py asserts the KeyError).  Position: k_tilde = the relation-transformed keys of the
query stream (tgt_tilde_k, hgt.py:330), the one-statement fix the fixture applies; that term is therefore pinned to the FIXED
reference only.
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

CET = Tuple[str, str, str]


def _edge_softmax(score: torch.Tensor, dst: torch.Tensor, n_dst: int) -> torch.Tensor:
    """dgl.ops.edge_softmax(norm_by='dst'): softmax over the in-edges of every destination, per head."""
    mx = torch.full((n_dst,) + score.shape[1:], -float("inf"), dtype=score.dtype, device=score.device)
    mx = mx.scatter_reduce(0, dst.view(-1, 1).expand_as(score), score, "amax", include_self=True)
    ex = torch.exp(score - mx[dst])
    den = torch.zeros_like(mx).index_add_(0, dst, ex)
    return ex / den[dst]


def _aggregate(v, att, src, dst, n_dst):
    m = v[src] * att.unsqueeze(-1)                                    # fn.u_mul_e
    return torch.zeros((n_dst,) + m.shape[1:], dtype=m.dtype, device=m.device).index_add_(0, dst, m)      # fn.sum


class _Layer:
    def __init__(self, sd, prefix, ntype2idx, etype2idx, n_heads):
        self.sd, self.p, self.nid, self.rid, self.H = sd, prefix, ntype2idx, etype2idx, n_heads

    def lin(self, name, ntype, x):
        return F.linear(x, self.sd[f"{self.p}{name}_linears.{self.nid[ntype]}.weight"], self.sd[f"{self.p}{name}_linears.{self.nid[ntype]}.bias"])

    def rel(self, name, etype):
        return self.sd[f"{self.p}relation_{name}"][self.rid[etype]]

    def out(self, ntype, t, h_res):
        o = self.lin("a", ntype, t) + h_res                          # hgt.py:401-403 (dropout is identity in eval)
        d = o.shape[-1]
        return F.layer_norm(o, (d,), self.sd[f"{self.p}norms.{self.nid[ntype]}.weight"], self.sd[f"{self.p}norms.{self.nid[ntype]}.bias"])


def hgt_layer_hetero(sd, prefix: str, h: Dict[str, torch.Tensor], edges: Dict[CET, Tuple[torch.Tensor, torch.Tensor]],
                     num_nodes: Dict[str, int], ntype2idx: Dict[str, int], etype2idx: Dict[str, int], n_heads: int,
                     etypes: Optional[List[CET]] = None, two_stream: bool = False) -> Dict[str, torch.Tensor]:
    """hgt.py:299-420 for arbitrary node / edge types.  `h` may carry 'tgt_tilde' (two_stream, layers >= 1)."""
    Lr = _Layer(sd, prefix, ntype2idx, etype2idx, n_heads)
    ntypes = sorted(num_nodes)                                        # G.ntypes (DGL sorts node types)
    etypes = list(etypes or edges.keys())
    d = h[ntypes[0]].shape[1]
    dk = d // n_heads
    view = lambda x: x.view(-1, n_heads, dk)
    K = {t: view(Lr.lin("k", t, h[t])) for t in ntypes}              # :320-322
    V = {t: view(Lr.lin("v", t, h[t])) for t in ntypes}
    Q = {t: view(Lr.lin("q", t, h[t])) for t in ntypes}
    if two_stream:                                                    # :324-330 (k_linear / v_linear there are the loop's last: 'tgt' sorts last)
        h_tilde = h["tgt_tilde"] if "tgt_tilde" in h else h["tgt"]
        last = ntypes[-1]
        Qt, Kt = view(Lr.lin("q", "tgt", h_tilde)), view(Lr.lin("k", last, h_tilde))
    agg: Dict[str, List[torch.Tensor]] = {}
    tilde: Dict[CET, torch.Tensor] = {}
    for (s, r, t) in etypes:
        src, dst = edges[(s, r, t)]
        k = torch.einsum("bij,ijk->bik", K[s], Lr.rel("att", r))     # :347
        v = torch.einsum("bij,ijk->bik", V[s], Lr.rel("msg", r))     # :348
        pri = Lr.rel("pri", r) / math.sqrt(dk)
        att = _edge_softmax((Q[t][dst] * k[src]).sum(-1) * pri, dst, num_nodes[t])       # :354-356
        agg.setdefault(t, []).append(_aggregate(v, att, src, dst, num_nodes[t]))          # :383-386
        if two_stream and (s, r, t) in (("tgt", "intra", "tgt"), ("src", "intra", "tgt")):    # :360-381
            score = (Qt[dst] * k[src]).sum(-1)
            if (s, r, t) == ("tgt", "intra", "tgt"):                  # self loops see the query stream's own key (Q11)
                k_tilde = torch.einsum("bij,ijk->bik", Kt, Lr.rel("att", r))
                loop = src == dst
                score = torch.where(loop.unsqueeze(-1), (Qt[dst] * k_tilde[src]).sum(-1), score)
            tilde[(s, r, t)] = _aggregate(v, _edge_softmax(score * pri, dst, num_nodes[t]), src, dst, num_nodes[t])   # :388-394
    new_h = {}
    for t in ntypes:
        if t not in agg:
            raise KeyError("t")                                       # G.nodes[t].data['t'] does not exist (:399)
        new_h[t] = Lr.out(t, torch.stack(agg[t], 0).mean(0).view(-1, d), h[t])            # cross_reducer='mean'
        if t == "tgt" and two_stream:                                 # :407-416
            tt = (tilde[("tgt", "intra", "tgt")] + tilde[("src", "intra", "tgt")]).view(-1, d) / 2
            new_h["tgt_tilde"] = Lr.out("tgt", tt, h["tgt_tilde"] if "tgt_tilde" in h else h["tgt"])
    return new_h


def hgt_forward_hetero(sd, feats: Dict[str, torch.Tensor], edges, num_nodes, ntype2idx, etype2idx, n_heads, n_layers,
                       prefix: str = "", etypes=None, two_stream: bool = False):
    """hgt.py:494-513 (in_dim == hidden_dim == out_dim)."""
    h = dict(feats)
    for l in range(n_layers):
        h = hgt_layer_hetero(sd, f"{prefix}gcs.{l}.", h, edges, num_nodes, ntype2idx, etype2idx, n_heads, etypes, two_stream)
    return h


def hgt_layer_infer(sd, prefix: str, h: Dict[str, torch.Tensor], edges, num_nodes, ntype2idx, etype2idx, n_heads, etypes,
                    state: dict, max_len: int = 512) -> Dict[str, torch.Tensor]:
    """HGTLayer.infer (hgt.py:81-297): `state` is this layer's buffer ({} before the first step).  h['tgt'] = [bsz, d] features of
    the current position only; the other types' features are complete."""
    Lr = _Layer(sd, prefix, ntype2idx, etype2idx, n_heads)
    ntypes = sorted(num_nodes)
    bsz, d = h["tgt"].shape
    dk = d // n_heads
    first = not state
    step = 0 if first else int(state["step"][0]) + 1                  # :92 / :208
    idx = torch.full((bsz,), step, dtype=torch.long) + torch.arange(bsz) * max_len      # :93-94 / :209-210
    if first:
        state["step"] = torch.zeros(bsz, dtype=torch.long)
        for t in ntypes:                                              # :103-123
            for a in "kqv":
                x = Lr.lin(a, t, h[t])
                if t == "tgt":
                    full = torch.zeros(num_nodes[t], d, dtype=x.dtype)
                    full[idx] = x
                    x = full
                state[f"{t}_{a}"] = x.view(bsz, num_nodes[t] // bsz, -1)
    else:
        state["step"] = state["step"] + 1
        for a in "kqv":                                               # :217-233: cached rows, the current position overwritten
            state[f"tgt_{a}"].view(-1, d)[idx] = Lr.lin(a, "tgt", h["tgt"])
    agg: Dict[str, List[torch.Tensor]] = {}
    for (s, r, t) in etypes:
        if not first and t != "tgt":                                  # :236-237
            continue
        src, dst = edges[(s, r, t)]
        k = torch.einsum("bij,ijk->bik", state[f"{s}_k"].view(-1, n_heads, dk), Lr.rel("att", r))
        v = torch.einsum("bij,ijk->bik", state[f"{s}_v"].view(-1, n_heads, dk), Lr.rel("msg", r))
        q = state[f"{t}_q"].view(-1, n_heads, dk)
        att = _edge_softmax((q[dst] * k[src]).sum(-1) * Lr.rel("pri", r) / math.sqrt(dk), dst, num_nodes[t])
        agg.setdefault(t, []).append(_aggregate(v, att, src, dst, num_nodes[t]))
    new_h = {}
    for t in ntypes:
        if t == "tgt":                                                # :177-178,186-189 / :274-290
            out = Lr.out(t, torch.stack(agg[t], 0).mean(0).view(-1, d)[idx], h[t])
            if first:
                state["tgt_out_feat"] = torch.zeros(bsz, num_nodes[t] // bsz, d, dtype=out.dtype)
            state["tgt_out_feat"].view(-1, d)[idx] = out
        elif first:
            out = Lr.out(t, torch.stack(agg[t], 0).mean(0).view(-1, d), h[t])
            state[f"{t}_out_feat"] = out.view(bsz, num_nodes[t] // bsz, -1)
        else:
            out = state[f"{t}_out_feat"].view(-1, d)                  # :291-294
        new_h[t] = out
    return new_h


def reorder_state(state: dict, new_order: torch.Tensor) -> dict:
    """reorder_incremental_state (hgt.py:422-438): every buffer is [bsz, ...]."""
    return {k: v.index_select(0, new_order) for k, v in state.items()}
