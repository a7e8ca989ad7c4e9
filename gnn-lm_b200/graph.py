"""TokenGraph: the per-batch kNN token graph as plain CSR arrays in HBM.

Stands where the reference passes a batched `dgl.DGLHeteroGraph` (built by
GraphTokenBlockDataset.new_build_graph, fairseq/data/token_block_dataset.py:338-412, batched by
dgl.batch, fairseq/data/monolingual_dataset.py:261, moved by utils.move_to_cuda,
fairseq/utils.py:43-67).  Node numbering equals the reference's: tgt id = b*L + t; ntgt ids in
creation order.  The ('tgt','intra','tgt') edges are implicit (causal inside each block) unless
`materialise_tt()` is called for a parity check.
"""
from typing import Optional

import torch

from . import _lib as L

ETYPES = [("tgt", "intra", "tgt"), ("ntgt", "inter", "tgt"), ("ntgt", "intra", "ntgt")]


class _NodeView:
    def __init__(self):
        self.data = {}


class TokenGraph:
    """Device-resident graph of one batch of B blocks x L target tokens with k neighbours each."""

    canonical_etypes = ETYPES
    ntypes = ["ntgt", "tgt"]

    def __init__(self, B: int, L_: int, k: int, left_ctx: int, right_ctx: int, intra_ctx: int, n_datastore: int):
        self.B, self.L, self.k = B, L_, k
        self.left_ctx, self.right_ctx, self.intra_ctx = left_ctx, right_ctx, intra_ctx
        self.n_datastore = n_datastore
        self.T = B * L_
        self.w = 1 + left_ctx + right_ctx
        self.node_cap = self.T * k * self.w
        self.nodes = {"tgt": _NodeView(), "ntgt": _NodeView()}
        # filled by build()
        self.node_base = self.valid_base = None
        self.ntgt_row = self.ntgt_owner = self.ntgt_dist = None
        self.nn_indptr = self.nn_indices = self.inter_indptr = self.inter_indices = None
        self.tt_indptr = self.tt_indices = None
        self.cluster_nl = None
        self.nbr = self.tgt_pos = None        # inputs kept so that token chunks can be re-assembled (chunked ntgt side)
        self.invalid_ctx = 0
        self._counts_host = None
        self.dedup = False                    # --deprecated builder: one node per distinct datastore row, general CSR
        self._n_ntgt = None

    # ---- device-side counts (no host sync needed by the kernels) ----
    @property
    def n_ntgt_dev(self) -> torch.Tensor:
        return self._n_ntgt if self.dedup else self.node_base[-1:]

    @property
    def n_valid_dev(self) -> torch.Tensor:
        return self.valid_base[-1:]

    def counts(self):
        """(n_ntgt, n_valid) on the host -- synchronises; only tests / full-mode outputs use it."""
        if self._counts_host is None:
            self._counts_host = (int(self.n_ntgt_dev.item()), int(self.valid_base[-1].item()))
        return self._counts_host

    def num_nodes(self, ntype: str) -> int:
        return self.T if ntype == "tgt" else self.counts()[0]

    def to(self, device):
        for name, val in list(vars(self).items()):
            if isinstance(val, torch.Tensor):
                setattr(self, name, val.to(device, non_blocking=True))
        for nv in self.nodes.values():
            nv.data = {k_: v.to(device, non_blocking=True) for k_, v in nv.data.items()}
        return self

    def local_scope(self):
        import contextlib
        return contextlib.nullcontext()

    def materialise_tt(self):
        """Explicit ('tgt','intra','tgt') CSR (parity checks only)."""
        lib = L.load()
        E = lib.gnnlm_graph_tt_num_edges(self.B, self.L, self.intra_ctx)
        dev = self.node_base.device
        self.tt_indptr = torch.empty(self.T + 1, dtype=torch.int32, device=dev)
        self.tt_indices = torch.empty(E, dtype=torch.int32, device=dev)
        L.call("gnnlm_graph_tt_csr", self.B, self.L, self.intra_ctx, L.ptr(self.tt_indptr), L.ptr(self.tt_indices),
               L.stream_ptr())
        return self.tt_indptr, self.tt_indices


def build_token_graph(nbr: torch.Tensor, n_datastore: int, left_ctx: int, right_ctx: int, *,
                      tgt_pos: Optional[torch.Tensor] = None, invalid_ctx: int = 0, intra_ctx: int = 0,
                      with_owner: bool = False, reach: Optional[int] = None, dedup: bool = False) -> TokenGraph:
    """nbr [B, L, k] int64 on the device (= neighbor_offsets[offsets], token_block_dataset.py:309).

    `reach`: keep only context nodes within `reach` chain hops of their centre.  With NL graph layers a tgt node sees its
    neighbour clusters through the centre after NL-1 ntgt-intra-ntgt hops, so nodes further than NL-1 from the centre
    cannot influence any tgt output (the decoder reads tgt rows only, transformer.py:1053): reach = NL-1 gives the same
    tgt features as the reference's full clusters with fewer rows to decode and project (e.g. w = 7 -> 5 at c = 3, NL = 3).
    The ntgt numbering then differs from the reference's, so leave it None when ntgt outputs are wanted."""
    if reach is not None and not dedup:      # de-duplicated context rows can be adjacent to OTHER centres: no pruning there
        left_ctx, right_ctx = min(left_ctx, reach), min(right_ctx, reach)
    assert nbr.is_cuda and nbr.dtype == torch.int64 and nbr.dim() == 3 and nbr.is_contiguous()
    B, Lb, k = nbr.shape
    g = TokenGraph(B, Lb, k, left_ctx, right_ctx, intra_ctx, n_datastore)
    g.nbr, g.tgt_pos, g.invalid_ctx = nbr, tgt_pos, invalid_ctx
    dev = nbr.device
    n = g.T * k
    i32 = dict(dtype=torch.int32, device=dev)
    g.node_base = torch.empty(n + 1, **i32)
    g.valid_base = torch.empty(n + 1, **i32)
    lib = L.load()
    ws_bytes = lib.gnnlm_graph_workspace_bytes(n)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
    if tgt_pos is not None:
        tgt_pos = tgt_pos.reshape(-1).contiguous()
        assert tgt_pos.dtype == torch.int64 and tgt_pos.numel() == g.T
    st = L.stream_ptr()
    L.call("gnnlm_graph_count", L.ptr(nbr), L.ptr(tgt_pos), g.T, k, n_datastore, left_ctx, right_ctx, invalid_ctx,
           L.ptr(g.node_base), L.ptr(g.valid_base), L.ptr(ws), ws.numel(), st)
    cap = g.node_cap
    if dedup:
        return _build_dedup(g, nbr, tgt_pos, n, cap, st)
    g.ntgt_row = torch.empty(cap, dtype=torch.int64, device=dev)
    g.ntgt_dist = torch.empty(cap, **i32)
    g.ntgt_owner = torch.empty(cap, **i32) if with_owner else None
    g.nn_indptr = torch.empty(cap + 1, **i32)
    g.nn_indices = torch.empty(3 * cap, **i32)
    g.inter_indptr = torch.empty(g.T + 1, **i32)
    g.inter_indices = torch.empty(n, **i32)
    g.cluster_nl = torch.empty(n, **i32)
    L.call("gnnlm_graph_fill", L.ptr(nbr), L.ptr(tgt_pos), g.T, k, n_datastore, left_ctx, right_ctx, invalid_ctx,
           L.ptr(g.node_base), L.ptr(g.valid_base), L.ptr(g.ntgt_row), L.ptr(g.ntgt_owner), L.ptr(g.ntgt_dist),
           L.ptr(g.nn_indptr), L.ptr(g.nn_indices), L.ptr(g.inter_indptr), L.ptr(g.inter_indices), L.ptr(g.cluster_nl), st)
    return g


def _build_dedup(g: TokenGraph, nbr, tgt_pos, n: int, cap: int, st) -> TokenGraph:
    """`--deprecated` builder (token_block_dataset.py:414-479) -> gnnlm_graph_dedup; valid_base is already counted."""
    dev = nbr.device
    i32 = dict(dtype=torch.int32, device=dev)
    g.dedup = True
    g.ntgt_row = torch.empty(cap, dtype=torch.int64, device=dev)
    g._n_ntgt = torch.empty(1, **i32)
    g.nn_indptr = torch.empty(cap + 1, **i32)
    g.nn_indices = torch.empty(3 * cap, **i32)
    g.inter_indptr = torch.empty(g.T + 1, **i32)
    g.inter_indices = torch.empty(n, **i32)
    ws_bytes = L.load().gnnlm_graph_dedup_workspace_bytes(g.T, g.k, g.w)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    L.call("gnnlm_graph_dedup", L.ptr(nbr), L.ptr(tgt_pos), g.T, g.L, g.k, g.n_datastore, g.left_ctx, g.right_ctx,
           g.invalid_ctx, L.ptr(g.valid_base), L.ptr(g.ntgt_row), L.ptr(g._n_ntgt), L.ptr(g.nn_indptr), L.ptr(g.nn_indices),
           L.ptr(g.inter_indptr), L.ptr(g.inter_indices), L.ptr(ws), ws.numel(), st)
    g.node_base = None                        # no cluster structure in this mode
    return g


def token_chunk_graph(G: TokenGraph, t0: int, t1: int) -> TokenGraph:
    """The graph of tokens [t0, t1) of the flattened batch on their own (ntgt clusters never cross tokens, so the ntgt
    side of a chunk is independent of the rest; tgt-intra-tgt is not used on chunk graphs)."""
    nbr = G.nbr.view(G.T, G.k)[t0:t1].reshape(1, t1 - t0, G.k)
    pos = None if G.tgt_pos is None else G.tgt_pos.reshape(-1)[t0:t1]
    return build_token_graph(nbr.contiguous(), G.n_datastore, G.left_ctx, G.right_ctx, tgt_pos=pos,
                             invalid_ctx=G.invalid_ctx, intra_ctx=G.intra_ctx)


def unique_centre_graph(G: TokenGraph):
    """(G_u, inv, n_unique): the graph of the DISTINCT centre rows of G's valid (token, neighbour) pairs -- one "token" per
    distinct row with that row as its only neighbour -- and, per compact valid pair of G, the compact centre index in G_u.
    Clusters with the same centre row are identical in every layer (they only exchange messages inside themselves), so the ntgt
    side can run on G_u and be gathered back (HGT.forward_tgt_shared).  No host synchronisation: capacity-sized arrays."""
    assert not G.dedup
    n = G.T * G.k
    dev = G.nbr.device
    uniq = torch.empty(n, dtype=torch.int64, device=dev)
    inv = torch.empty(n, dtype=torch.int32, device=dev)
    n_unique = torch.empty(1, dtype=torch.int32, device=dev)
    ws = torch.empty(L.load().gnnlm_unique_workspace_bytes(n), dtype=torch.uint8, device=dev)
    L.call("gnnlm_unique_centres", L.ptr(G.nbr), L.ptr(G.valid_base), n, L.ptr(uniq), L.ptr(inv), L.ptr(n_unique), L.ptr(ws),
           ws.numel(), L.stream_ptr())
    # validity (missing ids, invalid_ctx) was decided pair by pair in G: every id in uniq is a real row
    G_u = build_token_graph(uniq.view(1, n, 1), G.n_datastore, G.left_ctx, G.right_ctx)
    return G_u, inv, n_unique
