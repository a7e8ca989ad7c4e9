"""Product-quantisation codec -- host-side mirror of knn/pq_wrapper.py:87-203 (TorchPQCodec) with
the same buffer names (`A`, `b`, `centroids_torch`, `norm2_centroids_torch`, `sdc_table_torch`;
fairseq_cli/convert_ckpt.py:40-45 injects exactly these under decoder.tgt_quantizer.*), so the codec
loads from the checkpoint and faiss is not needed at evaluation time.

decode() runs on libgnnlm_sm100.so: TMA-staged codebook gather (pq_decode.cu) + the OPQ inverse
rotation as a GEMM with W = A^T."""
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops


class TorchPQCodec(nn.Module):
    def __init__(self, centroids=None, A=None, b=None, index=None, metric: str = "ip"):
        """centroids [M, 256, dsub]; optional OPQ pre-transform (A [d_out, d_in], b [d_out] or empty).
        `index` (a faiss IndexPreTransform / IndexPQ) is accepted for signature parity when faiss is
        importable; the arrays are extracted exactly as pq_wrapper.py:20-37 does."""
        super().__init__()
        if index is not None:
            centroids, A, b = _arrays_from_faiss(index)
        centroids = torch.as_tensor(np.asarray(centroids), dtype=torch.float32)
        assert centroids.dim() == 3 and centroids.shape[1] == 256, "8-bit PQ expected (pq_wrapper.py:36)"
        self.metric = metric
        self.pre_torch = A is not None
        if self.pre_torch:
            self.register_buffer("A", torch.as_tensor(np.asarray(A), dtype=torch.float32))
            self.register_buffer("b", torch.as_tensor(np.asarray(b if b is not None else np.zeros(0)), dtype=torch.float32))
        self.register_buffer("centroids_torch", centroids)
        self.register_buffer("norm2_centroids_torch", (centroids ** 2).sum(2))
        if metric == "l2":
            sdc = -torch.sqrt(((centroids.unsqueeze(2) - centroids.unsqueeze(1)) ** 2).sum(3))
        else:
            sdc = torch.matmul(centroids, centroids.transpose(1, 2))
        self.register_buffer("sdc_table_torch", sdc)
        self._rot = None
        self._rot_key = None
        self._split = None
        self._split_key = None
        self._q8cb, self._q8cb_key = None, None
        self._rot32, self._rot32_key = None, None

    @property
    def M(self):
        return self.centroids_torch.shape[0]

    @property
    def dsub(self):
        return self.centroids_torch.shape[2]

    def _rotation(self, math_mode):
        """W = A^T so that x @ A == linear(x, W) (pq_wrapper.py:202)."""
        from .hgt import _Weight
        key = (math_mode, self.A.device, int(self.A._version))
        if self._rot is None or self._rot_key != key:
            self._rot, self._rot_key = _Weight(self.A.t().contiguous(), None, math_mode), key
        return self._rot

    def _split_codebook(self):
        """(hi, lo) fp16 halves of `centroid - b` [M, 256, dsub]: the split-fp16 activation format applied to the
        codebook once, so that decoding into that format needs no arithmetic (same values as splitting after the lookup)."""
        cen = self.centroids_torch
        key = (cen.device, int(cen._version), int(self.b._version) if self.pre_torch else -1)
        if self._split is None or self._split_key != key:
            v = cen
            if self.pre_torch and self.b.numel() > 0:
                v = cen - self.b.view(self.M, 1, self.dsub)
            v = v.clamp(-65504.0, 65504.0)
            hi = v.half()
            lo = (v - hi.float()).half()
            self._split, self._split_key = (hi.contiguous(), lo.contiguous()), key
        return self._split

    def rotation_weight(self) -> torch.Tensor:
        """rot = A^T [d, d_dec] fp32 (x @ A == x @ rot^T), a persistent tensor: HGTLayer.prepare folds it into layer 0 and keys
        its cache on it."""
        key = (self.A.device, int(self.A._version))
        if self._rot32 is None or self._rot32_key != key:
            self._rot32, self._rot32_key = self.A.detach().t().contiguous().float(), key
        return self._rot32

    def _q8_codebook(self):
        """e4m3 companion of the split codebook, [M, 256, 16] bytes (8 B hi8 | 8 B lo8 per centroid)."""
        hi, lo = self._split_codebook()
        key = (hi.data_ptr(), lo.data_ptr())
        if self._q8cb is None or self._q8cb_key != key:
            both = ops.Split(torch.cat([hi.reshape(-1, 8), lo.reshape(-1, 8)], 1).contiguous(), 8)
            self._q8cb, self._q8cb_key = ops.to_q8(both).q8.view(self.M, 256, 16), key
        return self._q8cb

    @torch.no_grad()
    def gather_decode(self, codes_table: torch.Tensor, rows: torch.Tensor, *, row_ids: Optional[torch.Tensor] = None,
                      n_cap: Optional[int] = None, n_dev: Optional[torch.Tensor] = None,
                      math_mode: int = L.MATH_FP32_SIMT, rotate: bool = True, hi_only: bool = False, q8: bool = True) -> torch.Tensor:
        """Fused `quant_neighbor_feats[rows]` + decode: codes_table [N_d, M] uint8 in HBM.
        rotate=False (MATH_F16F8, dsub == 8): the caller folds `x @ A` into its own weights (rotation_weight()) -- returns the
        un-rotated features (with their e4m3 companion unless q8=False), as fp16 hi + companion only when hi_only."""
        from .hgt import act_dtype, gemm_act
        b = self.b if self.pre_torch and self.b.numel() > 0 else None
        act = act_dtype(math_mode)
        if not rotate:
            assert math_mode == L.MATH_F16F8 and self.dsub == 8 and self.pre_torch
            hi, lo = self._split_codebook()
            if hi_only:
                return ops.pq_gather_decode_hiq8(codes_table, hi, self._q8_codebook(), rows, row_ids=row_ids, n_cap=n_cap, n_dev=n_dev)
            return ops.pq_gather_decode_presplit(codes_table, hi, lo, rows, row_ids=row_ids, n_cap=n_cap, n_dev=n_dev, q8=q8)
        if act == ops.SPLIT and self.dsub == 8:        # pre-split codebook: pure byte movement (pq_decode_presplit_kernel)
            hi, lo = self._split_codebook()
            n = (row_ids.shape[0] if row_ids is not None else rows.shape[0]) if n_cap is None else n_cap
            x = ops.pq_gather_decode_presplit(codes_table, hi, lo, rows, row_ids=row_ids, n_cap=n_cap, n_dev=n_dev,
                                              q8=self.pre_torch and gemm_act(math_mode, n, self.M * 8) == ops.SPLIT_Q8)
        else:
            x, _, _ = ops.pq_gather_decode(codes_table, self.centroids_torch, rows, bias=b, row_ids=row_ids, n_cap=n_cap,
                                           n_dev=n_dev, out_dtype=act)
        if self.pre_torch:
            from .hgt import _lin, with_q8
            x = _lin(with_q8(x, math_mode, n_dev), self._rotation(math_mode), math_mode, m_dev=n_dev, out_dtype=act)
        return x

    @torch.no_grad()
    def encode(self, x: torch.Tensor, math_mode: int = L.MATH_FP32_SIMT) -> torch.Tensor:
        """x [n, D] fp32 -> codes [n, M] uint8 (pq_wrapper.py:131-167): OPQ pre-rotation `x @ A.T + b`, then the
        per-subspace nearest centroid under ||c||^2 - 2 x.c."""
        x = x.float().contiguous()
        if self.pre_torch:
            b = self.b if self.b.numel() > 0 else None
            x = ops.linear(x, self.A.contiguous(), b, math=L.MATH_FP32_SIMT)       # F.linear(x, A) == x @ A.T
        n = x.shape[0]
        codes = torch.empty((n, self.M), device=x.device, dtype=torch.uint8)
        L.call("gnnlm_pq_encode", L.ptr(x), x.stride(0), n, self.M, self.dsub, L.ptr(self.centroids_torch),
               L.ptr(self.norm2_centroids_torch), L.ptr(codes), L.stream_ptr())
        return codes

    @torch.no_grad()
    def decode(self, codes: torch.Tensor, math_mode: int = L.MATH_FP32_SIMT) -> torch.Tensor:
        """codes [n, M] uint8 -> [n, M*dsub] (pq_wrapper.py:169-203)."""
        n, MM = codes.shape
        assert MM == self.M, f"input codes have {MM} subspace, but quantizer have {self.M} subspace"
        rows = torch.arange(n, device=codes.device, dtype=torch.int64)
        x = self.gather_decode(codes.contiguous(), rows, math_mode=math_mode)
        return x.float() if isinstance(x, ops.Split) else x


def _arrays_from_faiss(index):
    import faiss  # noqa: only when the caller really has a faiss index
    A = b = None
    if isinstance(index, faiss.IndexPreTransform):
        vt = faiss.downcast_VectorTransform(index.chain.at(0))
        b = faiss.vector_to_array(vt.b)
        A = faiss.vector_to_array(vt.A).reshape(vt.d_out, vt.d_in)
        index = faiss.downcast_index(index.index)
    pq = index.pq
    cen = faiss.vector_to_array(pq.centroids).reshape(pq.M, pq.ksub, pq.dsub)
    return cen, A, b
