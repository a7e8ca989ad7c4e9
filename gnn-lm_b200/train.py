"""Training forward / backward of the HGT fine-tuning step.

The reference fine-tunes with `--freeze` (fairseq/models/transformer_lm.py:183-186: every parameter without "hgt" in its
name is frozen) under `--criterion adaptive_loss` (fairseq/criterions/adaptive_loss.py:31-83): the loss is the summed
cross-entropy of the adaptive-softmax head and tail clusters on the HGT's tgt outputs, and gradients reach
decoder.hgt_decoder.* only.  Its backward is autograd over DGL ops (fairseq/models/hgt.py:299-420).

Here every differentiable stage is a `torch.autograd.Function` whose forward AND backward are C-ABI kernels of the library:
projections (gnnlm_linear; dX = dY W and dW = dY^T X through the same GEMM on transposed copies, gnnlm_transpose_f32;
biases by gnnlm_colsum_f32), the three edge types' segmented softmax + aggregation (forward gnnlm_hgt_edge_attn /
gnnlm_hgt_causal_attn, backward gnnlm_hgt_edge_attn_bwd), residual + LayerNorm (gnnlm_layernorm / gnnlm_layernorm_bwd), the
centre-row gather (gnnlm_gather_rows / gnnlm_scatter_add_rows) and the adaptive loss (logits by gnnlm_linear, softmax
cross-entropy and its gradient by gnnlm_xent_fwd_bwd).  torch.autograd only chains them and differentiates the fold of the
relation transforms into the projection weights (relation_att / relation_msg / relation_pri -> K' / V' weights: d x d matrices).

Scope: fp32 activations, projections in fp32 FMA, 3xTF32 or 3xFP16 (`f16x3`: pre-split operands, gradients scaled into the fp16
range by a power of two, dW by split-K through one batched launch), graphs of either builder (general CSR kernels for the ntgt
edges, so `--deprecated` graphs train too).  Fast forms, chosen by shape (each checked against the generic kernel and fp64 autograd):
causal edges in GEMM form on the tensor cores from 256-token blocks (forward with the dropout multiplier inside the softmax-split
pass, backward through gnnlm_causal_softmax_bwd_split), the ntgt-intra-ntgt chains one warp per (cluster, head) (forward under
dropout and backward, no atomics) and the inter edges one warp per (token, head) from d_k = 32.  Dropout (model.train()): hgt.py's `drop` on the output
projections (:401), `attn_drop` on the edge-softmax weights (:356, one draw per edge and head) and the adaptive softmax's input /
tail dropouts (adaptive_softmax.py:156,101) are applied with masks that are pure functions of (seed, element)
(gnnlm_dropout_f32, the p_drop / seed arguments of the attention kernels): forward and backward regenerate them, and the tests replay
them through the oracle.  The draws are not torch's generator's, so a run is not sample-for-sample the reference's (neither are
two reference runs on different GPUs); the distribution and the gradient of the masked network are.  The last layer's ntgt
side is skipped as in evaluation (nothing reads it; its parameters get zero gradients in the reference too).
"""
import math
import os
from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import ops
from .graph import TokenGraph


def _st():
    return L.stream_ptr()


def _zeros(*shape, like):
    return torch.zeros(shape, device=like.device, dtype=torch.float32)


def _transpose(x: torch.Tensor, rows_pad: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[rows, cols] fp32 -> [cols, rows_pad] (zero-padded columns): an operand of dW = dY^T X."""
    rows, cols = x.shape
    rows_pad = rows if rows_pad is None else rows_pad
    if out is None:
        out = torch.empty((cols, rows_pad), device=x.device, dtype=torch.float32)
    assert out.shape == (cols, rows_pad) and out.stride(1) == 1
    L.call("gnnlm_transpose_f32", L.ptr(x), x.stride(0), rows, None, cols, L.ptr(out), out.stride(0), rows_pad, _st())
    return out


class _Last:
    """One-entry cache keyed by tensor identity + version: q / k / v (and the inter K' / V') project the SAME activation in
    consecutive calls, forward and backward, so its split-fp16 copies are made once.  The entry keeps the source tensor alive,
    which is what makes the identity check sound."""

    def __init__(self):
        self.src, self.version, self.value = None, -1, None

    def get(self, x: torch.Tensor, make):
        if self.src is not x or self.version != x._version:
            self.src, self.version, self.value = x, x._version, make(x)
        return self.value

    def clear(self):
        self.src, self.version, self.value = None, -1, None


_SPLIT_A, _SPLIT_XT = _Last(), _Last()


def release_caches():
    """Drop the cached operand copies (a step's activations stay referenced until the next step otherwise)."""
    _SPLIT_A.clear()
    _SPLIT_XT.clear()


def _gemm(a: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], mode: int, a_scale: float = 1.0, a_prescale: float = 1.0) -> torch.Tensor:
    """a [M, K] @ w [N, K]^T (+ b), fp32 in / out, in fp32 FMA, 3xTF32 or 3xFP16 (tcgen05; the weight operand split on the fly).
    Row strides may exceed K (padded buffers).  a_scale: `a` has been multiplied by it (the power of two that brings a gradient
    operand into the fp16 range, _pow2_scaled); the product divides it out."""
    if w.stride(1) != 1:
        w = w.contiguous()
    if mode == L.MATH_F16X3:
        hi, lo, sc = ops.split_f16(w.contiguous())
        if a_prescale != 1.0:                             # `a` is still UN-scaled: a_prescale (== a_scale) is applied by the split pass
            assert a_prescale == a_scale and a.shape[1] % 8 == 0 and a.stride(1) == 1
            a = ops.to_split(a, scale=a_prescale)
        elif a.shape[1] % 8 == 0 and a.stride(1) == 1:    # pre-split activation operand: the kernel without the in-loop operand split
            a = _SPLIT_A.get(a, ops.to_split) if a.shape[0] >= SPLIT_K_MIN_ROWS else ops.to_split(a)
        return ops.linear(a, hi, b, W_lo=lo, w_scale=sc * a_scale, math=mode)
    assert a_scale == 1.0
    if mode == L.MATH_TF32X3:
        hi, lo = ops.split_tf32(w)
        return ops.linear(a, hi, b, W_lo=lo, math=mode)
    return ops.linear(a, w, b, math=mode)


def _pow2_scaled(x: torch.Tensor, target: float = 1024.0):
    """(s x, s) with s the power of two that brings max |x| just below `target`: gradients can be far below the fp16 range the
    3xFP16 products split their operands into (one host read of the maximum per call)."""
    amax = float(x.abs().max())
    if not math.isfinite(amax):
        raise FloatingPointError("non-finite gradient")
    if amax == 0.0:
        return x, 1.0
    s = 2.0 ** math.floor(math.log2(target / amax))
    y = _zeros(*x.shape, like=x)
    L.call("gnnlm_axpy_f32", L.ptr(y), y.stride(0), L.ptr(x), x.stride(0), x.shape[0], None, x.shape[1], float(s), _st())
    return y, s


SPLIT_K_MIN_ROWS = 32768


def _dw_split_k(g: torch.Tensor, x: torch.Tensor, g_scale: float, n_split: int = 8) -> torch.Tensor:
    """dW = g^T x for a tall pair (g [R, N], x [R, K], R >> N, K) by split-K in the 3xFP16 arithmetic: the R rows are cut into
    `n_split` chunks, ONE batched launch (gnnlm_linear_batched_f16x3) computes the partial products and they are summed.  A
    [N, K] output alone is 32 tiles for N = K = 1024 -- a fifth of the SMs -- which is what bounded dW before (139 TFLOP/s).
    The operands come straight from the row-major tensors (gnnlm_transpose_split_f16: transpose + split, g scaled by the power
    of two g_scale on the way, divided out by the product)."""
    R, N = g.shape
    K = x.shape[1]
    Kc = ((R + n_split - 1) // n_split + 31) // 32 * 32
    S = (R + Kc - 1) // Kc
    dev = g.device
    f16 = dict(device=dev, dtype=torch.float16)
    a = torch.empty((S, N, 2 * Kc), **f16)

    def x_operand(x_):
        hi_, lo_ = torch.empty((S, K, Kc), **f16), torch.empty((S, K, Kc), **f16)
        for b in range(S):
            r0, r1 = b * Kc, min(R, (b + 1) * Kc)
            L.call("gnnlm_transpose_split_f16", L.ptr(x_[r0:r1]), x_.stride(0), r1 - r0, K, 1.0, Kc, 0, L.ptr(hi_[b]), L.ptr(lo_[b]), _st())
        return hi_, lo_, S, Kc
    hi, lo, s_c, kc_c = _SPLIT_XT.get(x, x_operand)
    assert (s_c, kc_c) == (S, Kc)
    for b in range(S):
        r0, r1 = b * Kc, min(R, (b + 1) * Kc)
        L.call("gnnlm_transpose_split_f16", L.ptr(g[r0:r1]), g.stride(0), r1 - r0, N, float(g_scale), Kc, 1, L.ptr(a[b]), None, _st())
    part = torch.empty((S, N, K), device=dev, dtype=torch.float32)
    L.call("gnnlm_linear_batched_f16x3", L.ptr(a), 2 * Kc, N * 2 * Kc, L.ptr(hi), L.ptr(lo), Kc, K * Kc, float(g_scale), None, 0, 0,
           L.ptr(part), K, N * K, S, N, K, Kc, 0, _st(), tag="dw_split_k", work=(N, K, S * Kc))
    dW = part[0]
    for b in range(1, S):
        L.call("gnnlm_axpy_f32", L.ptr(dW), dW.stride(0), L.ptr(part[b]), K, N, None, K, 1.0, _st())
    return dW


def _pow2_scale_of(x: torch.Tensor, target: float = 1024.0) -> float:
    """The power of two that brings max |x| just below `target` (one host read of the maximum)."""
    amax = float(x.abs().max())
    if not math.isfinite(amax):
        raise FloatingPointError("non-finite gradient")
    return 1.0 if amax == 0.0 else 2.0 ** math.floor(math.log2(target / amax))


class _Linear(torch.autograd.Function):
    """y = x W^T + b through gnnlm_linear, forward and backward."""

    @staticmethod
    def forward(ctx, x, W, b, mode):
        x = x.contiguous()
        ctx.save_for_backward(x, W)
        ctx.mode, ctx.has_b = mode, b is not None
        return _gemm(x, W.detach(), None if b is None else b.detach().contiguous(), mode)

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dy = dy.contiguous()
        mode = ctx.mode
        dx = dW = db = None
        split_k = mode == L.MATH_F16X3 and ctx.needs_input_grad[1] and x.shape[0] >= SPLIT_K_MIN_ROWS and x.shape[1] % 8 == 0
        # 3xFP16: dY is scaled into the fp16 range by a power of two INSIDE the pass that writes its operand format (split /
        # transposing split); the products divide it out.  (Widths that are not a multiple of 8 take a scaled copy first.)
        g, s, pre = dy, 1.0, 1.0
        if mode == L.MATH_F16X3:
            s = _pow2_scale_of(dy)
            m_pad8 = (x.shape[0] + 31) // 32 * 32
            if dy.shape[1] % 8 == 0 and m_pad8 % 8 == 0:
                pre = s
            else:
                g, s = _pow2_scaled(dy)
        if ctx.needs_input_grad[0]:
            dx = _gemm(g, _transpose(W.detach().contiguous()), None, mode, s, pre)                     # dX = dY W
        if ctx.needs_input_grad[1]:
            m_pad = (x.shape[0] + 31) // 32 * 32                                                       # k of the product, zero-padded
            if split_k:
                dW = _dw_split_k(dy, x, s)
            else:
                dW = _gemm(_transpose(g, m_pad), _transpose(x, m_pad), None, mode, s, pre)             # dW = dY^T X
        if ctx.has_b and ctx.needs_input_grad[2]:
            db = _zeros(dy.shape[1], like=dy)
            L.call("gnnlm_colsum_f32", L.ptr(dy), dy.stride(0), dy.shape[0], None, dy.shape[1], L.ptr(db), _st())
        return dx, dW, db, None


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ids):
        ctx.save_for_backward(ids)
        ctx.n = x.shape[0]
        return ops.gather_rows(x.contiguous(), ids)

    @staticmethod
    def backward(ctx, dy):
        (ids,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = _zeros(ctx.n, dy.shape[1], like=dy)
        L.call("gnnlm_scatter_add_rows", L.ptr(dx), dx.stride(0), L.ptr(dy), dy.stride(0), L.ptr(ids), dy.shape[0], None, dy.shape[1], _st())
        return dx, None


SITE_STRIDE = 0x632BE59BD9B4E019


def site_seed(seed: int, site: int) -> int:
    """Seed of one dropout site of a step: layer l -> 16 l + {0: tgt-intra-tgt attention, 1: inter, 2: ntgt-intra-ntgt, 3: `drop`
    on the tgt output projection, 4: on the ntgt one}; 1000: adaptive-softmax input, 1001 + i: tail i."""
    return (int(seed) + site * SITE_STRIDE) & 0xFFFFFFFFFFFFFFFF


class _Dropout(torch.autograd.Function):
    """nn.Dropout in training mode with a (seed, element)-addressed mask: y = x m / (1 - p); backward applies the same mask."""

    @staticmethod
    def forward(ctx, x, p, seed):
        x = x.contiguous()
        ctx.p, ctx.seed = p, seed
        y = torch.empty_like(x)
        L.call("gnnlm_dropout_f32", L.ptr(x), x.stride(0), L.ptr(y), y.stride(0), x.shape[0], x.shape[1], float(p), seed, _st())
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        L.call("gnnlm_dropout_f32", L.ptr(dy), dy.stride(0), L.ptr(dx), dx.stride(0), dy.shape[0], dy.shape[1], float(ctx.p), ctx.seed, _st())
        return dx, None, None


def _dropout(x, p, seed):
    return x if p <= 0 else _Dropout.apply(x, p, seed)


def _attn_fwd_train(q, k, v, H, scale, out, accumulate, p, seed, *, indptr=None, indices=None, causal=(0, 0)):
    """Attention forward with dropout on the softmax weights (gnnlm_hgt_edge_attn_train_fwd)."""
    d = q.shape[1]
    L.call("gnnlm_hgt_edge_attn_train_fwd", L.ptr(q), q.stride(0), L.ptr(k), k.stride(0), L.ptr(v), v.stride(0), L.ptr(indptr), L.ptr(indices),
           q.shape[0], causal[0], causal[1], H, d // H, float(scale), int(accumulate), L.ptr(out), out.stride(0), float(p), seed, _st())
    return out


_GEMM_BWD_BUFFERS = {}
# ntgt-intra-ntgt backward without atomics (gnnlm_hgt_edge_attn_bwd_sym: chains + self loops are a symmetric edge set): bit-reproducible
# gradients, but measured SLOWER than the atomic form on the Wiki103 shape (28 vs 21 ms per layer: two passes of uncoalesced
# 128 B-per-lane row reads; profiles/r2_train_probe.log), so it is opt-in (GNNLM_TRAIN_SYM_BWD=1).
SYMMETRIC_NN_BWD = os.environ.get("GNNLM_TRAIN_SYM_BWD", "0") == "1"


def causal_bwd_gemm_supported(d: int, H: int, Lb: int) -> bool:
    """Shape envelope of the GEMM form of the causal-edge backward (GNNLM_TRAIN_GEMM_BWD=0 switches it off for A/B timing)."""
    return os.environ.get("GNNLM_TRAIN_GEMM_BWD", "1") != "0" and Lb % 64 == 0 and Lb >= 256 and (d // H) % 8 == 0


def _causal_bwd_gemm(q, k, v, dout, Lb: int, intra_ctx: int, H: int, scale: float, p: float, seed: int):
    """Backward of out = scale * softmax_causal(Q K'^T) V' on the tensor cores at fp32 parity: five batched 3xFP16 products per
    block (all heads in one launch each) around gnnlm_causal_softmax_bwd_split -- instead of streaming every K' / V' / Q / dOut row
    of the L (L + 1) / 2 edges through L2 three times (gnnlm_hgt_causal_attn_bwd: 115 ms per layer at L = 3072)."""
    T, d = q.shape
    dk_ = d // H
    dev = q.device
    st = L.stream_ptr
    f16 = dict(device=dev, dtype=torch.float16)
    # gradients can be tiny: scale dOut into the fp16 range by a power of two (undone by the closing products' 1 / w_scale)
    amax = float(dout.abs().max())
    if not math.isfinite(amax):
        raise FloatingPointError("non-finite gradient reached the attention backward")
    s = 2.0 ** math.floor(math.log2(4.0 / amax)) if amax > 0 else 1.0
    g = _zeros(T, d, like=dout)
    L.call("gnnlm_axpy_f32", L.ptr(g), g.stride(0), L.ptr(dout), dout.stride(0), T, None, d, float(s), st())
    key = (H, Lb, dev)
    if key not in _GEMM_BWD_BUFFERS:        # split operands: zero above the diagonal once, only lower tiles are ever rewritten
        _GEMM_BWD_BUFFERS.clear()
        _GEMM_BWD_BUFFERS[key] = tuple(torch.zeros((H, Lb, 2 * Lb), **f16) for _ in range(3)) + \
            tuple(torch.empty((H, Lb, Lb), device=dev, dtype=torch.float32) for _ in range(2)) + \
            (torch.empty((H, Lb, 3), device=dev, dtype=torch.float32),)
    dS, WT, dST, S, G, stats = _GEMM_BWD_BUFFERS[key]
    dq, dkk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)

    def a_style(x):                          # [H, Lb, 2 d_k] split rows: A operand of a product contracting over d_k
        o = torch.empty((H, Lb, 2 * dk_), **f16)
        L.call("gnnlm_heads_split_f16", L.ptr(x), x.stride(0), Lb, H, dk_, 1, L.ptr(o), None, st())
        return o

    def w_style(x):                          # hi, lo [H, Lb, d_k]: W operand contracting over d_k
        hi, lo = torch.empty((H, Lb, dk_), **f16), torch.empty((H, Lb, dk_), **f16)
        L.call("gnnlm_heads_split_f16", L.ptr(x), x.stride(0), Lb, H, dk_, 0, L.ptr(hi), L.ptr(lo), st())
        return hi, lo

    def w_transposed(x):                     # hi, lo [H, d_k, Lb]: W operand contracting over the block's tokens
        hi, lo = torch.empty((H, dk_, Lb), **f16), torch.empty((H, dk_, Lb), **f16)
        L.call("gnnlm_heads_transpose_split_f16", L.ptr(x), x.stride(0), Lb, H, dk_, L.ptr(hi), L.ptr(lo), st())
        return hi, lo

    def scores(a, w, out, tag):              # out[h] = a_h w_h^T, tiles above the diagonal skipped
        L.call("gnnlm_linear_batched_f16x3", L.ptr(a), 2 * dk_, Lb * 2 * dk_, L.ptr(w[0]), L.ptr(w[1]), dk_, Lb * dk_, 1.0, None, 0, 0,
               L.ptr(out), Lb, Lb * Lb, H, Lb, Lb, dk_, 1, st(), tag=tag)

    def close(a, w, out, causal, tag):       # out[:, h] = (1 / s) a_h w_h^T, contraction over the block's tokens
        L.call("gnnlm_linear_batched_f16x3", L.ptr(a), 2 * Lb, Lb * 2 * Lb, L.ptr(w[0]), L.ptr(w[1]), Lb, dk_ * Lb, float(s), None, 0, 0,
               L.ptr(out), out.stride(0), dk_, H, Lb, dk_, Lb, causal, st(), tag=tag)

    for b in range(T // Lb):
        rows = slice(b * Lb, (b + 1) * Lb)
        qb, kb, vb, gb = q[rows], k[rows], v[rows], g[rows]
        scores(a_style(qb), w_style(kb), S, "attn_bwd_qk")
        scores(a_style(gb), w_style(vb), G, "attn_bwd_gv")
        L.call("gnnlm_causal_softmax_bwd_split", L.ptr(S), L.ptr(G), Lb, intra_ctx, H, b * Lb, float(scale), float(p), seed, L.ptr(stats),
               L.ptr(dS), L.ptr(WT), L.ptr(dST), st())
        close(dS, w_transposed(kb), dq[rows], 2, "attn_bwd_dq")       # row pair r contracts over k < 256 (r + 1): dS is causal
        close(WT, w_transposed(gb), dv[rows], 0, "attn_bwd_dv")
        close(dST, w_transposed(qb), dkk[rows], 0, "attn_bwd_dk")
    return dq, dkk, dv


def _attn_bwd(q, k, v, dout, H, scale, *, indptr=None, indices=None, causal=(0, 0), atomics: bool = False, p: float = 0.0, seed: int = 0,
              symmetric: bool = False):
    """symmetric: the CSR's edge set is symmetric with n_dst == n_src (ntgt-intra-ntgt of the non-deduplicating builder): the
    atomic-free two-pass form (gnnlm_hgt_edge_attn_bwd_sym)."""
    d = q.shape[1]
    if symmetric and not atomics and causal[0] == 0 and indices is not None and q.shape[0] == k.shape[0]:
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        stats = torch.empty((q.shape[0], H, 3), device=q.device, dtype=torch.float32)
        L.call("gnnlm_hgt_edge_attn_bwd_sym", L.ptr(q), q.stride(0), L.ptr(k), k.stride(0), L.ptr(v), v.stride(0), L.ptr(dout), dout.stride(0),
               L.ptr(indptr), L.ptr(indices), q.shape[0], H, d // H, float(scale), L.ptr(dq), dq.stride(0), L.ptr(dk), dk.stride(0),
               L.ptr(dv), dv.stride(0), L.ptr(stats), float(p), seed, _st())
        return dq, dk, dv
    if causal[0] > 0 and not atomics and causal_bwd_gemm_supported(d, H, causal[0]):
        return _causal_bwd_gemm(q, k, v, dout, causal[0], causal[1], H, scale, p, seed)
    dq = torch.empty_like(q)
    if causal[0] > 0 and not atomics:       # by-destination + by-source passes, no atomics (gnnlm_hgt_causal_attn_bwd)
        dk, dv = torch.empty_like(k), torch.empty_like(v)
        stats = torch.empty((q.shape[0], H, 3), device=q.device, dtype=torch.float32)
        L.call("gnnlm_hgt_causal_attn_bwd", L.ptr(q), q.stride(0), L.ptr(k), k.stride(0), L.ptr(v), v.stride(0), L.ptr(dout), dout.stride(0),
               q.shape[0] // causal[0], causal[0], causal[1], H, d // H, float(scale), L.ptr(dq), dq.stride(0), L.ptr(dk), dk.stride(0),
               L.ptr(dv), dv.stride(0), L.ptr(stats), float(p), seed, _st())
        return dq, dk, dv
    dk, dv = torch.zeros_like(k), torch.zeros_like(v)
    L.call("gnnlm_hgt_edge_attn_bwd", L.ptr(q), q.stride(0), L.ptr(k), k.stride(0), L.ptr(v), v.stride(0), L.ptr(dout), dout.stride(0),
           L.ptr(indptr), L.ptr(indices), None, q.shape[0], None, causal[0], causal[1], H, d // H, float(scale), L.ptr(dq), dq.stride(0),
           L.ptr(dk), dk.stride(0), L.ptr(dv), dv.stride(0), float(p), seed, _st())
    return dq, dk, dv


class _EdgeAttention(torch.autograd.Function):
    """out = softmax-by-destination attention over one CSR edge type (hgt.py:350-358)."""

    @staticmethod
    def forward(ctx, q, k, v, indptr, indices, H, p, seed, symmetric=False, chains=None):
        """chains = (node_base, cluster_nl, n_clusters) of a non-deduplicated TokenGraph: the edge type is ntgt-intra-ntgt and its
        backward runs per chain (gnnlm_hgt_cluster_attn_bwd) instead of over the CSR."""
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        ctx.save_for_backward(q, k, v, indptr, indices)
        ctx.H, ctx.p, ctx.seed, ctx.symmetric, ctx.chains = H, p, seed, symmetric, chains
        out = torch.empty_like(q)
        if p > 0:
            if chains is not None and (q.shape[1] // H) in (32, 64, 128) and os.environ.get("GNNLM_TRAIN_CHAIN_BWD", "1") != "0":
                node_base, cluster_nl, n_clusters = chains
                L.call("gnnlm_hgt_cluster_attn_train_fwd", L.ptr(q), q.stride(0), L.ptr(k), k.stride(0), L.ptr(v), v.stride(0),
                       L.ptr(node_base), L.ptr(cluster_nl), n_clusters, H, q.shape[1] // H, 1.0, L.ptr(out), out.stride(0), float(p), seed, _st())
                return out
            return _attn_fwd_train(q, k, v, H, 1.0, out, False, p, seed, indptr=indptr, indices=indices)
        return ops.edge_attn(q, k, v, indptr, indices, H, out)

    @staticmethod
    def backward(ctx, dout):
        q, k, v, indptr, indices = ctx.saved_tensors
        dout = dout.contiguous()
        if ctx.chains is not None and (q.shape[1] // ctx.H) in (32, 64, 128) and os.environ.get("GNNLM_TRAIN_CHAIN_BWD", "1") != "0":
            node_base, cluster_nl, n_clusters = ctx.chains
            dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
            L.call("gnnlm_hgt_cluster_attn_bwd", L.ptr(q), q.stride(0), L.ptr(k), k.stride(0), L.ptr(v), v.stride(0), L.ptr(dout),
                   dout.stride(0), L.ptr(node_base), L.ptr(cluster_nl), n_clusters, ctx.H, q.shape[1] // ctx.H, 1.0, L.ptr(dq), dq.stride(0),
                   L.ptr(dk), dk.stride(0), L.ptr(dv), dv.stride(0), float(ctx.p), ctx.seed, _st())
            return dq, dk, dv, None, None, None, None, None, None, None
        dq, dk, dv = _attn_bwd(q, k, v, dout, ctx.H, 1.0, indptr=indptr, indices=indices, p=ctx.p, seed=ctx.seed,
                               symmetric=ctx.symmetric)
        return dq, dk, dv, None, None, None, None, None, None, None


class _TgtAttention(torch.autograd.Function):
    """mean over the two edge types into tgt (hgt.py:383-386, cross_reducer='mean'): 0.5 * inter(q, k_i, v_i) + 0.5 * causal(q, k_t, v_t)."""

    @staticmethod
    def forward(ctx, q, k_i, v_i, k_t, v_t, inter_indptr, B, Lb, intra_ctx, H, p, seed_inter, seed_tt):
        q, k_i, v_i, k_t, v_t = (t.contiguous() for t in (q, k_i, v_i, k_t, v_t))
        ctx.save_for_backward(q, k_i, v_i, k_t, v_t, inter_indptr)
        ctx.cfg = (B, Lb, intra_ctx, H, p, seed_inter, seed_tt)
        out = torch.empty_like(q)
        d = q.shape[1]
        gemm = ops.causal_attn_gemm_supported(d, H, Lb) and os.environ.get("GNNLM_TRAIN_GEMM_FWD", "1") != "0"     # tensor cores, fp32 parity
        if p > 0:
            _attn_fwd_train(q, k_i, v_i, H, 0.5, out, False, p, seed_inter, indptr=inter_indptr)
            if gemm:
                ops.causal_attn_gemm(q, k_t, v_t, B, Lb, intra_ctx, H, out, out_scale=0.5, accumulate=True, drop=(p, seed_tt))
            else:
                _attn_fwd_train(q, k_t, v_t, H, 0.5, out, True, p, seed_tt, causal=(Lb, intra_ctx))
            return out
        ops.edge_attn(q, k_i, v_i, inter_indptr, None, H, out, out_scale=0.5, tag="inter")
        if gemm:
            ops.causal_attn_gemm(q, k_t, v_t, B, Lb, intra_ctx, H, out, out_scale=0.5, accumulate=True)
        else:
            ops.causal_attn(q, k_t, v_t, B, Lb, intra_ctx, H, out, out_scale=0.5, accumulate=True)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k_i, v_i, k_t, v_t, inter_indptr = ctx.saved_tensors
        B, Lb, intra_ctx, H, p, seed_inter, seed_tt = ctx.cfg
        dout = dout.contiguous()
        dq, dk_i, dv_i = _attn_bwd(q, k_i, v_i, dout, H, 0.5, indptr=inter_indptr, p=p, seed=seed_inter)
        dq2, dk_t, dv_t = _attn_bwd(q, k_t, v_t, dout, H, 0.5, causal=(Lb, intra_ctx), p=p, seed=seed_tt)
        L.call("gnnlm_axpy_f32", L.ptr(dq), dq.stride(0), L.ptr(dq2), dq2.stride(0), dq.shape[0], None, dq.shape[1], 1.0, _st())
        return dq, dk_i, dv_i, dk_t, dv_t, None, None, None, None, None, None, None, None


class _AddLayerNorm(torch.autograd.Function):
    """LayerNorm(o + h) * gamma + beta (hgt.py:403-405)."""

    @staticmethod
    def forward(ctx, o, h, gamma, beta, eps):
        o, h = o.contiguous(), h.contiguous()
        ctx.save_for_backward(o, h, gamma)
        ctx.eps = eps
        return ops.layernorm(o, gamma.detach().contiguous(), beta.detach().contiguous(), eps, residual=h)

    @staticmethod
    def backward(ctx, dy):
        o, h, gamma = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(o)
        dg, db = _zeros(o.shape[1], like=o), _zeros(o.shape[1], like=o)
        L.call("gnnlm_layernorm_bwd", L.ptr(o), o.stride(0), L.ptr(h), h.stride(0), L.ptr(gamma.detach().contiguous()), float(ctx.eps),
               L.ptr(dy), dy.stride(0), o.shape[0], None, o.shape[1], L.ptr(dx), dx.stride(0), L.ptr(dg), L.ptr(db), _st())
        return dx, dx, dg, db, None


class _Ctx:
    """Stand-in ctx for calling an autograd Function's forward outside the graph (frozen stages)."""


class _AdaptiveLoss(torch.autograd.Function):
    """sum_t -log p(target_t | x_t) under the (frozen) adaptive softmax: adaptive_softmax.py:147-168 + adaptive_loss.py:62-70.
    The gradient w.r.t. x is produced together with the loss (the softmax - onehot of every cluster's logits, pushed back
    through the cluster's frozen projections) and scaled in backward."""

    @staticmethod
    def forward(ctx, x, target, soft, mode, p_drop, seed, split_k=False):
        x = x.contiguous()
        T, d = x.shape
        drop = lambda a, site: a if p_drop <= 0 else _Dropout.forward(_Ctx(), a, p_drop, site_seed(seed, site))    # frozen part: no graph
        x_in = x
        x = drop(x, 1000)                                              # F.dropout(input) (adaptive_softmax.py:156)
        dev = x.device
        W = soft.train_weights()
        loss = torch.zeros(1, device=dev, dtype=torch.float64)
        pad4 = lambda n_: (n_ + 3) // 4 * 4

        def logits(a, w):          # a w^T into a buffer whose row stride is a multiple of 16 B (it is a GEMM operand on the way back)
            buf = torch.empty((a.shape[0], pad4(w.shape[0])), device=dev, dtype=torch.float32)
            lg = buf[:, :w.shape[0]]
            if mode == L.MATH_TF32X3:
                hi, lo = ops.split_tf32(w)
                return ops.linear(a, hi, None, W_lo=lo, math=mode, out=lg)
            return ops.linear(a, w, None, math=mode, out=lg)

        def back(dlg, w):          # dlogits w: the cluster's frozen weight transposed, k-extent = the cluster size
            if split_k and w.shape[0] >= SPLIT_K_MIN_ROWS and dlg.shape[0] > 0:
                return _back_split_k(dlg, w)
            wt = _transpose(w, pad4(w.shape[0]))                       # [k_in, N padded to a 16 B row stride]
            if mode == L.MATH_TF32X3:
                hi, lo = ops.split_tf32(wt)
                return ops.linear(dlg, hi[:, :w.shape[0]], None, W_lo=lo[:, :w.shape[0]], math=mode)
            return ops.linear(dlg, wt[:, :w.shape[0]], None, math=mode)
        xent = lambda lg, tg: L.call("gnnlm_xent_fwd_bwd", L.ptr(lg), lg.stride(0), L.ptr(tg), lg.shape[0], lg.shape[1], 1.0,
                                     L.ptr(loss), _st())
        if W.get("plain") is not None:                                 # plain softmax (C2): one cluster
            lg = logits(x, W["plain"])
            xent(lg, target.reshape(-1).long().contiguous())
            dx = back(lg, W["plain"])
        else:
            head_pick, tail_rows, tail_pick, tail_count = ops.adapt_target(target.reshape(-1).contiguous(), soft.cutoff)
            lg = logits(x, W["head"])
            xent(lg, head_pick.long())
            dx = back(lg, W["head"])
            counts = tail_count.tolist()                               # host sync: the training path sizes tail batches exactly
            for i, n_i in enumerate(counts[:len(W["proj"])]):
                if n_i == 0:
                    continue
                rows = tail_rows[i, :n_i].contiguous()
                xi = ops.gather_rows(x, rows)
                pi = drop(_gemm(xi, W["proj"][i], None, mode), 1001 + i)    # the tail's nn.Dropout (:101)
                lg = logits(pi, W["out"][i])
                xent(lg, tail_pick[i, :n_i].long().contiguous())
                dxi = _gemm(drop(back(lg, W["out"][i]), 1001 + i), _transpose(W["proj"][i]), None, mode)
                L.call("gnnlm_scatter_add_rows", L.ptr(dx), dx.stride(0), L.ptr(dxi), dxi.stride(0), L.ptr(rows), n_i, None, d, _st())
        dx = drop(dx, 1000)                                            # back through the input dropout
        del x_in
        ctx.save_for_backward(dx)
        return loss.float().squeeze(0)

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g, None, None, None, None, None, None


def _back_split_k(dlg: torch.Tensor, w: torch.Tensor, n_split: int = 8) -> torch.Tensor:
    """dlogits [M, V_c] @ w [V_c, k_in] for a wide cluster (V_c = 207,744 and k_in = 64 on the Wiki103 tail: ONE tile column, 19 CTAs
    as a plain product) by split-K over V_c in the 3xFP16 arithmetic: dlogits column chunks as split-fp16 rows, the weight's row
    chunks transposed + split in one pass, one batched launch, partial sums.  dlogits = softmax - onehot lies in [-1, 1]."""
    M, V = dlg.shape
    k_in = w.shape[1]
    Kc = ((V + n_split - 1) // n_split + 31) // 32 * 32
    S = (V + Kc - 1) // Kc
    dev = dlg.device
    f16 = dict(device=dev, dtype=torch.float16)
    a = torch.empty((S, M, 2 * Kc), **f16)
    hi, lo = torch.empty((S, k_in, Kc), **f16), torch.empty((S, k_in, Kc), **f16)
    for b in range(S):
        c0, c1 = b * Kc, min(V, (b + 1) * Kc)
        blk = dlg[:, c0:c1]
        L.call("gnnlm_transpose_split_f16", L.ptr(w[c0:c1]), w.stride(0), c1 - c0, k_in, 1.0, Kc, 0, L.ptr(hi[b]), L.ptr(lo[b]), _st())
        if c1 - c0 == Kc:                    # hi | lo halves of the chunk written in place, each Kc wide
            L.call("gnnlm_to_split_f16", L.ptr(blk), L.dtype_code(blk.dtype), blk.stride(0), L.ptr(a[b]), 2 * Kc, M, None, Kc, _st())
        else:                                # the ragged last chunk: zero padding up to Kc in both halves
            t = ops.to_split(blk)
            a[b].zero_()
            a[b, :, :c1 - c0].copy_(t.data[:, :c1 - c0])
            a[b, :, Kc:Kc + c1 - c0].copy_(t.data[:, c1 - c0:])
    part = torch.empty((S, M, k_in), device=dev, dtype=torch.float32)
    L.call("gnnlm_linear_batched_f16x3", L.ptr(a), 2 * Kc, M * 2 * Kc, L.ptr(hi), L.ptr(lo), Kc, k_in * Kc, 1.0, None, 0, 0,
           L.ptr(part), k_in, M * k_in, S, M, k_in, Kc, 0, _st(), tag="softmax_back_split_k", work=(M, k_in, S * Kc))
    out = part[0]
    for b in range(1, S):
        L.call("gnnlm_axpy_f32", L.ptr(out), out.stride(0), L.ptr(part[b]), k_in, M, None, k_in, 1.0, _st())
    return out


def fold_layer(layer, t: int, n: int) -> Dict[str, torch.Tensor]:
    """Differentiable fold of the relation transforms into the projections (hgt.py:339-348,354-355):
    K' = (h W_k^T + b_k) blockdiag(R_att[r]) pri[r] / sqrt(d_k),  V' = (h W_v^T + b_v) blockdiag(R_msg[r]).
    Returns fp32 weights per (node type, relation); autograd carries their gradients back to k/v_linears, relation_*."""
    H, dk = layer.n_heads, layer.d_k
    d = H * dk
    out = {}

    def fold(lin, R, scale):
        W3, b2 = lin.weight.view(H, dk, d), lin.bias.view(H, dk)
        Rs = R if scale is None else R * scale[:, None, None]
        return torch.einsum("hjk,hjc->hkc", Rs, W3).reshape(d, d), torch.einsum("hj,hjk->hk", b2, Rs).reshape(d)
    for tau, rels in ((t, (0,)), (n, (0, 1))):                       # tgt is a source of intra only; ntgt of intra and inter
        for r in rels:
            out[f"k{tau}{r}"] = fold(layer.k_linears[tau], layer.relation_att[r], layer.relation_pri[r] / math.sqrt(dk))
            out[f"v{tau}{r}"] = fold(layer.v_linears[tau], layer.relation_msg[r], None)
    return out


def hgt_forward_train(hgt, G: TokenGraph, h_t: torch.Tensor, h_n: torch.Tensor, mode: int = L.MATH_FP32_SIMT, training: bool = False,
                      seed: int = 0) -> torch.Tensor:
    """tgt outputs of HGT.forward (hgt.py:494-513) with autograd through the library's kernels.  h_t [T, d], h_n [n_ntgt, d] fp32
    (exact row counts: the caller sized them from G.counts())."""
    assert hgt.in_dim == hgt.hidden_dim == hgt.out_dim, "training path: plain HGT stack (no input adapters / output projection)"
    t, n = hgt.ntype2idx["tgt"], hgt.ntype2idx["ntgt"]
    n_ntgt, n_valid = G.counts()
    inter_ids = G.inter_indices[:n_valid].contiguous()
    nn_indptr, nn_indices = G.nn_indptr[:n_ntgt + 1].contiguous(), G.nn_indices
    NL = hgt.n_layers
    for l, layer in enumerate(hgt.gcs):
        p_feat, p_att = (float(layer.drop.p), float(layer.attn_drop.p)) if training else (0.0, 0.0)
        sd = lambda site: site_seed(seed, 16 * l + site)
        H = layer.n_heads
        F_ = fold_layer(layer, t, n)
        lin = lambda x, wb: _Linear.apply(x, wb[0], wb[1], mode)
        plain = lambda mods, tau: (mods[tau].weight, mods[tau].bias)
        # ---- tgt: mean of inter (centre ntgt -> tgt) and causal intra attention, output projection, residual + LayerNorm
        hc = _GatherRows.apply(h_n, inter_ids)
        q_t = lin(h_t, plain(layer.q_linears, t))
        agg = _TgtAttention.apply(q_t, lin(hc, F_[f"k{n}1"]), lin(hc, F_[f"v{n}1"]), lin(h_t, F_[f"k{t}0"]), lin(h_t, F_[f"v{t}0"]),
                                  G.inter_indptr, G.B, G.L, G.intra_ctx, H, p_att, sd(1), sd(0))
        new_t = _AddLayerNorm.apply(_dropout(lin(agg, plain(layer.a_linears, t)), p_feat, sd(3)), h_t, layer.norms[t].weight,
                                    layer.norms[t].bias, layer.norms[t].eps)
        # ---- ntgt (not needed after the last layer: the decoder reads tgt rows only, transformer.py:1053)
        if l < NL - 1:
            agg_n = _EdgeAttention.apply(lin(h_n, plain(layer.q_linears, n)), lin(h_n, F_[f"k{n}0"]), lin(h_n, F_[f"v{n}0"]),
                                         nn_indptr, nn_indices, H, p_att, sd(2), SYMMETRIC_NN_BWD and not G.dedup,
                                         None if G.dedup else (G.node_base, G.cluster_nl, G.T * G.k))
            h_n = _AddLayerNorm.apply(_dropout(lin(agg_n, plain(layer.a_linears, n)), p_feat, sd(4)), h_n, layer.norms[n].weight,
                                      layer.norms[n].bias, layer.norms[n].eps)
        h_t = new_t
    return h_t


def train_step_loss(model, sample: dict, mode: str = "fp32", seed: int = 0) -> torch.Tensor:
    """AdaptiveLoss.forward (adaptive_loss.py:31-83, reduce=True) for a model built with --freeze: the summed cross-entropy of the
    batch, differentiable w.r.t. decoder.hgt_decoder.* (call .backward() on it).  `sample` as eval: net_input.graph (TokenGraph
    with codes_table and tgt features), target."""
    dec = model.decoder
    G: TokenGraph = sample["net_input"]["graph"]
    m = L.MATH_NAMES[mode]
    assert m in (L.MATH_FP32_SIMT, L.MATH_TF32X3, L.MATH_F16X3), "training runs the projections in fp32 FMA, 3xTF32 or 3xFP16"
    if dec.orig_prob_ratio > 0:
        raise NotImplementedError("orig_prob_ratio > 0 in training (adaptive_loss.py:55-59) is not used by the shipped scripts")
    feats = G.nodes["tgt"].data["h"]
    h_t = feats.float().contiguous()
    n_ntgt, _ = G.counts()
    with torch.no_grad():                                             # PQ decode + OPQ rotation: inputs, no parameters (pq_wrapper.py:169-203)
        # the OPQ rotation in 3xTF32 whenever the projections run on the tensor cores (fp32 rows out either way)
        h_n = dec.tgt_quantizer.gather_decode(G.codes_table, G.ntgt_row, n_cap=G.node_cap, n_dev=G.n_ntgt_dev,
                                              math_mode=L.MATH_FP32_SIMT if m == L.MATH_FP32_SIMT else L.MATH_TF32X3)
        h_n = h_n[:n_ntgt].contiguous()
    training = bool(model.training)                                    # dropout masks only in train mode; `seed`: one per update
    x = hgt_forward_train(dec.hgt_decoder, G, h_t, h_n, m, training, seed)
    soft = dec.adaptive_softmax if dec.adaptive_softmax is not None else _PlainSoftmax(dec.embed_out)
    p_soft = float(getattr(soft, "dropout", 0.0)) if training else 0.0
    # the (frozen) output layer's logits keep 3xTF32 in the 3xFP16 mode: its gradient operand is produced inside one Function
    out = _AdaptiveLoss.apply(x, sample["target"], soft, L.MATH_TF32X3 if m == L.MATH_F16X3 else m, p_soft, seed, m == L.MATH_F16X3)
    return out


class _PlainSoftmax:
    dropout = 0.0

    def __init__(self, embed_out):
        self.w = embed_out

    def train_weights(self):
        return {"plain": self.w.detach().float().contiguous()}


class AdaptiveLoss:
    """fairseq/criterions/adaptive_loss.py:14-83 (`--criterion adaptive_loss`): forward(model, sample) -> (loss, sample_size,
    logging_output) with the reference's keys; the loss is differentiable w.r.t. the trainable (HGT) parameters."""

    def __init__(self, args=None, task=None, math: str = "fp32", seed: int = 1):
        self.args, self.math = args, math
        self.sentence_avg = bool(getattr(args, "sentence_avg", False))
        self.seed, self.calls = int(getattr(args, "seed", seed) or seed), 0       # dropout masks: a fresh seed per forward

    def forward(self, model, sample, reduce=True):
        if not reduce:
            raise NotImplementedError("reduce=False (per-token losses) is not used by the trainer (adaptive_loss.py:64-69)")
        self.calls += 1
        loss = train_step_loss(model, sample, self.math, seed=(self.seed * 0x9E3779B97F4A7C15 + self.calls) & 0xFFFFFFFFFFFFFFFF)
        ntokens = int(sample["target"].numel())                     # graph LM blocks carry no padding (transformer.py:975)
        nsentences = int(sample["target"].shape[0])
        sample_size = nsentences if self.sentence_avg else ntokens
        return loss, sample_size, {"loss": loss.detach(), "ntokens": ntokens, "nsentences": nsentences, "sample_size": sample_size}

    __call__ = forward


def train_step(model, sample, optimizer, criterion: Optional[AdaptiveLoss] = None, clip_norm: float = 0.0) -> dict:
    """One update as fairseq's trainer performs it for this model (fairseq/trainer.py train_step: forward + backward, gradients
    divided by the sample size, --clip-norm, optimizer step); the optimizer is any torch.optim instance over the trainable
    parameters (the scripts use Adam, hgt_lm_wiki103_reproduce.sh:27-30)."""
    criterion = criterion or AdaptiveLoss()
    model.train()
    optimizer.zero_grad(set_to_none=True)
    loss, sample_size, log = criterion(model, sample)
    loss.backward()
    release_caches()
    params = [p for p in model.parameters() if p.requires_grad and p.grad is not None]
    for p in params:
        p.grad.div_(float(sample_size))
    gnorm = torch.nn.utils.clip_grad_norm_(params, clip_norm if clip_norm > 0 else float("inf"))
    optimizer.step()
    log.update(gnorm=float(gnorm), loss_per_token_base2=float(loss.detach()) / log["ntokens"] / math.log(2))
    return log
