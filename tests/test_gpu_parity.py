"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures
generated from the reference's own code.  Integer / byte / index work is bit-exact; floating point
is held to the tolerance BASELINE.json's north_star states (per-token log-probs within 1e-4
relative in fp32) -- each tolerance is written next to its assert."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
GRAPH_CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLD, "graph_*.npz")))


def _sd(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


# ------------------------------------------------------------------------------------------ graph
@pytest.mark.parametrize("case", GRAPH_CASES)
def test_graph_matches_reference_builder(case, dev):
    from gnnlm_b200 import ops
    from gnnlm_b200.graph import build_token_graph
    from oracle import graph_oracle as go
    z = np.load(os.path.join(GOLD, f"graph_{case}.npz"))
    nbr = torch.from_numpy(z["nbr"])[None].contiguous().to(dev)
    pos = torch.from_numpy(z["offsets"])[None].contiguous().to(dev)
    g = build_token_graph(nbr, int(z["n_d"]), int(z["cl"]), int(z["cr"]), tgt_pos=pos,
                          invalid_ctx=int(z["invalid_ctx"]), intra_ctx=int(z["intra_ctx"]), with_owner=True)
    n_ntgt, n_valid = g.counts()
    L = int(z["L"])
    assert n_ntgt == len(z["ntgt_labels"]) and n_valid == len(z["inter_src"])
    # canonical CSRs from the reference's insertion-ordered COO: bit-exact
    nn_ip, nn_ix = go.canonical_csr(z["nn_src"], z["nn_dst"], n_ntgt)
    in_ip, in_ix = go.canonical_csr(z["inter_src"], z["inter_dst"], L)
    tt_ip, tt_ix = go.canonical_csr(z["tt_src"], z["tt_dst"], L)
    assert (g.nn_indptr[:n_ntgt + 1].cpu().numpy() == nn_ip).all()
    assert (g.nn_indices[:len(nn_ix)].cpu().numpy() == nn_ix).all()
    assert (g.inter_indptr.cpu().numpy() == in_ip).all()
    assert (g.inter_indices[:n_valid].cpu().numpy() == in_ix).all()
    ip, ix = g.materialise_tt()
    assert (ip.cpu().numpy() == tt_ip).all() and (ix.cpu().numpy() == tt_ix).all()
    assert (g.ntgt_owner[:n_valid and n_ntgt].cpu().numpy() >= 0).all()
    # gathered code rows and labels (`ntgt.h`, `ntgt.labels`): bit-exact
    codes = torch.from_numpy(z["codes"]).to(dev)
    vals = torch.from_numpy(z["vals"].reshape(-1)).to(dev)
    cen = torch.zeros(codes.shape[1], 256, 4, device=dev)
    _, labels, codes_out = ops.pq_gather_decode(codes, cen, g.ntgt_row, n_cap=n_ntgt, labels_table=vals,
                                                want_codes=True, decode=False)
    assert (codes_out.cpu().numpy() == z["ntgt_codes"]).all()
    assert (labels.cpu().numpy() == z["ntgt_labels"].reshape(-1)).all()


@pytest.mark.parametrize("B,L,k,cl,cr,stress", [(2, 256, 8, 1, 1, False), (3, 128, 32, 2, 2, True), (1, 3072, 32, 1, 1, True),
                                                (2, 64, 16, 3, 0, True), (1, 5, 1, 0, 0, False)])
def test_graph_vs_oracle_random(B, L, k, cl, cr, stress, dev):
    from gnnlm_b200.graph import build_token_graph
    from oracle import graph_oracle as go
    rng = np.random.RandomState(B * 1000 + L + k)
    n_d = 1 << 20
    nbr = rng.randint(0, n_d, size=(B, L, k)).astype(np.int64)
    if stress:
        nbr[rng.rand(B, L, k) < 0.01] = -1
        e = rng.rand(B, L, k) < 0.01
        nbr[e] = rng.choice([0, 1, 2, n_d - 1, n_d - 2, n_d - 3], size=int(e.sum()))
        nbr[0, L // 2] = -1
    off = np.arange(B * L, dtype=np.int64).reshape(B, L)
    ref = go.build_batch_vectorised(nbr, off, n_d, cl, cr)
    g = build_token_graph(torch.from_numpy(nbr).to(dev), n_d, cl, cr)
    n_ntgt, n_valid = g.counts()
    assert n_ntgt == ref["n_ntgt"]
    assert (g.ntgt_row[:n_ntgt].cpu().numpy() == ref["ntgt_offsets"]).all()
    assert (g.nn_indptr[:n_ntgt + 1].cpu().numpy() == ref["nn_csr"][0]).all()
    assert (g.nn_indices[:len(ref["nn_csr"][1])].cpu().numpy() == ref["nn_csr"][1]).all()
    assert (g.inter_indptr.cpu().numpy() == ref["inter_csr"][0]).all()
    assert (g.inter_indices[:n_valid].cpu().numpy() == ref["inter_csr"][1]).all()


def test_graph_all_invalid_block(dev):
    """a block whose neighbours are all -1 yields an empty ntgt set."""
    from gnnlm_b200.graph import build_token_graph
    nbr = torch.full((1, 16, 4), -1, dtype=torch.int64, device=dev)
    g = build_token_graph(nbr, 1000, 1, 1)
    assert g.counts() == (0, 0)
    assert (g.inter_indptr.cpu().numpy() == 0).all()


# ------------------------------------------------------------------------------------------ PQ
@pytest.mark.parametrize("case", ["m8", "m16b"])
def test_pq_decode_golden(case, dev):
    from gnnlm_b200.pq_codec import TorchPQCodec
    z = np.load(os.path.join(GOLD, f"pq_{case}.npz"))
    codes = torch.from_numpy(z["codes"]).to(dev)
    raw = TorchPQCodec(centroids=z["cen"]).to(dev).decode(codes)
    assert (raw.cpu().numpy() == z["x_nopre"]).all()                      # pure gather: bit-exact
    full = TorchPQCodec(centroids=z["cen"], A=z["A"], b=z["b"]).to(dev).decode(codes)
    np.testing.assert_allclose(full.cpu().numpy(), z["x_torch"], rtol=0, atol=2e-5)   # fp32 rotation, d<=128


@pytest.mark.parametrize("M,dsub,n", [(128, 8, 5000), (64, 8, 3000), (128, 4, 1000), (16, 2, 300)])
def test_pq_gather_decode_vs_oracle(M, dsub, n, dev):
    from gnnlm_b200 import ops
    from oracle import model_oracle as mo
    rng = np.random.RandomState(M + dsub)
    n_d = 20000
    codes = rng.randint(0, 256, size=(n_d, M)).astype(np.uint8)
    cen = rng.randn(M, 256, dsub).astype(np.float32)
    rows = rng.randint(0, n_d, size=n).astype(np.int64)
    x, _, _ = ops.pq_gather_decode(torch.from_numpy(codes).to(dev), torch.from_numpy(cen).to(dev),
                                   torch.from_numpy(rows).to(dev))
    ref = mo.pq_decode(codes[rows], cen)
    assert (x.cpu().numpy() == ref).all()                                  # bit-exact
    # row_ids indirection + device-side count
    ids = rng.permutation(n)[: n // 2].astype(np.int32)
    cnt = torch.tensor([n // 3], dtype=torch.int32, device=dev)
    x2, _, _ = ops.pq_gather_decode(torch.from_numpy(codes).to(dev), torch.from_numpy(cen).to(dev),
                                    torch.from_numpy(rows).to(dev), row_ids=torch.from_numpy(ids).to(dev), n_dev=cnt)
    assert (x2[: n // 3].cpu().numpy() == ref[ids[: n // 3]]).all()


# ------------------------------------------------------------------------------------------ GEMM / rows
@pytest.mark.parametrize("M,N,K", [(300, 200, 64), (1000, 20002, 128), (129, 77, 4), (4096, 1024, 1024)])
def test_linear_simt(M, N, K, dev):
    from gnnlm_b200 import ops
    torch.manual_seed(0)
    A, W, b = torch.randn(M, K), torch.randn(N, K) / K ** 0.5, torch.randn(N)
    R = torch.randn(M, N)
    ref = (A.double() @ W.double().t() + b.double() + R.double())
    out = ops.linear(A.to(dev), W.to(dev), b.to(dev), residual=R.to(dev))
    np.testing.assert_allclose(out.cpu().double().numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)   # fp32 FMA vs fp64
    pick = torch.randint(0, N, (M,), dtype=torch.int32)
    pm, ps, pk, nt = ops.linear_lse(A.to(dev), W.to(dev), pick.to(dev))
    lp = torch.empty(M, device=dev)
    ops.lse_finish(pm, ps, pk, nt, lp)
    logits = A.double() @ W.double().t()
    ref_lp = torch.log_softmax(logits, 1).gather(1, pick.long()[:, None]).squeeze(1)
    np.testing.assert_allclose(lp.cpu().double().numpy(), ref_lp.numpy(), rtol=1e-5, atol=2e-5)


def test_layernorm_and_gather(dev):
    from gnnlm_b200 import ops
    torch.manual_seed(1)
    for d in (32, 512, 1024):
        x, g, b = torch.randn(700, d) * 3 + 1, torch.randn(d), torch.randn(d)
        y = ops.layernorm(x.to(dev), g.to(dev), b.to(dev))
        ref = torch.nn.functional.layer_norm(x, (d,), g, b)
        np.testing.assert_allclose(y.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)
        r = torch.randn(700, d)
        ref_r = torch.nn.functional.layer_norm(x + r, (d,), g, b)
        y_r = ops.layernorm(x.to(dev), g.to(dev), b.to(dev), residual=r.to(dev))
        np.testing.assert_allclose(y_r.cpu().numpy(), ref_r.numpy(), rtol=1e-5, atol=1e-5)
        y_s = ops.layernorm(x.to(dev), g.to(dev), b.to(dev), residual=ops.to_split(r.to(dev)), out_dtype=ops.SPLIT)
        np.testing.assert_allclose(y_s.float().cpu().numpy(), ref_r.numpy(), rtol=1e-5, atol=1e-5)
        ids = torch.randint(0, 700, (333,), dtype=torch.int32)
        assert (ops.gather_rows(x.to(dev), ids.to(dev)).cpu() == x[ids.long()]).all()


# ------------------------------------------------------------------------------------------ HGT
def _hgt_from_golden(z, dev):
    from gnnlm_b200.hgt import HGT
    d, hidden, d_out = z["h_tgt"].shape[1], z["sd.gcs.0.k_linears.0.weight"].shape[0], z["out_tgt"].shape[1]
    m = HGT({"tgt": 0, "ntgt": 1}, {"intra": 0, "inter": 1}, d, hidden, d_out, int(z["n_layers"]), int(z["H"]))
    m.load_state_dict(_sd(z, "sd."), strict=True)
    return m.to(dev).eval()


@pytest.mark.parametrize("case", ["l2_c1", "l3_c2", "l2_adapt"])      # l2_adapt: in_dim 24 -> hidden 32 -> out 24 (adapt_ws / out)
def test_hgt_golden(case, dev):
    """Reference hgt.py executed under the DGL stub (tests/golden/make_golden.py) vs the CUDA path."""
    from gnnlm_b200.graph import build_token_graph
    z = np.load(os.path.join(GOLD, f"hgt_{case}.npz"))
    m = _hgt_from_golden(z, dev)
    g = build_token_graph(torch.from_numpy(z["nbr"]).to(dev), int(z["n_d"]), int(z["cl"]), int(z["cr"]))
    h_t, h_n = torch.from_numpy(z["h_tgt"]).to(dev), torch.from_numpy(z["h_ntgt"]).to(dev)
    assert g.counts()[0] == h_n.shape[0]
    out = m(g, features={"tgt": h_t, "ntgt": h_n})
    # fp32 end to end; tolerance: 1e-4 relative (north_star) on O(1) LayerNorm outputs
    np.testing.assert_allclose(out["tgt"].cpu().numpy(), z["out_tgt"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(out["ntgt"].cpu().numpy(), z["out_ntgt"], rtol=1e-4, atol=1e-4)
    # tgt-only fast path (dead-work elimination, capacity-sized arrays, device-side counts)
    h_cap = torch.zeros(g.node_cap, h_n.shape[1], device=dev)
    h_cap[: h_n.shape[0]] = h_n
    fast = m.forward_tgt(g, h_t, h_cap)
    np.testing.assert_allclose(fast.cpu().numpy(), z["out_tgt"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("d,H,L,k,c,NL", [(512, 8, 64, 8, 1, 1), (512, 8, 96, 4, 1, 3), (1024, 8, 48, 4, 2, 2), (128, 4, 40, 3, 0, 2),
                                          (1024, 8, 40, 4, 3, 3), (256, 4, 70, 5, 1, 2)])
def test_hgt_vs_oracle(d, H, L, k, c, NL, dev):
    from gnnlm_b200.graph import build_token_graph
    from gnnlm_b200.hgt import HGT
    from oracle import graph_oracle as go, model_oracle as mo
    torch.manual_seed(d + L)
    rng = np.random.RandomState(L)
    B, n_d = 2, 5000
    nbr = rng.randint(0, n_d, size=(B, L, k)).astype(np.int64)
    nbr[rng.rand(B, L, k) < 0.05] = -1
    nbr[1, 3] = -1
    off = np.arange(B * L).reshape(B, L)
    gref = go.build_batch_vectorised(nbr, off, n_d, c, c)
    m = HGT({"tgt": 0, "ntgt": 1}, {"intra": 0, "inter": 1}, d, d, d, NL, H)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1 or p.shape[-1] == H:
                p.add_(0.1 * torch.randn_like(p))
    h_t, h_n = torch.randn(B * L, d), torch.randn(gref["n_ntgt"], d)
    ref = mo.hgt_forward_csr({k_: v.double() for k_, v in m.state_dict().items()}, h_t.double(), h_n.double(), gref,
                             (B, L), H, NL)
    m = m.to(dev).eval()
    g = build_token_graph(torch.from_numpy(nbr).to(dev), n_d, c, c)
    h_cap = torch.zeros(g.node_cap, d, device=dev)
    h_cap[: h_n.shape[0]] = h_n.to(dev)
    fast = m.forward_tgt(g, h_t.to(dev), h_cap)
    np.testing.assert_allclose(fast.cpu().double().numpy(), ref["tgt"].numpy(), rtol=1e-4, atol=1e-4)
    full = m(g, features={"tgt": h_t.to(dev), "ntgt": h_n.to(dev)})
    np.testing.assert_allclose(full["ntgt"].cpu().double().numpy(), ref["ntgt"].numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(full["tgt"].cpu().double().numpy(), ref["tgt"].numpy(), rtol=1e-4, atol=1e-4)
    # generic CSR kernel instead of the cluster kernel for the ntgt-intra-ntgt edges: same results
    for layer in m.gcs:
        layer.use_cluster_kernel = False
    fast2 = m.forward_tgt(g, h_t.to(dev), h_cap)
    full2 = m(g, features={"tgt": h_t.to(dev), "ntgt": h_n.to(dev)})
    np.testing.assert_allclose(fast2.cpu().numpy(), fast.cpu().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(full2["ntgt"].cpu().numpy(), full["ntgt"].cpu().numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,L,H,d,ctx", [(2, 37, 4, 128, 5), (2, 200, 8, 1024, 0), (1, 333, 8, 512, 70), (1, 64, 2, 256, 0),
                                         (3, 129, 8, 1024, 33)])
def test_causal_attn_with_intra_context(B, L, H, d, ctx, dev):
    """warp-per-destination form (L < 64 / other d_k) and the tiled flash form (d_k in {64,128})."""
    from gnnlm_b200 import ops
    torch.manual_seed(3)
    q, k, v = torch.randn(B * L, d), torch.randn(B * L, d), torch.randn(B * L, d)
    out = torch.zeros(B * L, d, device=dev)
    out.fill_(1.0)
    ops.causal_attn(q.to(dev), k.to(dev), v.to(dev), B, L, ctx, H, out, out_scale=0.5, accumulate=True)
    out = (out - 1.0) * 2.0
    qh, kh, vh = (t.view(B, L, H, d // H).permute(0, 2, 1, 3).double() for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2)
    i = torch.arange(L)
    mask = (i[None, :] <= i[:, None]) & ((i[:, None] - i[None, :] < ctx) if ctx else True)
    s = s.masked_fill(~mask, -float("inf"))
    ref = (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B * L, d)
    # unnormalised N(0,1) q,k give |scores| up to ~40: fp32 dot-product rounding ~1e-5 on the logits
    np.testing.assert_allclose(out.cpu().double().numpy(), ref.numpy(), rtol=1e-4, atol=5e-5)


# ------------------------------------------------------------------------------------------ log-probs / kNN
@pytest.mark.parametrize("case", ["untied", "tied", "tied_noproj"])
def test_adaptive_softmax_golden(case, dev):
    """tied_noproj: --tie-adaptive-weights without --tie-adaptive-proj (tail projections stored [dim_i, d])."""
    from gnnlm_b200.model import AdaptiveSoftmax
    z = np.load(os.path.join(GOLD, f"adaptive_{case}.npz"))
    cutoff = z["cutoff"].tolist()
    m = AdaptiveSoftmax(cutoff[-1], z["x"].shape[-1], cutoff[:-1], tied=bool(z["tied"]),
                        tie_proj=bool(z["tie_proj"]) if "tie_proj" in z.files else None)
    sd = _sd(z, "sd.")
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    x, t = torch.from_numpy(z["x"]).to(dev), torch.from_numpy(z["target"]).to(dev)
    lp = m.target_log_prob(x, t)
    # per-token log-probs within 1e-4 relative (north_star)
    np.testing.assert_allclose(lp.cpu().numpy(), z["lp_target_mode_at_target"].reshape(-1), rtol=1e-4, atol=1e-5)
    full = m.get_log_prob(x, None)
    np.testing.assert_allclose(full.cpu().numpy(), z["lp_full"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("case", ["ip_t1", "l2_t001"])
def test_knn_prob_golden(case, dev):
    from gnnlm_b200.knn_model import KNNModel
    z = np.load(os.path.join(GOLD, f"knn_{case}.npz"))
    km = KNNModel(torch.from_numpy(z["vals"]).to(dev), vocab_size=int(z["V"]), metric_type=str(z["metric"]))
    d, i = torch.from_numpy(z["dists"]).to(dev), torch.from_numpy(z["ids"]).to(dev)
    km.set_search_results(d, i)
    p, rec = km.get_knn_prob(None, t=float(z["temp"]), targets=torch.from_numpy(z["targets"]).to(dev), return_recall=True)
    np.testing.assert_allclose(p.cpu().numpy(), z["p_target"], rtol=1e-4, atol=1e-7)
    assert (rec.cpu().numpy() == z["recall"]).all()                        # integer: exact
    km.set_search_results(d, i)
    full = km.get_knn_prob(None, t=float(z["temp"]))
    np.testing.assert_allclose(full.cpu().numpy(), z["p_full"], rtol=1e-4, atol=1e-6)


def test_scorer_golden_knn_mix(dev):
    """SequenceScorer.generate of the reference (scripted model + reference AdaptiveSoftmax + reference
    get_knn_prob) vs ours: per-position scores within 1e-4 relative."""
    from types import SimpleNamespace
    from gnnlm_b200.knn_model import KNNModel
    from gnnlm_b200.model import AdaptiveSoftmax
    from gnnlm_b200.sequence_scorer import SequenceScorer
    z = np.load(os.path.join(GOLD, "scorer_b1_knn.npz"))
    cutoff = z["cutoff"].tolist()
    feats = torch.from_numpy(z["feats"]).to(dev)
    soft = AdaptiveSoftmax(cutoff[-1], feats.shape[-1], cutoff[:-1])
    soft.load_state_dict(_sd(z, "sd."), strict=True)
    soft = soft.to(dev)

    class Dec:
        def target_log_probs(self, net_output, target):
            return soft.target_log_prob(net_output[0], target).view(target.shape)

    class Model:
        decoder = Dec()

        def eval(self):
            return self

        def __call__(self, **kw):
            return feats, {"inner_states": [feats.transpose(0, 1)]}

    d_ = SimpleNamespace(pad=lambda: 1, eos=lambda: 2)
    sc = SequenceScorer(d_, args=SimpleNamespace(lmbda=float(z["lmbda"]), knn_keytype=None))
    km = KNNModel(torch.from_numpy(z["vals"]).to(dev), vocab_size=cutoff[-1])
    km.set_search_results(torch.from_numpy(z["dists"]).to(dev), torch.from_numpy(z["ids"]).to(dev))
    sample = {"net_input": {}, "target": torch.from_numpy(z["target"]).to(dev),
              "start_indices": torch.from_numpy(z["start_indices"])}
    hy = sc.generate([Model()], sample, knn_dstore=km, temperature=float(z["temp"]))
    np.testing.assert_allclose(hy[0][0]["positional_scores"].cpu().numpy(), z["knn_pos_0"], rtol=1e-4, atol=1e-5)
    assert (hy[0][0]["knn_recall"].cpu().numpy() == z["knn_recall_0"]).all()
    np.testing.assert_allclose(float(hy[0][0]["score"]), float(z["knn_score_0"]), rtol=1e-4)


# ------------------------------------------------------------------------------------------ whole path
def test_whole_path_vs_oracle_c1(dev):
    """BASELINE.json configs[0] (tiny): 1-layer HGT, d=512, vocab 10k, 2x256 tokens, k=8, +-1 context,
    PQ M=64 -- graph assembly -> PQ decode -> HGT -> adaptive softmax -> kNN mix -> NLL, vs the fp32
    CPU oracle.  log-probs within 1e-4 relative; perplexity within 0.01 absolute."""
    from tests.synth import make_problem, run_gpu, run_oracle
    prob = make_problem("c1")
    ref = run_oracle(prob)
    out = run_gpu(prob, dev, math="fp32")
    np.testing.assert_allclose(out["logprob"], ref["logprob"].numpy(), rtol=1e-4, atol=1e-4)
    # "perplexity within 0.01 absolute" is stated for trained models (wiki103 GNN ppl 16.8, README.md:24);
    # d ppl = ppl * d nll, so the size-independent form is |d nll| < 0.01 / 16.8 (random-init ppl is ~ V).
    assert abs(out["nll"] - ref["nll"]) < 0.01 / 16.8
    assert out["count"] == ref["count"]
    assert (out["recall"] == ref["knn_recall"].numpy()).all()


# ------------------------------------------------------------------------------------------ tcgen05 GEMM
def _need_tc():
    from gnnlm_b200 import _lib
    if not _lib.load().gnnlm_has_tcgen05():
        pytest.fail("tcgen05 path unavailable on this device: the B200 product path must be present")


# tf32x3: the split recovers the operands to ~2^-21, what remains is the tensor core's fp32 accumulation
# rounding over 3*K/8 partial products (measured max |err| 4e-5 on O(3) outputs at K=1024) -> held to the
# north star's 1e-4; tf32: 10-bit mantissas; bf16: exact products of the bf16-rounded operands, fp32 accumulate.
# f16x3: operands split into two fp16 halves (22 significant bits), kind::f16 MMAs, same accumulation floor.
@pytest.mark.parametrize("math,rtol,atol", [("tf32x3", 1e-4, 1e-4), ("f16x3", 1e-4, 1e-4), ("tf32", 5e-3, 5e-3), ("bf16", 1e-4, 1e-4)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (300, 200, 64), (1000, 20002, 128), (4096, 3072, 1024), (5000, 1024, 1024), (129, 72, 256)])
def test_linear_tcgen05(math, rtol, atol, M, N, K, dev):
    _need_tc()
    from gnnlm_b200 import _lib as L, ops
    torch.manual_seed(M + N + K)
    mode = L.MATH_NAMES[math]
    ws = 1.0
    A, W, b = torch.randn(M, K), torch.randn(N, K) / K ** 0.5, torch.randn(N)
    R = torch.randn(M, N)
    if math == "bf16":
        Ad, Wd = A.to(dev).bfloat16(), W.to(dev).bfloat16()
        A64, W64 = Ad.double().cpu(), Wd.double().cpu()
        lo = None
    else:
        Ad, Wd = A.to(dev), W.to(dev)
        A64, W64 = A.double(), W.double()
        lo = None
        if math == "tf32x3":
            Wd, lo = ops.split_tf32(Wd)
        elif math == "f16x3":
            Wd, lo, ws = ops.split_f16(Wd)
    ref = A64 @ W64.t() + b.double() + R.double()
    out = ops.linear(Ad, Wd, b.to(dev), W_lo=lo, w_scale=ws, residual=R.to(dev), math=mode)
    np.testing.assert_allclose(out.cpu().double().numpy(), ref.numpy(), rtol=rtol, atol=atol)
    # device-side row count: rows >= m_dev must stay untouched
    cnt = torch.tensor([M // 2 + 1], dtype=torch.int32, device=dev)
    out2 = torch.full((M, N), 7.0, device=dev)
    ops.linear(Ad, Wd, b.to(dev), W_lo=lo, w_scale=ws, out=out2, m_dev=cnt, math=mode)
    live = M // 2 + 1
    np.testing.assert_allclose(out2[:live].cpu().double().numpy(), (ref - R.double())[:live].numpy(), rtol=rtol, atol=atol)
    assert (out2[live:] == 7.0).all()
    # fused log-sum-exp epilogue
    pick = torch.randint(0, N, (M,), dtype=torch.int32)
    pm, ps, pk, nt = ops.linear_lse(Ad, Wd, pick.to(dev), W_lo=lo, w_scale=ws, math=mode)
    lp = torch.empty(M, device=dev)
    ops.lse_finish(pm, ps, pk, nt, lp)
    ref_lp = torch.log_softmax(A64 @ W64.t(), 1).gather(1, pick.long()[:, None]).squeeze(1)
    np.testing.assert_allclose(lp.cpu().double().numpy(), ref_lp.numpy(), rtol=rtol, atol=max(atol, 2e-5))


@pytest.mark.parametrize("M,N,K,K2", [(128, 256, 64, 0), (300, 200, 128, 0), (5000, 1024, 1024, 0), (4097, 3072, 1024, 0), (2500, 1024, 512, 512),
                                       (129, 72, 256, 48), (70000, 136, 1024, 1024)])
def test_linear_f16f8(M, N, K, K2, dev):
    """MATH_F16F8 product (fp16 main product + the two 2^-11-sized correction products as FP8 e4m3 MMAs into the same TMEM
    accumulator): within 1e-4 of the fp64 product like the other fp32-parity modes (measured ~1e-5 of max|C|), for one and two
    k-concatenated sources, ragged M / N, device-side row counts; the e4m3 companions against their definition bit for bit."""
    _need_tc()
    from gnnlm_b200 import ops
    g = torch.Generator(device=dev).manual_seed(M + N + K + K2)
    A = torch.randn((M, K), generator=g, device=dev)
    A2 = torch.randn((M, K2), generator=g, device=dev) * 3 if K2 else None
    W = torch.randn((N, K + K2), generator=g, device=dev) / (K + K2) ** 0.5
    b = torch.randn((N,), generator=g, device=dev)
    Wh, Wl, sc = ops.split_f16(W)
    W8 = ops.quant_w8(Wh, Wl)
    As = ops.to_q8(ops.to_split(A))
    A2s = ops.to_q8(ops.to_split(A2)) if K2 else None
    # companions: hi8 = e4m3(hi), lo8 = e4m3(2^10 lo); weights lo8 = e4m3(lo), hi8 = e4m3(2^-10 hi) (round to nearest even, saturating)
    e4 = lambda t: t.float().clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8)
    assert torch.equal(As.q8[:, :K], e4(As.data[:, :K])) and torch.equal(As.q8[:, K:], e4(As.data[:, K:].float() * 1024))
    assert torch.equal(W8[:, :K + K2], e4(Wl)) and torch.equal(W8[:, K + K2:], e4(Wh.float() / 1024))
    Af = A if not K2 else torch.cat([A, A2], 1)
    ref = Af.double() @ W.double().T + b.double()
    out = ops.linear_f16f8(As, Wh, W8, b, A2=A2s, w_scale=sc)
    # errors are relative to the magnitude of the accumulated terms, not of a (possibly cancelling) result: 3e-5 of max|C|
    # (measured 8e-6; 3xFP16 3e-6; a single fp16 pass 2e-4)
    assert (out.double() - ref).abs().max().item() <= 3e-5 * ref.abs().max().item()
    cnt = torch.tensor([M // 2 + 1], dtype=torch.int32, device=dev)
    out2 = torch.full((M, N), 7.0, device=dev)
    ops.linear_f16f8(As, Wh, W8, b, A2=A2s, w_scale=sc, out=out2, m_dev=cnt)
    live = M // 2 + 1
    assert torch.equal(out2[:live], out[:live]) and (out2[live:] == 7.0).all()
    # split-fp16 output epilogue
    o3 = ops.linear_f16f8(As, Wh, W8, b, A2=A2s, w_scale=sc, out_dtype=ops.SPLIT)
    assert (o3.float() - out).abs().max().item() <= 2e-7 * out.abs().max().item()


def test_e4m3_companions_written_by_the_producers(dev):
    """LayerNorm, cluster attention (all nodes / centre only) and the pre-split PQ decode write the e4m3 companion of their
    split-fp16 output themselves in MATH_F16F8: bit-identical to the standalone gnnlm_split_to_q8 pass over the same output,
    and the split output itself unchanged."""
    _need_tc()
    from gnnlm_b200 import ops, synth
    from gnnlm_b200.graph import build_token_graph
    g = torch.Generator(device=dev).manual_seed(11)
    n, d, H = 3000, 1024, 8
    x = torch.randn((n, d), generator=g, device=dev) * 3
    res = ops.to_split(torch.randn((n, d), generator=g, device=dev))
    gam, bet = torch.randn(d, generator=g, device=dev), torch.randn(d, generator=g, device=dev)
    cnt = torch.tensor([n - 7], dtype=torch.int32, device=dev)
    plain = ops.layernorm(x, gam, bet, out_dtype=ops.SPLIT, residual=res, n_dev=cnt)
    both = ops.layernorm(x, gam, bet, out_dtype=ops.SPLIT_Q8, residual=res, n_dev=cnt)
    live = n - 7
    assert torch.equal(plain.data[:live], both.data[:live])
    want = ops.to_q8(ops.Split(both.data.clone(), d), cnt).q8
    assert torch.equal(both.q8[:live], want[:live])
    # cluster attention over a small graph (k = 4, c = 1)
    cfg = dict(synth.CONFIGS["c3mini"], L=96, k=4)
    tables = synth.make_tables(cfg, device=dev, n_d=1 << 14)
    batch = synth.make_batch(cfg, tables, device=dev)
    G = build_token_graph(batch["nbr"], tables["n_d"], 1, 1)
    n_ntgt, n_valid = G.counts()
    qkv = torch.randn((G.node_cap, 3 * d), generator=g, device=dev)
    for centre in (False, True):
        rows = n_valid if centre else G.node_cap
        q = qkv[:rows, :d].contiguous() if centre else qkv[:, :d]
        a = ops.empty_act(rows, d, ops.SPLIT, dev)
        b = ops.empty_act(rows, d, ops.SPLIT_Q8, dev)
        b.q8.zero_()
        ops.cluster_attn(q, qkv[:, d:2 * d], qkv[:, 2 * d:], G, H, a, centre_only=centre)
        ops.cluster_attn(q, qkv[:, d:2 * d], qkv[:, 2 * d:], G, H, b, centre_only=centre)
        m = n_valid if centre else n_ntgt
        assert torch.equal(a.data[:m], b.data[:m])
        assert torch.equal(b.q8[:m], ops.to_q8(ops.Split(b.data.clone(), d)).q8[:m]) and int(b.q8[:m].max()) > 0
    # PQ decode from the pre-split codebook
    model = synth.make_model(cfg)
    qz = model.decoder.tgt_quantizer.to(dev)
    hi, lo = qz._split_codebook()
    rows = G.ntgt_row
    p0 = ops.pq_gather_decode_presplit(tables["codes"], hi, lo, rows, n_cap=G.node_cap, n_dev=G.n_ntgt_dev)
    p1 = ops.pq_gather_decode_presplit(tables["codes"], hi, lo, rows, n_cap=G.node_cap, n_dev=G.n_ntgt_dev, q8=True)
    assert torch.equal(p0.data[:n_ntgt], p1.data[:n_ntgt])
    assert torch.equal(p1.q8[:n_ntgt], ops.to_q8(ops.Split(p1.data.clone(), d)).q8[:n_ntgt])
    # ... and the hi + companion only form from the pre-quantised codebook
    p2 = ops.pq_gather_decode_hiq8(tables["codes"], hi, qz._q8_codebook(), rows, n_cap=G.node_cap, n_dev=G.n_ntgt_dev)
    assert not p2.has_lo and torch.equal(p2.data[:n_ntgt], p1.data[:n_ntgt, :d]) and torch.equal(p2.q8[:n_ntgt], p1.q8[:n_ntgt])
    # cluster attention, hi + companion only
    c = ops.empty_act(G.node_cap, d, ops.HI_Q8, dev)
    ops.cluster_attn(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], G, H, c)
    full = ops.empty_act(G.node_cap, d, ops.SPLIT_Q8, dev)
    ops.cluster_attn(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], G, H, full)
    assert torch.equal(c.data[:n_ntgt], full.data[:n_ntgt, :d]) and torch.equal(c.q8[:n_ntgt], full.q8[:n_ntgt])


@pytest.mark.parametrize("name,NL", [("c3mini", 3), ("c3mini", 2), ("c1", 1)])
def test_f16f8_rotation_fold_matches_explicit_rotation(name, NL, dev):
    """MATH_F16F8 folds the OPQ rotation `x @ A` (pq_wrapper.py:202) into HGT layer 0 (Q|K'|V' = x (W A^T)^T, the residual of
    hgt.py:403 inside the output projection [t | x] [W_a | A^T]^T, the inter K' / V' on the token side) and decodes straight into
    fp16 hi + e4m3 companion: same log-probs as the explicit rotation GEMM of the same mode to fp32 re-association, and both
    within 1e-4 of the oracle; 1, 2 and 3 layers (centre-only decode / centre-only layer 0 / full layer 0)."""
    _need_tc()
    import copy
    from gnnlm_b200 import synth
    from tests.synth import make_problem, run_oracle
    cfg, model, data = make_problem(name)
    cfg = dict(cfg, NL=NL)
    model = synth.make_model(cfg)
    ref = run_oracle((cfg, model, data))
    outs = {}
    for fold in (True, False):
        m = copy.deepcopy(model)
        m.decoder.fold_rotation = fold
        outs[fold] = synth.run_gpu(cfg, m, data, dev, "f16f8")
        np.testing.assert_allclose(outs[fold]["logprob"], ref["logprob"].numpy(), rtol=1e-4, atol=1e-4)
    assert np.abs(outs[True]["logprob"] - outs[False]["logprob"]).max() < 2e-5 * np.abs(outs[False]["logprob"]).max()
    assert not np.array_equal(outs[True]["gcn_feat"], outs[False]["gcn_feat"])          # two different evaluation orders did run


@pytest.mark.parametrize("math", ["tf32x3", "f16x3", "f16f8"])
@pytest.mark.parametrize("name", ["c1", "c3mini"])
def test_whole_path_tf32x3(name, math, dev):
    """fp32-parity modes on tensor cores (3xTF32 / 3xFP16 splits): same 1e-4 bar as the CUDA-core fp32 mode."""
    _need_tc()
    from tests.synth import make_problem, run_gpu, run_oracle
    prob = make_problem(name)
    ref = run_oracle(prob)
    out = run_gpu(prob, dev, math=math)
    np.testing.assert_allclose(out["logprob"], ref["logprob"].numpy(), rtol=1e-4, atol=1e-4)
    assert abs(out["nll"] - ref["nll"]) < 0.01 / 16.8
    assert (out["recall"] == ref["knn_recall"].numpy()).all()


def test_whole_path_fp32_c3mini(dev):
    from tests.synth import make_problem, run_gpu, run_oracle
    prob = make_problem("c3mini")
    ref = run_oracle(prob)
    out = run_gpu(prob, dev, math="fp32")
    np.testing.assert_allclose(out["logprob"], ref["logprob"].numpy(), rtol=1e-4, atol=1e-4)
    assert abs(out["nll"] - ref["nll"]) < 0.01 / 16.8


# ------------------------------------------------------------------------------------------ dataset + eval loop
def test_eval_lm_dataset_vs_oracle(dev):
    """GraphTokenBlockDataset -> collater -> move_to_cuda (on-device graph assembly) -> evaluate(), over a token
    stream with a ragged last block and a gcn-context-window, against the CPU oracle run block by block."""
    from types import SimpleNamespace
    import copy
    from gnnlm_b200 import synth
    from gnnlm_b200.dataset import DeviceDatastore, GraphTokenBlockDataset
    from gnnlm_b200.eval_lm import evaluate
    from gnnlm_b200.knn_model import KNNModel
    from gnnlm_b200.sequence_scorer import SequenceScorer
    from oracle import model_oracle as mo
    from tests.synth import oracle_model
    cfg = dict(synth.CONFIGS["c1"], NL=2, k=4, n_d=1 << 16, k_nn=16)
    model = synth.make_model(cfg)
    rng = np.random.RandomState(5)
    n_tok, blk, cw = 600, 128, 8
    tables = synth.make_tables(cfg, device="cpu")
    tokens = rng.randint(4, cfg["V"], size=n_tok).astype(np.int64)
    nbr = rng.randint(1, cfg["n_d"] - 1, size=(n_tok, cfg["k"])).astype(np.int64)
    nbr[rng.rand(n_tok, cfg["k"]) < 0.05] = -1
    feats = rng.randn(n_tok, cfg["d"]).astype(np.float16)
    kd = rng.randn(n_tok, cfg["k_nn"]).astype(np.float32)
    ki = rng.randint(0, cfg["n_d"], size=(n_tok, cfg["k_nn"])).astype(np.int64)
    ds = GraphTokenBlockDataset(tokens, blk, pad=1, eos=2, neighbor_offsets=nbr, n_datastore=cfg["n_d"], neighbor_context=1,
                                precompute_feats=feats, context_window=cw, knn_dists=kd, knn_ids=ki)
    dstore = DeviceDatastore(tables["codes"].to(dev), tables["vals"].to(dev))
    scorer = SequenceScorer(synth.Dictionary(cfg["V"]), args=SimpleNamespace(lmbda=cfg["lmbda"], knn_keytype=None))
    knn = KNNModel(dstore.vals, vocab_size=cfg["V"])
    res = evaluate(copy.deepcopy(model).to(dev).set_math("fp32"), ds, dstore, scorer, knn_dstore=knn, temperature=1.0,
                   max_sentences=2, device=dev)
    # oracle, block by block (context tokens are scored by the model but excluded from the loss)
    om = oracle_model(cfg, model)
    tot, cnt = 0.0, 0
    for i in range(len(ds)):
        it = ds[i]
        cs, e = it["offsets"]
        batch = {"nbr": nbr[cs:e][None], "offsets": np.arange(cs, e)[None], "tgt_feats": torch.from_numpy(feats[cs:e]).float(),
                 "target": it["target"], "codes": tables["codes"].numpy(), "cl": 1, "cr": 1, "n_d": cfg["n_d"]}
        k_ = {"dists": torch.from_numpy(kd[cs:e]), "ids": torch.from_numpy(ki[cs:e]), "vals": tables["vals"].long(),
              "lmbda": cfg["lmbda"], "temperature": 1.0}
        out = mo.eval_batch(om, batch, k_)
        lp = out["logprob"][it["start_idx"]:]
        tot += float(lp.double().sum())
        cnt += lp.numel()
    assert res["count"] == cnt == n_tok
    assert abs(res["score_sum"] - tot) / abs(tot) < 1e-5
    # --save-knnlm-dstore with gcn_feat keys: same files / dtypes as eval_lm.py:178-244
    import tempfile, json as _json
    from gnnlm_b200.eval_lm import DstoreWriter
    with tempfile.TemporaryDirectory() as tmp:
        w = DstoreWriter(tmp, "valid", n_tok, cfg["d"], cfg["V"], dstore_fp16=True, knn_keytype="gcn_feat")
        m2 = copy.deepcopy(model).to(dev).set_math("fp32")
        evaluate(m2, ds, dstore, scorer, knn_dstore=knn, max_sentences=2, device=dev, dstore_writer=w, knn_keytype="gcn_feat")
        assert w.close() == n_tok
        d_ = os.path.join(tmp, "valid_dstore-gcn_feat")
        info = _json.load(open(os.path.join(d_, "info.json")))
        assert info == {"dstore_size": n_tok, "hidden_size": cfg["d"], "vocab_size": cfg["V"], "dstore_fp16": True, "val_size": 1}
        keys = np.memmap(os.path.join(d_, "keys.npy"), dtype=np.float16, mode="r", shape=(n_tok, cfg["d"]))
        vals = np.memmap(os.path.join(d_, "vals.npy"), dtype=np.int16, mode="r", shape=(n_tok, 1))
        assert (vals.reshape(-1) == tokens).all()
        # keys of the first block == oracle gcn features, fp16-rounded
        it = ds[0]
        batch0 = {"nbr": nbr[:blk][None], "offsets": np.arange(blk)[None], "tgt_feats": torch.from_numpy(feats[:blk]).float(),
                  "target": it["target"], "codes": tables["codes"].numpy(), "cl": 1, "cr": 1, "n_d": cfg["n_d"]}
        o0 = mo.eval_batch(om, batch0, None)
        np.testing.assert_allclose(np.asarray(keys[:blk], dtype=np.float32), o0["gcn_feat"].numpy(), rtol=2e-3, atol=2e-3)
    assert abs(res["ppl"] - mo.perplexity(tot, cnt)[1]) / res["ppl"] < 1e-4
    # whole-step CUDA-graph replay (one capture per batch shape: 2-block batches, then the ragged tail) == eager
    for math_ in ("fp32", "f16x3"):
        if math_ != "fp32":
            _need_tc()
        m3 = copy.deepcopy(model).to(dev).set_math(math_)
        eager = evaluate(m3, ds, dstore, scorer, knn_dstore=knn, max_sentences=2, device=dev)
        graphed = evaluate(m3, ds, dstore, scorer, knn_dstore=knn, max_sentences=2, device=dev, cuda_graph=True)
        assert graphed["count"] == eager["count"] == n_tok
        assert abs(graphed["score_sum"] - eager["score_sum"]) <= 1e-9 * abs(eager["score_sum"])


def test_inter_attn_many_tokens_split_launches(dev):
    """More tokens than one launch of the tensor-core inter kernel stages edge ranges for (148 SMs x 512): the library splits the
    token range over several launches.  Scoring the range in one call == scoring it in slices of 9,000 tokens (bit-identical), and
    sampled tokens match the fp64 statement of hgt.py:339-358."""
    from gnnlm_b200 import ops
    _need_tc()
    d, H, T = 512, 8, 80000
    dk = d // H
    g = torch.Generator().manual_seed(1)
    degs = (torch.arange(T) % 4 == 0).long() * 0 + (torch.arange(T) % 4).clamp(max=2)          # 0, 1, 2, 2, 0, 1, ...
    degs[-1] = 19                                                                             # a two-tile token at the very end
    indptr = torch.zeros(T + 1, dtype=torch.int32)
    indptr[1:] = torch.cumsum(degs, 0)
    n_c = int(indptr[-1])
    q = torch.randn(T, d, generator=g).to(dev)
    hc = torch.randn(n_c, d, generator=g).to(dev)
    Wk = (torch.randn(d, d, generator=g) / d ** 0.5 * 0.3).to(dev)
    Wv = (torch.randn(d, d, generator=g) / d ** 0.5).to(dev)
    bv = torch.randn(d, generator=g).to(dev)
    wk_t = ops.split_f16(Wk.view(H, dk, d).transpose(1, 2).contiguous().view(H * d, dk))
    wv = ops.split_f16(Wv.contiguous())
    hs, ip = ops.to_split(hc), indptr.to(dev)
    whole = torch.full((T, d), float("nan"), device=dev)
    ops.inter_attn_fused(q, [(0, T, ip, hs)], H, whole, wk_t, wv, bv, out_scale=0.5)
    sliced = torch.full((T, d), float("nan"), device=dev)
    ops.inter_attn_fused(q, [(t0, min(9000, T - t0), ip[t0:], hs) for t0 in range(0, T, 9000)], H, sliced, wk_t, wv, bv, out_scale=0.5)
    assert torch.isfinite(whole).all() and torch.equal(whole, sliced)
    for t in (1, 2, 3, 75775, 75776, 75777, 79998, T - 1):
        e0, e1 = int(indptr[t]), int(indptr[t + 1])
        K = (hc[e0:e1].double() @ Wk.double().T).view(-1, H, dk)
        V = (hc[e0:e1].double() @ Wv.double().T + bv.double()).view(-1, H, dk)
        s_ = torch.einsum("hj,chj->ch", q[t].double().view(H, dk), K)
        ref = 0.5 * torch.einsum("ch,chj->hj", torch.softmax(s_, 0), V).reshape(d)
        assert (whole[t].double() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()      # a token without centres: 0 == 0


def test_eval_lm_main_from_checkpoint_and_data_dir(dev, tmp_path):
    """eval_lm.main: reference command line + a fairseq-format checkpoint ({'args', 'model'} with the bypassed base-transformer
    keys still in it) + a reference data directory -> the same scores as evaluate() on the in-memory model and arrays; with
    --knnlm the neighbours come from neighbors.mmap.{k} and the similarities are recomputed from the PQ codes; two shards
    (--num-shards 2) partition the blocks."""
    from argparse import Namespace
    from types import SimpleNamespace
    import copy, json as _json
    from gnnlm_b200 import synth
    from gnnlm_b200.dataset import DeviceDatastore, GraphTokenBlockDataset
    from gnnlm_b200.eval_lm import evaluate, main
    from gnnlm_b200.knn_model import KNNModel
    from gnnlm_b200.sequence_scorer import SequenceScorer
    from tests.test_formats import write_mmap_indexed
    cfg = dict(synth.CONFIGS["c1"], NL=2, k=4, n_d=1 << 14, V=1000, cutoff=[200, 600], k_nn=8)
    model = synth.make_model(cfg)
    tables = synth.make_tables(cfg, device="cpu")
    rng = np.random.RandomState(11)
    lens = [64, 64, 64, 40]
    sents = [np.concatenate([rng.randint(4, cfg["V"], size=n - 1), [2]]).astype(np.int64) for n in lens]
    n_tok = sum(lens)
    root = str(tmp_path / "data-bin")
    os.makedirs(os.path.join(root, "valid_dstore"))
    os.makedirs(os.path.join(root, "train_dstore"))
    with open(os.path.join(root, "dict.txt"), "w") as f:
        for i in range(4, cfg["V"]):
            f.write(f"w{i} {cfg['V'] - i}\n")
    write_mmap_indexed(os.path.join(root, "valid"), sents, np.uint16)
    nbr = rng.randint(1, cfg["n_d"] - 1, size=(n_tok, cfg["k"])).astype(np.int64)
    knn_ids = rng.randint(0, cfg["n_d"], size=(n_tok, cfg["k_nn"])).astype(np.int64)
    feats = rng.randn(n_tok, cfg["d"]).astype(np.float16)
    nbr.tofile(os.path.join(root, "valid_dstore", f"neighbors.mmap.{cfg['k']}"))
    knn_ids.tofile(os.path.join(root, "valid_dstore", f"neighbors.mmap.{cfg['k_nn']}"))
    feats.tofile(os.path.join(root, "valid_dstore", "keys.npy"))
    info = {"hidden_size": cfg["d"], "vocab_size": cfg["V"], "dstore_fp16": True, "val_size": 1}
    _json.dump(dict(info, dstore_size=n_tok), open(os.path.join(root, "valid_dstore", "info.json"), "w"))
    _json.dump(dict(info, dstore_size=cfg["n_d"]), open(os.path.join(root, "train_dstore", "info.json"), "w"))
    tables["vals"].numpy().astype(np.int16).reshape(-1, 1).tofile(os.path.join(root, "train_dstore", "vals.npy"))
    np.save(os.path.join(root, "train_dstore", "quantized-keys.npy"), tables["codes"].numpy())
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["decoder.layers.0.self_attn.k_proj.weight"] = torch.zeros(4, 4)          # the bypassed 16-layer transformer: ignored
    sd["decoder.embed_tokens.embeddings.0.0.weight"] = torch.zeros(4, 4)
    sd["decoder.version"] = torch.tensor([3.0])
    ckpt_args = Namespace(arch="transformer_lm", decoder_embed_dim=cfg["d"], decoder_attention_heads=cfg["H"], graph_layer=2,
                          decoder_gcn_dim=cfg["d"], adaptive_softmax_cutoff="200,600", tie_adaptive_weights=False,
                          quantizer_path="/somewhere/else/quantizer", task="language_modeling", tokens_per_sample=3072)
    ckpt = str(tmp_path / "checkpoint_best.pt")
    torch.save({"args": ckpt_args, "model": sd}, ckpt)
    argv = [root, "--path", ckpt, "--graph", "--use-precompute-feat", "--gen-subset", "valid", "--tokens-per-sample", "64",
            "--gcn-k", str(cfg["k"]), "--neighbor-context", "1", "--math", "fp32", "--max-sentences", "2"]
    lines = []
    got = main(argv, device=dev, log=lines.append)
    assert any(l.startswith("Loss (base 2):") for l in lines) and any("examples" in l for l in lines)
    flat = np.concatenate(sents)
    ds_mem = GraphTokenBlockDataset(flat, 64, pad=1, eos=2, neighbor_offsets=nbr, n_datastore=cfg["n_d"], neighbor_context=1,
                                    precompute_feats=feats, knn_dists=np.zeros((n_tok, cfg["k_nn"]), np.float32), knn_ids=knn_ids)
    dstore = DeviceDatastore(tables["codes"].to(dev), tables["vals"].to(torch.int16).to(dev))
    m = copy.deepcopy(model).to(dev).set_math("fp32")
    plain = SequenceScorer(synth.Dictionary(cfg["V"]), args=SimpleNamespace(lmbda=0.0, knn_keytype=None))
    want = evaluate(m, ds_mem, dstore, plain, max_sentences=2, device=dev)
    assert got["count"] == want["count"] == n_tok and got["score_sum"] == want["score_sum"]
    # kNN-LM from the precomputed neighbour file, similarities recomputed against the PQ-decoded keys
    got_knn = main(argv + ["--knnlm", "--k", str(cfg["k_nn"]), "--lmbda", "0.25", "--knn-sim-func", "ip", "--temperature", "10"],
                   device=dev, log=lines.append)
    knn = KNNModel(dstore.vals, vocab_size=cfg["V"], metric_type="ip", k=cfg["k_nn"], pq_codes=dstore.codes,
                   quantizer=m.decoder.tgt_quantizer)
    mixed = SequenceScorer(synth.Dictionary(cfg["V"]), args=SimpleNamespace(lmbda=0.25, knn_keytype=None))
    want_knn = evaluate(m, ds_mem, dstore, mixed, knn_dstore=knn, temperature=10.0, max_sentences=2, device=dev)
    assert got_knn["score_sum"] == want_knn["score_sum"] != want["score_sum"]
    with pytest.raises(ValueError):               # distances are not on disk and the metric does not recompute them
        main(argv + ["--knnlm", "--k", str(cfg["k_nn"]), "--lmbda", "0.25"], device=dev, log=lines.append)
    # --num-shards / --shard-id: contiguous block ranges, scores add up
    parts = [main(argv + ["--num-shards", "2", "--shard-id", str(i)], device=dev, log=lines.append) for i in range(2)]
    assert parts[0]["count"] + parts[1]["count"] == n_tok
    assert abs(parts[0]["score_sum"] + parts[1]["score_sum"] - want["score_sum"]) < 1e-9 * abs(want["score_sum"])
    # --save-knnlm-dstore is shard-aware: two separately launched shards fill the row ranges of ONE keys.npy / vals.npy
    # (fairseq_cli/eval_lm.py:178-244), identical to the single-process datastore
    one, two = str(tmp_path / "ds1"), str(tmp_path / "ds2")
    r1 = main(argv + ["--save-knnlm-dstore", "--dstore-mmap", one, "--knn-keytype", "gcn_feat", "--dstore-fp16"], device=dev,
              log=lines.append)
    rs = [main(argv + ["--save-knnlm-dstore", "--dstore-mmap", two, "--knn-keytype", "gcn_feat", "--dstore-fp16", "--num-shards", "2",
                       "--shard-id", str(i)], device=dev, log=lines.append) for i in (1, 0)]
    assert r1["dstore_items"] == n_tok and rs[0]["dstore_items"] + rs[1]["dstore_items"] == n_tok
    for name, dt in (("keys.npy", np.float16), ("vals.npy", np.int16)):
        a = np.fromfile(os.path.join(one, "valid_dstore-gcn_feat", name), dtype=dt)
        b = np.fromfile(os.path.join(two, "valid_dstore-gcn_feat", name), dtype=dt)
        assert a.shape == b.shape and (a == b).all() and np.abs(a.astype(np.float64)).sum() > 0
    assert (np.fromfile(os.path.join(one, "valid_dstore-gcn_feat", "vals.npy"), dtype=np.int16) == flat).all()
    # a plain-PQ checkpoint (--index PQ64: no OPQ transform, convert_ckpt.py:40-45 writes neither A nor b) loads and runs
    sd_pq = {k: v for k, v in sd.items() if k not in ("decoder.tgt_quantizer.A", "decoder.tgt_quantizer.b")}
    ckpt_pq = str(tmp_path / "checkpoint_pq.pt")
    torch.save({"args": ckpt_args, "model": sd_pq}, ckpt_pq)
    got_pq = main([root, "--path", ckpt_pq] + argv[3:], device=dev, log=lines.append)
    from gnnlm_b200.pq_codec import TorchPQCodec
    m_pq = copy.deepcopy(model)
    m_pq.decoder.tgt_quantizer = TorchPQCodec(centroids=model.decoder.tgt_quantizer.centroids_torch.numpy())
    want_pq = evaluate(m_pq.to(dev).set_math("fp32"), ds_mem, dstore, plain, max_sentences=2, device=dev)
    assert got_pq["count"] == n_tok and got_pq["score_sum"] == want_pq["score_sum"] != want["score_sum"]
    # a checkpoint convert_ckpt.py was never run on (no decoder.tgt_quantizer.* at all) + the faiss `quantizer` file its
    # --quantizer_path names (transformer.py:936-937): the codec is parsed from the file (formats.read_faiss_quantizer)
    from gnnlm_b200.formats import write_faiss_quantizer
    q_ = model.decoder.tgt_quantizer
    qfile = str(tmp_path / "quantizer")
    write_faiss_quantizer(qfile, q_.centroids_torch.numpy(), q_.A.numpy(), q_.b.numpy())
    sd_raw = {k: v for k, v in sd.items() if not k.startswith("decoder.tgt_quantizer.")}
    ckpt_raw = str(tmp_path / "checkpoint_raw.pt")
    raw_args = Namespace(**dict(vars(ckpt_args), quantizer_path=qfile))
    torch.save({"args": raw_args, "model": sd_raw}, ckpt_raw)
    got_raw = main([root, "--path", ckpt_raw] + argv[3:], device=dev, log=lines.append)
    assert got_raw["count"] == n_tok and got_raw["score_sum"] == want["score_sum"]


def test_knn_model_precomputed_arrays(dev):
    """KNNModel(knns=..., dists=...) -- the whole-split constructor: get_knns(positions=[B, L]) returns one row of search results
    per token ([B*L, k], fp32 / int64), and get_knn_prob equals the set_search_results route."""
    from gnnlm_b200.knn_model import KNNModel
    g = torch.Generator().manual_seed(4)
    n_d, V, n_split, k, B, Lb = 5000, 300, 64, 16, 2, 8
    vals = torch.randint(4, V, (n_d,), generator=g, dtype=torch.int32).to(dev)
    dists = torch.randn((n_split, k), generator=g).half().to(dev)            # stored fp16: must come back fp32
    knns = torch.randint(0, n_d, (n_split, k), generator=g, dtype=torch.int32).to(dev)
    knns[3, 5] = -1
    pos = torch.randperm(n_split, generator=g)[:B * Lb].view(B, Lb).to(dev)
    tgt = torch.randint(4, V, (B, Lb), generator=g).to(dev)
    whole = KNNModel(vals, vocab_size=V, dists=dists, knns=knns, k=k)
    d_, i_ = whole.get_knns(None, positions=pos)
    assert d_.shape == i_.shape == (B * Lb, k) and d_.dtype == torch.float32 and i_.dtype == torch.int64
    p1, r1 = whole.get_knn_prob(None, targets=tgt, return_recall=True, positions=pos)
    step = KNNModel(vals, vocab_size=V, k=k)
    step.set_search_results(dists[pos.reshape(-1)].float(), knns[pos.reshape(-1)].long())
    p2, r2 = step.get_knn_prob(None, targets=tgt, return_recall=True)
    assert p1.shape == (B * Lb,) and torch.equal(p1, p2) and torch.equal(r1, r2)


def test_evaluate_host_pipeline_matches_reference_shaped_calls(dev, tmp_path):
    """evaluate()'s producer thread (HostBatcher: memmap slices copied once into reusable pinned buffers,
    GraphTokenBlockDataset.collate_into) against the reference-shaped per-batch calls dataset[i] + collater
    (token_block_dataset.py:287-333, monolingual_dataset.py:237-262): the same device inputs, hence the same scores -- with a
    --gcn-context-window, B = 2 batches, a ragged last block and kNN-LM arrays; a foreign neighbour id surfaces as IndexError."""
    from types import SimpleNamespace
    import copy
    from gnnlm_b200 import synth
    from gnnlm_b200.dataset import DeviceDatastore, GraphTokenBlockDataset
    from gnnlm_b200.eval_lm import evaluate
    from gnnlm_b200.knn_model import KNNModel
    from gnnlm_b200.sequence_scorer import SequenceScorer
    cfg = dict(synth.CONFIGS["c1"], NL=2, k=4, n_d=1 << 14, V=1000, cutoff=[200, 600], k_nn=8)
    model = synth.make_model(cfg)
    tables = synth.make_tables(cfg, device="cpu")
    rng = np.random.RandomState(7)
    n_tok = 64 * 7 + 23
    tokens = rng.randint(4, cfg["V"], size=n_tok).astype(np.uint16)
    nbr = rng.randint(1, cfg["n_d"] - 1, size=(n_tok, cfg["k"])).astype(np.int64)
    nbr[rng.rand(n_tok, cfg["k"]) < 0.05] = -1
    feats = rng.randn(n_tok, cfg["d"]).astype(np.float16)
    kid = rng.randint(0, cfg["n_d"], size=(n_tok, cfg["k_nn"])).astype(np.int64)
    kd = rng.randn(n_tok, cfg["k_nn"]).astype(np.float32)
    dstore = DeviceDatastore(tables["codes"].to(dev), tables["vals"].to(dev))
    scorer = SequenceScorer(synth.Dictionary(cfg["V"]), args=SimpleNamespace(lmbda=0.25, knn_keytype=None))
    m = copy.deepcopy(model).to(dev).set_math("fp32")
    for cw in (0, 16):
        ds = GraphTokenBlockDataset(tokens, 64, pad=1, eos=2, neighbor_offsets=nbr, n_datastore=cfg["n_d"], neighbor_context=1,
                                    precompute_feats=feats, context_window=cw, knn_dists=kd, knn_ids=kid)
        res = {}
        for threads in (True, False, "workers", "mmap"):
            knn = KNNModel(dstore.vals, vocab_size=cfg["V"], k=cfg["k_nn"])
            d_ = ds
            if threads == "mmap":            # read-only file mappings (what a data directory gives)
                def mm(name, a):
                    a.tofile(str(tmp_path / f"{name}{cw}"))
                    return np.memmap(str(tmp_path / f"{name}{cw}"), dtype=a.dtype, mode="r", shape=a.shape)
                d_ = GraphTokenBlockDataset(mm("t", tokens), 64, pad=1, eos=2, neighbor_offsets=mm("n", nbr), n_datastore=cfg["n_d"],
                                            neighbor_context=1, precompute_feats=mm("f", feats), context_window=cw,
                                            knn_dists=mm("kd", kd), knn_ids=mm("ki", kid))
            res[threads] = evaluate(m, d_, dstore, scorer, knn_dstore=knn, max_sentences=2, device=dev, host_threads=bool(threads),
                                    host_workers=2 if threads == "workers" else 0)
        for k_ in (True, "workers", "mmap"):
            assert res[k_]["count"] == res[False]["count"] == n_tok
            assert abs(res[k_]["score_sum"] - res[False]["score_sum"]) <= 1e-12 * abs(res[False]["score_sum"])
        # the producer's tensors are the collater's
        ids = [2, 3]
        spec = ds.batch_spec(ids)
        bufs = {n: torch.empty(sh, dtype=dt) for n, sh, dt in spec}
        fast = ds.collate_into(ids, bufs)["host"]
        from gnnlm_b200.eval_lm import host_inputs
        slow = host_inputs(ds.collater([ds[i] for i in ids]))
        assert set(fast) == set(slow)
        for k_ in slow:
            assert torch.equal(fast[k_], slow[k_].reshape(fast[k_].shape)), k_
    bad = nbr.copy()
    bad[200, 1] = cfg["n_d"]
    ds = GraphTokenBlockDataset(tokens, 64, pad=1, eos=2, neighbor_offsets=bad, n_datastore=cfg["n_d"], neighbor_context=1,
                                precompute_feats=feats)
    with pytest.raises(IndexError):
        evaluate(m, ds, dstore, SequenceScorer(synth.Dictionary(cfg["V"]), args=SimpleNamespace(lmbda=0.0, knn_keytype=None)),
                 max_sentences=2, device=dev)


def test_eval_lm_through_the_registration_face(dev, tmp_path):
    """Reference-style command line -> registry.eval_lm_parser -> (stand-in) fairseq registries -> task.setup_task /
    load_dataset / load_datastore -> ARCH_MODEL_REGISTRY[arch].build_model -> evaluate(): the same score as the same model
    evaluated over the arrays in memory (`--reinit-nfeat` run: no quantizer file needed)."""
    from types import SimpleNamespace
    from gnnlm_b200 import registry
    from gnnlm_b200.dataset import DeviceDatastore, GraphTokenBlockDataset
    from gnnlm_b200.eval_lm import evaluate
    from gnnlm_b200.sequence_scorer import SequenceScorer
    from tests.test_formats import _write_data_dir
    from tests.test_multi_rank_cpu import _fake_fairseq
    fm, ft = _fake_fairseq()
    registry.register_with_fairseq(fm, ft, override=True)
    root = str(tmp_path / "data-bin")
    n_tok, feats, nbr = _write_data_dir(root, k=4, hidden=64, n_d=500)
    vals = np.random.RandomState(3).randint(4, 48, size=(500, 1)).astype(np.int16)
    vals.tofile(os.path.join(root, "train_dstore", "vals.npy"))
    argv = [root, "--graph", "--use-precompute-feat", "--reinit-nfeat", "--graph_layer", "2", "--decoder_gcn_dim", "64",
            "--decoder-embed-dim", "64", "--decoder-attention-heads", "4", "--adaptive-softmax-cutoff", "10,20", "--adaptive-input",
            "--adaptive-input-cutoff", "10,20", "--tokens-per-sample", "8", "--gcn-k", "4", "--neighbor-context", "1",
            "--gen-subset", "valid"]
    args = registry.eval_lm_parser().parse_args(argv)
    task = ft.TASK_REGISTRY["language_modeling"].setup_task(args)
    ds = task.load_dataset(args.gen_subset)
    dstore = task.load_datastore(dev)
    assert dstore.codes is None and dstore.vals.dtype == torch.int16
    torch.manual_seed(0)
    model = fm.ARCH_MODEL_REGISTRY[args.arch].build_model(args, task).eval().to(dev).set_math("fp32")
    assert "decoder.embed_tokens.embeddings.2.1.weight" in model.state_dict()
    assert model.load_reference_state_dict({k: v.clone() for k, v in model.state_dict().items()}) == []     # strict round trip
    scorer = SequenceScorer(task.target_dictionary, args.softmax_batch, args=args)
    got = evaluate(model, ds, dstore, scorer, max_sentences=2, device=dev)
    z = np.load(os.path.join(GOLD, "fmt.npz"))
    ds_mem = GraphTokenBlockDataset(z["uint16_flat"], 8, pad=1, eos=2, neighbor_offsets=nbr, n_datastore=500, neighbor_context=1,
                                    precompute_feats=feats)
    want = evaluate(model, ds_mem, DeviceDatastore(None, torch.from_numpy(vals.reshape(-1)).to(dev)), scorer, max_sentences=2, device=dev)
    assert got["count"] == want["count"] == n_tok and np.isfinite(got["ppl"])
    assert got["score_sum"] == want["score_sum"]


def test_adaptive_input_mirror_golden(dev):
    """model.AdaptiveInput (projected-table form) against the reference's AdaptiveInput.forward: strict state_dict load with the
    reference's keys, every band and cutoff edge."""
    from gnnlm_b200.model import AdaptiveInput
    z = np.load(os.path.join(GOLD, "adaptive_input_v300.npz"))
    cutoff = [int(c) for c in z["cutoff"]]
    m = AdaptiveInput(int(z["V"]), 1, int(z["d"]), 4, int(z["d"]), cutoff[:-1])
    m.load_state_dict(_sd(z, "sd."), strict=True)
    m = m.to(dev)
    out = m(torch.from_numpy(z["tokens"]).to(dev))
    np.testing.assert_allclose(out.cpu().numpy(), z["out"], rtol=1e-5, atol=1e-6)     # fp32 GEMM re-association only
    assert m.table().shape == (int(z["V"]), int(z["d"]))


@pytest.mark.parametrize("math,adaptive,NL", [("fp32", True, 2), ("f16x3", True, 3), ("fp32", False, 1)])
def test_eval_lm_reinit_nfeat(math, adaptive, NL, dev):
    """--reinit-nfeat (language_modeling.py:273-278, transformer.py:1046-1048): no code rows, ntgt features =
    embed_tokens(neighbour tokens) -- adaptive input and plain embedding -- through evaluate() against the oracle."""
    from types import SimpleNamespace
    import copy
    from gnnlm_b200 import synth
    from gnnlm_b200.dataset import DeviceDatastore, GraphTokenBlockDataset
    from gnnlm_b200.eval_lm import evaluate
    from gnnlm_b200.model import TransformerLanguageModel, default_args
    from gnnlm_b200.sequence_scorer import SequenceScorer
    from oracle import model_oracle as mo
    from tests.synth import oracle_model
    if math != "fp32":
        _need_tc()
    cfg = dict(synth.CONFIGS["c1"], NL=NL, k=4, n_d=1 << 14, V=1000, cutoff=[200, 600])
    torch.manual_seed(3)
    args = default_args(decoder_embed_dim=cfg["d"], decoder_attention_heads=cfg["H"], graph_layer=NL,
                        adaptive_softmax_cutoff=cfg["cutoff"], reinit_nfeat=True, adaptive_input=adaptive,
                        adaptive_input_cutoff="200,600", adaptive_input_factor=4)
    model = TransformerLanguageModel.build_model(args, dictionary=synth.Dictionary(cfg["V"])).eval()
    assert model.decoder.tgt_quantizer is None
    sd_keys = set(model.state_dict())
    want_key = "decoder.embed_tokens.embeddings.1.1.weight" if adaptive else "decoder.embed_tokens.weight"
    assert want_key in sd_keys
    tables = synth.make_tables(cfg, device="cpu")
    rng = np.random.RandomState(5)
    n_tok, blk = 200, 64
    tokens = rng.randint(4, cfg["V"], size=n_tok).astype(np.int64)
    nbr = rng.randint(1, cfg["n_d"] - 1, size=(n_tok, cfg["k"])).astype(np.int64)
    nbr[rng.rand(n_tok, cfg["k"]) < 0.05] = -1
    feats = rng.randn(n_tok, cfg["d"]).astype(np.float16)
    ds = GraphTokenBlockDataset(tokens, blk, pad=1, eos=2, neighbor_offsets=nbr, n_datastore=cfg["n_d"], neighbor_context=1,
                                precompute_feats=feats)
    dstore = DeviceDatastore(None, tables["vals"].to(dev))
    scorer = SequenceScorer(synth.Dictionary(cfg["V"]), args=SimpleNamespace(lmbda=0.0, knn_keytype=None))
    res = evaluate(copy.deepcopy(model).to(dev).set_math(math), ds, dstore, scorer, max_sentences=1, device=dev)
    # oracle
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    om = {"sd": {k[len("decoder.hgt_decoder."):]: v for k, v in sd.items() if k.startswith("decoder.hgt_decoder.")},
          "n_heads": cfg["H"], "n_layers": NL, "cutoff": list(cfg["cutoff"]) + [cfg["V"]],
          "softmax": mo.adaptive_weights({k[len("decoder.adaptive_softmax."):]: v for k, v in sd.items()
                                          if k.startswith("decoder.adaptive_softmax.")})}
    if adaptive:
        om["embed_cutoff"] = [200, 600, cfg["V"]]
        om["embed_bands"] = [(sd[f"decoder.embed_tokens.embeddings.{i}.0.weight"], sd[f"decoder.embed_tokens.embeddings.{i}.1.weight"])
                             for i in range(3)]
    else:
        om["embed_cutoff"], om["embed_bands"] = [cfg["V"]], [(sd["decoder.embed_tokens.weight"], None)]
    tot, cnt = 0.0, 0
    for i in range(len(ds)):
        it = ds[i]
        cs, e = it["offsets"]
        batch = {"nbr": nbr[cs:e][None], "offsets": np.arange(cs, e)[None], "tgt_feats": torch.from_numpy(feats[cs:e]).float(),
                 "target": it["target"], "vals": tables["vals"].numpy(), "reinit_nfeat": True, "cl": 1, "cr": 1, "n_d": cfg["n_d"]}
        out = mo.eval_batch(om, batch, None)
        tot += float(out["logprob"].double().sum())
        cnt += out["logprob"].numel()
    assert res["count"] == cnt == n_tok
    assert abs(res["score_sum"] - tot) / abs(tot) < 1e-5        # summed log-probs; per-token parity is covered by the kernels' own tests


def test_eval_lm_from_reference_data_dir(dev, tmp_path):
    """A data directory laid out as the reference's task expects it (dict.txt, valid.bin/.idx, {valid,train}_dstore/ with
    info.json, raw keys / vals / neighbour memmaps, quantized-keys.npy; knn/path_utils.py:13-41) -> formats.load_graph_lm_dataset
    + DeviceDatastore.from_dir -> evaluate(): identical to evaluating the same arrays handed over in memory."""
    from types import SimpleNamespace
    import copy, json as _json
    from gnnlm_b200 import synth
    from gnnlm_b200.dataset import DeviceDatastore, GraphTokenBlockDataset
    from gnnlm_b200.eval_lm import evaluate
    from gnnlm_b200.formats import load_graph_lm_dataset
    from gnnlm_b200.sequence_scorer import SequenceScorer
    from tests.test_formats import write_mmap_indexed
    cfg = dict(synth.CONFIGS["c1"], NL=2, k=4, n_d=1 << 14, V=1000, cutoff=[200, 600])
    model = synth.make_model(cfg)
    tables = synth.make_tables(cfg, device="cpu")
    rng = np.random.RandomState(11)
    lens = [40, 7, 64, 1, 90, 33, 65]
    sents = [np.concatenate([rng.randint(4, cfg["V"], size=n - 1), [2]]).astype(np.int64) for n in lens]
    n_tok = sum(lens)
    root = str(tmp_path / "data-bin")
    os.makedirs(os.path.join(root, "valid_dstore"))
    os.makedirs(os.path.join(root, "train_dstore"))
    with open(os.path.join(root, "dict.txt"), "w") as f:
        for i in range(4, cfg["V"]):
            f.write(f"w{i} {cfg['V'] - i}\n")
    write_mmap_indexed(os.path.join(root, "valid"), sents, np.uint16)
    nbr = rng.randint(1, cfg["n_d"] - 1, size=(n_tok, cfg["k"])).astype(np.int64)
    nbr[rng.rand(n_tok, cfg["k"]) < 0.05] = -1
    feats = rng.randn(n_tok, cfg["d"]).astype(np.float16)
    nbr.tofile(os.path.join(root, "valid_dstore", f"neighbors.mmap.{cfg['k']}"))
    feats.tofile(os.path.join(root, "valid_dstore", "keys.npy"))
    info = {"hidden_size": cfg["d"], "vocab_size": cfg["V"], "dstore_fp16": True, "val_size": 1}
    _json.dump(dict(info, dstore_size=n_tok), open(os.path.join(root, "valid_dstore", "info.json"), "w"))
    _json.dump(dict(info, dstore_size=cfg["n_d"]), open(os.path.join(root, "train_dstore", "info.json"), "w"))
    tables["vals"].numpy().astype(np.int16).reshape(-1, 1).tofile(os.path.join(root, "train_dstore", "vals.npy"))   # fp16 & V < 2^15
    np.save(os.path.join(root, "train_dstore", "quantized-keys.npy"), tables["codes"].numpy())
    ds, dictionary = load_graph_lm_dataset(root, "valid", tokens_per_sample=64, gcn_k=cfg["k"], neighbor_context="1")
    assert len(dictionary) == cfg["V"]
    dstore = DeviceDatastore.from_dir(root, len(dictionary), dev)
    assert dstore.vals.dtype == torch.int16 and (dstore.codes.cpu() == tables["codes"]).all()
    scorer = SequenceScorer(dictionary, args=SimpleNamespace(lmbda=0.0, knn_keytype=None))
    m = copy.deepcopy(model).to(dev).set_math("fp32")
    got = evaluate(m, ds, dstore, scorer, max_sentences=2, device=dev)
    flat = np.concatenate(sents)
    ds_mem = GraphTokenBlockDataset(flat, 64, pad=1, eos=2, neighbor_offsets=nbr, n_datastore=cfg["n_d"], neighbor_context=1,
                                    precompute_feats=feats)
    dstore_mem = DeviceDatastore(tables["codes"].to(dev), tables["vals"].to(dev))
    want = evaluate(m, ds_mem, dstore_mem, scorer, max_sentences=2, device=dev)
    assert got["count"] == want["count"] == n_tok
    assert got["score_sum"] == want["score_sum"]              # same kernels, same inputs: bit-identical


@pytest.mark.parametrize("name", ["c1", "c3mini"])
def test_whole_path_bf16(name, dev):
    """bf16 mode (bf16 activations / operands on the ntgt side, fp32 accumulation, fp32 tgt-side attention,
    fp32 pre-norm sums): per-token log-probs within 1e-2 relative of the fp32 oracle (north_star)."""
    _need_tc()
    from tests.synth import make_problem, run_gpu, run_oracle
    prob = make_problem(name)
    ref = run_oracle(prob)
    out = run_gpu(prob, dev, math="bf16")
    np.testing.assert_allclose(out["logprob"], ref["logprob"].numpy(), rtol=1e-2, atol=1e-2)
    assert abs(out["nll"] - ref["nll"]) < 1e-2
    assert (out["recall"] == ref["knn_recall"].numpy()).all()


@pytest.mark.parametrize("M,N,K", [(300, 200, 64), (4096, 1024, 1024), (1000, 3072, 256), (129, 72, 32)])
def test_linear_f16x3_split_format(M, N, K, dev):
    """MATH_F16X3 with the split-fp16 activation format (GNNLM_F16X2) on A, the residual and C: same 1e-4 bar."""
    _need_tc()
    from gnnlm_b200 import _lib as L, ops
    torch.manual_seed(M + N)
    A, W, b, R = torch.randn(M, K), torch.randn(N, K) / K ** 0.5, torch.randn(N), torch.randn(M, N)
    ref = A.double() @ W.double().t() + b.double() + R.double()
    As, Rs = ops.to_split(A.to(dev)), ops.to_split(R.to(dev))
    # the format itself: hi + lo reproduces fp32 to ~2^-22
    np.testing.assert_allclose(As.float().cpu().numpy(), A.numpy(), rtol=1e-6, atol=1e-7)
    Wh, Wl, ws = ops.split_f16(W.to(dev))
    out = ops.linear(As, Wh, b.to(dev), W_lo=Wl, w_scale=ws, residual=Rs, out_dtype=ops.SPLIT, math=L.MATH_F16X3)
    assert isinstance(out, ops.Split)
    np.testing.assert_allclose(out.float().cpu().double().numpy(), ref.numpy(), rtol=1e-4, atol=1e-4)
    out32 = ops.linear(As, Wh, b.to(dev), W_lo=Wl, w_scale=ws, residual=R.to(dev), math=L.MATH_F16X3)
    np.testing.assert_allclose(out32.cpu().double().numpy(), ref.numpy(), rtol=1e-4, atol=1e-4)
    pick = torch.randint(0, N, (M,), dtype=torch.int32)
    pm, ps, pk, nt = ops.linear_lse(As, Wh, pick.to(dev), W_lo=Wl, w_scale=ws, math=L.MATH_F16X3)
    lp = torch.empty(M, device=dev)
    ops.lse_finish(pm, ps, pk, nt, lp)
    ref_lp = torch.log_softmax(A.double() @ W.double().t(), 1).gather(1, pick.long()[:, None]).squeeze(1)
    np.testing.assert_allclose(lp.cpu().double().numpy(), ref_lp.numpy(), rtol=1e-4, atol=1e-4)
    # LayerNorm -> split, row gather of a split matrix
    g, be = torch.randn(N), torch.randn(N)
    y = ops.layernorm(out32, g.to(dev), be.to(dev), out_dtype=ops.SPLIT)
    np.testing.assert_allclose(y.float().cpu().numpy(), torch.nn.functional.layer_norm(out32.cpu(), (N,), g, be).numpy(),
                               rtol=1e-5, atol=1e-5)
    ids = torch.randint(0, M, (77,), dtype=torch.int32)
    assert (ops.gather_rows(y, ids.to(dev)).data.cpu() == y.data.cpu()[ids.long()]).all()


@pytest.mark.parametrize("B,L,H,d,ctx", [(1, 256, 8, 1024, 0), (2, 320, 8, 512, 0), (1, 512, 8, 1024, 100), (1, 1288, 8, 1024, 0),
                                          (1, 3072, 8, 1024, 0), (1, 1024, 4, 512, 300)])
def test_causal_attn_gemm_form(B, L, H, d, ctx, dev):
    """Tensor-core (3xFP16 GEMM) form of the tgt-intra-tgt attention vs the fp64 definition and vs the flash kernel."""
    _need_tc()
    from gnnlm_b200 import ops
    torch.manual_seed(7)
    q, k, v = torch.randn(B * L, d) * 0.3, torch.randn(B * L, d) * 0.3, torch.randn(B * L, d)
    base = torch.randn(B * L, d)
    out = base.clone().to(dev)
    ops.causal_attn_gemm(q.to(dev), k.to(dev), v.to(dev), B, L, ctx, H, out, out_scale=0.5, accumulate=True)
    out2 = base.clone().to(dev)
    ops.causal_attn(q.to(dev), k.to(dev), v.to(dev), B, L, ctx, H, out2, out_scale=0.5, accumulate=True)
    qh, kh, vh = (t.view(B, L, H, d // H).permute(0, 2, 1, 3).double() for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2)
    i = torch.arange(L)
    mask = (i[None, :] <= i[:, None]) & ((i[:, None] - i[None, :] < ctx) if ctx else True)
    ref = base.double() + 0.5 * (torch.softmax(s.masked_fill(~mask, -float("inf")), -1) @ vh).permute(0, 2, 1, 3).reshape(B * L, d)
    np.testing.assert_allclose(out.cpu().double().numpy(), ref.numpy(), rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(out.cpu().numpy(), out2.cpu().numpy(), rtol=1e-4, atol=5e-5)


@pytest.mark.parametrize("B,L,H,d,ctx", [(1, 256, 8, 1024, 0), (2, 320, 8, 512, 0), (1, 512, 8, 1024, 100), (1, 1288, 8, 1024, 0),
                                          (1, 3072, 8, 1024, 0), (1, 1024, 4, 512, 300), (3, 200, 8, 1024, 0), (2, 77, 8, 512, 5),
                                          (1, 33, 8, 1024, 1)])
@pytest.mark.parametrize("split_out", [False, True])
def test_causal_flash_kernel(B, L, H, d, ctx, split_out, dev):
    """gnnlm_hgt_causal_flash (one mma.sync flash kernel, K' / V' as split fp16) vs the fp64 definition of hgt.py:350-358 over the
    edges of auto_regressive_edges, incl. ragged L, several blocks, a context window and the split-fp16 output."""
    from gnnlm_b200 import ops
    torch.manual_seed(11)
    q, k, v = torch.randn(B * L, d) * 0.3, torch.randn(B * L, d) * 0.3, torch.randn(B * L, d)
    base = torch.randn(B * L, d)
    qkv = torch.cat([q, k, v], 1).to(dev)                      # q as a column slice (row stride 3d), as the layer passes it
    kv = ops.to_split(qkv[:, d:].contiguous())
    out = base.clone().to(dev)
    if split_out:
        res = ops.Split.empty(B * L, d, dev)
        ops.causal_attn_flash(qkv[:, :d], kv, B, L, ctx, H, out, out_scale=0.5, accumulate=True, out_split=res)
        assert torch.equal(out.cpu(), base)                    # only read
        got = res.float().cpu().double()
    else:
        ops.causal_attn_flash(qkv[:, :d], kv, B, L, ctx, H, out, out_scale=0.5, accumulate=True)
        got = out.cpu().double()
        out0 = torch.full((B * L, d), float("nan"), device=dev)
        ops.causal_attn_flash(qkv[:, :d], kv, B, L, ctx, H, out0, out_scale=1.0, accumulate=False)     # overwrite form
    qh, kh, vh = (t.view(B, L, H, d // H).permute(0, 2, 1, 3).double() for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2)
    i = torch.arange(L)
    mask = (i[None, :] <= i[:, None]) & ((i[:, None] - i[None, :] < ctx) if ctx else True)
    att = (torch.softmax(s.masked_fill(~mask, -float("inf")), -1) @ vh).permute(0, 2, 1, 3).reshape(B * L, d)
    ref = base.double() + 0.5 * att
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=2e-5, atol=2e-5)
    if not split_out:
        np.testing.assert_allclose(out0.cpu().double().numpy(), att.numpy(), rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("B,L,H,d,ctx", [(1, 256, 8, 1024, 0), (1, 64, 8, 1024, 0), (1, 512, 8, 1024, 100), (1, 1288, 8, 1024, 0),
                                          (1, 3072, 8, 1024, 0), (3, 200, 8, 1024, 0), (2, 72, 4, 512, 5), (1, 40, 8, 1024, 1)])
@pytest.mark.parametrize("split_out", [False, True])
def test_causal_flash_tc_kernel(B, L, H, d, ctx, split_out, dev):
    """gnnlm_hgt_causal_flash_tc (tcgen05 / TMEM / TMA flash kernel, d_k = 128) vs the fp64 definition of hgt.py:350-358 over the
    edges of auto_regressive_edges: ragged L, several blocks, a context window, fp32 and split-fp16 outputs."""
    _need_tc()
    from gnnlm_b200 import ops
    torch.manual_seed(13)
    q, k, v = torch.randn(B * L, d) * 0.3, torch.randn(B * L, d) * 0.3, torch.randn(B * L, d)
    base = torch.randn(B * L, d)
    qkv = torch.cat([q, k, v], 1).to(dev)
    out = base.clone().to(dev)
    if split_out:
        res = ops.Split.empty(B * L, d, dev)
        ops.causal_attn_flash_tc(qkv, B, L, ctx, H, out, out_scale=0.5, accumulate=True, out_split=res)
        assert torch.equal(out.cpu(), base)
        got = res.float().cpu().double()
    else:
        ops.causal_attn_flash_tc(qkv, B, L, ctx, H, out, out_scale=0.5, accumulate=True)
        got = out.cpu().double()
    qh, kh, vh = (t.view(B, L, H, d // H).permute(0, 2, 1, 3).double() for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2)
    i = torch.arange(L)
    mask = (i[None, :] <= i[:, None]) & ((i[:, None] - i[None, :] < ctx) if ctx else True)
    att = (torch.softmax(s.masked_fill(~mask, -float("inf")), -1) @ vh).permute(0, 2, 1, 3).reshape(B * L, d)
    np.testing.assert_allclose(got.numpy(), (base.double() + 0.5 * att).numpy(), rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("math", ["fp32", "f16x3"])
def test_token_chunked_ntgt_side_matches(math, dev):
    """forward_tgt_chunked (ntgt side in token chunks, inter attention per chunk) == forward_tgt."""
    if math != "fp32":
        _need_tc()
    import copy
    from gnnlm_b200 import synth
    from gnnlm_b200.graph import build_token_graph
    cfg = dict(synth.CONFIGS["c3mini"], B=2, L=96, NL=3)
    model = copy.deepcopy(synth.make_model(cfg)).to(dev).set_math(math)
    tables = synth.make_tables(cfg, device=dev)
    batch = synth.make_batch(cfg, tables, device=dev)
    dec, hgt = model.decoder, model.decoder.hgt_decoder
    g = build_token_graph(batch["nbr"], tables["n_d"], cfg["c"], cfg["c"])
    h_t = batch["feats"].float()
    mode = dec.math_mode
    decode = lambda gg, centre: dec.tgt_quantizer.gather_decode(
        tables["codes"], gg.ntgt_row, row_ids=gg.inter_indices if centre else None,
        n_cap=None if centre else gg.node_cap, n_dev=gg.n_valid_dev if centre else gg.n_ntgt_dev, math_mode=mode)
    from gnnlm_b200.hgt import as_float
    ref = as_float(hgt.forward_tgt(g, h_t, decode(g, False)))
    for chunk in (64, 80, 192):
        out = as_float(hgt.forward_tgt_chunked(g, h_t, decode, chunk))
        np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-5)


def test_unique_centres_kernel(dev):
    """gnnlm_unique_centres: distinct valid ids (-1 padded), inverse map per compact valid pair, device-side count."""
    from gnnlm_b200.graph import build_token_graph, unique_centre_graph
    torch.manual_seed(3)
    nbr = torch.randint(5, 60, (2, 40, 6), dtype=torch.int64)
    nbr[torch.rand(nbr.shape) < 0.1] = -1
    G = build_token_graph(nbr.to(dev), 1000, 1, 1)
    G_u, inv, n_u = unique_centre_graph(G)
    flat = nbr.reshape(-1)
    valid = flat[flat >= 0]
    want = torch.unique(valid)
    uniq = G_u.nbr.reshape(-1).cpu()
    n = int(n_u.item())
    assert n == want.numel() and (uniq[n:] == -1).all()
    assert torch.equal(torch.sort(uniq[:n]).values, want)
    assert torch.equal(uniq[inv.cpu()[:valid.numel()].long()], valid)
    assert G_u.counts()[1] == n


@pytest.mark.parametrize("math,NL", [("fp32", 3), ("f16x3", 3), ("f16f8", 3), ("f16f8", 2), ("f16x3", 1), ("bf16", 3)])
def test_shared_centres_on_duplicated_ids(math, NL, dev):
    """Neighbour ids with real-graph locality (repeated centres, overlapping clusters): the ntgt side run once per distinct centre
    (share_centres) scores like the oracle's duplicated clusters of new_build_graph, and like the plain CUDA path."""
    if math != "fp32":
        _need_tc()
    import copy
    from gnnlm_b200 import synth
    from tests.synth import run_oracle
    cfg = dict(synth.CONFIGS["c3mini"], L=128, NL=NL)
    model = synth.make_model(cfg)
    data = synth.make_data(cfg, device="cpu", locality=(0.5, 64))
    flat = data["nbr"].reshape(-1)
    n_valid, n_uni = int((flat >= 0).sum()), int(torch.unique(flat[flat >= 0]).numel())
    assert n_uni < 0.8 * n_valid                                   # the workload does repeat centres
    ref = run_oracle((cfg, model, data))
    outs = {}
    for share in (False, True):
        m = copy.deepcopy(model)
        m.decoder.share_centres = share
        outs[share] = synth.run_gpu(cfg, m, data, dev, math)
    tol = 1e-2 if math == "bf16" else 1e-4
    lp = ref["logprob"].numpy()
    for share in (False, True):
        rel = np.abs(outs[share]["logprob"] - lp) / np.abs(lp)
        assert rel.max() < tol, (share, rel.max())
    np.testing.assert_allclose(outs[True]["logprob"], outs[False]["logprob"], rtol=1e-6 if math != "bf16" else 1e-2, atol=1e-6)


@pytest.mark.parametrize("c,NL", [(3, 2), (3, 3), (2, 1), (2, 4)])
def test_unreachable_context_pruning(c, NL, dev):
    """Context nodes further than NL-1 chain hops from their centre cannot reach a tgt node: the pruned graph
    (build_token_graph reach=NL-1, the default of Runner / evaluate) scores exactly like the reference's full
    clusters -- checked against the oracle on the FULL graph and against the unpruned CUDA path."""
    import copy
    from gnnlm_b200 import synth
    from tests.synth import run_oracle
    cfg = dict(synth.CONFIGS["c3mini"], c=c, NL=NL, L=64, k=6, n_d=1 << 14)
    model = synth.make_model(cfg)
    data = synth.make_data(cfg, seed=3, device="cpu")
    data["nbr"][0, :4, 0] = torch.tensor([0, 1, cfg["n_d"] - 1, cfg["n_d"] - 2])       # clusters clipped at both ends
    ref = run_oracle((cfg, model, data))
    d = synth.to_device({k: data[k] for k in ("nbr", "feats", "target", "knn_dists", "knn_ids")}, dev)
    outs = []
    for prune in (True, False):
        r = synth.Runner(cfg, copy.deepcopy(model), data, dev, "fp32", prune_unreachable=prune)
        lp = r.step_resident(d)[0].reshape(-1).cpu().numpy()
        np.testing.assert_allclose(lp, ref["logprob"].numpy(), rtol=1e-4, atol=1e-4)
        outs.append(lp)
        g = r.sample_from(d["nbr"], d["feats"], d["target"], d["knn_dists"], d["knn_ids"])["net_input"]["graph"]
        assert g.left_ctx == (min(c, NL - 1) if prune else c)
    np.testing.assert_allclose(outs[0], outs[1], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("M,dsub,with_pre", [(128, 8, True), (64, 8, False), (32, 4, True), (16, 16, False)])
def test_pq_encode_vs_oracle(M, dsub, with_pre, dev):
    """PQ encode (knn/pq_wrapper.py:131-167).  Codes equal the oracle's except where the two best
    centroids are closer than fp32 summation-order noise (the reference's matmul order is unspecified)."""
    from gnnlm_b200.pq_codec import TorchPQCodec
    from oracle import model_oracle as mo
    rng = np.random.RandomState(M)
    d, n = M * dsub, 3000
    cen = rng.randn(M, 256, dsub).astype(np.float32)
    A = np.linalg.qr(rng.randn(d, d))[0].astype(np.float32) if with_pre else None
    b = rng.randn(d).astype(np.float32) if with_pre else None
    x = rng.randn(n, d).astype(np.float32)
    codec = TorchPQCodec(centroids=cen, A=A, b=b).to(dev)
    codes = codec.encode(torch.from_numpy(x).to(dev)).cpu().numpy()
    ref_codes, dist = mo.pq_encode(x, cen, A, b)
    diff = codes != ref_codes
    assert diff.mean() < 2e-3
    if diff.any():      # every disagreement is a numerical tie
        nn_, mm = np.nonzero(diff)
        gap = np.abs(dist[nn_, mm, codes[nn_, mm]] - dist[nn_, mm, ref_codes[nn_, mm]])
        assert (gap <= 1e-4 * np.maximum(1.0, np.abs(dist[nn_, mm, ref_codes[nn_, mm]]))).all()
    # round trip: decode(encode(x)) is the nearest-centroid reconstruction
    rec = codec.decode(torch.from_numpy(codes).to(dev)).cpu().numpy()
    ref_rec = mo.pq_decode(ref_codes, cen, A, b)
    assert np.abs(rec - ref_rec).mean() < 1e-3


# ------------------------------------------------------------------------------------------ kNN similarity recompute (8f-3)
@pytest.mark.parametrize("case", ["recomp_l2", "recomp_ip", "recomp_cos"])
def test_knn_sims_keys_golden(case, dev, golden_dir):
    """metric_type l2 / ip (knn/knn_model.py:159-177) against vectors produced by the reference's get_knn_prob."""
    from gnnlm_b200.knn_model import KNNModel
    z = np.load(os.path.join(golden_dir, f"knn_{case}.npz"))
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dev)
    metric, cosine = str(z["metric"]), bool(z["cosine"])
    m = KNNModel(t(z["vals"]), vocab_size=int(z["V"]), metric_type=metric, keys=t(z["keys"]),
                 index_file="faiss_store.cosine" if cosine else "faiss_store." + metric)
    sims = m.similarities(t(z["queries"]), None, t(z["ids"])).cpu().numpy()
    live = z["ids"] != -1
    np.testing.assert_allclose(sims[live], z["sims"][live], rtol=1e-5, atol=1e-4)
    m.set_search_results(None, t(z["ids"]))
    p, rec = m.get_knn_prob(t(z["queries"]), t=float(z["temp"]), targets=t(z["targets"]), return_recall=True)
    np.testing.assert_allclose(p.cpu().numpy(), z["p_target"], rtol=2e-4, atol=1e-7)
    assert (rec.cpu().numpy() == z["recall"]).all()


@pytest.mark.parametrize("fp16,norm,opq", [(True, False, True), (False, True, False)])
def test_quantize_features_script(fp16, norm, opq, dev, tmp_path):
    """gnnlm_b200.quantize_features (knn/quantize_features.py:46-72,115-152, --pretrained_quantizer): keys.npy -> quantized-keys.npy
    through the parsed faiss quantizer file, several GPU batches with a ragged tail; codes == the oracle's encode of the same keys
    (up to numerical ties), reconstruction error as the reference computes it."""
    import json as _json
    from gnnlm_b200 import quantize_features as qf
    from gnnlm_b200.formats import write_faiss_quantizer
    from oracle import model_oracle as mo
    rng = np.random.RandomState(3)
    M, dsub, n = 16, 8, 2500
    d = M * dsub
    cen = rng.randn(M, 256, dsub).astype(np.float32) * (0.1 if norm else 1.0)
    A = np.linalg.qr(rng.randn(d, d))[0].astype(np.float32) if opq else None
    keys = rng.randn(n, d).astype(np.float16 if fp16 else np.float32)
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "train_dstore"))
    keys.tofile(os.path.join(root, "train_dstore", "keys.npy"))
    _json.dump({"dstore_size": n, "hidden_size": d, "vocab_size": 10, "dstore_fp16": fp16, "val_size": 1},
               open(os.path.join(root, "train_dstore", "info.json"), "w"))
    write_faiss_quantizer(qf.quantizer_path(root, norm=norm), cen, A)
    with pytest.raises(NotImplementedError):                     # training a quantizer is faiss's job
        qf.main(["--data-dir", root], device=dev, log=lambda *_: None)
    argv = ["--data-dir", root, "--subset", "train", "--code-size", str(M), "--pretrained_quantizer", "--compute-error", "--batch-size", "1024"]
    res = qf.main(argv + (["--norm"] if norm else []), device=dev, log=lambda *_: None)
    codes = np.load(os.path.join(root, "train_dstore", "quantized-keys.npy"))
    assert codes.shape == (n, M) and codes.dtype == np.uint8 and res["n"] == n and res["M"] == M
    x = keys.astype(np.float32)
    if norm:
        x = x / np.sqrt((x ** 2).sum(-1, keepdims=True))
    ref_codes, dist = mo.pq_encode(x, cen, A, None)
    diff = codes != ref_codes
    assert diff.mean() < 2e-3
    if diff.any():
        nn_, mm = np.nonzero(diff)
        gap = np.abs(dist[nn_, mm, codes[nn_, mm]] - dist[nn_, mm, ref_codes[nn_, mm]])
        assert (gap <= 1e-4 * np.maximum(1.0, np.abs(dist[nn_, mm, ref_codes[nn_, mm]]))).all()
    rec = mo.pq_decode(ref_codes, cen, A, None)
    want_err = float(((x - rec) ** 2).sum() / (x ** 2).sum())
    assert abs(res["avg_error"] - want_err) < 2e-2 * want_err     # the reference averages per-batch ratios
    with pytest.raises(ValueError):
        qf.main(["--data-dir", root, "--code-size", "64", "--pretrained_quantizer"] + (["--norm"] if norm else []), device=dev, log=lambda *_: None)


@pytest.mark.parametrize("M,dsub,opq,with_b", [(128, 8, True, True), (64, 8, True, False), (16, 4, False, False), (32, 16, True, True)])
@pytest.mark.parametrize("metric,cosine", [("l2", False), ("ip", False), ("ip", True), ("l2", True)])
def test_knn_sims_pq_vs_oracle(M, dsub, opq, with_b, metric, cosine, dev):
    """Similarities against the PQ-decoded keys (asymmetric distance computation on the codes) == the oracle's
    recompute on pq_decode(codes); with a cosine index the queries are normalised first (knn_model.py:181-184) and `ip`
    normalises the (decoded) keys (:171-172)."""
    from gnnlm_b200.knn_model import KNNModel
    from gnnlm_b200.pq_codec import TorchPQCodec
    from oracle import model_oracle as mo
    rng = np.random.RandomState(M + dsub)
    d, n_d, T, k = M * dsub, 4000, 37, 50
    cen = rng.randn(M, 256, dsub).astype(np.float32)
    A = np.linalg.qr(rng.randn(d, d))[0].astype(np.float32) if opq else None
    b = (rng.randn(d).astype(np.float32) if with_b else np.zeros(0, np.float32)) if opq else None
    codes = rng.randint(0, 256, size=(n_d, M)).astype(np.uint8)
    q = rng.randn(T, d).astype(np.float32)
    ids = rng.randint(0, n_d, size=(T, k)).astype(np.int64)
    ids[rng.rand(T, k) < 0.05] = -1
    keys_hat = mo.pq_decode(codes, cen, A, b if (b is not None and b.size) else None)
    ref = mo.knn_sims(None, metric, mo.knn_queries(torch.from_numpy(q), cosine), keys_hat, torch.from_numpy(ids), cosine).numpy()
    codec = TorchPQCodec(centroids=cen, A=A, b=b).to(dev)
    vals = torch.zeros(n_d, dtype=torch.int32, device=dev)
    m = KNNModel(vals, vocab_size=10, metric_type=metric, pq_codes=torch.from_numpy(codes).to(dev), quantizer=codec,
                 index_file="faiss_store.cosine" if cosine else "faiss_store")
    sims = m.similarities(torch.from_numpy(q).to(dev), None, torch.from_numpy(ids).to(dev)).cpu().numpy()
    scale = np.abs(ref).max()
    np.testing.assert_allclose(sims, ref, rtol=1e-4, atol=1e-5 * scale)


@pytest.mark.parametrize("source", ["keys", "pq"])
def test_whole_path_recomputed_similarities(source, dev):
    """The path with --knn-sim-func style recompute: no dists input at all; kNN queries are the block's precomputed
    features.  Per-token log-probs vs the oracle (which recomputes from the same keys)."""
    import copy
    from types import SimpleNamespace
    from gnnlm_b200 import synth
    from gnnlm_b200.knn_model import KNNModel
    from gnnlm_b200.sequence_scorer import SequenceScorer
    from oracle import model_oracle as mo
    from tests.synth import oracle_model
    cfg = dict(synth.CONFIGS["c1"], n_d=1 << 14, k_nn=32, L=96)
    model = synth.make_model(cfg)
    data = synth.make_data(cfg, seed=11, device="cpu")
    q_ = model.decoder.tgt_quantizer
    if source == "keys":
        keys = (0.05 * torch.randn(cfg["n_d"], cfg["d"])).half()
    else:       # the datastore keys as the PQ codes decode them
        keys = torch.from_numpy(mo.pq_decode(data["codes"].numpy(), q_.centroids_torch.numpy(), q_.A.numpy(), None)).float() * 1.0
    om = oracle_model(cfg, model)
    batch = {"nbr": data["nbr"].numpy(), "offsets": data["positions"].numpy(), "tgt_feats": data["feats"].float(),
             "target": data["target"], "codes": data["codes"].numpy(), "cl": cfg["c"], "cr": cfg["c"], "n_d": data["n_d"]}
    temp = 20.0
    knn = {"dists": torch.zeros(data["knn_ids"].shape), "ids": data["knn_ids"], "vals": data["vals"].long(), "lmbda": cfg["lmbda"],
           "temperature": temp, "metric": "ip", "keys": keys.numpy(), "queries": "inner"}
    ref = mo.eval_batch(om, batch, knn)
    r = synth.Runner(cfg, copy.deepcopy(model), data, dev, "fp32")
    kw = dict(keys=keys.to(dev)) if source == "keys" else dict(pq_codes=r.dstore.codes, quantizer=r.model.decoder.tgt_quantizer)
    r.knn = KNNModel(r.dstore.vals, vocab_size=cfg["V"], metric_type="ip", k=cfg["k_nn"], **kw)
    d = synth.to_device({k: data[k] for k in ("nbr", "feats", "target", "knn_ids")}, dev)
    s = r.sample_from(d["nbr"], d["feats"], d["target"], None, d["knn_ids"])
    lp, p, rec, _ = r.scorer.score_tokens(r.model, s, r.knn, temp, want_knn=True)
    np.testing.assert_allclose(p.cpu().numpy(), ref["knn_prob"].numpy(), rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(lp.reshape(-1).cpu().numpy(), ref["logprob"].numpy(), rtol=1e-4, atol=1e-4)
    assert (rec.cpu().numpy() == ref["knn_recall"].numpy()).all()


# ------------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties_wiki103_shape(dev):
    """BASELINE.json's headline shape (d=1024, H=8, V=267744, L=3072, k=32, c=1, M=128, 3 layers; datastore cut to 2^22
    rows so that the test fits next to others) through size-independent properties, in addition to the direct comparison
    with the fp64 oracle at this size (tests/test_full_size_parity.py): CSR structure, PQ decode -> encode round trip,
    log-prob normalisation, the kNN distribution summing to one, agreement of the CUDA-core fp32 path with the
    tensor-core parity mode, and invariance to the order of a token's neighbours."""
    _need_tc()
    import copy
    from gnnlm_b200 import synth
    cfg = dict(synth.CONFIGS["c3"], n_d=1 << 22)
    model = synth.make_model(cfg)
    tables = synth.make_tables(cfg, device=dev)
    batch = synth.make_batch(cfg, tables, device=dev)
    T, k, w, d, V = cfg["L"], cfg["k"], 3, cfg["d"], cfg["V"]

    # (1) graph structure in closed form (token_block_dataset.py:338-412,545-584)
    g = synth.build_token_graph(batch["nbr"], tables["n_d"], cfg["c"], cfg["c"])
    n_ntgt, n_valid = g.counts()
    nbr = batch["nbr"].view(-1)
    assert n_valid == int((nbr >= 0).sum())
    size = (g.node_base[1:] - g.node_base[:-1]).long()
    exp_size = torch.where(nbr >= 0, 1 + torch.minimum(nbr, torch.tensor(cfg["c"], device=dev)).clamp(min=0)
                           + torch.minimum(tables["n_d"] - 1 - nbr, torch.tensor(cfg["c"], device=dev)).clamp(min=0), 0)
    assert torch.equal(size, exp_size)
    assert n_ntgt == int(exp_size.sum())
    indptr = g.nn_indptr[:n_ntgt + 1].long()
    assert int(indptr[-1]) == 3 * n_ntgt - 2 * n_valid                     # chain with self loops: 3w - 2 edges per cluster
    deg = indptr[1:] - indptr[:-1]
    assert int(deg.min()) >= 1 and int(deg.max()) <= 3
    src = g.nn_indices[:int(indptr[-1])].long()
    dst = torch.repeat_interleave(torch.arange(n_ntgt, device=dev), deg)
    off = g.ntgt_row[:n_ntgt]
    assert int((off[src] - off[dst]).abs().max()) <= 1                     # context = 1: datastore rows at distance <= 1
    inter = g.inter_indptr.long()
    assert torch.equal(inter[1:] - inter[:-1], (batch["nbr"].view(T, k) >= 0).sum(1))
    centre = g.inter_indices[:n_valid].long()
    assert torch.equal(off[centre], nbr[nbr >= 0])                         # tgt2ntgt edges start at the centre node

    # (2) PQ decode -> encode round trip on every ntgt node (pq_wrapper.py:131-203): bit-exact codes
    q = copy.deepcopy(model.decoder.tgt_quantizer).to(dev)
    rows = g.ntgt_row[:n_ntgt]
    x = q.gather_decode(tables["codes"], rows)
    assert torch.equal(q.encode(x if torch.is_tensor(x) else x.float()), tables["codes"][rows])

    # (3) whole path: CUDA-core fp32 vs 3xFP16 tensor-core mode
    d_in = {k_: batch[k_] for k_ in ("nbr", "feats", "target", "knn_dists", "knn_ids")}
    outs = {}
    for math_ in ("fp32", "f16x3"):
        r = synth.Runner(cfg, copy.deepcopy(model), tables, dev, math_)
        lp, p, rec, dec_out = r.step_resident(d_in, want_knn=True)
        outs[math_] = (lp.reshape(-1).float().cpu().numpy(), p.cpu().numpy(), rec.cpu().numpy(), r)
        feat = dec_out[0]
    np.testing.assert_allclose(outs["f16x3"][0], outs["fp32"][0], rtol=1e-4, atol=1e-4)
    assert (outs["f16x3"][2] == outs["fp32"][2]).all()
    nll = {m: -o[0].astype(np.float64).mean() for m, o in outs.items()}
    assert abs(nll["f16x3"] - nll["fp32"]) < 0.01 / 16.8

    # (4) normalisation: the full adaptive-softmax rows exponentiate to one and contain the fused target log-prob
    r = outs["f16x3"][3]
    dec = r.model.decoder
    xs = feat.reshape(-1, d)[:16].float().contiguous()
    tg = batch["target"].reshape(-1)[:16]
    full = dec.adaptive_softmax.get_log_prob(xs.view(1, 16, d), None, dec.math_mode).view(16, V)
    np.testing.assert_allclose(torch.logsumexp(full.double(), 1).cpu().numpy(), 0.0, atol=1e-4)
    fused = dec.adaptive_softmax.target_log_prob(xs.view(1, 16, d), tg.view(1, 16), dec.math_mode).reshape(-1)
    np.testing.assert_allclose(fused.cpu().numpy(), full[torch.arange(16), tg].cpu().numpy(), rtol=1e-4, atol=1e-4)

    # (5) the kNN distribution over the vocabulary sums to one and its target column is p_knn
    r.knn.set_search_results(batch["knn_dists"][:64], batch["knn_ids"][:64])
    pf = r.knn.get_knn_prob(None, t=cfg["temp"], output_size=V)
    np.testing.assert_allclose(pf.sum(1).cpu().numpy(), 1.0, atol=1e-5)
    np.testing.assert_allclose(pf[torch.arange(64), batch["target"].reshape(-1)[:64]].cpu().numpy(), outs["f16x3"][1][:64],
                               rtol=1e-4, atol=1e-7)

    # (6) a token's neighbours are a set: permuting them changes nothing but summation order
    perm = torch.randperm(k, device=dev)
    d_perm = dict(d_in, nbr=batch["nbr"][:, :, perm].contiguous())
    lp2 = r.step_resident(d_perm)[0].reshape(-1).float().cpu().numpy()
    np.testing.assert_allclose(lp2, outs["f16x3"][0], rtol=1e-4, atol=1e-4)


def test_eval_lm_sentence_blocks_eos_mode(dev):
    """--sample-break-mode eos (the one_billion scripts): one sentence per block, lengths 1..24.  Blocks are batched by
    equal length over the whole shard (never padded) and replayed through per-shape CUDA graphs; the summed score equals
    the CPU oracle run sentence by sentence."""
    from types import SimpleNamespace
    import copy
    from gnnlm_b200 import synth
    from gnnlm_b200.dataset import DeviceDatastore, GraphTokenBlockDataset
    from gnnlm_b200.eval_lm import evaluate
    from gnnlm_b200.knn_model import KNNModel
    from gnnlm_b200.sequence_scorer import SequenceScorer
    from oracle import model_oracle as mo
    from tests.synth import oracle_model
    cfg = dict(synth.CONFIGS["c1"], NL=2, k=4, n_d=1 << 16, k_nn=16)
    model = synth.make_model(cfg)
    rng = np.random.RandomState(9)
    sizes = np.concatenate([[1, 2], rng.randint(3, 25, size=38)])
    n_tok = int(sizes.sum())
    tables = synth.make_tables(cfg, device="cpu")
    tokens = rng.randint(4, cfg["V"], size=n_tok).astype(np.int64)
    nbr = rng.randint(1, cfg["n_d"] - 1, size=(n_tok, cfg["k"])).astype(np.int64)
    nbr[rng.rand(n_tok, cfg["k"]) < 0.05] = -1
    feats = rng.randn(n_tok, cfg["d"]).astype(np.float16)
    kd = rng.randn(n_tok, cfg["k_nn"]).astype(np.float32)
    ki = rng.randint(0, cfg["n_d"], size=(n_tok, cfg["k_nn"])).astype(np.int64)
    ds = GraphTokenBlockDataset(tokens, 3072, pad=1, eos=2, neighbor_offsets=nbr, n_datastore=cfg["n_d"], neighbor_context=1,
                                precompute_feats=feats, knn_dists=kd, knn_ids=ki, break_mode="eos", sizes=sizes)
    assert len(ds) == len(sizes)
    dstore = DeviceDatastore(tables["codes"].to(dev), tables["vals"].to(dev))
    scorer = SequenceScorer(synth.Dictionary(cfg["V"]), args=SimpleNamespace(lmbda=cfg["lmbda"], knn_keytype=None))
    knn = KNNModel(dstore.vals, vocab_size=cfg["V"])
    om = oracle_model(cfg, model)
    tot = 0.0
    for i in range(len(ds)):
        cs, e = ds[i]["offsets"]
        batch = {"nbr": nbr[cs:e][None], "offsets": np.arange(cs, e)[None], "tgt_feats": torch.from_numpy(feats[cs:e]).float(),
                 "target": ds[i]["target"], "codes": tables["codes"].numpy(), "cl": 1, "cr": 1, "n_d": cfg["n_d"]}
        k_ = {"dists": torch.from_numpy(kd[cs:e]), "ids": torch.from_numpy(ki[cs:e]), "vals": tables["vals"].long(),
              "lmbda": cfg["lmbda"], "temperature": 1.0}
        tot += float(mo.eval_batch(om, batch, k_)["logprob"].double().sum())
    m = copy.deepcopy(model).to(dev).set_math("fp32")
    for kw in (dict(), dict(bucket_by_length=True, max_tokens=96), dict(bucket_by_length=True, cuda_graph=True)):
        res = evaluate(m, ds, dstore, scorer, knn_dstore=knn, temperature=1.0, max_sentences=8, device=dev, **kw)
        assert res["count"] == n_tok
        assert abs(res["score_sum"] - tot) / abs(tot) < 1e-5, kw


@pytest.mark.parametrize("M,with_b", [(128, True), (64, False), (24, True)])
def test_pq_decode_presplit_codebook(M, with_b, dev):
    """Decoding into the split-fp16 format from the pre-split codebook is bit-identical to splitting after the fp32
    lookup (and so inherits its parity with pq_wrapper.py:169-201), incl. row indirection, a device-side count and M not
    a multiple of the 16-subspace chunk."""
    from gnnlm_b200 import ops
    from gnnlm_b200.pq_codec import TorchPQCodec
    rng = np.random.RandomState(M)
    d, n_d, n = M * 8, 5000, 3001
    cen = (rng.randn(M, 256, 8) * 3).astype(np.float32)
    A = np.linalg.qr(rng.randn(d, d))[0].astype(np.float32)
    b = rng.randn(d).astype(np.float32) if with_b else np.zeros(0, np.float32)
    codec = TorchPQCodec(centroids=cen, A=A, b=b).to(dev)
    codes = torch.from_numpy(rng.randint(0, 256, size=(n_d, M)).astype(np.uint8)).to(dev)
    rows = torch.from_numpy(rng.randint(0, n_d, size=n).astype(np.int64)).to(dev)
    ids = torch.from_numpy(rng.permutation(n)[:1777].astype(np.int32)).to(dev)
    n_dev = torch.tensor([1500], dtype=torch.int32, device=dev)
    hi, lo = codec._split_codebook()
    bias = codec.b if with_b else None
    for kw in (dict(), dict(row_ids=ids), dict(row_ids=ids, n_cap=1777, n_dev=n_dev)):
        ref, _, _ = ops.pq_gather_decode(codes, codec.centroids_torch, rows, bias=bias, out_dtype=ops.SPLIT, **kw)
        out = ops.pq_gather_decode_presplit(codes, hi, lo, rows, **kw)
        live = 1500 if "n_dev" in kw else ref.data.shape[0]
        assert torch.equal(out.data[:live], ref.data[:live])


@pytest.mark.parametrize("math", ["fp32", "f16x3"])
def test_whole_path_heads_not_power_of_two(math, dev):
    """d=768, H=12 (d_k=64), M=96: shapes outside the cluster-attention envelope fall back to the CSR edge kernel;
    everything else (GEMM tiles with ragged N / K, PQ chunks, causal attention) must cope too."""
    if math != "fp32":
        _need_tc()
    import copy
    from gnnlm_b200 import synth
    from tests.synth import run_oracle
    cfg = dict(synth.CONFIGS["c3mini"], d=768, H=12, M=96, L=80, k=6, c=2, NL=3, n_d=1 << 14, V=5000, cutoff=[1000, 3000])
    model = synth.make_model(cfg)
    data = synth.make_data(cfg, seed=21, device="cpu")
    ref = run_oracle((cfg, model, data))
    out = synth.run_gpu(cfg, copy.deepcopy(model), data, dev, math)
    np.testing.assert_allclose(out["logprob"], ref["logprob"].numpy(), rtol=1e-4, atol=1e-4)
    assert abs(out["nll"] - ref["nll"]) < 0.01 / 16.8


@pytest.mark.parametrize("cl,cr,NL,intra,cw,math", [(2, 0, 2, 0, 0, "fp32"), (0, 2, 3, 5, 4, "fp32"), (1, 2, 3, 0, 0, "f16x3"),
                                                    (3, 1, 1, 7, 0, "fp32"), (0, 0, 2, 0, 3, "f16x3"), (2, 2, 4, 9, 0, "f16x3")])
def test_eval_lm_option_matrix(cl, cr, NL, intra, cw, math, dev):
    """Asymmetric neighbour context (--neighbor-context "(l, r)"), --intra-context, --gcn-context-window and 1..4 graph
    layers through dataset -> evaluate(), against the CPU oracle block by block (unpruned reference graph)."""
    if math != "fp32":
        _need_tc()
    from types import SimpleNamespace
    import copy
    from gnnlm_b200 import synth
    from gnnlm_b200.dataset import DeviceDatastore, GraphTokenBlockDataset
    from gnnlm_b200.eval_lm import evaluate
    from gnnlm_b200.knn_model import KNNModel
    from gnnlm_b200.sequence_scorer import SequenceScorer
    from oracle import model_oracle as mo
    from tests.synth import oracle_model
    cfg = dict(synth.CONFIGS["c1"], NL=NL, k=5, n_d=1 << 12, k_nn=8)
    model = synth.make_model(cfg)
    rng = np.random.RandomState(cl * 7 + cr * 3 + NL)
    n_tok, blk = 230, 64
    tables = synth.make_tables(cfg, device="cpu")
    tokens = rng.randint(4, cfg["V"], size=n_tok).astype(np.int64)
    nbr = rng.randint(0, cfg["n_d"], size=(n_tok, cfg["k"])).astype(np.int64)
    nbr[rng.rand(n_tok, cfg["k"]) < 0.1] = -1
    nbr[:3, 0] = [0, 1, cfg["n_d"] - 1]
    feats = rng.randn(n_tok, cfg["d"]).astype(np.float16)
    kd = rng.randn(n_tok, cfg["k_nn"]).astype(np.float32)
    ki = rng.randint(0, cfg["n_d"], size=(n_tok, cfg["k_nn"])).astype(np.int64)
    ds = GraphTokenBlockDataset(tokens, blk, pad=1, eos=2, neighbor_offsets=nbr, n_datastore=cfg["n_d"],
                                neighbor_context=(cl, cr), precompute_feats=feats, context_window=cw, intra_context=intra,
                                knn_dists=kd, knn_ids=ki)
    dstore = DeviceDatastore(tables["codes"].to(dev), tables["vals"].to(dev))
    scorer = SequenceScorer(synth.Dictionary(cfg["V"]), args=SimpleNamespace(lmbda=cfg["lmbda"], knn_keytype=None))
    knn = KNNModel(dstore.vals, vocab_size=cfg["V"])
    om = oracle_model(cfg, model)
    tot, cnt = 0.0, 0
    for i in range(len(ds)):
        it = ds[i]
        cs, e = it["offsets"]
        batch = {"nbr": nbr[cs:e][None], "offsets": np.arange(cs, e)[None], "tgt_feats": torch.from_numpy(feats[cs:e]).float(),
                 "target": it["target"], "codes": tables["codes"].numpy(), "cl": cl, "cr": cr, "n_d": cfg["n_d"],
                 "intra_ctx": intra}
        k_ = {"dists": torch.from_numpy(kd[cs:e]), "ids": torch.from_numpy(ki[cs:e]), "vals": tables["vals"].long(),
              "lmbda": cfg["lmbda"], "temperature": 1.0}
        lp = mo.eval_batch(om, batch, k_)["logprob"][it["start_idx"]:]
        tot += float(lp.double().sum())
        cnt += lp.numel()
    m = copy.deepcopy(model).to(dev).set_math(math)
    for kw in (dict(), dict(cuda_graph=True, prune_unreachable=False)):
        res = evaluate(m, ds, dstore, scorer, knn_dstore=knn, temperature=1.0, max_sentences=2, device=dev, **kw)
        assert res["count"] == cnt == n_tok
        assert abs(res["score_sum"] - tot) / abs(tot) < 2e-5, (kw, res["score_sum"], tot)


@pytest.mark.parametrize("math", ["fp32", "f16x3", "bf16"])
def test_whole_path_gcn_dim_differs(math, dev):
    """--decoder_gcn_dim != embedding width: HGT's input adapters gelu(adapt_ws[t](h)) and output projection
    (hgt.py:482-492,505-513) are live; d = 512 -> hidden 256 -> 512, two layers."""
    if math != "fp32":
        _need_tc()
    import copy
    from gnnlm_b200 import synth
    from tests.synth import run_oracle
    cfg = dict(synth.CONFIGS["c1"], NL=2, L=96, k=6, n_d=1 << 14, gcn_dim=256)
    model = synth.make_model(cfg)
    assert len(model.decoder.hgt_decoder.adapt_ws) == 2 and model.decoder.hgt_decoder.out.weight.shape == (512, 256)
    data = synth.make_data(cfg, seed=5, device="cpu")
    ref = run_oracle((cfg, model, data))
    out = synth.run_gpu(cfg, copy.deepcopy(model), data, dev, math)
    tol = 1e-2 if math == "bf16" else 1e-4
    np.testing.assert_allclose(out["logprob"], ref["logprob"].numpy(), rtol=tol, atol=tol)


# ------------------------------------------------------------------------------------------ --deprecated (dedup) graphs
_DEDUP = sorted(os.path.basename(p)[6:-4] for p in __import__("glob").glob(os.path.join(GOLD, "dedup_*.npz")))


def _check_dedup_graph(g, ref, n_tgt, dev):
    n_ntgt, n_valid = g.counts()
    assert n_ntgt == ref["n_ntgt"] and n_valid == len(ref["inter"][0])
    from oracle import graph_oracle as go
    nn_ip, nn_ix = go.canonical_csr(*ref["nn"], n_ntgt)
    in_ip, in_ix = go.canonical_csr(*ref["inter"], n_tgt)
    assert (g.ntgt_row[:n_ntgt].cpu().numpy() == ref["ntgt_offsets"]).all()             # first-appearance numbering
    assert (g.nn_indptr[:n_ntgt + 1].cpu().numpy() == nn_ip).all()
    assert (g.nn_indices[:len(nn_ix)].cpu().numpy() == nn_ix).all()
    assert (g.inter_indptr.cpu().numpy() == in_ip).all()
    assert (g.inter_indices[:n_valid].cpu().numpy() == in_ix).all()


@pytest.mark.parametrize("case", _DEDUP)
def test_graph_dedup_matches_reference_builder(case, dev):
    """gnnlm_graph_dedup vs deprecated_build_graph executed from the reference source: node numbering, canonical CSRs and
    the gathered code rows are bit-exact."""
    from gnnlm_b200 import ops
    from gnnlm_b200.graph import build_token_graph
    z = np.load(os.path.join(GOLD, f"dedup_{case}.npz"))
    nbr = torch.from_numpy(z["nbr"])[None].contiguous().to(dev)
    pos = torch.from_numpy(z["offsets"])[None].contiguous().to(dev)
    g = build_token_graph(nbr, int(z["n_d"]), int(z["cl"]), int(z["cr"]), tgt_pos=pos, invalid_ctx=int(z["invalid_ctx"]),
                          intra_ctx=int(z["intra_ctx"]), dedup=True)
    n = z["ntgt_codes"].shape[0]
    ref = {"n_ntgt": n, "inter": (z["inter_src"], z["inter_dst"]), "nn": (z["nn_src"], z["nn_dst"]),
           "ntgt_offsets": None}
    # rows of the reference's nodes: recover from the gathered code rows is ambiguous, so take them from the oracle
    from oracle import graph_oracle as go
    o = go.deprecated_build_graph(z["offsets"], z["nbr"], int(z["n_d"]), int(z["cl"]), int(z["cr"]), int(z["invalid_ctx"]))
    ref["ntgt_offsets"] = o["ntgt_offsets"]
    _check_dedup_graph(g, ref, int(z["L"]), dev)
    codes = torch.from_numpy(z["codes"]).to(dev)
    cen = torch.zeros(codes.shape[1], 256, 4, device=dev)
    _, _, codes_out = ops.pq_gather_decode(codes, cen, g.ntgt_row, n_cap=n, want_codes=True, decode=False)
    assert (codes_out.cpu().numpy() == z["ntgt_codes"]).all()


@pytest.mark.parametrize("B,L,k,cl,cr,span", [(2, 64, 8, 1, 1, 300), (3, 40, 16, 2, 0, 100), (1, 512, 32, 1, 1, 20000),
                                              (2, 33, 5, 0, 0, 50), (1, 96, 12, 3, 3, 400)])
def test_graph_dedup_vs_oracle_random(B, L, k, cl, cr, span, dev):
    """Heavier collision patterns (neighbours drawn from `span` rows), several blocks per batch, clipping at both ends."""
    from gnnlm_b200.graph import build_token_graph
    from oracle import graph_oracle as go
    rng = np.random.RandomState(B * 100 + L + k)
    n_d = span + 7
    nbr = rng.randint(0, n_d, size=(B, L, k)).astype(np.int64)
    nbr[rng.rand(B, L, k) < 0.05] = -1
    off = np.arange(B * L, dtype=np.int64).reshape(B, L)
    ref = go.build_batch(nbr, off, n_d, cl, cr, deprecated=True)
    g = build_token_graph(torch.from_numpy(nbr).to(dev), n_d, cl, cr, dedup=True)
    _check_dedup_graph(g, ref, B * L, dev)


@pytest.mark.parametrize("math,NL", [("fp32", 1), ("fp32", 3), ("f16x3", 2)])
def test_eval_lm_deprecated_graph(math, NL, dev):
    """The whole path over --deprecated graphs (shared ntgt nodes, general CSR attention) vs the CPU oracle."""
    if math != "fp32":
        _need_tc()
    from types import SimpleNamespace
    import copy
    from gnnlm_b200 import synth
    from gnnlm_b200.dataset import DeviceDatastore, GraphTokenBlockDataset
    from gnnlm_b200.eval_lm import evaluate
    from gnnlm_b200.knn_model import KNNModel
    from gnnlm_b200.sequence_scorer import SequenceScorer
    from oracle import model_oracle as mo
    from tests.synth import oracle_model
    cfg = dict(synth.CONFIGS["c1"], NL=NL, k=6, n_d=900, k_nn=8)
    model = synth.make_model(cfg)
    rng = np.random.RandomState(NL)
    n_tok, blk = 150, 48
    tables = synth.make_tables(cfg, device="cpu")
    tokens = rng.randint(4, cfg["V"], size=n_tok).astype(np.int64)
    nbr = rng.randint(0, 300, size=(n_tok, cfg["k"])).astype(np.int64)          # many repeated / adjacent rows
    nbr[rng.rand(n_tok, cfg["k"]) < 0.1] = -1
    feats = rng.randn(n_tok, cfg["d"]).astype(np.float16)
    kd = rng.randn(n_tok, cfg["k_nn"]).astype(np.float32)
    ki = rng.randint(0, cfg["n_d"], size=(n_tok, cfg["k_nn"])).astype(np.int64)
    ds = GraphTokenBlockDataset(tokens, blk, pad=1, eos=2, neighbor_offsets=nbr, n_datastore=cfg["n_d"], neighbor_context=1,
                                precompute_feats=feats, knn_dists=kd, knn_ids=ki, deprecated=True)
    dstore = DeviceDatastore(tables["codes"].to(dev), tables["vals"].to(dev))
    scorer = SequenceScorer(synth.Dictionary(cfg["V"]), args=SimpleNamespace(lmbda=cfg["lmbda"], knn_keytype=None))
    knn = KNNModel(dstore.vals, vocab_size=cfg["V"])
    om = oracle_model(cfg, model)
    tot = 0.0
    for i in range(len(ds)):
        cs, e = ds[i]["offsets"]
        batch = {"nbr": nbr[cs:e][None], "offsets": np.arange(cs, e)[None], "tgt_feats": torch.from_numpy(feats[cs:e]).float(),
                 "target": ds[i]["target"], "codes": tables["codes"].numpy(), "cl": 1, "cr": 1, "n_d": cfg["n_d"],
                 "deprecated": True}
        k_ = {"dists": torch.from_numpy(kd[cs:e]), "ids": torch.from_numpy(ki[cs:e]), "vals": tables["vals"].long(),
              "lmbda": cfg["lmbda"], "temperature": 1.0}
        tot += float(mo.eval_batch(om, batch, k_)["logprob"].double().sum())
    m = copy.deepcopy(model).to(dev).set_math(math)
    for kw in (dict(), dict(cuda_graph=True)):
        res = evaluate(m, ds, dstore, scorer, knn_dstore=knn, temperature=1.0, max_sentences=2, device=dev, **kw)
        assert res["count"] == n_tok
        assert abs(res["score_sum"] - tot) / abs(tot) < 2e-5, (kw, res["score_sum"], tot)


# ------------------------------------------------------------------------------------------ error behaviour of the C-ABI
@pytest.mark.parametrize("d,H,fmt", [(1024, 8, "split"), (1024, 8, "f32"), (1024, 8, "bf16"), (512, 8, "split"), (768, 8, "split"), (256, 4, "split"), (512, 8, "f32"),
                                     (768, 12, "split"), (1024, 16, "f32")])
def test_inter_attn_token_side_form(d, H, fmt, dev):
    """('ntgt','inter','tgt') attention with the K' / V' projections moved to the token side (inter_attn.cu: register-q~ kernel
    for H <= 8, shared-memory q~ kernel above) against hgt.py:339-358 evaluated in fp64: ragged degrees around the 16-row tile
    (0, 1, 15, 16, 17, 32, 33, 40, 48, 100, 512), bias b_v' only for tokens with at least one centre."""
    from gnnlm_b200 import ops
    dk = d // H
    g = torch.Generator().manual_seed(d + H)
    degs = torch.tensor([0, 1, 15, 16, 17, 32, 33, 40, 0, 3, 32, 32, 512, 0, 0, 100, 48], dtype=torch.int64)     # k up to 512 (configs[4])
    T, n_c = len(degs), int(degs.sum())
    indptr = torch.zeros(T + 1, dtype=torch.int32)
    indptr[1:] = torch.cumsum(degs, 0)
    q = torch.randn(T, d, generator=g)
    hc = torch.randn(n_c, d, generator=g)
    if fmt == "bf16":
        hc = hc.bfloat16().float()
    Wk = torch.randn(d, d, generator=g) / d ** 0.5 * 0.3
    Wv = torch.randn(d, d, generator=g) / d ** 0.5
    bk, bv = torch.randn(d, generator=g), torch.randn(d, generator=g)
    # fp64 statement of the reference: project every centre, score against q, softmax per (token, head), weighted V' sum
    K = (hc.double() @ Wk.double().T + bk.double()).view(n_c, H, dk)
    V = (hc.double() @ Wv.double().T + bv.double()).view(n_c, H, dk)
    ref = torch.zeros(T, H, dk, dtype=torch.float64)
    for t in range(T):
        e0, e1 = int(indptr[t]), int(indptr[t + 1])
        if e1 > e0:
            s = torch.einsum("hj,chj->ch", q[t].double().view(H, dk), K[e0:e1])
            ref[t] = torch.einsum("ch,chj->hj", torch.softmax(s, 0), V[e0:e1])
    ref = 0.5 * ref.view(T, d)
    assert ops.inter_fused_supported(d, H)
    Wk_d, Wv_d = Wk.to(dev), Wv.to(dev)
    wk_t = ops.split_f16(Wk_d.view(H, dk, d).transpose(1, 2).contiguous().view(H * d, dk))
    wv = ops.split_f16(Wv_d.contiguous())
    hc_d = hc.to(dev)
    hc_in = ops.to_split(hc_d) if fmt == "split" else (hc_d.bfloat16() if fmt == "bf16" else hc_d)
    t_agg = torch.full((T, d), float("nan"), device=dev)
    ops.inter_attn_fused(q.to(dev), [(0, T, indptr.to(dev), hc_in)], H, t_agg, wk_t, wv, bv.to(dev), out_scale=0.5)
    got = t_agg.cpu().double()
    assert torch.isfinite(got).all()
    # fp32 re-association + 3xFP16 GEMMs (22 significant bits): 1e-4 relative to the row scale, the path's fp32 bar
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-4, err
    assert (got[degs == 0] == 0).all()            # no centre: no message and no bias


def test_c_abi_rejects_bad_arguments(dev):
    """Argument / shape problems come back as a status code + message (GnnlmError), never as a crash or a silent
    fallback -- the places where the reference asserts or raises."""
    from gnnlm_b200 import _lib as L, ops
    from gnnlm_b200._lib import GnnlmError
    from gnnlm_b200.graph import build_token_graph
    x = torch.randn(64, 96, device=dev)
    w = torch.randn(32, 96, device=dev)
    with pytest.raises(GnnlmError, match="gnnlm_linear"):
        L.call("gnnlm_linear", None, 0, 96, L.ptr(w), None, 1.0, 96, None, None, 0, 0, L.ptr(x), 0, 32, 64, None, 32, 96, 0,
               L.stream_ptr())                                                         # null A
    with pytest.raises(GnnlmError, match="W_lo"):
        ops.linear(x, w, None, math=L.MATH_F16X3)                                      # 3xFP16 without the split weights
    with pytest.raises(GnnlmError, match="dsub must be 8"):
        L.call("gnnlm_pq_gather_decode_presplit", L.ptr(x), 10, 8, L.ptr(x), L.ptr(x), 4, L.ptr(x), None, 4, None, L.ptr(x), 128,
               L.stream_ptr())
    q = torch.randn(10, 60, device=dev)                                                # d_k = 30: no lane mapping
    with pytest.raises(GnnlmError, match="head layout"):
        ops.edge_attn(q, q, q, torch.zeros(11, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev), 2,
                      torch.empty(10, 60, device=dev))
    nbr = torch.zeros(1, 7, 2, dtype=torch.int64, device=dev)
    with pytest.raises(GnnlmError, match="multiple of the block length"):
        L.call("gnnlm_graph_dedup", L.ptr(nbr), None, 7, 3, 2, 100, 1, 1, 0, L.ptr(nbr), L.ptr(nbr), L.ptr(nbr), L.ptr(nbr),
               L.ptr(nbr), L.ptr(nbr), L.ptr(nbr), L.ptr(nbr), 1 << 20, L.stream_ptr())
    with pytest.raises(GnnlmError, match="metric"):
        L.call("gnnlm_knn_sims_keys", L.ptr(x), 96, L.ptr(x), 0, 64, 96, L.ptr(nbr), 2, 7, 0, L.ptr(x), 1, L.stream_ptr())
    # the library is still usable afterwards
    g = build_token_graph(torch.full((1, 4, 2), -1, dtype=torch.int64, device=dev), 100, 1, 1)
    assert g.counts() == (0, 0)


# ------------------------------------------------------------------------------------------------ training
def _train_problem(NL, cutoff, V, L_=48, d=64, H=4, k=4, c=1):
    from gnnlm_b200 import synth
    cfg = dict(d=d, H=H, V=V, cutoff=cutoff, tied=False, B=2, L=L_, k=k, c=c, M=8, NL=NL, n_d=1 << 12, k_nn=8, lmbda=0.25, temp=1.0)
    model = synth.make_model(cfg)
    data = synth.make_data(cfg, device="cpu")
    return cfg, model, data


@pytest.mark.parametrize("NL,cutoff,V,mode,deprecated", [(2, [40, 120], 300, "fp32", False), (3, None, 97, "fp32", False),
                                                         (1, [40, 120], 300, "fp32", False), (2, [40, 120], 300, "tf32x3", False),
                                                         (2, [40, 120], 300, "fp32", True), (2, [40, 120], 300, "f16x3", False),
                                                         (3, None, 97, "f16x3", False)])
def test_training_step_gradients_vs_oracle_autograd(NL, cutoff, V, mode, deprecated, dev):
    """train.train_step_loss: the adaptive loss (adaptive_loss.py:31-83) and its gradients w.r.t. every decoder.hgt_decoder.*
    parameter (the --freeze set, transformer_lm.py:183-186), every backward stage a kernel of the library, against
    torch.autograd over the oracle's fp64 statement of hgt.py:299-420 -- the way the reference computes them."""
    if mode != "fp32":
        _need_tc()
    import copy
    from gnnlm_b200 import synth, train
    from tests.synth import oracle_train
    cfg, model, data = _train_problem(NL, cutoff, V)
    if deprecated:                                                 # --deprecated graphs: shared ntgt nodes, general CSR
        cfg = dict(cfg, deprecated=True)
        data = dict(data, nbr=torch.randint(1, 200, data["nbr"].shape, dtype=torch.int64))     # overlapping clusters
    ref_loss, ref_g = oracle_train((cfg, model, data), deprecated=deprecated)
    m = copy.deepcopy(model).to(dev).train()
    for name, p in m.named_parameters():                           # --freeze
        p.requires_grad_("hgt" in name)
    r = synth.Runner(cfg, m, data, dev, "fp32")
    d_ = synth.to_device({k_: data[k_] for k_ in r.KEYS}, dev)
    sample = r.sample_from(d_["nbr"], d_["feats"], d_["target"], d_["knn_dists"], d_["knn_ids"])
    loss = train.train_step_loss(m, sample, mode)
    loss.backward()
    assert abs(float(loss.detach()) - ref_loss) < 2e-5 * abs(ref_loss)
    checked = 0
    gmax = max(float(g_.abs().max()) for g_ in ref_g.values() if g_ is not None)
    for name, p in m.decoder.hgt_decoder.named_parameters():
        g_ref = ref_g.get(name)
        if g_ref is None or name.endswith("skip"):                 # unused by the forward (hgt.py:74,399)
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        if p.grad is None:                                         # last layer's ntgt side / tgt-as-source-of-inter: zero in the reference too
            assert float(g_ref.abs().max()) == 0.0, name
            continue
        got = p.grad.detach().cpu().double()
        scale = float(g_ref.abs().max())                           # (K biases cancel inside the softmax: their gradients are ~0)
        err = float((got - g_ref).abs().max())
        assert err < 2e-4 * scale + 2e-6 * gmax, (name, err, scale)
        checked += 1
    assert checked >= 10 * NL


def test_training_backward_kernels_direct(dev):
    """gnnlm_hgt_edge_attn_bwd (CSR and implicit-causal edge sources), gnnlm_layernorm_bwd and gnnlm_xent_fwd_bwd against
    torch.autograd on fp64 statements of the same forward."""
    from gnnlm_b200 import train, ops
    from gnnlm_b200 import _lib as L
    torch.manual_seed(5)
    H, d, n_dst, n_src = 4, 128, 37, 90
    deg = torch.randint(0, 6, (n_dst,))
    indptr = torch.cat([torch.zeros(1, dtype=torch.long), deg.cumsum(0)]).int()
    indices = torch.randint(0, n_src, (int(indptr[-1]),)).int()
    q, k, v = torch.randn(n_dst, d), torch.randn(n_src, d) * 0.5, torch.randn(n_src, d)
    dout = torch.randn(n_dst, d)

    def ref(q_, k_, v_, src, dst):
        qh, kh, vh = q_.view(-1, H, d // H), k_.view(-1, H, d // H), v_.view(-1, H, d // H)
        s = (qh[dst] * kh[src]).sum(-1)
        mx = torch.full((n_dst, H), -float("inf"), dtype=s.dtype).scatter_reduce(0, dst[:, None].expand(-1, H), s, "amax")
        e = torch.exp(s - mx[dst])
        den = torch.zeros((n_dst, H), dtype=s.dtype).index_add_(0, dst, e)
        return torch.zeros_like(qh).index_add_(0, dst, vh[src] * (e / den[dst]).unsqueeze(-1)).reshape(-1, d)
    dst = torch.repeat_interleave(torch.arange(n_dst), deg)
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    out = ref(qd, kd, vd, indices.long(), dst)
    g = torch.autograd.grad(out, [qd, kd, vd], dout.double())
    dq, dk, dv = train._attn_bwd(q.to(dev), k.to(dev), v.to(dev), dout.to(dev), H, 1.0, indptr=indptr.to(dev), indices=indices.to(dev))
    for a, b in zip((dq, dk, dv), g):
        np.testing.assert_allclose(a.cpu().double().numpy(), b.numpy(), rtol=1e-4, atol=1e-5)
    # a symmetric edge set (chains with self loops, some isolated nodes): the atomic-free two-pass form, with and without dropout
    n = 61
    a = torch.arange(n - 1)
    keep = torch.rand(n - 1) < 0.7
    und = torch.stack([a[keep], a[keep] + 1])
    loops = torch.arange(0, n, 2)
    src_s = torch.cat([und[0], und[1], loops])
    dst_s = torch.cat([und[1], und[0], loops])
    order = torch.sort(dst_s, stable=True).indices
    ip_s = torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(dst_s, minlength=n).cumsum(0)]).int()
    ix_s = src_s[order].int()
    qs_, ks_, vs_, ds_ = (torch.randn(n, d) * 0.7 for _ in range(4))
    n_dst = n
    qd, kd, vd = (t.double().requires_grad_(True) for t in (qs_, ks_, vs_))
    g = torch.autograd.grad(ref(qd, kd, vd, ix_s.long(), dst_s[order]), [qd, kd, vd], ds_.double())
    args = (qs_.to(dev), ks_.to(dev), vs_.to(dev), ds_.to(dev), H, 1.0)
    sym = train._attn_bwd(*args, indptr=ip_s.to(dev), indices=ix_s.to(dev), symmetric=True)
    for a_, b in zip(sym, g):
        np.testing.assert_allclose(a_.cpu().double().numpy(), b.numpy(), rtol=1e-4, atol=1e-5)
    sym_p = train._attn_bwd(*args, indptr=ip_s.to(dev), indices=ix_s.to(dev), symmetric=True, p=0.3, seed=77)
    atom_p = train._attn_bwd(*args, indptr=ip_s.to(dev), indices=ix_s.to(dev), p=0.3, seed=77)
    for a_, b in zip(sym_p, atom_p):
        np.testing.assert_allclose(a_.cpu().numpy(), b.cpu().numpy(), rtol=1e-4, atol=1e-5)
    assert float((sym_p[0] - sym[0]).abs().max()) > 1e-3               # the masks do something
    # indices == None (source id = edge id, the inter edges): plain stores, rows past the last edge stay zero
    deg2 = torch.randint(0, 5, (n_dst,))
    ip2 = torch.cat([torch.zeros(1, dtype=torch.long), deg2.cumsum(0)]).int()
    n_e = int(ip2[-1])
    k3, v3 = torch.randn(n_e + 3, d) * 0.5, torch.randn(n_e + 3, d)
    dst2 = torch.repeat_interleave(torch.arange(n_dst), deg2)
    qd, kd, vd = (t.double().requires_grad_(True) for t in (qs_, k3, v3))
    g = torch.autograd.grad(ref(qd, kd, vd, torch.arange(n_e), dst2), [qd, kd, vd], ds_.double())
    ex = train._attn_bwd(qs_.to(dev), k3.to(dev), v3.to(dev), ds_.to(dev), H, 1.0, indptr=ip2.to(dev))
    for a_, b in zip(ex, g):
        np.testing.assert_allclose(a_.cpu().double().numpy(), b.numpy(), rtol=1e-4, atol=1e-5)
    # ... and with attention dropout == the generic (atomic) kernel given the same edges explicitly
    ex_p = train._attn_bwd(qs_.to(dev), k3.to(dev), v3.to(dev), ds_.to(dev), H, 0.5, indptr=ip2.to(dev), p=0.2, seed=3)
    at_p = train._attn_bwd(qs_.to(dev), k3.to(dev), v3.to(dev), ds_.to(dev), H, 0.5, indptr=ip2.to(dev),
                           indices=torch.arange(n_e, dtype=torch.int32, device=dev), p=0.2, seed=3)
    for a_, b in zip(ex_p, at_p):
        np.testing.assert_allclose(a_.cpu().numpy(), b.cpu().numpy(), rtol=1e-4, atol=1e-5)
    # the training forward over contiguous sources (one online-softmax pass per destination and head) == the generic kernel
    base = torch.randn(n_dst, d, device=dev)
    f_ranged, f_generic = base.clone(), base.clone()
    train._attn_fwd_train(qs_.to(dev), k3.to(dev), v3.to(dev), H, 0.5, f_ranged, True, 0.2, 3, indptr=ip2.to(dev))
    train._attn_fwd_train(qs_.to(dev), k3.to(dev), v3.to(dev), H, 0.5, f_generic, True, 0.2, 3, indptr=ip2.to(dev),
                          indices=torch.arange(n_e, dtype=torch.int32, device=dev))
    np.testing.assert_allclose(f_ranged.cpu().numpy(), f_generic.cpu().numpy(), rtol=1e-5, atol=1e-5)
    # implicit causal edges inside blocks of Lb tokens with a context window
    Lb, ctx, B = 19, 7, 2
    q2, k2, v2, do2 = (torch.randn(B * Lb, d) * 0.5 for _ in range(4))
    i = torch.arange(Lb)
    u, w_ = torch.nonzero((i[:, None] <= i[None, :]) & (i[None, :] - i[:, None] < ctx), as_tuple=True)
    src = torch.cat([u + b * Lb for b in range(B)])
    dsts = torch.cat([w_ + b * Lb for b in range(B)])
    n_dst = B * Lb
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q2, k2, v2))
    g = torch.autograd.grad(ref(qd, kd, vd, src, dsts), [qd, kd, vd], do2.double())
    for atomics in (False, True):          # the two-pass form without atomics, and the CSR kernel's causal mode
        dq, dk, dv = train._attn_bwd(q2.to(dev), k2.to(dev), v2.to(dev), do2.to(dev), H, 1.0, causal=(Lb, ctx), atomics=atomics)
        for a, b in zip((dq, dk, dv), g):
            np.testing.assert_allclose(a.cpu().double().numpy(), b.numpy(), rtol=1e-4, atol=1e-5)
    # LayerNorm(o + h)
    rows, dd = 23, 320
    o, h, gam, bet, dy = torch.randn(rows, dd), torch.randn(rows, dd), torch.randn(dd), torch.randn(dd), torch.randn(rows, dd)
    od, hd, gd, bd = (t.double().requires_grad_(True) for t in (o, h, gam, bet))
    g = torch.autograd.grad(torch.nn.functional.layer_norm(od + hd, (dd,), gd, bd, 1e-5), [od, gd, bd], dy.double())
    o_, h_, g_, b_ = (t.to(dev).requires_grad_(True) for t in (o, h, gam, bet))
    y = train._AddLayerNorm.apply(o_, h_, g_, b_, 1e-5)
    y.backward(dy.to(dev))
    for a, b in zip((o_.grad, g_.grad, b_.grad), g):
        np.testing.assert_allclose(a.cpu().double().numpy(), b.numpy(), rtol=1e-4, atol=1e-4)
    assert torch.equal(o_.grad, h_.grad)
    # softmax cross-entropy of one cluster, with an ignored row
    R, C = 9, 1000
    lg = torch.randn(R, C) * 3
    tg = torch.randint(0, C, (R,))
    tg[4] = -1
    ld = lg.double().requires_grad_(True)
    keep = tg >= 0
    ref_loss = torch.nn.functional.cross_entropy(ld[keep], tg[keep], reduction="sum")
    (gl,) = torch.autograd.grad(ref_loss, [ld])
    lg_dev, loss = lg.to(dev).clone(), torch.zeros(1, dtype=torch.float64, device=dev)
    L.call("gnnlm_xent_fwd_bwd", L.ptr(lg_dev), lg_dev.stride(0), L.ptr(tg.to(dev)), R, C, 1.0, L.ptr(loss), L.stream_ptr())
    assert abs(float(loss) - float(ref_loss.detach())) < 1e-5 * abs(float(ref_loss.detach()))
    np.testing.assert_allclose(lg_dev.cpu().double().numpy(), gl.numpy(), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("Lb,ctx,H,d,p", [(256, 0, 4, 256, 0.0), (320, 100, 8, 512, 0.0), (512, 0, 8, 1024, 0.1), (384, 37, 2, 128, 0.25)])
def test_causal_backward_gemm_form(Lb, ctx, H, d, p, dev, monkeypatch):
    """The tensor-core (GEMM) form of the causal-edge backward (train._causal_bwd_gemm: five 3xFP16 batched products around
    gnnlm_causal_softmax_bwd_split) == torch.autograd over the fp64 dense statement (no dropout), and == the streaming form
    (gnnlm_hgt_causal_attn_bwd, same (seed, edge)-addressed masks) with attention dropout; tiny gradients keep their precision."""
    _need_tc()
    from gnnlm_b200 import train
    torch.manual_seed(Lb + H)
    B, scale, seed = 2, 0.5, 991
    q, k, v = (torch.randn(B * Lb, d) * 0.4 for _ in range(3))
    for gscale in (1.0, 1e-6):
        dout = torch.randn(B * Lb, d) * gscale
        assert train.causal_bwd_gemm_supported(d, H, Lb)
        got = train._attn_bwd(q.to(dev), k.to(dev), v.to(dev), dout.to(dev), H, scale, causal=(Lb, ctx), p=p, seed=seed)
        monkeypatch.setenv("GNNLM_TRAIN_GEMM_BWD", "0")
        stream = train._attn_bwd(q.to(dev), k.to(dev), v.to(dev), dout.to(dev), H, scale, causal=(Lb, ctx), p=p, seed=seed)
        monkeypatch.delenv("GNNLM_TRAIN_GEMM_BWD")
        for a, b in zip(got, stream):
            ref = b.cpu().double()
            assert float((a.cpu().double() - ref).abs().max()) < 2e-5 * float(ref.abs().max())
        if gscale == 1.0:                     # the forward in GEMM form (with the dropout multiplier) == the streaming training forward
            from gnnlm_b200 import ops
            o_gemm, o_stream = (torch.zeros(B * Lb, d, device=dev) for _ in range(2))
            ops.causal_attn_gemm(q.to(dev), k.to(dev), v.to(dev), B, Lb, ctx, H, o_gemm, out_scale=scale, drop=(p, seed))
            train._attn_fwd_train(q.to(dev), k.to(dev), v.to(dev), H, scale, o_stream, False, p, seed, causal=(Lb, ctx))
            assert float((o_gemm - o_stream).abs().max()) < 2e-5 * float(o_stream.abs().max())
        if p == 0.0:
            qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
            i = torch.arange(Lb)
            allowed = (i[None, :] <= i[:, None]) & ((i[:, None] - i[None, :] < ctx) if ctx > 0 else torch.ones(Lb, Lb, dtype=torch.bool))
            dk_ = d // H
            outs = []
            for b in range(B):
                qh, kh, vh = (t[b * Lb:(b + 1) * Lb].view(Lb, H, dk_).transpose(0, 1) for t in (qd, kd, vd))
                sc = (qh @ kh.transpose(1, 2)).masked_fill(~allowed, -float("inf"))
                outs.append((torch.softmax(sc, -1) @ vh).transpose(0, 1).reshape(Lb, d))
            g = torch.autograd.grad(scale * torch.cat(outs), [qd, kd, vd], dout.double())
            for a, b in zip(got, g):
                assert float((a.cpu().double() - b).abs().max()) < 2e-5 * float(b.abs().max())


@pytest.mark.parametrize("d,H,c,p", [(128, 4, 1, 0.0), (256, 2, 2, 0.2), (512, 8, 3, 0.0), (64, 2, 0, 0.1), (1024, 8, 1, 0.1)])
def test_cluster_chain_backward(d, H, c, p, dev):
    """gnnlm_hgt_cluster_attn_bwd (one warp per chain and head, no atomics) == the CSR backward over nn_indptr / nn_indices
    (itself pinned to fp64 autograd above), with invalid neighbours, datastore-boundary chains and attention dropout."""
    from gnnlm_b200 import train
    from gnnlm_b200.graph import build_token_graph
    rng = np.random.RandomState(d + c)
    torch.manual_seed(d)
    B, Lb, k, n_d = 2, 24, 5, 400
    nbr = rng.randint(0, n_d, size=(B, Lb, k)).astype(np.int64)
    nbr[rng.rand(B, Lb, k) < 0.1] = -1
    nbr[0, 0, 0], nbr[0, 0, 1], nbr[1, 2] = 0, n_d - 1, -1
    G = build_token_graph(torch.from_numpy(nbr).to(dev), n_d, c, c)
    n = G.counts()[0]
    q, k_, v, do = (torch.randn(n, d, device=dev) * 0.6 for _ in range(4))
    ip, ix = G.nn_indptr[:n + 1].contiguous(), G.nn_indices
    ref = train._attn_bwd(q, k_, v, do, H, 1.0, indptr=ip, indices=ix, p=p, seed=5)
    qa, ka, va = (t.clone().requires_grad_(True) for t in (q, k_, v))
    out = train._EdgeAttention.apply(qa, ka, va, ip, ix, H, p, 5, False, (G.node_base, G.cluster_nl, G.T * G.k))
    fwd_ref = torch.empty_like(q)                                 # the forward per chain (dropout on) == the generic CSR forward
    if p > 0:
        train._attn_fwd_train(q, k_, v, H, 1.0, fwd_ref, False, p, 5, indptr=ip, indices=ix)
    else:
        from gnnlm_b200 import ops
        ops.edge_attn(q, k_, v, ip, ix, H, fwd_ref)
    assert float((out.detach() - fwd_ref).abs().max()) < 2e-5 * float(fwd_ref.abs().max())
    out.backward(do)
    for a, b in zip((qa.grad, ka.grad, va.grad), ref):       # (c = 0: one-node chains, dQ = dK' = 0 up to rounding noise)
        assert float((a - b).abs().max()) < 2e-5 * float(b.abs().max()) + 1e-6


@pytest.mark.parametrize("L_,mode,drop", [(48, "fp32", False), (48, "fp32", True), (256, "f16x3", True), (256, "tf32x3", False)])
def test_training_step_gradients_wide_heads(L_, mode, drop, dev):
    """The end-to-end loss / gradient check at d_k = 32 -- the width from which the ntgt-intra-ntgt and inter edges run their
    per-chain / per-token kernels (forward under dropout and backward) -- and at 256-token blocks, from which the causal edges run
    in GEMM form (forward with the dropout multiplier, backward); masks replayed through the oracle."""
    if mode != "fp32":
        _need_tc()
    import copy
    from gnnlm_b200 import synth, train
    from tests.synth import oracle_train
    cfg, model, data = _train_problem(2, [40, 120], 300, L_=L_, d=128, H=4)
    p_feat, p_att, p_soft, seed = 0.3, 0.1, 0.2, 4711
    ref_loss, ref_g = oracle_train((cfg, model, data), dropout=(seed, p_feat, p_att, p_soft) if drop else None)
    m = copy.deepcopy(model).to(dev).train()
    if drop:
        for layer in m.decoder.hgt_decoder.gcs:
            layer.drop.p, layer.attn_drop.p = p_feat, p_att
        m.decoder.adaptive_softmax.dropout = p_soft
    else:
        m.eval()
    for name, p_ in m.named_parameters():
        p_.requires_grad_("hgt" in name)
    r = synth.Runner(cfg, m, data, dev, "fp32")
    if drop:
        m.train()
    assert train.causal_bwd_gemm_supported(128, 4, L_) == (L_ >= 256)
    d_ = synth.to_device({k_: data[k_] for k_ in r.KEYS}, dev)
    sample = r.sample_from(d_["nbr"], d_["feats"], d_["target"], d_["knn_dists"], d_["knn_ids"])
    loss = train.train_step_loss(m, sample, mode, seed=seed)
    loss.backward()
    assert abs(float(loss.detach()) - ref_loss) < 2e-5 * abs(ref_loss)
    gmax = max(float(g_.abs().max()) for g_ in ref_g.values() if g_ is not None)
    checked = 0
    for name, p_ in m.decoder.hgt_decoder.named_parameters():
        g_ref = ref_g.get(name)
        if g_ref is None or name.endswith("skip") or p_.grad is None:
            continue
        err = float((p_.grad.detach().cpu().double() - g_ref).abs().max())
        assert err < 2e-4 * float(g_ref.abs().max()) + 2e-6 * gmax, (name, err)
        checked += 1
    assert checked >= 20


def test_training_dw_split_k(dev):
    """train._dw_split_k: dW = dY^T X of a tall pair by split-K in 3xFP16 (one batched launch + partial sums) against fp64, with
    1e-5-sized gradients pre-scaled by a power of two (train._pow2_scaled) and a ragged last chunk."""
    _need_tc()
    from gnnlm_b200 import train
    torch.manual_seed(9)
    R, N, K = 40000 + 13, 256, 128
    g, x = torch.randn(R, N) * 1e-5, torch.randn(R, K)
    ref = g.double().T @ x.double()
    gs, s = train._pow2_scaled(g.to(dev))
    assert 512.0 <= float(gs.abs().max()) <= 1024.0 and s == train._pow2_scale_of(g.to(dev))
    dW = train._dw_split_k(g.to(dev), x.to(dev), s)          # scaled on the way into the split operand
    assert dW.shape == (N, K)
    assert float((dW.cpu().double() - ref).abs().max()) < 2e-5 * float(ref.abs().max())
    # through the autograd Function (rows above the split-K threshold)
    W = (torch.randn(N, K) * 0.1).to(dev).requires_grad_(True)
    xd = x.to(dev).requires_grad_(True)
    y = train._Linear.apply(xd, W, None, train.L.MATH_F16X3)
    y.backward(g.to(dev))
    assert float((W.grad.cpu().double() - ref).abs().max()) < 2e-5 * float(ref.abs().max())
    ref_dx = g.double() @ W.detach().cpu().double()
    assert float((xd.grad.cpu().double() - ref_dx).abs().max()) < 2e-5 * float(ref_dx.abs().max())
    # two more projections of the SAME activation: the cached split operands (train._Last) give the same gradients
    W2 = (torch.randn(N // 2, K) * 0.1).to(dev).requires_grad_(True)
    for Wi in (W2, W2):
        Wi.grad = None
        yi = train._Linear.apply(xd, Wi, None, train.L.MATH_F16X3)
        yi.backward(g.to(dev)[:, :N // 2].contiguous())
        ref2 = g[:, :N // 2].double().T @ x.double()
        assert float((Wi.grad.cpu().double() - ref2).abs().max()) < 2e-5 * float(ref2.abs().max())
    assert train._SPLIT_XT.src is xd
    train.release_caches()
    assert train._SPLIT_XT.src is None
    # the adaptive softmax's wide back-product dlogits @ W (split-K over the cluster, ragged last chunk)
    M_, V_, k_in = 300, 40000 + 77, 64
    dlg = torch.softmax(torch.randn(M_, V_) * 3, -1)
    dlg[torch.arange(M_), torch.randint(0, V_, (M_,))] -= 1.0
    buf = torch.zeros(M_, (V_ + 3) // 4 * 4)
    buf[:, :V_] = dlg
    wv = torch.randn(V_, k_in) * 0.05
    got = train._back_split_k(buf.to(dev)[:, :V_], wv.to(dev))
    want = dlg.double() @ wv.double()
    assert got.shape == (M_, k_in) and float((got.cpu().double() - want).abs().max()) < 2e-5 * float(want.abs().max())


def test_training_steps_reduce_the_loss(dev):
    """train.train_step (criterion + backward + --clip-norm + Adam) on a --freeze model: only decoder.hgt_decoder.* moves and
    the loss of a fixed batch goes down."""
    import copy
    from gnnlm_b200 import synth, train
    cfg, model, data = _train_problem(2, [40, 120], 300)
    m = copy.deepcopy(model).to(dev)
    for name, p in m.named_parameters():
        p.requires_grad_("hgt" in name)
    frozen = {n_: p.detach().clone() for n_, p in m.named_parameters() if not p.requires_grad}
    r = synth.Runner(cfg, m, data, dev, "fp32")
    d_ = synth.to_device({k_: data[k_] for k_ in r.KEYS}, dev)
    sample = r.sample_from(d_["nbr"], d_["feats"], d_["target"], d_["knn_dists"], d_["knn_ids"])
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-3, betas=(0.9, 0.98))
    losses = [train.train_step(m, sample, opt, clip_norm=0.1)["loss_per_token_base2"] for _ in range(8)]
    assert losses[-1] < losses[0] - 0.01, losses
    for n_, p in m.named_parameters():
        if n_ in frozen:
            assert torch.equal(p.detach(), frozen[n_])


@pytest.mark.parametrize("NL,cutoff,V,mode", [(2, [40, 120], 300, "fp32"), (3, None, 97, "fp32"), (2, [40, 120], 300, "tf32x3"),
                                              (2, [40, 120], 300, "f16x3")])
def test_training_step_with_dropout_vs_oracle_autograd(NL, cutoff, V, mode, dev):
    """model.train(): hgt.py's `drop` (0.3) / `attn_drop` (0.1) and the adaptive softmax's input / tail dropout (0.2) -- the rates of
    transformer_lm_wiki103 -- with the library's (seed, element)-addressed masks replayed through the oracle: loss and every
    decoder.hgt_decoder.* gradient against torch.autograd over the masked fp64 statement."""
    if mode != "fp32":
        _need_tc()
    import copy
    from gnnlm_b200 import synth, train
    from tests.synth import oracle_train
    cfg, model, data = _train_problem(NL, cutoff, V)
    p_feat, p_att, p_soft, seed = 0.3, 0.1, 0.2, 20240917
    ref_loss, ref_g = oracle_train((cfg, model, data), dropout=(seed, p_feat, p_att, p_soft if cutoff is not None else 0.0))
    ref0, _ = oracle_train((cfg, model, data))
    assert abs(ref_loss - ref0) > 1e-6 * abs(ref0)                 # the masks do something
    m = copy.deepcopy(model).to(dev).train()
    for layer in m.decoder.hgt_decoder.gcs:
        layer.drop.p, layer.attn_drop.p = p_feat, p_att
    if m.decoder.adaptive_softmax is not None:
        m.decoder.adaptive_softmax.dropout = p_soft
    for name, p in m.named_parameters():
        p.requires_grad_("hgt" in name)
    r = synth.Runner(cfg, m, data, dev, "fp32")
    m.train()
    d_ = synth.to_device({k_: data[k_] for k_ in r.KEYS}, dev)
    sample = r.sample_from(d_["nbr"], d_["feats"], d_["target"], d_["knn_dists"], d_["knn_ids"])
    loss = train.train_step_loss(m, sample, mode, seed=seed)
    loss.backward()
    assert abs(float(loss.detach()) - ref_loss) < 2e-5 * abs(ref_loss), (float(loss.detach()), ref_loss)
    gmax = max(float(g_.abs().max()) for g_ in ref_g.values() if g_ is not None)
    checked = 0
    for name, p in m.decoder.hgt_decoder.named_parameters():
        g_ref = ref_g.get(name)
        if g_ref is None or p.grad is None or name.endswith("skip"):
            continue
        err = float((p.grad.detach().cpu().double() - g_ref).abs().max())
        assert err < 2e-4 * float(g_ref.abs().max()) + 2e-6 * gmax, (name, err)
        checked += 1
    assert checked >= 10 * NL
    m.eval()                                                       # eval mode: no masks, the deterministic loss
    assert abs(float(train.train_step_loss(m, sample, mode, seed=seed).detach()) - ref0) < 2e-5 * abs(ref0)


def test_f24_projection_output_and_cluster_attention(dev):
    """GNNLM_F24 (an fp32 value rounded to its top three bytes: 16-bit plane + byte plane): gnnlm_linear_f16f8 writes exactly the
    rounding of its fp32 output, and gnnlm_hgt_cluster_attn_hq on it equals gnnlm_hgt_cluster_attn on the same values as fp32
    (all-nodes and centre-only forms, chains of 3 and 5 nodes)."""
    _need_tc()
    from gnnlm_b200 import ops
    from gnnlm_b200.graph import build_token_graph
    torch.manual_seed(21)
    d, H, k = 1024, 8, 4
    for c in (1, 2):
        nbr = torch.randint(c, 5000 - c, (1, 24, k), dtype=torch.int64)
        nbr[0, 3, 1] = -1
        nbr[0, 5, 0] = 0                                    # clipped at the left boundary
        G = build_token_graph(nbr.to(dev), 5000, c, c)
        n_ntgt, n_valid = G.counts()
        x = ops.to_q8(ops.to_split(torch.randn(G.node_cap, d, device=dev)))
        W = torch.randn(3 * d, d, device=dev) / 32
        b = torch.randn(3 * d, device=dev)
        Wh, Wl, sc = ops.split_f16(W)
        W8 = ops.quant_w8(Wh, Wl)
        ref = ops.linear_f16f8(x, Wh, W8, b, w_scale=sc, m_dev=G.n_ntgt_dev)
        hq = ops.linear_f16f8(x, Wh, W8, b, w_scale=sc, m_dev=G.n_ntgt_dev, out_dtype=ops.HILO8)
        deq = hq.float()
        want = ((ref[:n_ntgt].view(torch.int32) + 0x80) & ~0xFF).view(torch.float32)       # round at bit 8, keep the top 24 bits
        assert torch.equal(deq[:n_ntgt], want)
        assert float(((deq[:n_ntgt] - ref[:n_ntgt]).abs() / ref[:n_ntgt].abs().clamp_min(1e-30)).max()) <= 2 ** -16
        for centre in (False, True):
            rows = n_valid if centre else G.node_cap
            qf = deq[:, :d] if not centre else ops.gather_rows(deq[:, :d].contiguous(), G.inter_indices, n_cap=n_valid)
            qh = hq[:, :d] if not centre else ops.HiLo8(ops.gather_rows(hq.hi[:, :d].contiguous(), G.inter_indices, n_cap=n_valid),
                                                        ops.gather_rows(hq.lo8[:, :d].contiguous().view(torch.float16), G.inter_indices,
                                                                        n_cap=n_valid).view(torch.uint8))
            o_ref = ops.Split.empty(rows, d, dev, q8=True)
            o_hq = ops.Split.empty(rows, d, dev, q8=True)
            ops.cluster_attn(qf, deq[:, d:2 * d], deq[:, 2 * d:], G, H, o_ref, centre_only=centre)
            ops.cluster_attn(qh, hq[:, d:2 * d], hq[:, 2 * d:], G, H, o_hq, centre_only=centre)
            n = n_valid if centre else n_ntgt
            np.testing.assert_allclose(o_hq.float()[:n].cpu().numpy(), o_ref.float()[:n].cpu().numpy(), rtol=2e-6, atol=2e-6)
            assert torch.equal(o_hq.q8[:n].cpu(), o_ref.q8[:n].cpu()) or (o_hq.q8[:n].cpu() != o_ref.q8[:n].cpu()).float().mean() < 1e-3


def test_deferred_layernorm_path(dev):
    """Layer 0's ntgt LayerNorm deferred into its consumers (HGT._ntgt_side: un-rotated pre-norm sum + per-node statistics, the
    normalisation applied inside the centre-only cluster kernel): same scores as the explicit rotation + LayerNorm path and within
    the parity bar of the oracle; the statistics kernel against torch."""
    _need_tc()
    import copy
    from gnnlm_b200 import ops, synth
    from tests.synth import run_oracle
    cfg = dict(synth.CONFIGS["c3mini"])
    model = synth.make_model(cfg)
    data = synth.make_data(cfg, device="cpu")
    ref = run_oracle((cfg, model, data))
    outs = {}
    for defer in (True, False):
        m = copy.deepcopy(model)
        for layer in m.decoder.hgt_decoder.gcs:
            layer.use_deferred_ln = defer
        outs[defer] = synth.run_gpu(cfg, m, data, dev, "f16f8")
        hgt_ = m.decoder.hgt_decoder
        assert ("defer" in hgt_.gcs[0]._prep) == defer            # the path under test was the one that ran
    lp = ref["logprob"].numpy()
    for defer in (True, False):
        assert (np.abs(outs[defer]["logprob"] - lp) / np.abs(lp)).max() < 1e-4
    np.testing.assert_allclose(outs[True]["logprob"], outs[False]["logprob"], rtol=2e-5, atol=2e-5)
    # gnnlm_rowstats_q8 vs torch: z' = o + x, statistics of z' rot^T
    torch.manual_seed(9)
    rows, d = 300, 1024
    rot = torch.linalg.qr(torch.randn(d, d, dtype=torch.float64))[0]
    o = torch.randn(rows, d, device=dev)
    x = ops.to_q8(ops.to_split(torch.randn(rows, d, device=dev)))
    xq = x.data[:, :d].float() + x.q8[:, d:].view(torch.float8_e4m3fn).float() / 1024.0         # hi + lo8 / 2^10: what the kernel reads
    u = (rot.sum(0) / d).float().to(dev)
    z, stats = ops.rowstats_q8(o, x, u, 1e-5)
    zp = (o + xq).double().cpu()
    zr = zp @ rot.T
    np.testing.assert_allclose(stats[:, 0].cpu().double().numpy(), zr.mean(1).numpy(), rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(stats[:, 1].cpu().double().numpy(), (1.0 / torch.sqrt(zr.var(1, unbiased=False) + 1e-5)).numpy(), rtol=1e-5)
    np.testing.assert_allclose(z.data.float().cpu().numpy(), zp.float().numpy(), rtol=2 ** -10, atol=1e-6)
