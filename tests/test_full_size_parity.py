"""BASELINE.json's configurations at FULL size against the oracle.

The oracle (oracle/model_oracle.py: the reference's torch statements, nothing from libgnnlm_sm100.so) is evaluated in
**fp64 on the GPU** here, which makes a 3072-token Wiki103-shape block (292 k ntgt nodes, ~9 TFLOP) a matter of
seconds; its tgt-intra-tgt attention runs in the DenseCausal form, asserted equal to the COO form of
auto_regressive_edges in tests/test_oracle_model.py.  The product path runs through the C ABI as everywhere else.

Bars (north_star): per-token log-probs within 1e-4 relative in the fp32-parity modes (1e-2 in bf16); perplexity within
0.01 absolute, taken in its size-independent form |d nll| < 0.01 / 16.8 (d ppl = ppl * d nll at the trained wiki103
perplexity; a random-init model's perplexity is ~V); kNN recall counts bit-exact."""
import copy
import gc

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL = {"fp32": 1e-4, "f16x3": 1e-4, "f16f8": 1e-4, "tf32x3": 1e-4, "bf16": 1e-2}


def _need_tc():
    from gnnlm_b200 import _lib
    if not _lib.load().gnnlm_has_tcgen05():
        pytest.fail("tcgen05 path unavailable on this device: the B200 product path must be present")


def _problem(name, dev, seed=0, **over):
    """Model on the host, datastore tables + one batch generated directly in HBM (the Wiki103 code table is 13 GB)."""
    from gnnlm_b200 import synth
    cfg = dict(synth.CONFIGS[name], **over)
    model = synth.make_model(cfg)
    data = synth.make_data(cfg, seed=seed, device=dev)
    return cfg, model, data


def _free():
    gc.collect()
    torch.cuda.empty_cache()


def _check(prob, dev, modes, budget_gb=None):
    from gnnlm_b200 import synth
    from tests.synth import run_oracle
    cfg, model, data = prob
    ref = run_oracle(prob, dtype=torch.float64, device=dev)
    ref_lp = ref["logprob"].cpu().numpy()
    ref_rec = ref["knn_recall"].cpu().numpy()
    ref_p = ref["knn_prob"].cpu().numpy()
    n_ntgt = ref["n_ntgt"]
    del ref["gcn_feat"]
    _free()
    report = {}
    for mode in modes:
        m = copy.deepcopy(model)
        if budget_gb is not None:
            m.decoder.args.ntgt_memory_budget_gb = budget_gb
        out = synth.run_gpu(cfg, m, data, dev, mode)
        assert out["count"] == ref["count"] == cfg["B"] * cfg["L"]
        rel = np.abs(out["logprob"].astype(np.float64) - ref_lp) / np.abs(ref_lp)
        report[mode] = (rel.max(), abs(out["nll"] - ref["nll"]))
        assert rel.max() < REL[mode], (mode, rel.max())
        assert abs(out["nll"] - ref["nll"]) < 0.01 / 16.8, (mode, out["nll"], ref["nll"])
        assert (out["recall"] == ref_rec).all()
        np.testing.assert_allclose(out["knn_prob"], ref_p, rtol=1e-4, atol=1e-9)
        del out, m
        _free()
    print(f"\n[{cfg['L']} tokens, k={cfg['k']}, c={cfg['c']}, V={cfg['V']}, {n_ntgt} ntgt nodes] "
          + ", ".join(f"{m}: max rel dlogp {r:.2e}, |d nll| {n:.2e}" for m, (r, n) in report.items()))
    return report


def test_c2_enwik8_shape_vs_fp64_oracle(dev):
    """BASELINE.json configs[1]: d=512, H=8, char vocab 204 with a plain softmax (transformer.py:1081-1085), 512-token
    blocks, k=32, c=1, M=64, 3 layers, 2^24-row datastore -- every arithmetic mode."""
    _need_tc()
    _check(_problem("c2", dev), dev, ["fp32", "f16x3", "f16f8", "tf32x3", "bf16"])


def test_c3_wiki103_shape_full_block_vs_fp64_oracle(dev):
    """BASELINE.json configs[2], the headline workload of bench.py: ONE full 3072-token block, d=1024, H=8, V=267,744 with
    the adaptive softmax (cutoffs 20000/60000, tied), k=32, c=1, M=128, 3 layers, k_nn=1024, the 103,227,021-row datastore
    resident in HBM."""
    _need_tc()
    _check(_problem("c3", dev), dev, ["f16x3", "f16f8", "bf16", "fp32"])


def test_c3e_reference_eval_script_setting_vs_fp64_oracle(dev):
    """The Wiki103 shape at the reference evaluation script's setting (hgt_lm_wiki103_reproduce.sh:127-147): 256-token sample,
    --gcn-k 128, --neighbor-context 2 (clusters of 5), lambda 0.1, temperature 0.01."""
    _need_tc()
    _check(_problem("c3e", dev), dev, ["f16x3", "f16f8", "bf16"])


def test_c5_cell_k128_c3_token_chunked_vs_fp64_oracle(dev, monkeypatch):
    """One cell of the edge-aggregation sweep (BASELINE.json configs[4]): k=128, c=3 (clusters of 7 in the oracle, 5 after
    unreachable-context pruning on the device), with the ntgt side forced into token chunks by a 4 GB activation budget
    (model.ntgt_memory_budget_gb) as the k >= 128 cells run at 3072 tokens.  512 tokens here: the fp64 oracle's COO
    intermediates of the 458 k-node graph are ~60 GB."""
    _need_tc()
    from gnnlm_b200.hgt import HGT
    chunks, inner = [], HGT.forward_tgt_chunked

    def spy(self, G, h_tgt, decode, chunk_tokens, **kw):
        chunks.append(chunk_tokens)
        return inner(self, G, h_tgt, decode, chunk_tokens, **kw)

    monkeypatch.setattr(HGT, "forward_tgt_chunked", spy)
    prob = _problem("c3", dev, k=128, c=3, L=512, n_d=1 << 24)
    _check(prob, dev, ["f16x3", "f16f8", "bf16"], budget_gb=4.0)
    assert chunks == [128, 128, 128]         # every mode ran the ntgt side in four 128-token chunks


def test_c4_one_billion_word_vocab_logprob_stage_vs_fp64_oracle(dev):
    """BASELINE.json configs[3]'s distinguishing stage: adaptive softmax over V=793,471 (cutoffs 60000/160000, tied; tail
    clusters of 100,000 x 256 and 633,471 x 64) on 3072 feature rows, fused target log-prob against the oracle's
    adaptive_softmax.py:170-206 restatement in fp64."""
    _need_tc()
    from gnnlm_b200 import _lib as L, synth
    from gnnlm_b200.model import AdaptiveSoftmax
    from oracle import model_oracle as mo
    cfg = synth.CONFIGS["c4"]
    torch.manual_seed(5)
    soft = AdaptiveSoftmax(cfg["V"], cfg["d"], cfg["cutoff"], tied=True).to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    T = 3072
    x = torch.randn((T, cfg["d"]), generator=g, device=dev)
    target = torch.randint(4, cfg["V"], (T,), generator=g, device=dev)
    target[:6] = torch.tensor([4, 59999, 60000, 159999, 160000, cfg["V"] - 1], device=dev)        # every cluster edge
    w = mo.adaptive_weights({k: v.detach().double() for k, v in soft.state_dict().items()})
    ref = mo.adaptive_target_logprob(w, list(cfg["cutoff"]) + [cfg["V"]], x.double(), target).cpu().numpy()
    del w
    _free()
    for mode in ("f16x3", "tf32x3", "fp32", "bf16"):
        lp = soft.target_log_prob(x.view(1, T, -1), target.view(1, T), L.MATH_NAMES[mode]).cpu().numpy().astype(np.float64)
        rel = np.abs(lp - ref) / np.abs(ref)
        print(f"\n[c4 log-prob stage] {mode}: max rel {rel.max():.2e}")
        assert rel.max() < REL[mode], (mode, rel.max())


@pytest.mark.parametrize("math", ["f16x3", "tf32x3", "bf16"])
@pytest.mark.parametrize("N,K,M", [(207744, 64, 1500), (633471, 64, 1100), (20002, 1024, 3072)])
def test_linear_lse_wide_vocab_vs_fp64_log_softmax(math, N, K, M, dev):
    """The log-sum-exp GEMM epilogue at the widths of the real tail clusters -- N = 207,744 (wiki103 tail 2: 812 partial
    (max, sum) tiles per row), N = 633,471 (one-billion-word tail 2), and the wiki103 head (20,002 x 1024) -- against
    torch.log_softmax of the fp64 product.  Logits are given a spread of ~+-12 so that the running-max merges matter."""
    _need_tc()
    from gnnlm_b200 import _lib as L, ops
    g = torch.Generator(device=dev).manual_seed(N + K)
    A = torch.randn((M, K), generator=g, device=dev)
    W = torch.randn((N, K), generator=g, device=dev) * (3.0 / K ** 0.5)
    pick = torch.randint(0, N, (M,), generator=g, device=dev, dtype=torch.int32)
    pick[:3] = torch.tensor([0, N - 1, N // 2], device=dev, dtype=torch.int32)
    mode, lo, ws = L.MATH_NAMES[math], None, 1.0
    if math == "bf16":
        Ad, Wd = A.bfloat16(), W.bfloat16()
        A64, W64 = Ad.double(), Wd.double()
    else:
        Ad, Wd, A64, W64 = A, W, A.double(), W.double()
        if math == "tf32x3":
            Wd, lo = ops.split_tf32(Wd)
        else:
            Wd, lo, ws = ops.split_f16(Wd)
    ref = torch.empty(M, dtype=torch.float64, device=dev)
    for r0 in range(0, M, 256):                                   # [256, N] fp64 slabs
        z = A64[r0:r0 + 256] @ W64.t()
        ref[r0:r0 + 256] = torch.log_softmax(z, 1).gather(1, pick[r0:r0 + 256].long()[:, None]).squeeze(1)
    pm, ps, pk, nt = ops.linear_lse(Ad, Wd, pick, W_lo=lo, w_scale=ws, math=mode)
    lp = torch.empty(M, device=dev)
    ops.lse_finish(pm, ps, pk, nt, lp)
    rel = ((lp.double() - ref).abs() / ref.abs()).max().item()
    print(f"\n[linear_lse N={N} K={K} M={M} {math}] tiles per row {nt}, max rel {rel:.2e}")
    assert rel < 1e-4
    # device-side row count: only the live rows are produced
    cnt = torch.tensor([M // 3], dtype=torch.int32, device=dev)
    pm2, ps2, pk2, _ = ops.linear_lse(Ad, Wd, pick, W_lo=lo, w_scale=ws, math=mode, m_dev=cnt)
    lp2 = torch.full((M,), 7.0, device=dev)
    ops.lse_finish(pm2, ps2, pk2, nt, lp2, m_dev=cnt)
    assert torch.equal(lp2[:M // 3], lp[:M // 3]) and (lp2[M // 3:] == 7.0).all()
