"""HGT on a general heterograph, `two_stream` and incremental `infer()` (fairseq/models/hgt.py:81-297,
299-420): the CUDA path against fixtures produced by executing the reference module under the DGL stub
(tests/golden/make_golden.py --hetero-only) and against oracle/hetero_oracle.py at larger, randomly drawn sizes.
Tolerance: 1e-4 (north_star's fp32 bar) on O(1) LayerNorm outputs, written at each assert."""
import os

import numpy as np
import pytest
import torch

from tests.test_oracle_model import load_hetero_case

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ETYPE2IDX = {"intra": 0, "inter": 1}


def _model(z, sd, ntype2idx, dev, two_stream=False, math="fp32"):
    from gnnlm_b200.hgt import HGT
    d = sd["gcs.0.k_linears.0.weight"].shape[0]
    m = HGT(ntype2idx, ETYPE2IDX, d, d, d, int(z["n_layers"]), int(z["H"]), two_stream=two_stream)
    m.load_state_dict(sd, strict=True)
    return m.to(dev).eval().set_math(math)


# ------------------------------------------------------------------------------------------ container (no GPU needed)
def test_hetero_graph_csr_is_stable_by_destination():
    from gnnlm_b200.hetero import HeteroGraph
    from oracle import graph_oracle as go
    rng = np.random.RandomState(0)
    src, dst = rng.randint(0, 9, 60), rng.randint(0, 7, 60)
    g = HeteroGraph({("a", "r", "b"): (src, dst), ("b", "r", "b"): (dst, dst)}, {"a": 9, "b": 8}, device="cpu")
    assert g.ntypes == ["a", "b"] and g.num_nodes("b") == 8 and g.num_edges(("a", "r", "b")) == 60
    ip, ix = g.csr(("a", "r", "b"))
    rp, rx = go.canonical_csr(src, dst, 8)
    assert ip.dtype == torch.int32 and (ip.numpy() == rp).all() and (ix.numpy() == rx).all()
    _, ix2 = g.csr(("b", "r", "b"), self_loops_shifted=True)
    assert (ix2.numpy() == np.sort(dst, kind="stable") + 8).all()          # every edge is a self loop: all shifted by n_src
    with pytest.raises(ValueError):
        HeteroGraph({("a", "r", "b"): ([0, 9], [0, 1])}, {"a": 9, "b": 8}, device="cpu")
    inferred = HeteroGraph({("a", "r", "b"): ([0, 4], [2, 1])}, device="cpu")
    assert inferred.num_nodes("a") == 5 and inferred.num_nodes("b") == 3
    with g.local_scope():
        g.nodes["a"].data["tmp"] = torch.zeros(1)
    assert "tmp" not in g.nodes["a"].data


def test_two_stream_on_the_token_graph_raises_like_the_reference():
    from gnnlm_b200.hgt import HGTLayer
    layer = HGTLayer(16, 16, {"tgt": 0, "ntgt": 1}, ETYPE2IDX, 2, two_stream=True)
    with pytest.raises(KeyError):
        layer(object(), {"tgt": torch.zeros(1, 16), "ntgt": torch.zeros(1, 16)})


def test_incremental_state_keys_are_per_module():
    from gnnlm_b200.hgt import HGT
    m = HGT({"tgt": 0, "ntgt": 1}, ETYPE2IDX, 16, 16, 16, 2, 2)
    inc = {}
    m.gcs[0].set_incremental_state(inc, "prev_g", {"x": 1})
    assert m.gcs[1].get_incremental_state(inc, "prev_g") is None and m.gcs[0].get_incremental_state(inc, "prev_g") == {"x": 1}
    assert m.gcs[0].get_incremental_state(None, "prev_g") is None


# ------------------------------------------------------------------------------------------ GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("math,tol", [("fp32", 1e-4)])
def test_hetero_golden(math, tol, dev):
    """The reference module's self-test topology (4 node types, 10 canonical edge types, zero-in-degree destinations)."""
    from gnnlm_b200.hetero import HeteroGraph
    z, sd, nn_, edges, feats = load_hetero_case(GOLD, "hetero4")
    m = _model(z, sd, {"src": 0, "nsrc": 1, "tgt": 2, "ntgt": 3}, dev, math=math)
    g = HeteroGraph(edges, nn_, device=dev)
    for t in nn_:
        g.nodes[t].data["h"] = feats[t].to(dev)
    out = m(g, features={"tgt": feats["tgt"].to(dev)})               # the other types from G.nodes[.].data['h'] (hgt.py:501-503)
    assert sorted(out) == sorted(nn_)
    for t in nn_:
        np.testing.assert_allclose(out[t].cpu().numpy(), z["out." + t], rtol=tol, atol=tol)


@pytest.mark.gpu
def test_two_stream_golden(dev):
    from gnnlm_b200.hetero import HeteroGraph
    z, sd, nn_, edges, feats = load_hetero_case(GOLD, "two_stream")
    m = _model(z, sd, {"src": 0, "tgt": 1, "ntgt": 2}, dev, two_stream=True)
    g = HeteroGraph(edges, nn_, device=dev)
    out = m(g, features={t: x.to(dev) for t, x in feats.items()})
    assert sorted(out) == ["ntgt", "src", "tgt", "tgt_tilde"]
    for t in out:
        np.testing.assert_allclose(out[t].cpu().numpy(), z["out." + t], rtol=1e-4, atol=1e-4)
    # a graph without ('src','intra','tgt') cannot run the query stream (hgt.py:392-393 raises in DGL)
    g2 = HeteroGraph({c: e for c, e in edges.items() if c != ("src", "intra", "tgt")}, nn_, device=dev)
    with pytest.raises(KeyError):
        m(g2, features={t: x.to(dev) for t, x in feats.items()})


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["infer_b2", "infer_b3_reorder"])
def test_incremental_infer_golden(case, dev):
    """HGTLayer.infer driven step by step through HGT.forward(incremental_state=...), incl. reorder_incremental_state."""
    from gnnlm_b200.hetero import HeteroGraph
    z, sd, nn_, edges, feats = load_hetero_case(GOLD, case)
    m = _model(z, sd, {"tgt": 0, "ntgt": 1}, dev)
    g = HeteroGraph(edges, nn_, device=dev)
    g.nodes["ntgt"].data["h"] = feats["ntgt"].to(dev)
    steps, bsz, reorder_at = int(z["steps"]), int(z["bsz"]), int(z["reorder_at"])
    h_steps = torch.from_numpy(z["h_steps"]).to(dev)
    inc: dict = {}
    order = torch.arange(bsz, device=dev)
    for s in range(steps):
        if s == reorder_at:
            order = order.flip(0)
            for layer in m.gcs:
                layer.reorder_incremental_state(inc, order)
        x = h_steps[s][order] if 0 <= reorder_at <= s else h_steps[s]
        out = m(g, features={"tgt": x}, etypes=list(edges), incremental_state=inc)
        assert out["tgt"].shape == (bsz, x.shape[1])
        np.testing.assert_allclose(out["tgt"].cpu().numpy(), z["out_steps"][s], rtol=1e-4, atol=1e-4)
        if s == 0:
            np.testing.assert_allclose(out["ntgt"].cpu().numpy(), z["out_ntgt_first"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(out["ntgt"].cpu().numpy(), z["out_ntgt_last"], rtol=1e-4, atol=1e-4)
    buf = m.gcs[0].get_incremental_state(inc, "prev_g")
    assert buf["step"].tolist() == [steps - 1] * bsz                 # hgt.py:211
    with pytest.raises(TypeError):
        m(object(), features={"tgt": x}, incremental_state={})


@pytest.mark.gpu
@pytest.mark.parametrize("d,H,NL,math,tol", [(256, 4, 2, "fp32", 1e-4), (512, 8, 2, "f16x3", 1e-4), (128, 2, 3, "tf32x3", 1e-4),
                                             (256, 8, 2, "f16f8", 1e-4), (256, 4, 2, "bf16", 6e-2)])
def test_hetero_vs_oracle_random(d, H, NL, math, tol, dev):
    """Larger random heterographs (duplicate edges, isolated destinations) against the fp64 oracle; two_stream on."""
    from gnnlm_b200.hetero import HeteroGraph
    from gnnlm_b200.hgt import HGT
    from oracle import hetero_oracle as ho
    rng = np.random.RandomState(d + NL)
    torch.manual_seed(d)
    nn_ = {"src": 300, "tgt": 257, "ntgt": 1000}
    ntype2idx = {"src": 0, "tgt": 1, "ntgt": 2}
    u, v = np.triu_indices(nn_["tgt"])
    keep = rng.rand(u.shape[0]) < 0.3
    keep |= u == v
    rnd = lambda s, t, n: (torch.from_numpy(rng.randint(0, nn_[s], n)), torch.from_numpy(rng.randint(3, nn_[t], n)))
    edges = {("src", "intra", "src"): rnd("src", "src", 3000), ("src", "intra", "tgt"): rnd("src", "tgt", 4000),
             ("tgt", "intra", "tgt"): (torch.from_numpy(u[keep]), torch.from_numpy(v[keep])),
             ("ntgt", "inter", "tgt"): rnd("ntgt", "tgt", 2000), ("ntgt", "intra", "ntgt"): rnd("ntgt", "ntgt", 9000)}
    m = HGT(ntype2idx, ETYPE2IDX, d, d, d, NL, H, two_stream=True)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1 or p.shape[-1] == H:
                p.add_(0.1 * torch.randn_like(p))
    feats = {t: torch.randn(n, d) for t, n in nn_.items()}
    ref = ho.hgt_forward_hetero({k: x.double() for k, x in m.state_dict().items()}, {t: x.double() for t, x in feats.items()},
                                edges, nn_, ntype2idx, ETYPE2IDX, H, NL, two_stream=True)
    m = m.to(dev).eval().set_math(math)
    out = m(HeteroGraph(edges, nn_, device=dev), features={t: x.to(dev) for t, x in feats.items()})
    for t in ref:
        np.testing.assert_allclose(out[t].cpu().numpy(), ref[t].float().numpy(), rtol=tol, atol=tol)


@pytest.mark.gpu
def test_incremental_equals_full_layer_on_the_prefix(dev):
    """Teacher-forced incremental decoding of one layer == the full layer on the same prefix at every decoded position (the
    causal tgt-intra-tgt edges and the zero rows of undecoded positions make the two coincide)."""
    from gnnlm_b200.hetero import HeteroGraph, MAX_LEN
    from gnnlm_b200.hgt import HGT
    rng = np.random.RandomState(3)
    torch.manual_seed(3)
    bsz, steps, n_per, d, H = 4, 12, 40, 256, 4
    nn_ = {"tgt": bsz * MAX_LEN, "ntgt": bsz * n_per}
    u, v = np.triu_indices(MAX_LEN)
    blk = lambda a, per: torch.from_numpy(np.concatenate([a + b * per for b in range(bsz)]))
    i_src = np.arange(nn_["ntgt"])
    edges = {("tgt", "intra", "tgt"): (blk(u, MAX_LEN), blk(v, MAX_LEN)),
             ("ntgt", "inter", "tgt"): (torch.from_numpy(i_src), torch.from_numpy((i_src // n_per) * MAX_LEN + rng.randint(0, steps, nn_["ntgt"]))),
             ("ntgt", "intra", "ntgt"): (blk(rng.randint(0, n_per, 200), n_per), blk(rng.randint(0, n_per, 200), n_per))}
    m = HGT({"tgt": 0, "ntgt": 1}, ETYPE2IDX, d, d, d, 1, H).to(dev).eval()
    g = HeteroGraph(edges, nn_, device=dev)
    h_n = torch.randn(nn_["ntgt"], d, device=dev)
    h_steps = torch.randn(steps, bsz, d, device=dev)
    full_in = torch.zeros(nn_["tgt"], d, device=dev)
    pos = torch.arange(bsz, device=dev) * MAX_LEN
    for s in range(steps):
        full_in[pos + s] = h_steps[s]
    full = m(g, features={"tgt": full_in, "ntgt": h_n})
    inc: dict = {}
    for s in range(steps):
        out = m(g, features={"tgt": h_steps[s], "ntgt": h_n}, incremental_state=inc)
        np.testing.assert_allclose(out["tgt"].cpu().numpy(), full["tgt"][pos + s].cpu().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(out["ntgt"].cpu().numpy(), full["ntgt"].cpu().numpy(), rtol=1e-5, atol=1e-5)
