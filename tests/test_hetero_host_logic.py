"""Host logic of gnnlm_b200.hetero on the CPU (`-m "not gpu"`): the three C-ABI calls the module makes -- gnnlm_linear,
gnnlm_hgt_edge_attn, gnnlm_layernorm -- are replaced by plain-torch stand-ins INSIDE THIS TEST ONLY (monkeypatch), so that what is
exercised is the module's own bookkeeping: folded-weight column layout per node type, CSR construction, the cross-type mean scales,
the doubled source table of the query stream, the in-place step cache and `dst_ids` of incremental decoding, reordering.  The
stand-ins are test infrastructure (the product path has no CPU fallback: tests/test_multi_rank_cpu.py checks that it fails loudly
without the extension); the kernels themselves are checked on the GPU (tests/test_gpu_hetero.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.test_oracle_model import load_hetero_case

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ETYPE2IDX = {"intra": 0, "inter": 1}


@pytest.fixture
def torch_kernels(monkeypatch):
    from gnnlm_b200 import ops

    def linear(A, W, bias=None, *, W_lo=None, residual=None, out=None, out_dtype=None, m_dev=None, math=0, tag=None, w_scale=1.0):
        y = F.linear(A.float(), W.float(), bias)
        if out is not None:
            out.copy_(y)
            return out
        return y

    def edge_attn(q, k, v, indptr, indices, H, out, *, dst_ids=None, n_dst=None, n_dst_dev=None, out_scale=1.0, accumulate=False, tag=None):
        d = q.shape[1]
        dk = d // H
        n = q.shape[0] if n_dst is None else n_dst
        for i in range(n):
            row = int(dst_ids[i]) if dst_ids is not None else i
            a, b = int(indptr[row]), int(indptr[row + 1])
            res = torch.zeros(d)
            if b > a:
                src = indices[a:b].long()
                att = torch.softmax((k[src].view(-1, H, dk) * q[i].view(1, H, dk)).sum(-1), 0)
                res = (att[..., None] * v[src].view(-1, H, dk)).sum(0).reshape(-1)
            out[i] = (out[i] if accumulate else 0) + out_scale * res
        return out

    def layernorm(x, gamma, beta, eps=1e-5, out=None, out_dtype=None, n_dev=None, residual=None):
        return F.layer_norm(x if residual is None else x + residual, (x.shape[1],), gamma, beta, eps)

    monkeypatch.setattr(ops, "linear", linear)
    monkeypatch.setattr(ops, "edge_attn", edge_attn)
    monkeypatch.setattr(ops, "layernorm", layernorm)


def _model(z, sd, ntype2idx, two_stream=False):
    from gnnlm_b200.hgt import HGT
    d = sd["gcs.0.k_linears.0.weight"].shape[0]
    m = HGT(ntype2idx, ETYPE2IDX, d, d, d, int(z["n_layers"]), int(z["H"]), two_stream=two_stream)
    m.load_state_dict(sd, strict=True)
    return m.eval()


def test_general_heterograph_host_logic(torch_kernels):
    from gnnlm_b200.hetero import HeteroGraph
    z, sd, nn_, edges, feats = load_hetero_case(GOLD, "hetero4")
    m = _model(z, sd, {"src": 0, "nsrc": 1, "tgt": 2, "ntgt": 3})
    g = HeteroGraph(edges, nn_, device="cpu")
    for t in nn_:
        g.nodes[t].data["h"] = feats[t]
    out = m(g, features={"tgt": feats["tgt"]})
    for t in nn_:
        np.testing.assert_allclose(out[t].numpy(), z["out." + t], rtol=1e-4, atol=1e-5)


def test_two_stream_host_logic(torch_kernels):
    from gnnlm_b200.hetero import HeteroGraph
    z, sd, nn_, edges, feats = load_hetero_case(GOLD, "two_stream")
    m = _model(z, sd, {"src": 0, "tgt": 1, "ntgt": 2}, two_stream=True)
    out = m(HeteroGraph(edges, nn_, device="cpu"), features=dict(feats))
    assert sorted(out) == ["ntgt", "src", "tgt", "tgt_tilde"]
    for t in out:
        np.testing.assert_allclose(out[t].numpy(), z["out." + t], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("case", ["infer_b2", "infer_b3_reorder"])
def test_incremental_infer_host_logic(case, torch_kernels):
    from gnnlm_b200.hetero import HeteroGraph
    z, sd, nn_, edges, feats = load_hetero_case(GOLD, case)
    m = _model(z, sd, {"tgt": 0, "ntgt": 1})
    g = HeteroGraph(edges, nn_, device="cpu")
    g.nodes["ntgt"].data["h"] = feats["ntgt"]
    steps, bsz, reorder_at = int(z["steps"]), int(z["bsz"]), int(z["reorder_at"])
    h_steps = torch.from_numpy(z["h_steps"])
    inc: dict = {}
    order = torch.arange(bsz)
    for s in range(steps):
        if s == reorder_at:
            order = order.flip(0)
            for layer in m.gcs:
                layer.reorder_incremental_state(inc, order)
        x = h_steps[s][order] if 0 <= reorder_at <= s else h_steps[s]
        out = m(g, features={"tgt": x}, etypes=list(edges), incremental_state=inc)
        np.testing.assert_allclose(out["tgt"].numpy(), z["out_steps"][s], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["ntgt"].numpy(), z["out_ntgt_last"], rtol=1e-4, atol=1e-5)
    assert m.gcs[0].get_incremental_state(inc, "prev_g")["step"].tolist() == [steps - 1] * bsz
