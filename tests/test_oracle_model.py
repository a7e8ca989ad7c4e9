"""Pin the model oracle to fixtures produced by executing the reference's own code
(tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import graph_oracle as go
from oracle import model_oracle as mo


def _sd(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


@pytest.mark.parametrize("case", ["m8", "m16b"])
def test_pq_decode(case, golden_dir):
    z = np.load(os.path.join(golden_dir, f"pq_{case}.npz"))
    x = mo.pq_decode(z["codes"], z["cen"], z["A"], z["b"])
    np.testing.assert_allclose(x, z["x_numpy"], rtol=0, atol=1e-6)   # tolerance of pq_wrapper.py:235-239
    np.testing.assert_allclose(x, z["x_torch"], rtol=0, atol=1e-5)
    raw = mo.pq_decode(z["codes"], z["cen"])
    assert (raw == z["x_nopre"]).all()                                # pure gather: bit-exact
    idx = mo.pq_decode_indices(z["codes"])
    assert (z["cen"].reshape(-1, z["cen"].shape[2])[idx].reshape(raw.shape) == raw).all()


@pytest.mark.parametrize("case", ["untied", "tied", "tied_noproj"])
def test_adaptive_softmax(case, golden_dir):
    z = np.load(os.path.join(golden_dir, f"adaptive_{case}.npz"))
    w = mo.adaptive_weights(_sd(z, "sd."))
    cutoff = z["cutoff"].tolist()
    x, t = torch.from_numpy(z["x"]), torch.from_numpy(z["target"])
    lp = mo.adaptive_target_logprob(w, cutoff, x, t)
    np.testing.assert_allclose(lp.numpy(), z["lp_target_mode_at_target"].reshape(-1), rtol=1e-5, atol=1e-5)
    full = mo.adaptive_full_logprob(w, cutoff, x)
    np.testing.assert_allclose(full.numpy(), z["lp_full"][0], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(torch.logsumexp(full, 1).numpy(), 0, atol=1e-4)


@pytest.mark.parametrize("case", ["ip_t1", "l2_t001"])
def test_knn_prob(case, golden_dir):
    z = np.load(os.path.join(golden_dir, f"knn_{case}.npz"))
    p, rec = mo.knn_target_prob(torch.from_numpy(z["dists"]), torch.from_numpy(z["ids"]),
                                torch.from_numpy(z["vals"]), torch.from_numpy(z["targets"]),
                                float(z["temp"]), str(z["metric"]))
    np.testing.assert_allclose(p.numpy(), z["p_target"], rtol=1e-5, atol=1e-7)
    assert (rec.numpy() == z["recall"]).all()
    full = mo.knn_full_prob(torch.from_numpy(z["dists"]), torch.from_numpy(z["ids"]), torch.from_numpy(z["vals"]),
                            int(z["V"]), float(z["temp"]), str(z["metric"]))
    np.testing.assert_allclose(full.numpy(), z["p_full"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("case", ["recomp_l2", "recomp_ip", "recomp_cos"])
def test_knn_sims_recompute(case, golden_dir):
    """metric_type l2 / ip (knn_model.py:159-177): similarities from keys[knns], fp16 and fp32 keys, cosine index."""
    z = np.load(os.path.join(golden_dir, f"knn_{case}.npz"))
    q, ids, cosine = torch.from_numpy(z["queries"]), torch.from_numpy(z["ids"]), bool(z["cosine"])
    sims = mo.knn_sims(None, str(z["metric"]), mo.knn_queries(q, cosine), z["keys"], ids, cosine)
    live = z["ids"] != -1
    np.testing.assert_allclose(sims.numpy()[live], z["sims"][live], rtol=1e-5, atol=1e-5)
    assert (z["sims"][~live] == np.float32(-1e10)).all()
    p, rec = mo.knn_target_prob(torch.zeros_like(sims), ids, torch.from_numpy(z["vals"]), torch.from_numpy(z["targets"]),
                                float(z["temp"]), str(z["metric"]), queries=q, keys=z["keys"], cosine=cosine)
    np.testing.assert_allclose(p.numpy(), z["p_target"], rtol=1e-4, atol=1e-7)
    assert (rec.numpy() == z["recall"]).all()


def test_scorer_knn_mix(golden_dir):
    z = np.load(os.path.join(golden_dir, "scorer_b1_knn.npz"))
    w = mo.adaptive_weights(_sd(z, "sd."))
    x, t = torch.from_numpy(z["feats"]), torch.from_numpy(z["target"])
    lm = mo.adaptive_target_logprob(w, z["cutoff"].tolist(), x, t)
    s0 = int(z["start_indices"][0, 0])
    np.testing.assert_allclose(lm.numpy()[s0:], z["lm_pos_0"], rtol=1e-5, atol=1e-5)
    p, rec = mo.knn_target_prob(torch.from_numpy(z["dists"]), torch.from_numpy(z["ids"]),
                                torch.from_numpy(z["vals"]), t.reshape(-1), float(z["temp"]))
    mix = mo.combine_knn_and_vocab_probs(p, lm, float(z["lmbda"]))
    np.testing.assert_allclose(mix.numpy()[s0:], z["knn_pos_0"], rtol=1e-5, atol=1e-5)
    assert (rec.numpy()[s0:] == z["knn_recall_0"]).all()


def test_scorer_lm_b2_start_indices(golden_dir):
    z = np.load(os.path.join(golden_dir, "scorer_b2_lm.npz"))
    w = mo.adaptive_weights(_sd(z, "sd."))
    x, t = torch.from_numpy(z["feats"]), torch.from_numpy(z["target"])
    lm = mo.adaptive_target_logprob(w, z["cutoff"].tolist(), x, t).view(t.shape)
    for i in range(t.shape[0]):
        s = int(z["start_indices"][i, 0])
        np.testing.assert_allclose(lm[i, s:].numpy(), z[f"lm_pos_{i}"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("case", ["l2_c1", "l3_c2", "l2_adapt"])
def test_hgt_matches_reference_layer_code(case, golden_dir):
    z = np.load(os.path.join(golden_dir, f"hgt_{case}.npz"))
    g = go.build_batch(z["nbr"], z["offsets"], int(z["n_d"]), int(z["cl"]), int(z["cr"]))
    sd = _sd(z, "sd.")
    out = mo.hgt_forward(sd, torch.from_numpy(z["h_tgt"]), torch.from_numpy(z["h_ntgt"]), g,
                         int(z["H"]), int(z["n_layers"]))
    np.testing.assert_allclose(out["tgt"].numpy(), z["out_tgt"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(out["ntgt"].numpy(), z["out_ntgt"], rtol=1e-5, atol=2e-6)
    B, L = z["nbr"].shape[:2]
    out2 = mo.hgt_forward_csr(sd, torch.from_numpy(z["h_tgt"]), torch.from_numpy(z["h_ntgt"]), g, (B, L),
                              int(z["H"]), int(z["n_layers"]))
    np.testing.assert_allclose(out2["tgt"].numpy(), z["out_tgt"], rtol=1e-5, atol=2e-6)


def test_zero_in_degree_tgt_gets_half_intra(golden_dir):
    """DGL semantics restated: a tgt with no valid neighbour still averages over
    two edge types, i.e. receives intra/2."""
    z = np.load(os.path.join(golden_dir, "hgt_l2_c1.npz"))
    assert (z["nbr"][0, 1] == -1).all()


def test_perplexity():
    nll2, ppl = mo.perplexity(-10.0 * np.log(2), 10)
    assert abs(nll2 - 1.0) < 1e-12 and abs(ppl - 2.0) < 1e-12


def test_adaptive_input_oracle_matches_reference(golden_dir):
    """oracle adaptive_input_forward == the reference's AdaptiveInput.forward (fixture generated by executing
    fairseq/modules/adaptive_input.py): every band, both edges of every cutoff."""
    from oracle import model_oracle as mo
    z = np.load(os.path.join(golden_dir, "adaptive_input_v300.npz"))
    cutoff = [int(c) for c in z["cutoff"]]
    bands = [(torch.from_numpy(z[f"sd.embeddings.{i}.0.weight"]), torch.from_numpy(z[f"sd.embeddings.{i}.1.weight"]))
             for i in range(len(cutoff))]
    out = mo.adaptive_input_forward(bands, cutoff, torch.from_numpy(z["tokens"]))
    np.testing.assert_allclose(out.numpy(), z["out"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name,intra_ctx", [("c1", 0), ("c3mini", 0), ("c3mini", 37)])
def test_dense_causal_form_equals_coo_form(name, intra_ctx):
    """The tgt-intra-tgt attention as a masked [L, L] softmax per block and head (DenseCausal -- what makes the oracle
    runnable on 3072-token blocks) against the COO edge list of auto_regressive_edges (token_block_dataset.py:586-594) it
    restates: same per-token log-probs to fp64 rounding, whole path, B = 2 included, with and without --intra-context."""
    from tests.synth import make_problem, oracle_model
    cfg, model, data = make_problem(name)
    batch = {"nbr": data["nbr"].numpy(), "offsets": data["positions"].numpy(), "tgt_feats": data["feats"].float(),
             "target": data["target"], "codes": data["codes"].numpy(), "cl": cfg["c"], "cr": cfg["c"], "n_d": data["n_d"],
             "intra_ctx": intra_ctx}
    om = oracle_model(cfg, model)
    coo = mo.eval_batch(om, batch, None, dtype=torch.float64, dense_tt=False)
    dense = mo.eval_batch(om, batch, None, dtype=torch.float64, dense_tt=True)
    assert (coo["logprob"] - dense["logprob"]).abs().max().item() < 1e-11
    assert (coo["gcn_feat"] - dense["gcn_feat"]).abs().max().item() < 1e-11
    # the tensor form of the code gather + decode used when the table lives on a device
    batch_t = dict(batch, codes=data["codes"])
    via_torch = mo.eval_batch(om, batch_t, None, dtype=torch.float64, dense_tt=True)
    assert (via_torch["logprob"] - dense["logprob"]).abs().max().item() < 1e-11
    if intra_ctx:
        assert (coo["logprob"] - mo.eval_batch(om, dict(batch, intra_ctx=0), None, dtype=torch.float64)["logprob"]).abs().max() > 1e-6


def test_oracle_adaptive_loss_equals_target_logprob_sum():
    """adaptive_loss.py:31-83 restated: the summed cross-entropy of the head and tail clusters equals -sum log p(target) of
    get_log_prob (adaptive_softmax.py:170-206), for the adaptive layouts of the fixtures and for a plain output layer."""
    torch.manual_seed(0)
    d, V, cutoff = 32, 60, [10, 30, 60]
    w = {"head": torch.randn(cutoff[0] + 2, d, dtype=torch.float64), "tail_proj": [torch.randn(8, d, dtype=torch.float64), torch.randn(4, d, dtype=torch.float64)],
         "tail_out": [torch.randn(20, 8, dtype=torch.float64), torch.randn(30, 4, dtype=torch.float64)]}
    x = torch.randn(50, d, dtype=torch.float64)
    target = torch.randint(0, V, (50,))
    loss = mo.adaptive_loss(w, cutoff, x, target)
    lp = mo.adaptive_target_logprob(w, cutoff, x, target)
    assert abs(float(loss) + float(lp.sum())) < 1e-9 * abs(float(loss))
    wp = {"plain": torch.randn(V, d, dtype=torch.float64)}
    assert abs(float(mo.adaptive_loss(wp, None, x, target)) + float(mo.plain_target_logprob(wp["plain"], x, target).sum())) < 1e-9


# ---------------------------------------------------------------------------------------------------------------------
# General heterograph / two_stream / incremental infer (hgt.py:81-297,324-330,360-394): oracle/hetero_oracle.py
# ---------------------------------------------------------------------------------------------------------------------
def load_hetero_case(golden_dir, name):
    """(z, sd, num_nodes, edges as int64 tensors, feats) of a tests/golden/hgt_{hetero4,two_stream,infer_*}.npz fixture."""
    z = np.load(os.path.join(golden_dir, f"hgt_{name}.npz"))
    num_nodes = {str(t): int(n) for t, n in zip(z["ntypes"], z["num_nodes"])}
    edges = {}
    for c in z["cets"]:
        cet = tuple(str(c).split("|"))
        if "src." + str(c) in z.files:
            edges[cet] = (torch.from_numpy(z["src." + str(c)]), torch.from_numpy(z["dst." + str(c)]))
        else:                                           # causal tgt-intra-tgt inside blocks of max_len (not stored)
            L_, B = int(z["max_len"]), int(z["bsz"])
            u, v = np.triu_indices(L_)
            edges[cet] = (torch.from_numpy(np.concatenate([u + b * L_ for b in range(B)])),
                          torch.from_numpy(np.concatenate([v + b * L_ for b in range(B)])))
    feats = {t: torch.from_numpy(z["h." + t]) for t in num_nodes if "h." + t in z.files}
    return z, _sd(z, "sd."), num_nodes, edges, feats


def test_hetero_layer_matches_reference(golden_dir):
    from oracle import hetero_oracle as ho
    z, sd, nn_, edges, feats = load_hetero_case(golden_dir, "hetero4")
    out = ho.hgt_forward_hetero(sd, feats, edges, nn_, {"src": 0, "nsrc": 1, "tgt": 2, "ntgt": 3}, {"intra": 0, "inter": 1},
                                int(z["H"]), int(z["n_layers"]))
    for t in nn_:
        np.testing.assert_allclose(out[t].numpy(), z["out." + t], rtol=1e-5, atol=2e-6)


def test_two_stream_matches_fixed_reference(golden_dir):
    from oracle import hetero_oracle as ho
    z, sd, nn_, edges, feats = load_hetero_case(golden_dir, "two_stream")
    out = ho.hgt_forward_hetero(sd, feats, edges, nn_, {"src": 0, "tgt": 1, "ntgt": 2}, {"intra": 0, "inter": 1},
                                int(z["H"]), int(z["n_layers"]), two_stream=True)
    assert sorted(out) == ["ntgt", "src", "tgt", "tgt_tilde"]
    for t in out:
        np.testing.assert_allclose(out[t].numpy(), z["out." + t], rtol=1e-5, atol=2e-6)
    assert np.abs(z["out.tgt_tilde"] - z["out.tgt"]).max() > 1e-2          # the query stream is a different function


@pytest.mark.parametrize("case", ["infer_b2", "infer_b3_reorder"])
def test_incremental_infer_matches_reference(case, golden_dir):
    from oracle import hetero_oracle as ho
    z, sd, nn_, edges, feats = load_hetero_case(golden_dir, case)
    H, NL, steps, bsz = int(z["H"]), int(z["n_layers"]), int(z["steps"]), int(z["bsz"])
    etypes = list(edges)
    states = [dict() for _ in range(NL)]
    h_steps = torch.from_numpy(z["h_steps"])
    order = torch.arange(bsz)
    for s in range(steps):
        if int(z["reorder_at"]) == s:
            order = torch.arange(bsz).flip(0)
            states = [ho.reorder_state(st, order) for st in states]
        h = {"tgt": h_steps[s][order] if 0 <= int(z["reorder_at"]) <= s else h_steps[s], "ntgt": feats["ntgt"]}
        for l in range(NL):
            h = ho.hgt_layer_infer(sd, f"gcs.{l}.", h, edges, nn_, {"tgt": 0, "ntgt": 1}, {"intra": 0, "inter": 1}, H, etypes,
                                   states[l])
        np.testing.assert_allclose(h["tgt"].numpy(), z["out_steps"][s], rtol=1e-5, atol=2e-6)
        if s == 0:
            np.testing.assert_allclose(h["ntgt"].numpy(), z["out_ntgt_first"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(h["ntgt"].numpy(), z["out_ntgt_last"], rtol=1e-5, atol=2e-6)
    if int(z["reorder_at"]) < 0:
        # incremental decoding with teacher forcing == the full layer on the same prefix: position s of every block
        full = ho.hgt_forward_hetero(sd, {"tgt": _prefix_feats(h_steps, int(z["max_len"])), "ntgt": feats["ntgt"]}, edges, nn_,
                                     {"tgt": 0, "ntgt": 1}, {"intra": 0, "inter": 1}, H, 1)
        idx = torch.arange(bsz) * int(z["max_len"])
        np.testing.assert_allclose(full["tgt"][idx].numpy(), _first_layer_step0(z, sd, nn_, edges, feats), rtol=1e-4, atol=1e-5)


def _prefix_feats(h_steps, max_len):
    steps, bsz, d = h_steps.shape
    x = torch.zeros(bsz * max_len, d)
    for s in range(steps):
        x[torch.arange(bsz) * max_len + s] = h_steps[s]
    return x


def _first_layer_step0(z, sd, nn_, edges, feats):
    from oracle import hetero_oracle as ho
    h = {"tgt": torch.from_numpy(z["h_steps"])[0], "ntgt": feats["ntgt"]}
    return ho.hgt_layer_infer(sd, "gcs.0.", h, edges, nn_, {"tgt": 0, "ntgt": 1}, {"intra": 0, "inter": 1}, int(z["H"]),
                              list(edges), {})["tgt"].numpy()
