"""Pin the graph oracle to fixtures produced by executing the reference's own builder
(tests/golden/make_golden.py; fairseq/data/token_block_dataset.py:338-412,545-594)."""
import glob
import os

import numpy as np
import pytest

from oracle import graph_oracle as go

CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "graph_*.npz")))


def test_doctest_vectors():
    # token_block_dataset.py:549-554
    o2i = {0: 0, 1: 1, 2: 2, 12: 3, 13: 4}
    assert go.build_ntgt_edges(o2i, 3) == ([0, 0, 1, 0, 1, 2, 3, 3, 4], [0, 1, 1, 2, 2, 2, 3, 4, 4])
    assert go.build_ntgt_edges(o2i, 0) == ([0, 1, 2, 3, 4], [0, 1, 2, 3, 4])
    assert go.build_ntgt_edges({}, 1) == ([], [])


def test_misc_edges(golden_dir):
    z = np.load(os.path.join(golden_dir, "edges_misc.npz"))
    s, d = go.build_ntgt_edges({7: 3, 5: 1, 6: 0, 9: 2}, 1, bidirect=True)
    assert s == z["bidirect_src"].tolist() and d == z["bidirect_dst"].tolist()
    u, v = go.auto_regressive_edges(6, 0)
    assert (u == z["ar6_u"]).all() and (v == z["ar6_v"]).all()
    u, v = go.auto_regressive_edges(6, 3)
    assert (u == z["ar6c3_u"]).all() and (v == z["ar6c3_v"]).all()


@pytest.mark.parametrize("case", CASES)
def test_new_build_graph_matches_reference(case, golden_dir):
    z = np.load(os.path.join(golden_dir, f"graph_{case}.npz"))
    g = go.new_build_graph(z["offsets"], z["nbr"], int(z["n_d"]), int(z["cl"]), int(z["cr"]),
                           int(z["invalid_ctx"]), int(z["intra_ctx"]), z["codes"], z["vals"])
    for et in ("tt", "inter", "nn"):
        assert (g[et][0] == z[et + "_src"]).all(), et
        assert (g[et][1] == z[et + "_dst"]).all(), et
    assert (g["ntgt_codes"] == z["ntgt_codes"]).all()
    assert (g["ntgt_labels"] == z["ntgt_labels"].reshape(-1)).all()


@pytest.mark.parametrize("case", CASES)
def test_vectorised_matches_loop(case, golden_dir):
    z = np.load(os.path.join(golden_dir, f"graph_{case}.npz"))
    nbr, off = z["nbr"][None], z["offsets"][None]
    ref = go.build_batch(nbr, off, int(z["n_d"]), int(z["cl"]), int(z["cr"]), int(z["invalid_ctx"]))
    vec = go.build_batch_vectorised(nbr, off, int(z["n_d"]), int(z["cl"]), int(z["cr"]), int(z["invalid_ctx"]))
    assert vec["n_ntgt"] == ref["n_ntgt"]
    assert (vec["ntgt_offsets"] == ref["ntgt_offsets"]).all()
    for name in ("nn_csr", "inter_csr"):
        assert (vec[name][0] == ref[name][0]).all(), name
        assert (vec[name][1] == ref[name][1]).all(), name


def test_batching_offsets():
    rng = np.random.RandomState(0)
    nbr = rng.randint(0, 50, size=(3, 5, 2)).astype(np.int64)
    nbr[1, 2] = -1
    off = np.arange(15).reshape(3, 5)
    g = go.build_batch(nbr, off, 50, 1, 1)
    assert g["n_tgt"] == 15
    # every inter edge's dst is the owner tgt, in [b*L, (b+1)*L)
    assert g["inter"][1].max() < 15 and (np.diff(g["inter"][1]) >= 0).all()
    # tgt-intra-tgt never crosses a block
    assert ((g["tt"][0] // 5) == (g["tt"][1] // 5)).all()
    assert g["inter_csr"][0][-1] == len(g["inter"][0])


def test_slice_indices_all_break_modes(golden_dir):
    """Block boundaries of every --sample-break-mode against the reference's (de-cythonised) slicing function."""
    z = np.load(os.path.join(golden_dir, "slices.npz"))
    n = 0
    for key in z.files:
        if key.endswith(".sizes"):
            continue
        name, mode, bs = key.split(".")
        got = go.slice_indices(z[f"{name}.sizes"], mode, int(bs))
        assert np.array_equal(got, z[key]), key
        n += 1
    assert n == 32


DEDUP = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dedup_*.npz")))


@pytest.mark.parametrize("case", DEDUP)
def test_deprecated_build_graph_matches_reference(case, golden_dir):
    """`--deprecated` (de-duplicating) builder, token_block_dataset.py:414-479, executed from the reference source."""
    z = np.load(os.path.join(golden_dir, f"dedup_{case}.npz"))
    g = go.deprecated_build_graph(z["offsets"], z["nbr"], int(z["n_d"]), int(z["cl"]), int(z["cr"]), int(z["invalid_ctx"]),
                                  int(z["intra_ctx"]), quant_feats=z["codes"])
    assert g["n_ntgt"] == z["ntgt_codes"].shape[0]
    for name in ("tt", "inter", "nn"):
        assert np.array_equal(g[name][0], z[f"{name}_src"]), name
        assert np.array_equal(g[name][1], z[f"{name}_dst"]), name
    assert np.array_equal(g["ntgt_codes"], z["ntgt_codes"])
    assert len(set(g["ntgt_offsets"].tolist())) == g["n_ntgt"]            # one node per distinct datastore row
