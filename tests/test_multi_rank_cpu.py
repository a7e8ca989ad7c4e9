"""N>1 host logic on CPU with world_size-2 gloo: contiguous block sharding and the path's only
collective -- one all-reduce of {sum log p, n_tokens}.  Per-block scores come from the CPU oracle (test
infrastructure); the sharded + reduced perplexity must equal the single-process one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _block_scores(n_blocks, seed=0):
    """Deterministic stand-in for per-block (score_sum, count): what evaluate() accumulates per batch."""
    rng = np.random.RandomState(seed)
    counts = rng.randint(50, 100, size=n_blocks)
    sums = -rng.rand(n_blocks) * counts * 5.0
    return sums, counts


def _worker(rank, world, port, n_blocks, out):
    import importlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ev = importlib.import_module("gnnlm_b200.eval_lm")
    lo, hi = ev.shard_range(n_blocks, rank, world)
    sums, counts = _block_scores(n_blocks)
    acc = torch.tensor([sums[lo:hi].sum(), counts[lo:hi].sum()], dtype=torch.float64)
    dist.all_reduce(acc)
    out[rank] = (lo, hi, acc.tolist())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_blocks", [1, 7, 16])
def test_sharded_nll_allreduce_gloo(n_blocks):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_blocks, out), nprocs=world, join=True)
    sums, counts = _block_scores(n_blocks)
    ranges = sorted((out[r][0], out[r][1]) for r in range(world))
    assert ranges[0][0] == 0 and ranges[-1][1] == n_blocks
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))        # contiguous, disjoint, complete
    for r in range(world):
        s, c = out[r][2]
        assert c == counts.sum() and abs(s - sums.sum()) < 1e-9


def test_shard_range_properties():
    from gnnlm_b200.eval_lm import shard_range
    for nb in (0, 1, 5, 8, 33, 1000):
        for R in (1, 2, 3, 4, 8):
            rs = [shard_range(nb, r, R) for r in range(R)]
            assert rs[0][0] == 0 and rs[-1][1] == nb
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def test_batches_group_equal_lengths():
    from gnnlm_b200.dataset import GraphTokenBlockDataset
    from gnnlm_b200.eval_lm import batches
    n = 1000
    ds = GraphTokenBlockDataset(np.arange(n) % 50 + 4, 128, pad=1, eos=2, neighbor_offsets=np.zeros((n, 2), np.int64),
                                n_datastore=10)
    got = list(batches(ds, 0, len(ds), 3))
    assert got == [[0, 1, 2], [3, 4, 5], [6], [7]]        # 7 full blocks (3+3+1) + the ragged last block alone
    it = ds[0]
    assert it["source"][0] == 2 and (it["source"][1:] == it["target"][:-1]).all()     # eos-shifted source (:310-311)
    it3 = ds[3]
    assert (it3["source"] == torch.from_numpy(ds.tokens[3 * 128 - 1:4 * 128 - 1])).all()   # buffer[s-1:e-1] (:319)
    dsw = GraphTokenBlockDataset(np.arange(n) % 50 + 4, 128, pad=1, eos=2, neighbor_offsets=np.zeros((n, 2), np.int64),
                                 n_datastore=10, context_window=16)
    assert dsw[2]["start_idx"] == 16 and len(dsw[2]["target"]) == 144 and dsw[0]["start_idx"] == 0


def test_c_abi_exports_every_declared_symbol(built_library):
    """include/gnnlm_sm100.h <-> libgnnlm_sm100.so <-> ctypes table (no compute calls without a GPU)."""
    import re
    from gnnlm_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "gnnlm_sm100.h")).read()
    declared = set(re.findall(r"\b(gnnlm_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gnnlm_version() >= 100
    assert lib.gnnlm_graph_tt_num_edges(2, 6, 0) == 2 * 21 and lib.gnnlm_graph_tt_num_edges(1, 6, 3) == 15
    assert lib.gnnlm_graph_workspace_bytes(5000) > 0
    assert lib.gnnlm_lse_num_tiles(20002, 0) == 157


def test_product_never_imports_oracle():
    root = os.path.join(os.path.dirname(__file__), "..", "gnn-lm_b200")
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            src = open(os.path.join(root, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_break_mode_slicing_matches_reference():
    """dataset.get_slice_indices (host logic) against the reference's slicing function for every --sample-break-mode."""
    import os
    import numpy as np
    from gnnlm_b200.dataset import get_slice_indices
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "slices.npz"))
    for key in z.files:
        if key.endswith(".sizes"):
            continue
        name, mode, bs = key.split(".")
        assert np.array_equal(get_slice_indices(z[f"{name}.sizes"], mode, int(bs)), z[key]), key


def test_batches_bucket_by_length():
    import numpy as np
    from gnnlm_b200.dataset import GraphTokenBlockDataset
    from gnnlm_b200.eval_lm import batches
    rng = np.random.RandomState(0)
    sizes = rng.randint(1, 9, size=200)
    n = int(sizes.sum())
    ds = GraphTokenBlockDataset(np.arange(n), 3072, pad=1, eos=2, neighbor_offsets=np.zeros((n, 2), np.int64), n_datastore=10,
                                break_mode="eos", sizes=sizes)
    assert len(ds) == 200 and [e - s for s, e in ds.slice_indices] == sizes.tolist()
    seen = []
    for b in batches(ds, 10, 190, max_sentences=16, max_tokens=64, bucket_by_length=True):
        lens = {int(ds.sizes[i]) for i in b}
        assert len(lens) == 1 and len(b) <= 16 and len(b) * lens.pop() <= 64
        seen += b
    assert sorted(seen) == list(range(10, 190))
    in_order = [i for b in batches(ds, 10, 190, max_sentences=16) for i in b]
    assert in_order == list(range(10, 190))


# ---- the reference's own known-answer tests for the host logic of the path
def _blocks(data, block_size, mode):
    import numpy as np
    from gnnlm_b200.dataset import GraphTokenBlockDataset
    from oracle import graph_oracle as go
    tokens = np.concatenate([np.asarray(x) for x in data])
    sizes = np.array([len(x) for x in data])
    ds = GraphTokenBlockDataset(tokens, block_size, pad=0, eos=1, neighbor_offsets=np.zeros((len(tokens), 1), np.int64),
                                n_datastore=4, break_mode=mode, sizes=sizes)
    got = [ds[i]["target"].tolist() for i in range(len(ds))]
    assert got == [tokens[s:e].tolist() for s, e in go.slice_indices(sizes, mode, block_size)]       # oracle agrees
    return got


def test_token_block_known_answers():
    """The vectors of the reference's tests/test_token_block_dataset.py:22-78 (eos / none / complete break modes)."""
    assert _blocks([[5, 4, 3, 2, 1], [1], [8, 7, 6, 1]], 10 ** 9, "eos") == [[5, 4, 3, 2, 1], [1], [8, 7, 6, 1]]
    assert _blocks([[5, 4, 3, 2, 1], [8, 7, 6, 1], [1]], 10 ** 9, "eos") == [[5, 4, 3, 2, 1], [8, 7, 6, 1], [1]]
    assert _blocks([[5, 4, 3, 2, 1], [8, 7, 6, 1], [9, 1]], 3, "none") == [[5, 4, 3], [2, 1, 8], [7, 6, 1], [9, 1]]
    assert _blocks([[5, 4, 3, 2, 1], [8, 7, 6, 1], [9, 1]], 6, "complete") == [[5, 4, 3, 2, 1], [8, 7, 6, 1, 9, 1]]
    assert _blocks([[4, 3, 2, 1], [5, 1], [1], [6, 1]], 3, "complete") == [[4, 3, 2, 1], [5, 1, 1], [6, 1]]


def test_sequence_scorer_known_answers():
    """The vectors of the reference's tests/test_sequence_scorer.py:17-91: a scripted model emits the per-step
    distributions, the scorer must return the target tokens (pad stripped), their positional log-probs and the mean
    score.  Runs without the kNN stage, so no kernel is involved: this is SequenceScorer.generate's host logic."""
    import math
    from types import SimpleNamespace
    import torch
    from gnnlm_b200.sequence_scorer import SequenceScorer
    pad, eos, w1, w2 = 1, 2, 4, 5
    targets = [[w1, w2, w1, eos], [w2, w1, eos], [w2, eos]]
    beam_probs = [torch.tensor([[0.0, 0.0, 0.6, 0.4], [0.0, 0.0, 0.4, 0.6], [0.0, 0.0, 0.7, 0.3]]),
                  torch.tensor([[0.0, 0.0, 0.2, 0.7], [0.0, 0.0, 0.8, 0.2], [0.7, 0.0, 0.1, 0.2]]),
                  torch.tensor([[0.10, 0.0, 0.50, 0.4], [0.15, 0.0, 0.15, 0.7], [0.0, 0.0, 0.0, 0.0]]),
                  torch.tensor([[0.9, 0.0, 0.05, 0.05], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]])]
    expected = [[0.6, 0.7, 0.5, 0.9], [0.6, 0.8, 0.15], [0.3, 0.7]]
    col = {eos: 0, 3: 1, w1: 2, w2: 3}                                   # columns of beam_probs: eos, unk, w1, w2
    L = 4
    tgt = torch.full((3, L), pad, dtype=torch.long)
    for i, t in enumerate(targets):
        tgt[i, :len(t)] = torch.tensor(t)

    class Decoder:
        def target_log_probs(self, net_output, target):
            probs = torch.stack(beam_probs, 1)                            # [bsz, step, 4]
            idx = torch.tensor([[col.get(int(v), 1) for v in row] for row in target])
            return torch.log(probs.gather(2, idx.unsqueeze(-1)).squeeze(-1))

    class Model:
        decoder = Decoder()

        def eval(self):
            return self

        def __call__(self, **net_input):
            feat = torch.zeros(L, 3, 2)
            return feat.transpose(0, 1), {"inner_states": [feat]}

    d = SimpleNamespace(pad=lambda: pad, eos=lambda: eos)
    scorer = SequenceScorer(d)
    hypos = scorer.generate([Model()], {"net_input": {"src_tokens": tgt}, "target": tgt})
    for i, h in enumerate(hypos):
        assert h[0]["tokens"].tolist() == targets[i]
        want = torch.log(torch.tensor(expected[i]))
        assert (h[0]["positional_scores"] - want).abs().max() < 1e-4
        assert abs(float(h[0]["score"]) - float(want.sum()) / len(expected[i])) < 1e-6
