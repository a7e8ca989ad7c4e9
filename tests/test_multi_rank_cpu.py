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


# ------------------------------------------------------------------------------------------ registration face
def _registry_golden():
    import json
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "registry.json")))


@pytest.mark.parametrize("group", ["model", "task", "eval"])
def test_flags_match_reference_parsers(group):
    """Every flag of the reference's transformer_lm.add_args / LanguageModelingTask.add_args / add_eval_lm_args exists here with
    the same spellings, kind, type and default (golden produced by executing the reference's own add_args)."""
    import argparse
    from gnnlm_b200 import registry
    g = _registry_golden()
    p = argparse.ArgumentParser()
    {"model": registry.add_model_args, "task": registry.add_task_args, "eval": registry.add_eval_lm_args}[group](p)
    got = vars(p.parse_args(["DATA"] if group == "task" else []))
    want = g[group + "_flags"]
    assert got == want
    acts = {a.dest: a for a in p._actions if a.dest != "help"}
    for dest, o in g[group + "_options"].items():
        a = acts[dest]
        assert list(a.option_strings) == o["flags"], dest
        assert type(a).__name__ == o["action"] and getattr(a.type, "__name__", None) == o["type"], dest
        assert (list(a.choices) if a.choices else None) == o["choices"] and a.nargs == o["nargs"] and a.const == o["const"], dest


def test_architecture_presets_match_reference():
    """apply_architecture == the reference's `transformer_lm*` architecture functions on an empty Namespace, on an old-checkpoint
    Namespace, and with a pre-set attribute winning over the preset."""
    import argparse
    from gnnlm_b200 import registry
    g = _registry_golden()
    assert set(registry.ARCHITECTURES) == set(g["archs"])
    for arch, want in g["archs"].items():
        assert vars(registry.apply_architecture(argparse.Namespace(), arch)) == want, arch
    old = registry.apply_architecture(argparse.Namespace(no_tie_adaptive_proj=False, decoder_final_norm=False), "transformer_lm")
    assert vars(old) == g["archs_old_checkpoint"]
    a = registry.apply_architecture(argparse.Namespace(decoder_embed_dim=768, dropout=0.0), "transformer_lm_wiki103")
    assert a.decoder_embed_dim == 768 and a.decoder_input_dim == 768 and a.dropout == 0.0 and a.decoder_layers == 16
    with pytest.raises(ValueError):
        registry.apply_architecture(argparse.Namespace(), "transformer_lm_nope")


def _fake_fairseq():
    """Stand-in for fairseq.models / fairseq.tasks with the decorator semantics of fairseq/models/__init__.py:51-119 and
    fairseq/tasks/__init__.py (duplicate names and wrong base classes raise ValueError)."""
    import types
    import torch.nn as nn
    fm, ft = types.SimpleNamespace(), types.SimpleNamespace()
    fm.MODEL_REGISTRY, fm.ARCH_MODEL_REGISTRY, fm.ARCH_MODEL_INV_REGISTRY, fm.ARCH_CONFIG_REGISTRY = {}, {}, {}, {}

    class BaseFairseqModel(nn.Module):
        def __init__(self):
            super().__init__()
            self._is_generation_fast = False

    fm.BaseFairseqModel = BaseFairseqModel

    def register_model(name):
        def deco(cls):
            if name in fm.MODEL_REGISTRY:
                raise ValueError("Cannot register duplicate model ({})".format(name))
            if not issubclass(cls, BaseFairseqModel):
                raise ValueError("Model ({}: {}) must extend BaseFairseqModel".format(name, cls.__name__))
            fm.MODEL_REGISTRY[name] = cls
            return cls
        return deco

    def register_model_architecture(model_name, arch_name):
        def deco(fn):
            if model_name not in fm.MODEL_REGISTRY:
                raise ValueError("Cannot register model architecture for unknown model type ({})".format(model_name))
            if arch_name in fm.ARCH_MODEL_REGISTRY:
                raise ValueError("Cannot register duplicate model architecture ({})".format(arch_name))
            fm.ARCH_MODEL_REGISTRY[arch_name] = fm.MODEL_REGISTRY[model_name]
            fm.ARCH_MODEL_INV_REGISTRY.setdefault(model_name, []).append(arch_name)
            fm.ARCH_CONFIG_REGISTRY[arch_name] = fn
            return fn
        return deco

    fm.register_model, fm.register_model_architecture = register_model, register_model_architecture

    class FairseqTask:
        def __init__(self, args):
            self.args, self.datasets = args, {}

    ft.FairseqTask, ft.TASK_REGISTRY = FairseqTask, {}

    def register_task(name):
        def deco(cls):
            if name in ft.TASK_REGISTRY:
                raise ValueError("Cannot register duplicate task ({})".format(name))
            if not issubclass(cls, FairseqTask):
                raise ValueError("Task ({}: {}) must extend FairseqTask".format(name, cls.__name__))
            ft.TASK_REGISTRY[name] = cls
            return cls
        return deco

    ft.register_task = register_task

    # the stock entries a real fairseq already holds
    class StockLM(BaseFairseqModel):
        @classmethod
        def build_model(cls, args, task):
            return "stock model"

    class StockTask(FairseqTask):
        @classmethod
        def setup_task(cls, args, **kw):
            return "stock task"

    register_model("transformer_lm")(StockLM)
    register_model_architecture("transformer_lm", "transformer_lm")(lambda a: None)
    register_model_architecture("transformer_lm", "transformer_lm_wiki103")(lambda a: None)
    register_task("language_modeling")(StockTask)
    return fm, ft


def test_fairseq_registration_adds_hgt_lm_and_overrides_on_request():
    import argparse
    from gnnlm_b200 import registry
    from gnnlm_b200.model import TransformerLanguageModel
    fm, ft = _fake_fairseq()
    stock = fm.MODEL_REGISTRY["transformer_lm"]
    names = registry.register_with_fairseq(fm, ft, override=False)
    assert "hgt_lm" in fm.MODEL_REGISTRY and "hgt_lm_wiki103" in names and "hgt_lm_baevski_gbw" in names
    assert fm.MODEL_REGISTRY["transformer_lm"] is stock and ft.TASK_REGISTRY["language_modeling"].__name__ == "StockTask"
    assert issubclass(fm.MODEL_REGISTRY["hgt_lm"], fm.BaseFairseqModel) and issubclass(fm.MODEL_REGISTRY["hgt_lm"], TransformerLanguageModel)
    a = argparse.Namespace()
    fm.ARCH_CONFIG_REGISTRY["hgt_lm_wiki103"](a)
    assert a.decoder_embed_dim == 1024 and a.adaptive_softmax_cutoff == "20000,60000" and a.decoder_attention_heads == 8
    with pytest.raises(ValueError):                               # fairseq's own duplicate check fires on a second import
        registry.register_with_fairseq(fm, ft, override=False)
    # override: existing checkpoints (arch = transformer_lm_wiki103) and scripts (--task language_modeling) resolve here ...
    fm2, ft2 = _fake_fairseq()
    registry.register_with_fairseq(fm2, ft2, override=True)
    cls = fm2.ARCH_MODEL_REGISTRY["transformer_lm_wiki103"]
    assert cls is fm2.MODEL_REGISTRY["hgt_lm"] is fm2.MODEL_REGISTRY["transformer_lm"]
    # ... only when they are graph runs: everything else goes back to the stock classes
    assert cls.build_model(argparse.Namespace(graph_layer=0), None) == "stock model"
    assert ft2.TASK_REGISTRY["language_modeling"].setup_task(argparse.Namespace(graph=False)) == "stock task"


def test_registered_model_and_task_build_from_a_command_line(tmp_path):
    """The reference's eval command line -> parser -> task.setup_task / load_dataset -> ARCH_MODEL_REGISTRY[arch].build_model:
    the model carries the reference's state_dict keys and the dataset reads the reference's on-disk layout."""
    from gnnlm_b200 import registry
    from tests.test_formats import _write_data_dir
    fm, ft = _fake_fairseq()
    registry.register_with_fairseq(fm, ft, override=True)
    root = str(tmp_path / "data-bin")
    _write_data_dir(root, k=4, hidden=64)
    argv = [root, "--graph", "--use-precompute-feat", "--graph_layer", "2", "--decoder_gcn_dim", "64", "--decoder-embed-dim", "64",
            "--decoder-attention-heads", "4", "--adaptive-softmax-cutoff", "10,20", "--tokens-per-sample", "8", "--gcn-k", "4",
            "--neighbor-context", "1", "--gen-subset", "valid", "--lmbda", "0.25", "--knn-keytype", "gcn_feat", "--softmax-batch", "1024"]
    args = registry.eval_lm_parser().parse_args(argv)
    assert args.graph and args.neighbor_context == "1" and args.k == 1024 and args.softmax_batch == 1024
    fm.ARCH_CONFIG_REGISTRY[args.arch](args) if args.arch in fm.ARCH_CONFIG_REGISTRY else None
    task = ft.TASK_REGISTRY["language_modeling"].setup_task(args)
    assert isinstance(task, ft.FairseqTask) and len(task.source_dictionary) == 48
    ds = task.load_dataset(args.gen_subset)
    assert task.dataset("valid") is ds and ds.block_size == 8 and (ds.left_neighbor_context, ds.right_neighbor_context) == (1, 1)
    model = fm.ARCH_MODEL_REGISTRY[args.arch].build_model(args, task)
    assert isinstance(model, fm.BaseFairseqModel) and model._is_generation_fast is False
    keys = set(model.state_dict())
    for k in ("decoder.hgt_decoder.gcs.1.k_linears.0.weight", "decoder.hgt_decoder.gcs.0.relation_att", "decoder.hgt_decoder.gcs.0.skip",
              "decoder.adaptive_softmax.head.weight", "decoder.adaptive_softmax.tail.1.2.weight"):
        assert k in keys, k
    assert args.decoder_input_dim == 64 and args.decoder_normalize_before is True and args.max_target_positions == 8


def test_adaptive_softmax_key_layouts_match_reference(golden_dir):
    """state_dict keys and shapes of model.AdaptiveSoftmax against the reference module's own (fixtures written by
    tests/golden/make_golden.py from fairseq/modules/adaptive_softmax.py): untied, tied + tied projections, and tied weights
    with untied projections (nn.Linear(d, dim_i), adaptive_softmax.py:96-101)."""
    from gnnlm_b200.model import AdaptiveSoftmax
    for case in ("untied", "tied", "tied_noproj"):
        z = np.load(os.path.join(golden_dir, f"adaptive_{case}.npz"))
        cutoff = z["cutoff"].tolist()
        m = AdaptiveSoftmax(cutoff[-1], z["x"].shape[-1], cutoff[:-1], tied=bool(z["tied"]),
                            tie_proj=bool(z["tie_proj"]) if "tie_proj" in z.files else None)
        want = {k[3:]: z[k].shape for k in z.files if k.startswith("sd.")}
        got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert got == want, case


def test_dataset_rejects_neighbour_ids_of_another_datastore():
    """neighbors.mmap is a raw memmap without a header: ids outside [-1, N_d) raise IndexError when the block is sliced, as the
    reference does at `quant_neighbor_feats[o]` (token_block_dataset.py:369-370), instead of reaching the device gathers."""
    from gnnlm_b200.dataset import GraphTokenBlockDataset
    tokens = np.arange(4, 36).astype(np.int64)
    nbr = np.random.RandomState(0).randint(-1, 100, size=(32, 4)).astype(np.int64)
    ok = GraphTokenBlockDataset(tokens, 8, pad=1, eos=2, neighbor_offsets=nbr, n_datastore=100)
    assert ok[3]["nbr"].shape == (8, 4)
    for bad in (100, -2):
        nb2 = nbr.copy()
        nb2[17, 2] = bad
        ds = GraphTokenBlockDataset(tokens, 8, pad=1, eos=2, neighbor_offsets=nb2, n_datastore=100)
        assert ds[1]["nbr"].shape == (8, 4)          # other blocks are unaffected
        with pytest.raises(IndexError):
            ds[2]
    ids = np.zeros((32, 3), np.int64)
    ids[5, 1] = 100
    ds = GraphTokenBlockDataset(tokens, 8, pad=1, eos=2, neighbor_offsets=nbr, n_datastore=100, knn_ids=ids,
                                knn_dists=np.zeros((32, 3), np.float32))
    with pytest.raises(IndexError):
        ds[0]


def test_dstore_writer_shards_compose(tmp_path):
    """--save-knnlm-dstore under sharding: every shard writes its own row range of the same keys.npy / vals.npy, in any
    launch order, without clearing rows another shard wrote; the clip at dstore_size (eval_lm.py:227-230) applies per shard."""
    from gnnlm_b200.eval_lm import DstoreWriter, shard_range
    rng = np.random.RandomState(0)
    sizes = np.array([5, 7, 3, 9, 4])
    n, d = int(sizes.sum()), 6
    feats, toks = rng.randn(n, d).astype(np.float32), rng.randint(4, 50, size=n)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    written = 0
    for rank in (2, 0, 1):                        # any order
        lo, hi = shard_range(len(sizes), rank, 3)
        w = DstoreWriter(str(tmp_path), "valid", n, d, 50, dstore_fp16=False, knn_keytype="gcn_feat", offset=int(starts[lo]),
                         limit=int(starts[hi] - starts[lo]), write_info=rank == 0)
        for b in range(lo, hi):
            w.add(torch.from_numpy(feats[starts[b]:starts[b + 1]]), torch.from_numpy(toks[starts[b]:starts[b + 1]]))
        w.add(torch.zeros(2, d), torch.zeros(2, dtype=torch.int64))          # past the shard's range: dropped
        written += w.close()
    assert written == n
    out = os.path.join(str(tmp_path), "valid_dstore-gcn_feat")
    assert (np.fromfile(os.path.join(out, "keys.npy"), np.float32).reshape(n, d) == feats).all()
    assert (np.fromfile(os.path.join(out, "vals.npy"), np.int32) == toks).all()
    import json
    assert json.load(open(os.path.join(out, "info.json")))["dstore_size"] == n


def test_host_batcher_fills_the_collaters_tensors():
    """evaluate()'s fast host path: GraphTokenBlockDataset.collate_into (memmap slices copied once into preallocated batch
    buffers) produces exactly the tensors of the reference-shaped dataset[i] + collater (token_block_dataset.py:287-333,
    monolingual_dataset.py:237-262), with and without --gcn-context-window, for full, ragged and single-block batches; the
    HostBatcher producer thread delivers the batches of eval_lm.batches in order and recycles its buffer sets."""
    from gnnlm_b200.dataset import GraphTokenBlockDataset
    from gnnlm_b200.eval_lm import HostBatcher, batches, host_inputs
    rng = np.random.RandomState(0)
    n_tok = 64 * 9 + 13
    tokens = rng.randint(4, 900, size=n_tok).astype(np.uint16)
    nbr = rng.randint(-1, 5000, size=(n_tok, 4)).astype(np.int64)
    feats = rng.randn(n_tok, 32).astype(np.float16)
    kid, kd = rng.randint(0, 5000, size=(n_tok, 8)).astype(np.int64), rng.randn(n_tok, 8).astype(np.float32)
    for cw in (0, 16):
        ds = GraphTokenBlockDataset(tokens, 64, pad=1, eos=2, neighbor_offsets=nbr, n_datastore=5000, neighbor_context=1,
                                    precompute_feats=feats, context_window=cw, knn_dists=kd, knn_ids=kid)
        want = list(batches(ds, 0, len(ds), 2))
        got = [item["ids"] for item in HostBatcher(ds, want, depth=len(want), workers=0)]     # inline: calling thread, native copies
        assert got == want
        got = []
        for item in HostBatcher(ds, want, depth=2, workers=2):       # 3 buffer sets for 5-6 batches: recycling is exercised
            slow = host_inputs(ds.collater([ds[i] for i in item["ids"]]))
            assert set(item["host"]) == set(slow)
            for k_ in slow:
                assert torch.equal(item["host"][k_], slow[k_].reshape(item["host"][k_].shape)), (cw, item["ids"], k_)
            assert item["ntokens"] == sum(len(ds[i]["target"]) for i in item["ids"])
            got.append(item["ids"])
            HostBatcher.release(item)
        assert got == want
    bad = nbr.copy()
    bad[70, 0] = 5000
    ds = GraphTokenBlockDataset(tokens, 64, pad=1, eos=2, neighbor_offsets=bad, n_datastore=5000)
    with pytest.raises(IndexError):
        list(HostBatcher(ds, [[0], [1]]))
