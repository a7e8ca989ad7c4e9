"""On-disk formats of a reference data directory (gnnlm_b200.formats): the readers against files written AND read back by
the reference's own code (tests/golden/fmt_* from make_golden.py: MMapIndexedDatasetBuilder / MMapIndexedDataset /
Dictionary), and load_graph_lm_dataset -- the `--graph` branch of LanguageModelingTask.load_dataset -- over a directory laid
out as knn/path_utils.py:13-41 names it.  CPU only."""
import json
import os
import shutil

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


from gnnlm_b200.formats import write_mmap_indexed      # writer of {prefix}.bin/.idx, checked byte for byte below


@pytest.mark.parametrize("tag,dtype", [("uint16", np.uint16), ("int32", np.int32)])
def test_test_side_writer_is_byte_identical_to_the_reference_builder(tmp_path, tag, dtype):
    z = np.load(os.path.join(GOLD, "fmt.npz"))
    write_mmap_indexed(str(tmp_path / "w"), [z[f"{tag}_sent{i}"] for i in range(6)], dtype)
    for ext in ("bin", "idx"):
        assert (tmp_path / f"w.{ext}").read_bytes() == open(os.path.join(GOLD, f"fmt_{tag}.{ext}"), "rb").read()


@pytest.mark.parametrize("tag", ["uint16", "int32"])
def test_mmap_indexed_dataset_reads_reference_files(tag):
    from gnnlm_b200.formats import MMapIndexedDataset
    z = np.load(os.path.join(GOLD, "fmt.npz"))
    ds = MMapIndexedDataset(os.path.join(GOLD, f"fmt_{tag}"))
    assert len(ds) == len(z[f"{tag}_sizes"]) == 6
    assert (np.asarray(ds.sizes) == z[f"{tag}_sizes"]).all() and ds.sizes.dtype == np.int32
    for i in range(len(ds)):
        s = ds[i]
        assert s.dtype == np.int64 and (s == z[f"{tag}_sent{i}"]).all()
    flat = ds.tokens()
    assert flat.dtype == np.dtype(tag) and (flat.astype(np.int64) == z[f"{tag}_flat"]).all()


def test_mmap_indexed_dataset_errors(tmp_path):
    from gnnlm_b200.formats import MMapIndexedDataset
    with pytest.raises(FileNotFoundError):
        MMapIndexedDataset(str(tmp_path / "absent"))
    (tmp_path / "bad.idx").write_bytes(b"TNTIDX\x00\x00" + b"\x00" * 32)      # the legacy (non-mmap) index magic
    (tmp_path / "bad.bin").write_bytes(b"")
    with pytest.raises(ValueError):
        MMapIndexedDataset(str(tmp_path / "bad"))


def test_dictionary_numbering_matches_reference():
    from gnnlm_b200.formats import Dictionary
    z = np.load(os.path.join(GOLD, "fmt.npz"))
    d = Dictionary.load(os.path.join(GOLD, "fmt_dict.txt"))
    assert len(d) == int(z["dict_len"])
    assert (d.bos(), d.pad(), d.eos(), d.unk()) == (int(z["dict_bos"]), int(z["dict_pad"]), int(z["dict_eos"]), int(z["dict_unk"]))
    assert (d.bos(), d.pad(), d.eos(), d.unk()) == (0, 1, 2, 3)
    assert d.symbols == [str(s) for s in z["dict_symbols"]]
    assert d.index("w7") == int(z["dict_index_w7"]) and d.index("nope") == int(z["dict_index_missing"]) == d.unk()
    assert d.index("with space") == d.symbols.index("with space")               # split at the LAST space


def _write_data_dir(root, *, k=4, hidden=16, n_d=500, fp16=True, seed=0):
    """A reference data directory around the golden uint16 token file: dict.txt, valid.{bin,idx}, {valid,train}_dstore/."""
    rng = np.random.RandomState(seed)
    os.makedirs(root, exist_ok=True)
    shutil.copy(os.path.join(GOLD, "fmt_dict.txt"), os.path.join(root, "dict.txt"))
    for ext in ("bin", "idx"):
        shutil.copy(os.path.join(GOLD, f"fmt_uint16.{ext}"), os.path.join(root, f"valid.{ext}"))
    n_tok = int(np.load(os.path.join(GOLD, "fmt.npz"))["uint16_flat"].shape[0])
    os.makedirs(os.path.join(root, "valid_dstore"))
    os.makedirs(os.path.join(root, "train_dstore"))
    feats = rng.randn(n_tok, hidden).astype(np.float16 if fp16 else np.float32)
    nbr = rng.randint(0, n_d, size=(n_tok, k)).astype(np.int64)
    nbr[0, 0] = -1
    feats.tofile(os.path.join(root, "valid_dstore", "keys.npy"))              # raw, despite the suffix
    nbr.tofile(os.path.join(root, "valid_dstore", f"neighbors.mmap.{k}"))
    json.dump({"dstore_size": n_tok, "hidden_size": hidden, "vocab_size": 48, "dstore_fp16": fp16, "val_size": 1},
              open(os.path.join(root, "valid_dstore", "info.json"), "w"))
    json.dump({"dstore_size": n_d, "hidden_size": hidden, "vocab_size": 48, "dstore_fp16": fp16, "val_size": 1},
              open(os.path.join(root, "train_dstore", "info.json"), "w"))
    return n_tok, feats, nbr


@pytest.mark.parametrize("fp16", [True, False])
def test_load_graph_lm_dataset_from_reference_layout(tmp_path, fp16):
    from gnnlm_b200.dataset import GraphTokenBlockDataset
    from gnnlm_b200.formats import load_graph_lm_dataset
    root = str(tmp_path / "data-bin")
    n_tok, feats, nbr = _write_data_dir(root, fp16=fp16)
    z = np.load(os.path.join(GOLD, "fmt.npz"))
    ds, d = load_graph_lm_dataset(root, "valid", tokens_per_sample=8, gcn_k=4, neighbor_context="(2,0)", gcn_context_window=3,
                                  invalid_neighbor_context=100)
    assert len(d) == int(z["dict_len"]) and ds.n_datastore == 500
    assert (ds.left_neighbor_context, ds.right_neighbor_context) == (2, 0)
    assert ds.invalid_neighbor_context == 0                     # only the train split keeps it (language_modeling.py:295)
    direct = GraphTokenBlockDataset(z["uint16_flat"], 8, pad=1, eos=2, neighbor_offsets=nbr, n_datastore=500,
                                    neighbor_context=(2, 0), precompute_feats=feats, context_window=3)
    assert len(ds) == len(direct) == -(-n_tok // 8)
    for i in range(len(ds)):
        a, b = ds[i], direct[i]
        assert a["offsets"] == b["offsets"] and a["start_idx"] == b["start_idx"]
        for key in ("source", "target", "nbr", "feats"):
            assert a[key].dtype == b[key].dtype and (a[key] == b[key]).all(), key
    # sentence-per-block slicing uses the .idx sizes
    ds_eos, _ = load_graph_lm_dataset(root, "valid", tokens_per_sample=8, gcn_k=4, sample_break_mode="eos")
    assert [e - s for s, e in ds_eos.slice_indices] == list(z["uint16_sizes"])


def test_load_graph_lm_dataset_errors(tmp_path):
    from gnnlm_b200.formats import load_graph_lm_dataset
    root = str(tmp_path / "data-bin")
    _write_data_dir(root)
    with pytest.raises(FileNotFoundError):                      # split without token files
        load_graph_lm_dataset(root, "test", tokens_per_sample=8, gcn_k=4)
    with pytest.raises(FileNotFoundError):                      # neighbours were searched with another k
        load_graph_lm_dataset(root, "valid", tokens_per_sample=8, gcn_k=32)
    with open(os.path.join(root, "valid_dstore", "neighbors.mmap.4"), "r+b") as f:
        f.truncate(64)
    with pytest.raises(ValueError):                             # truncated neighbour file
        load_graph_lm_dataset(root, "valid", tokens_per_sample=8, gcn_k=4)
