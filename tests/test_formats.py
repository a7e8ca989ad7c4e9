"""On-disk formats of a reference data directory (gnnlm_b200.formats): the readers against files written AND read back by
the reference's own code (tests/golden/fmt_* from make_golden.py: MMapIndexedDatasetBuilder / MMapIndexedDataset /
Dictionary), and load_graph_lm_dataset -- the `--graph` branch of LanguageModelingTask.load_dataset -- over a directory laid
out as knn/path_utils.py:13-41 names it.  CPU only."""
import json
import os
import shutil

import numpy as np
import pytest
import torch

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


# ---------------------------------------------------------------------------------------------------------------------
# faiss `quantizer` file (transformer.py:936-937, pq_wrapper.py:20-37) -- layout restated from faiss's index_write.cpp;
# faiss is absent here, so this pins the reader to the writer of the same restatement and to hand-assembled bytes.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("opq,with_b", [(True, False), (True, True), (False, False)])
def test_faiss_quantizer_round_trip(tmp_path, opq, with_b):
    from gnnlm_b200.formats import read_faiss_quantizer, write_faiss_quantizer
    rng = np.random.RandomState(0)
    M, dsub, d_in = 8, 4, 40
    cen = rng.randn(M, 256, dsub).astype(np.float32)
    A = rng.randn(M * dsub, d_in).astype(np.float32) if opq else None
    b = rng.randn(M * dsub).astype(np.float32) if with_b else None
    path = str(tmp_path / "quantizer")
    write_faiss_quantizer(path, cen, A, b)
    cen2, A2, b2 = read_faiss_quantizer(path)
    assert (cen2 == cen).all()
    if opq:
        assert A2.shape == (M * dsub, d_in) and (A2 == A).all() and b2.size == (M * dsub if with_b else 0)
        assert not with_b or (b2 == b).all()
    else:
        assert A2 is None and b2 is None


def test_faiss_quantizer_hand_assembled_bytes(tmp_path):
    """The byte layout spelled out field by field (index_write.cpp: write_index_header, generic LinearTransform 'LTra',
    write_ProductQuantizer, IndexPQ 'IxPq'), independent of write_faiss_quantizer."""
    import struct
    from gnnlm_b200.formats import read_faiss_quantizer
    M, dsub = 4, 2
    d = M * dsub
    cen = np.arange(M * 256 * dsub, dtype=np.float32)
    A = np.arange(d * d, dtype=np.float32)
    hdr = lambda dim: struct.pack("<i", dim) + struct.pack("<q", 0) + struct.pack("<qq", 1 << 20, 1 << 20) + b"\x01" + struct.pack("<i", 0)
    raw = b"IxPT" + hdr(d) + struct.pack("<i", 1)
    raw += b"LTra" + b"\x00" + struct.pack("<Q", d * d) + A.tobytes() + struct.pack("<Q", 0) + struct.pack("<ii", d, d) + b"\x01"
    raw += b"IxPq" + hdr(d) + struct.pack("<QQQ", d, M, 8) + struct.pack("<Q", cen.size) + cen.tobytes()
    raw += struct.pack("<Q", 0) + struct.pack("<i", 0) + b"\x00" + struct.pack("<i", 0)
    path = tmp_path / "quantizer"
    path.write_bytes(raw)
    cen2, A2, b2 = read_faiss_quantizer(str(path))
    assert cen2.shape == (M, 256, dsub) and (cen2.reshape(-1) == cen).all() and (A2.reshape(-1) == A).all() and b2.size == 0
    for bad in (raw[:100], b"IxFl" + raw[4:], raw.replace(b"LTra", b"PcAm")):
        path.write_bytes(bad)
        with pytest.raises(ValueError):
            read_faiss_quantizer(str(path))
    nine_bit = raw.replace(struct.pack("<QQQ", d, M, 8), struct.pack("<QQQ", d, M, 9))
    path.write_bytes(nine_bit)
    with pytest.raises(ValueError):
        read_faiss_quantizer(str(path))


def test_decoder_builds_codec_from_quantizer_path(tmp_path):
    """--quantizer_path (transformer_lm.py:122-139; transformer.py:936-937) without faiss: same buffers as the arrays give."""
    from gnnlm_b200.formats import write_faiss_quantizer
    from gnnlm_b200.model import TransformerLanguageModel, default_args
    from gnnlm_b200.pq_codec import TorchPQCodec
    rng = np.random.RandomState(1)
    d, M = 64, 16
    cen = rng.randn(M, 256, d // M).astype(np.float32)
    A = np.linalg.qr(rng.randn(d, d))[0].astype(np.float32)
    path = str(tmp_path / "quantizer")
    write_faiss_quantizer(path, cen, A)
    args = default_args(decoder_embed_dim=d, decoder_attention_heads=4, graph_layer=1, quantizer_path=path)

    class _Dict:
        def __len__(self):
            return 50

        def pad(self):
            return 1
    m = TransformerLanguageModel.build_model(args, None, dictionary=_Dict())
    want = TorchPQCodec(centroids=cen, A=A, b=np.zeros(0, np.float32)).state_dict()
    got = m.decoder.tgt_quantizer.state_dict()
    assert sorted(got) == sorted(want) and all(torch.equal(got[k], want[k]) for k in want)


def test_quantize_features_command_line_matches_the_reference_script():
    """Flag names / defaults of knn/quantize_features.py:31-46 (+ our --batch-size)."""
    from gnnlm_b200.quantize_features import build_parser, quantizer_path
    a = build_parser().parse_args(["--data-dir", "X"])
    assert (a.prefix, a.index, a.subset, a.code_size, a.chunk_size) == ("de-en", "OPQ64_512,PQ64", "train", 64, 10000000)
    assert not (a.compute_error or a.use_gpu or a.norm or a.pretrained_quantizer)
    assert quantizer_path("D") == os.path.join("D", "quantizer") and quantizer_path("D", norm=True) == os.path.join("D", "quantizer-norm")


@pytest.mark.parametrize("opq", [True, False])
def test_convert_ckpt_injects_the_codec_buffers(tmp_path, opq):
    """gnnlm_b200.convert_ckpt == fairseq_cli/convert_ckpt.py:36-51: decoder.tgt_quantizer.{centroids_torch, norm2_centroids_torch,
    sdc_table_torch[, A, b]} from the quantizer file, args.graph = True, everything else untouched."""
    from argparse import Namespace
    from gnnlm_b200 import convert_ckpt
    from gnnlm_b200.formats import write_faiss_quantizer
    from gnnlm_b200.pq_codec import TorchPQCodec
    rng = np.random.RandomState(5)
    M, dsub = 8, 4
    cen = rng.randn(M, 256, dsub).astype(np.float32)
    A = rng.randn(M * dsub, M * dsub).astype(np.float32) if opq else None
    qfile = str(tmp_path / "quantizer")
    write_faiss_quantizer(qfile, cen, A)
    src, out = str(tmp_path / "in.pt"), str(tmp_path / "sub" / "out.pt")
    w = torch.randn(3, 3)
    torch.save({"args": Namespace(arch="transformer_lm", graph=False), "model": {"decoder.layers.0.fc1.weight": w}, "extra_state": {"epoch": 7}}, src)
    convert_ckpt.main(["--ckpt", src, "--out", out, "--quantizer", qfile], log=lambda *_: None)
    st = torch.load(out, map_location="cpu", weights_only=False)
    assert st["args"].graph is True and st["extra_state"] == {"epoch": 7} and torch.equal(st["model"]["decoder.layers.0.fc1.weight"], w)
    want = TorchPQCodec(centroids=cen, A=A, b=None if A is None else np.zeros(0, np.float32)).state_dict()
    got = {k[len("decoder.tgt_quantizer."):]: v for k, v in st["model"].items() if k.startswith("decoder.tgt_quantizer.")}
    assert sorted(got) == sorted(want) == (["A", "b"] if opq else []) + ["centroids_torch", "norm2_centroids_torch", "sdc_table_torch"]
    assert all(torch.equal(got[k], want[k]) for k in want)
