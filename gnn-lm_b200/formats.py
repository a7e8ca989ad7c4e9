"""On-disk formats either side of the hot path -- host-side readers for the files the reference's task opens
(fairseq/tasks/language_modeling.py:216-319), so that `evaluate()` runs on a reference data directory as it is:

  {split}.bin / {split}.idx      fairseq MMapIndexedDataset (fairseq/data/indexed_dataset.py:351-494)
  dict.txt                       fairseq Dictionary (fairseq/data/dictionary.py:18-60,183-228)
  {split}_dstore/info.json       {dstore_size, hidden_size, vocab_size, dstore_fp16, val_size} (knn/data_store.py:78-88)
  {split}_dstore/keys.npy        RAW memmap fp16/fp32 [N, hidden]   (not NPY despite the name; knn/data_store.py)
  {split}_dstore/vals.npy        RAW memmap int16/int32 [N, 1]
  {split}_dstore/neighbors.mmap.{k}   RAW memmap int64 [N_split, k], -1 = missing (knn/find_knn.py:45-66)
  train_dstore/quantized-keys.npy     real NPY uint8 [N_d, M] (knn/quantize_features.py:151-152; read by DeviceDatastore)
  quantizer                      faiss index file `IndexPreTransform(OPQMatrix -> IndexPQ)` or a bare `IndexPQ`
                                 (knn/quantize_features.py:92,108-109; read at transformer.py:936-937 through faiss.read_index)

Nothing here touches the GPU; arrays stay memory-mapped and `GraphTokenBlockDataset` slices them per block."""
import json
import os
import struct
from typing import Optional, Tuple, Union

import numpy as np

from .dataset import GraphTokenBlockDataset, dstore_path, feature_path, neighbor_path

_HDR_MAGIC = b"MMIDIDX\x00\x00"
# element type codes of the index header (indexed_dataset.py:83-92; code 6 is the reference's `np.float` = float64)
_DTYPES = {1: np.uint8, 2: np.int8, 3: np.int16, 4: np.int32, 5: np.int64, 6: np.float64, 7: np.float64, 8: np.uint16}


class MMapIndexedDataset:
    """Reader of fairseq's mmap token storage: `prefix.idx` = magic, version (u64 = 1), dtype code (u8), sentence count
    (u64), sizes int32[n], byte pointers int64[n]; `prefix.bin` = the sentences back to back (indexed_dataset.py:395-417)."""

    def __init__(self, path_prefix: str):
        idx, data = path_prefix + ".idx", path_prefix + ".bin"
        if not (os.path.exists(idx) and os.path.exists(data)):
            raise FileNotFoundError("Dataset not found: {}".format(path_prefix))      # language_modeling.py:232-235
        with open(idx, "rb") as f:
            if f.read(9) != _HDR_MAGIC:
                raise ValueError("Index file doesn't match expected format. Make sure that --dataset-impl is configured properly.")
            (version,) = struct.unpack("<Q", f.read(8))
            if version != 1:
                raise ValueError(f"unsupported MMapIndexedDataset index version {version}")
            (code,) = struct.unpack("<B", f.read(1))
            if code not in _DTYPES:
                raise ValueError(f"unknown element type code {code} in {idx}")
            self.dtype = np.dtype(_DTYPES[code])
            (self._len,) = struct.unpack("<Q", f.read(8))
            offset = f.tell()
        index = np.memmap(idx, mode="r", order="C")
        self.sizes = np.frombuffer(index, dtype=np.int32, count=self._len, offset=offset)
        self._pointers = np.frombuffer(index, dtype=np.int64, count=self._len, offset=offset + self.sizes.nbytes)
        self._bin = np.memmap(data, mode="r", order="C")

    def __len__(self):
        return self._len

    def __getitem__(self, i) -> np.ndarray:                     # :469-476 (always int64 to the caller)
        a = np.frombuffer(self._bin, dtype=self.dtype, count=int(self.sizes[i]), offset=int(self._pointers[i]))
        return a.astype(np.int64) if self.dtype != np.int64 else a

    def tokens(self) -> np.ndarray:
        """The flat token stream (all sentences back to back) as a zero-copy view of the .bin file.  The builder writes
        sentences contiguously (pointers are the running byte sum of the sizes, :371-380), which is checked."""
        total = int(self.sizes.astype(np.int64).sum())
        if self._len:
            expect = np.concatenate([[0], np.cumsum(self.sizes[:-1].astype(np.int64) * self.dtype.itemsize)])
            if not np.array_equal(expect, self._pointers):
                raise ValueError("non-contiguous MMapIndexedDataset: sentence pointers are not the running sum of the sizes")
        return np.frombuffer(self._bin, dtype=self.dtype, count=total, offset=0)


def write_mmap_indexed(prefix: str, sentences, dtype) -> None:
    """Writer side of the same storage -- what MMapIndexedDatasetBuilder.add_item / finalize produce
    (indexed_dataset.py:496-527 with the index layout of :357-393): `prefix.bin` = the sentences back to back in `dtype`,
    `prefix.idx` = magic, version, dtype code, count, int32 sizes, int64 byte pointers.  Byte-identical to files written by the
    reference's builder (tests/test_formats.py)."""
    code = {np.dtype(v): k for k, v in _DTYPES.items() if k != 7}[np.dtype(dtype)]
    sizes = np.array([len(s_) for s_ in sentences], dtype=np.int32)
    with open(prefix + ".bin", "wb") as f:
        for s_ in sentences:
            f.write(np.asarray(s_, dtype=dtype).tobytes(order="C"))
    pointers = np.concatenate([[0], np.cumsum(sizes[:-1].astype(np.int64) * np.dtype(dtype).itemsize)]).astype(np.int64)
    with open(prefix + ".idx", "wb") as f:
        f.write(_HDR_MAGIC + struct.pack("<Q", 1) + struct.pack("<B", code) + struct.pack("<Q", len(sizes)))
        f.write(sizes.tobytes(order="C"))
        f.write(pointers.tobytes(order="C"))


class Dictionary:
    """dict.txt reader with the reference's numbering: <s>=0, <pad>=1, </s>=2, <unk>=3, then the file's symbols in
    order (dictionary.py:18-39,183-228).  A line is '<symbol> <count>', split at the LAST space."""

    def __init__(self, pad="<pad>", eos="</s>", unk="<unk>", bos="<s>"):
        self.symbols, self.count, self.indices = [], [], {}
        self.unk_word, self.pad_word, self.eos_word = unk, pad, eos
        self.bos_index = self.add_symbol(bos)
        self.pad_index = self.add_symbol(pad)
        self.eos_index = self.add_symbol(eos)
        self.unk_index = self.add_symbol(unk)
        self.nspecial = len(self.symbols)

    def add_symbol(self, word, n=1):                            # dictionary.py:108-120
        if word in self.indices:
            idx = self.indices[word]
            self.count[idx] += n
            return idx
        idx = len(self.symbols)
        self.indices[word] = idx
        self.symbols.append(word)
        self.count.append(n)
        return idx

    @classmethod
    def load(cls, path: str) -> "Dictionary":
        d = cls()
        with open(path, "r", encoding="utf-8") as f:
            for line in f.readlines():
                i = line.rfind(" ")
                if i == -1:
                    raise ValueError("Incorrect dictionary format, expected '<token> <cnt>'")
                # the reference appends unconditionally (a duplicate symbol keeps both slots; `indices` points at the last)
                d.indices[line[:i]] = len(d.symbols)
                d.symbols.append(line[:i])
                d.count.append(int(line[i + 1:]))
        return d

    def __len__(self):
        return len(self.symbols)

    def __getitem__(self, idx):
        return self.symbols[idx] if idx < len(self.symbols) else self.unk_word

    def index(self, sym):
        return self.indices.get(sym, self.unk_index)

    def bos(self): return self.bos_index
    def pad(self): return self.pad_index
    def eos(self): return self.eos_index
    def unk(self): return self.unk_index


class MmapDataset:
    """Raw memmap with a shape and dtype (fairseq/data/mmap_dataset.py:28-58)."""

    def __init__(self, path, shape, dtype, warmup=False, verbose=False):
        need = int(np.prod(shape)) * np.dtype(dtype).itemsize
        have = os.path.getsize(path)                             # FileNotFoundError if absent, like np.memmap
        if have < need:
            raise ValueError(f"{path}: {have} bytes on disk, {need} needed for shape {tuple(shape)} of {np.dtype(dtype)}")
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self._mmap = np.memmap(path, mode="r", shape=self.shape, dtype=self.dtype)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, item):
        return self._mmap[item]

    def array(self) -> np.ndarray:
        return self._mmap


def load_graph_lm_dataset(data_path: str, split: str, *, tokens_per_sample: int, gcn_k: int = 32,
                          neighbor_context: Union[int, Tuple[int, int], str] = 1, use_precompute_feat: bool = True,
                          invalid_neighbor_context: int = 0, gcn_context_window: int = 0, intra_context: int = 0,
                          sample_break_mode: Optional[str] = "none", deprecated: bool = False,
                          knn_dists: Optional[np.ndarray] = None, knn_ids: Optional[np.ndarray] = None):
    """The `--graph` branch of LanguageModelingTask.load_dataset (language_modeling.py:216-232,265-303) over a reference
    data directory; keyword names are the task's flags.  Returns (GraphTokenBlockDataset, Dictionary).
    `neighbor_context` may be the flag's string form ("1" or "(2,0)"; the task evals it, :289)."""
    dictionary = Dictionary.load(os.path.join(data_path, "dict.txt"))
    sentences = MMapIndexedDataset(os.path.join(data_path, split))
    tokens = sentences.tokens()
    num_tokens = int(tokens.shape[0])
    neighbor_info = json.load(open(os.path.join(dstore_path(data_path, "train"), "info.json")))
    info = json.load(open(os.path.join(dstore_path(data_path, split), "info.json")))
    if isinstance(neighbor_context, str):
        import ast
        neighbor_context = ast.literal_eval(neighbor_context)
    neighbors = MmapDataset(neighbor_path(data_path, split, gcn_k), shape=(num_tokens, gcn_k), dtype=np.int64)
    feats = None
    if use_precompute_feat:                                       # :291-294
        feats = MmapDataset(feature_path(data_path, split), shape=(num_tokens, info["hidden_size"]),
                            dtype=np.float16 if info["dstore_fp16"] else np.float32).array()
    ds = GraphTokenBlockDataset(
        tokens, tokens_per_sample, pad=dictionary.pad(), eos=dictionary.eos(), neighbor_offsets=neighbors.array(),
        n_datastore=int(neighbor_info["dstore_size"]), neighbor_context=neighbor_context, precompute_feats=feats,
        invalid_neighbor_context=invalid_neighbor_context if split == "train" else 0,      # :295
        context_window=gcn_context_window, intra_context=intra_context, knn_dists=knn_dists, knn_ids=knn_ids,
        break_mode=sample_break_mode, deprecated=deprecated, sizes=np.asarray(sentences.sizes))
    return ds, dictionary


# ----------------------------------------------------------------------------------------------------------------------
# faiss `quantizer` file (transformer.py:936-937 + knn/pq_wrapper.py:20-37 need exactly three arrays out of it)
# ----------------------------------------------------------------------------------------------------------------------
# faiss (>= 1.5.3, README.md:46) is a third-party dependency that is NOT under /root/reference and is not installed here: the layout
# below restates faiss/impl/index_write.cpp (write_index_header, write_VectorTransform's generic-LinearTransform branch that an
# OPQMatrix takes, write_ProductQuantizer, the IndexPQ branch of write_index) as published for 1.5.3 - 1.7.x.  No faiss-written
# file is available offline, so this reader is "parity unpinned": it checks every fourcc and every redundant size the format
# carries and raises ValueError on the first mismatch instead of guessing.
def _fourcc(tag: str) -> int:
    return struct.unpack("<I", tag.encode("ascii"))[0]


class _Reader:
    def __init__(self, f, path):
        self.f, self.path = f, path

    def take(self, fmt: str):
        size = struct.calcsize(fmt)
        raw = self.f.read(size)
        if len(raw) != size:
            raise ValueError(f"{self.path}: truncated faiss index file")
        return struct.unpack(fmt, raw)[0]

    def vector(self, dtype) -> np.ndarray:
        n = self.take("<Q")
        raw = self.f.read(n * np.dtype(dtype).itemsize)
        if len(raw) != n * np.dtype(dtype).itemsize:
            raise ValueError(f"{self.path}: truncated faiss index file (vector of {n} x {np.dtype(dtype).name})")
        return np.frombuffer(raw, dtype=dtype).copy()

    def header(self) -> dict:
        """write_index_header: d (int32), ntotal (int64), two dummies (int64 = 1 << 20), is_trained (bool), metric_type (int32)
        [, metric_arg (float32) when metric_type > 1]."""
        h = {"d": self.take("<i"), "ntotal": self.take("<q")}
        self.take("<q"), self.take("<q")
        h["is_trained"] = bool(self.take("<B"))
        h["metric_type"] = self.take("<i")
        if h["metric_type"] > 1:
            h["metric_arg"] = self.take("<f")
        return h


def read_faiss_quantizer(path: str) -> Tuple[np.ndarray, Optional[np.ndarray], Optional[np.ndarray]]:
    """(centroids [M, 256, dsub] fp32, A [d_out, d_in] or None, b [d_out] / empty or None) -- what NumpyPQCodec / TorchPQCodec
    extract from `faiss.read_index(path)` (knn/pq_wrapper.py:20-37), without faiss."""
    with open(path, "rb") as f:
        r = _Reader(f, path)
        tag = r.take("<I")
        A = b = None
        if tag == _fourcc("IxPT"):
            outer = r.header()
            if not outer["is_trained"]:
                raise ValueError(f"{path}: index is not trained (pq_wrapper.py:15)")
            nt = r.take("<i")
            if nt < 1:
                raise ValueError(f"{path}: IndexPreTransform without a transform")
            for i in range(nt):
                vt = r.take("<I")
                if vt != _fourcc("LTra"):
                    raise ValueError(f"{path}: transform {i} is not a plain LinearTransform / OPQMatrix (pq_wrapper.py:21-22)")
                have_bias = bool(r.take("<B"))
                A_i, b_i = r.vector(np.float32), r.vector(np.float32)
                d_in, d_out, trained = r.take("<i"), r.take("<i"), bool(r.take("<B"))
                if A_i.size != d_in * d_out or (have_bias and b_i.size != d_out) or not trained:
                    raise ValueError(f"{path}: inconsistent LinearTransform ({A_i.size} coefficients for {d_out} x {d_in})")
                if i == 0:                                        # index.chain.at(0) is the only transform the codec reads (:21)
                    A, b = A_i.reshape(d_out, d_in), b_i
            if nt != 1:
                raise ValueError(f"{path}: {nt} chained transforms; the reference codec applies only the first (pq_wrapper.py:21)")
            tag = r.take("<I")
        if tag != _fourcc("IxPq"):
            raise ValueError(f"{path}: expected an IndexPQ (pq_wrapper.py:32), found {struct.pack('<I', tag)!r}")
        inner = r.header()
        d, M, nbits = r.take("<Q"), r.take("<Q"), r.take("<Q")
        cen = r.vector(np.float32)
        if nbits != 8:
            raise ValueError(f"{path}: {nbits}-bit PQ; 8-bit expected (pq_wrapper.py:36)")
        if M == 0 or d % M or cen.size != d * 256 or d != inner["d"] or (A is not None and A.shape[0] != d):
            raise ValueError(f"{path}: inconsistent ProductQuantizer (d={d}, M={M}, {cen.size} centroid values)")
        return cen.reshape(M, 256, d // M), A, b


def write_faiss_quantizer(path: str, centroids: np.ndarray, A: Optional[np.ndarray] = None, b: Optional[np.ndarray] = None,
                          metric_type: int = 1) -> None:
    """The same layout written back (empty code store): synthetic data directories and the reader's round-trip test."""
    cen = np.ascontiguousarray(centroids, dtype=np.float32)
    M, ksub, dsub = cen.shape
    assert ksub == 256
    d = M * dsub

    def header(f, dim):
        f.write(struct.pack("<iqqqBi", dim, 0, 1 << 20, 1 << 20, 1, metric_type))

    def vector(f, a, dtype):
        a = np.ascontiguousarray(a, dtype=dtype).reshape(-1)
        f.write(struct.pack("<Q", a.size))
        f.write(a.tobytes())

    with open(path, "wb") as f:
        if A is not None:
            A = np.ascontiguousarray(A, dtype=np.float32)
            assert A.shape[0] == d
            bias = np.zeros(0, np.float32) if b is None else np.asarray(b, np.float32)
            f.write(struct.pack("<I", _fourcc("IxPT")))
            header(f, A.shape[1])
            f.write(struct.pack("<i", 1))
            f.write(struct.pack("<IB", _fourcc("LTra"), int(bias.size > 0)))
            vector(f, A, np.float32)
            vector(f, bias, np.float32)
            f.write(struct.pack("<iiB", A.shape[1], A.shape[0], 1))
        f.write(struct.pack("<I", _fourcc("IxPq")))
        header(f, d)
        f.write(struct.pack("<QQQ", d, M, 8))
        vector(f, cen, np.float32)
        vector(f, np.zeros(0, np.uint8), np.uint8)
        f.write(struct.pack("<iBi", 0, 0, 0))                     # search_type, encode_signs, polysemous_ht
