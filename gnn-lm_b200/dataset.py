"""Block dataset + batch assembly -- host-side mirror of the reference's data path for the graph LM:

  GraphTokenBlockDataset   fairseq/data/token_block_dataset.py:172-333 (block slicing :246-285,
                           source/target shift :304-319, neighbour-id gather :309, features :327-329)
  GraphMonolingualDataset  fairseq/data/monolingual_dataset.py:209-266 (collater, no shuffling)
  on-disk layout           knn/path_utils.py:13-41, fairseq/tasks/language_modeling.py:265-303

What changes: the reference builds a DGL graph per block in Python inside DataLoader workers and
ships codes/edges over PCIe; here __getitem__ only slices the *inputs* of graph assembly (neighbour
ids, fp16 features, tokens -- contiguous slices in `none` break mode) and the graph is assembled on
the device from the HBM-resident datastore (`DeviceDatastore`) by graph_build.cu / pq_decode.cu.
Every `--sample-break-mode` is sliced (`none` for the wiki103 / enwik8 scripts, `eos` for one_billion); blocks of
different lengths are never padded into one batch (the reference's graph decoder cannot take padded batches either,
transformer.py:975) -- eval_lm.batches groups equal lengths, across the whole shard when asked to."""
import json
import os
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from .graph import TokenGraph, build_token_graph


# ---- knn/path_utils.py:13-41
def feature_path(data_dir, mode): return os.path.join(data_dir, f"{mode}_dstore", "keys.npy")
def value_path(data_dir, mode): return os.path.join(data_dir, f"{mode}_dstore", "vals.npy")
def dstore_path(data_dir, subset): return os.path.join(data_dir, f"{subset}_dstore")
def quantized_feature_path(data_dir, mode): return os.path.join(data_dir, f"{mode}_dstore", "quantized-keys.npy")
def neighbor_path(data_dir, mode, k=32): return os.path.join(data_dir, f"{mode}_dstore", f"neighbors.mmap.{k}")


class DeviceDatastore:
    """The quantised datastore replicated in one GPU's HBM: codes [N_d, M] uint8
    (train_dstore/quantized-keys.npy) and values [N_d] int16/int32 (train_dstore/vals.npy)."""

    def __init__(self, codes: Optional[torch.Tensor], vals: torch.Tensor):
        """codes=None is `--reinit-nfeat` (language_modeling.py:273-278: no quantized features are loaded; the ntgt features are
        embeddings of the datastore VALUES)."""
        self.vals = vals.reshape(-1).contiguous()
        assert self.vals.dtype in (torch.int16, torch.int32)
        if codes is not None:
            assert codes.dtype == torch.uint8 and codes.dim() == 2 and self.vals.numel() == codes.shape[0]
            codes = codes.contiguous()
        self.codes = codes
        self.size = self.vals.numel()

    @classmethod
    def from_dir(cls, data_dir: str, vocab_size: int, device, reinit_nfeat: bool = False) -> "DeviceDatastore":
        info = json.load(open(os.path.join(dstore_path(data_dir, "train"), "info.json")))
        n = info["dstore_size"]
        vdt = np.int16 if info.get("dstore_fp16") and vocab_size < 2 ** 15 else np.int32   # language_modeling.py:272
        vals = np.memmap(value_path(data_dir, "train"), dtype=vdt, mode="r", shape=(n, 1))
        codes = None
        if not reinit_nfeat:
            codes = torch.from_numpy(np.array(np.load(quantized_feature_path(data_dir, "train"), mmap_mode="r"))).to(device)
        return cls(codes, torch.from_numpy(np.array(vals)).to(device))


def get_slice_indices(sizes, break_mode: Optional[str], block_size: int, document_sep_len: int = 1) -> np.ndarray:
    """Block boundaries [n_blocks, 2] over the flat token stream (token_block_utils_fast.pyx:22-105):
    none = fixed block_size windows; eos = one sentence per block; complete = whole sentences packed greedily up to
    block_size (an oversize sentence is its own block); complete_doc = the same inside documents, where a sentence of
    document_sep_len tokens separates documents and blocks of <= 1 token are dropped."""
    sizes = np.asarray(sizes, dtype=np.int64).reshape(-1)
    cum = np.concatenate([np.zeros(1, np.int64), np.cumsum(sizes)])
    total = int(cum[-1])
    if break_mode is None or break_mode == "none":
        starts = np.arange(0, total, block_size, dtype=np.int64)
        return np.stack([starts, np.minimum(starts + block_size, total)], 1)
    if break_mode == "eos":
        return np.stack([cum[:-1], cum[1:]], 1)
    if break_mode not in ("complete", "complete_doc"):
        raise ValueError("Invalid break_mode: " + str(break_mode))
    n = len(sizes)
    if break_mode == "complete":
        docs, keep = [(0, n)], 0
    else:
        seps = np.flatnonzero(sizes == document_sep_len)
        edges = np.concatenate([[-1], seps, [n]])
        docs, keep = [(int(a) + 1, int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a + 1], 1
    out = []
    for a, b in docs:
        i = a
        while i < b:
            j = int(np.searchsorted(cum, cum[i] + block_size, side="right")) - 1      # last boundary within block_size
            j = min(max(j, i + 1), b)
            if cum[j] - cum[i] > keep:
                out.append((int(cum[i]), int(cum[j])))
            i = j
    return np.asarray(out, dtype=np.int64).reshape(-1, 2)


class GraphTokenBlockDataset:
    """token_block_dataset.py:172-333 over a flat token stream.  `sizes` (sentence lengths, summing to len(tokens)) is
    needed by every break mode except `none`."""

    def __init__(self, tokens: np.ndarray, block_size: int, pad: int, eos: int, neighbor_offsets: np.ndarray,
                 n_datastore: int, neighbor_context: Union[int, Tuple[int, int]] = 1,
                 precompute_feats: Optional[np.ndarray] = None, invalid_neighbor_context: int = 0,
                 context_window: int = 0, intra_context: int = 0, knn_dists: Optional[np.ndarray] = None,
                 knn_ids: Optional[np.ndarray] = None, break_mode: str = "none", deprecated: bool = False,
                 sizes: Optional[np.ndarray] = None, document_sep_len: int = 1):
        if break_mode not in (None, "none") and sizes is None:
            raise ValueError(f"--sample-break-mode {break_mode} needs the sentence lengths (sizes=...)")
        self.deprecated = deprecated                # --deprecated: de-duplicating builder (token_block_dataset.py:305,414-479)
        self.tokens = tokens
        self.block_size, self.pad, self.eos = block_size, pad, eos
        self.neighbor_offsets = neighbor_offsets
        self.n_datastore = n_datastore
        if isinstance(neighbor_context, int):                           # :232-236
            self.left_neighbor_context = self.right_neighbor_context = neighbor_context
        else:
            self.left_neighbor_context, self.right_neighbor_context = neighbor_context
        self.precompute_feats = precompute_feats
        self.invalid_neighbor_context = invalid_neighbor_context
        self.context_window = context_window
        self.max_intra_context = intra_context
        self.knn_dists, self.knn_ids = knn_dists, knn_ids
        self.host_copy_threads = 4              # native threads of one batch-slice copy (collate_into)
        n = len(tokens)
        if sizes is not None and int(np.sum(sizes)) != n:
            raise ValueError("sizes must sum to the number of tokens")
        sl = get_slice_indices(np.array([n]) if sizes is None else sizes, break_mode, block_size, document_sep_len)
        self.break_mode = break_mode or "none"
        self.slice_indices = [(int(s), int(e)) for s, e in sl]
        self.sizes = np.array([e - s for s, e in self.slice_indices])

    def __len__(self):
        return len(self.slice_indices)

    def __getitem__(self, index):
        s, e = self.slice_indices[index]
        cs = s if (self.context_window == 0 or index == 0) else max(0, s - self.context_window)    # :246-285
        item = torch.from_numpy(np.asarray(self.tokens[cs:e]).astype(np.int64))
        if cs == 0:                                                                                 # :310-319
            source = torch.cat([item.new_tensor([self.eos]), item[:-1]])
        else:
            source = torch.from_numpy(np.asarray(self.tokens[cs - 1:e - 1]).astype(np.int64))
        nbr = np.array(self.neighbor_offsets[cs:e])
        # the files are raw memmaps without a header: a neighbors.mmap built against another datastore would index past
        # the code table.  The reference raises IndexError at `quant_neighbor_feats[o]` (token_block_dataset.py:369-370)
        if nbr.size and (int(nbr.min()) < -1 or int(nbr.max()) >= self.n_datastore):
            raise IndexError(f"block {index}: neighbour id outside [-1, {self.n_datastore}) -- neighbors.mmap does not "
                             "belong to this train_dstore")
        out = {"id": index, "source": source, "target": item, "offsets": (cs, e), "start_idx": s - cs,
               "nbr": torch.from_numpy(nbr)}
        if self.precompute_feats is not None:
            out["feats"] = torch.from_numpy(np.array(self.precompute_feats[cs:e]))
        if self.knn_ids is not None:
            out["knn_dists"] = torch.from_numpy(np.array(self.knn_dists[cs:e]))
            ids = np.array(self.knn_ids[cs:e])
            if ids.size and (int(ids.min()) < -1 or int(ids.max()) >= self.n_datastore):
                raise IndexError(f"block {index}: kNN-LM neighbour id outside [-1, {self.n_datastore})")
            out["knn_ids"] = torch.from_numpy(ids)
        return out

    def collater(self, samples: List[dict]) -> dict:
        """monolingual_dataset.py:13-53,237-262 for equal-length blocks (the reference's own
        `x.view(bsz, tgt_len, -1)` assumes this, transformer.py:975 'todo: fix padding cases');
        a ragged last block must be its own batch."""
        if not samples:
            return {}
        Ls = {len(s["target"]) for s in samples}
        assert len(Ls) == 1, "batch blocks of equal length only (SURVEY.md Q6)"
        pin = lambda t: t.pin_memory() if torch.cuda.is_available() else t
        st = lambda key: pin(torch.stack([s[key] for s in samples]))
        batch = {
            "id": torch.LongTensor([s["id"] for s in samples]),
            "nsentences": len(samples),
            "ntokens": sum(len(s["source"]) for s in samples),
            "net_input": {"src_tokens": st("source"),
                          "src_lengths": torch.LongTensor([s["source"].numel() for s in samples])},
            "target": st("target"),
            "start_indices": torch.LongTensor([[s["start_idx"]] for s in samples]),
            "nbr": st("nbr"),
            "positions": pin(torch.stack([torch.arange(s["offsets"][0], s["offsets"][1]) for s in samples])),
        }
        for k in ("feats", "knn_dists", "knn_ids"):
            if k in samples[0]:
                batch[k] = st(k)
        return batch

    # ---- fast host path of evaluate(): block slices copied ONCE, straight from the memmaps into reusable pinned batch buffers
    def batch_spec(self, ids: List[int]):
        """(name, shape, dtype) of every tensor the device step consumes for the equal-length blocks `ids` (see collate_into)."""
        s, e = self.slice_indices[ids[0]]
        cs = s if (self.context_window == 0 or ids[0] == 0) else max(0, s - self.context_window)
        B, Lb = len(ids), e - cs
        spec = [("nbr", (B, Lb, self.neighbor_offsets.shape[1]), torch.int64), ("positions", (B, Lb), torch.int64),
                ("src_tokens", (B, Lb), torch.int64), ("target", (B, Lb), torch.int64), ("start_indices", (B,), torch.int32)]
        if self.precompute_feats is not None:
            spec.append(("feats", (B, Lb, self.precompute_feats.shape[1]), torch.from_numpy(np.zeros(0, self.precompute_feats.dtype)).dtype))
        if self.knn_ids is not None:
            spec.append(("knn_dists", (B, Lb, self.knn_dists.shape[1]), torch.float32))
            spec.append(("knn_ids", (B, Lb, self.knn_ids.shape[1]), torch.int64))
        return spec

    def collate_into(self, ids: List[int], out: dict) -> dict:
        """collater([self[i] for i in ids]) without the intermediate copies: the same tensors (eval_lm.host_inputs names),
        written into the preallocated (pinned) buffers `out` -- one memcpy per array from the page cache.  Same id checks as
        __getitem__ (range check of the int64 ids fused into the copy)."""
        import ctypes
        from . import _lib as L
        lib = L.load()
        n_tok = 0
        bad = ctypes.c_int32(0)

        def take(name, b, src, cs, e, ids_check=False):
            """slice [cs, e) of a per-token array -> the batch buffer: one native multi-threaded copy (gnnlm_host_copy), int64 id
            arrays range-checked in the same pass"""
            sl = src[cs:e]
            dst = out[name][b]
            if sl.flags["C_CONTIGUOUS"] and dst.is_contiguous() and dst.numel() * dst.element_size() == sl.nbytes:
                rc = lib.gnnlm_host_copy(dst.data_ptr(), sl.ctypes.data, sl.nbytes, int(ids_check), -1, self.n_datastore,
                                         self.host_copy_threads, ctypes.byref(bad))
                if rc != 0:
                    raise L.GnnlmError(lib.gnnlm_last_error().decode())
            else:                                                   # broadcast / strided sources
                np.copyto(dst.numpy(), sl)
                if ids_check and sl.size and (int(sl.min()) < -1 or int(sl.max()) >= self.n_datastore):
                    bad.value = 1
            if bad.value:
                raise IndexError(f"block {ids[b]}: neighbour id outside [-1, {self.n_datastore}) -- "
                                 f"{'neighbors.mmap' if name == 'nbr' else 'the kNN-LM neighbour file'} does not belong to this "
                                 "train_dstore")

        for b, i in enumerate(ids):
            s, e = self.slice_indices[i]
            cs = s if (self.context_window == 0 or i == 0) else max(0, s - self.context_window)
            take("nbr", b, self.neighbor_offsets, cs, e, ids_check=True)
            tgt = out["target"][b].numpy()
            np.copyto(tgt, self.tokens[cs:e], casting="unsafe")
            src = out["src_tokens"][b].numpy()
            if cs == 0:
                src[0] = self.eos
                src[1:] = tgt[:-1]
            else:
                np.copyto(src, self.tokens[cs - 1:e - 1], casting="unsafe")
            out["positions"][b].copy_(torch.arange(cs, e))
            out["start_indices"][b] = s - cs
            if self.precompute_feats is not None:
                take("feats", b, self.precompute_feats, cs, e)
            if self.knn_ids is not None:
                take("knn_dists", b, self.knn_dists, cs, e)
                take("knn_ids", b, self.knn_ids, cs, e, ids_check=True)
            n_tok += e - cs
        return {"ids": list(ids), "nsentences": len(ids), "ntokens": n_tok, "host": out}

    def ordered_indices(self):
        return np.arange(len(self))                                       # monolingual_dataset.py:264-266


def sample_from_inputs(inp: dict, dataset: GraphTokenBlockDataset, dstore: DeviceDatastore, reach=None) -> dict:
    """Device tensors of one batch -> the fairseq `sample` dict with net_input.graph = TokenGraph.  Device work only (no
    host synchronisation), so the whole step can be captured in a CUDA graph (eval_lm.GraphedScorer)."""
    pos = inp["positions"]
    g = build_token_graph(inp["nbr"], dstore.size, dataset.left_neighbor_context, dataset.right_neighbor_context,
                          tgt_pos=pos if dataset.invalid_neighbor_context > 0 else None,
                          invalid_ctx=dataset.invalid_neighbor_context, intra_ctx=dataset.max_intra_context, reach=reach,
                          dedup=getattr(dataset, "deprecated", False))
    g.codes_table = dstore.codes
    g.labels_table = dstore.vals
    if "feats" in inp:
        g.nodes["tgt"].data["h"] = inp["feats"].view(-1, inp["feats"].shape[-1])
    target = inp["target"]
    sample = {"ntokens": target.numel(), "nsentences": target.shape[0],
              "net_input": {"src_tokens": inp["src_tokens"], "graph": g},
              "target": target, "start_indices": inp["start_indices"], "positions": pos}
    if "knn_ids" in inp:
        sample["knn_dists"] = inp["knn_dists"].view(-1, inp["knn_dists"].shape[-1])
        sample["knn_ids"] = inp["knn_ids"].view(-1, inp["knn_ids"].shape[-1])
    return sample


def move_to_cuda(batch: dict, dataset: GraphTokenBlockDataset, dstore: DeviceDatastore, device="cuda") -> dict:
    """utils.move_to_cuda (fairseq/utils.py:43-67) + on-device graph assembly.  Returns the fairseq
    `sample` dict with net_input.graph = TokenGraph."""
    nb = lambda t: t.to(device, non_blocking=True)
    inp = {"nbr": nb(batch["nbr"]), "positions": nb(batch["positions"]), "src_tokens": nb(batch["net_input"]["src_tokens"]),
           "target": nb(batch["target"]), "start_indices": batch["start_indices"]}
    for k in ("feats", "knn_dists", "knn_ids"):
        if k in batch:
            inp[k] = nb(batch[k])
    sample = sample_from_inputs(inp, dataset, dstore)
    sample.update({"id": batch["id"], "nsentences": batch["nsentences"], "ntokens": batch["ntokens"]})
    sample["net_input"]["src_lengths"] = batch["net_input"]["src_lengths"]
    return sample
