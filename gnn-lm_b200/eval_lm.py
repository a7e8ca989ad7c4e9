"""Evaluation loop -- mirror of fairseq_cli/eval_lm.py:208-331 for the graph LM (+ kNN-LM).

Differences that do not change results: the per-hypothesis `score_sum += pos_scores.sum().cpu()`
(one device->host sync per sample, eval_lm.py:273) becomes an fp64 accumulation on the device inside
the scoring kernel with a single read-back at the end; blocks are sharded over ranks as contiguous
ranges and the two scalars {sum log p, n_tokens} are all-reduced once (the reference's sharded
eval never combines its shards)."""
import json
import math
import os
import time
from typing import Optional

import numpy as np
import torch

from . import ops

from .dataset import DeviceDatastore, GraphTokenBlockDataset, move_to_cuda
from .sequence_scorer import SequenceScorer


def shard_range(n_blocks: int, rank: int, world: int):
    """Contiguous block range of `rank`: [ceil(nb*r/R), ceil(nb*(r+1)/R))."""
    lo = -(-n_blocks * rank // world)
    hi = -(-n_blocks * (rank + 1) // world)
    return lo, hi


def batches(dataset: GraphTokenBlockDataset, lo: int, hi: int, max_sentences: int):
    """Equal-length blocks batched together in order; a ragged last block is its own batch."""
    cur, cur_len = [], None
    for i in range(lo, hi):
        n = int(dataset.sizes[i]) + (0 if (dataset.context_window == 0 or i == 0) else
                                     min(dataset.context_window, dataset.slice_indices[i][0]))
        if cur and (n != cur_len or len(cur) >= max_sentences):
            yield cur
            cur = []
        cur.append(i)
        cur_len = n
    if cur:
        yield cur


class DstoreWriter:
    """--save-knnlm-dstore (fairseq_cli/eval_lm.py:178-244): keys.npy / vals.npy raw memmaps + info.json, same names,
    dtypes and shapes; keys are the features selected by --knn-keytype (e.g. `gcn_feat`), written in token order.
    The fp32 -> fp16 cast runs on the device (gnnlm_convert) before the D2H copy."""

    def __init__(self, dstore_mmap: str, subset: str, dstore_size: int, hidden: int, vocab: int, dstore_fp16: bool,
                 knn_keytype: Optional[str] = None):
        suffix = "" if not knn_keytype else f"-{knn_keytype}"
        self.dir = os.path.join(dstore_mmap, f"{subset}_dstore{suffix}")
        os.makedirs(self.dir, exist_ok=True)
        self.fp16, self.size, self.idx = dstore_fp16, int(dstore_size), 0
        info = {"dstore_size": int(dstore_size), "hidden_size": hidden, "vocab_size": vocab, "dstore_fp16": dstore_fp16,
                "val_size": 1}
        json.dump(info, open(os.path.join(self.dir, "info.json"), "w"), indent=4, sort_keys=True)
        self.keys = np.memmap(os.path.join(self.dir, "keys.npy"), dtype=np.float16 if dstore_fp16 else np.float32, mode="w+",
                              shape=(self.size, hidden))
        vdt = np.int16 if dstore_fp16 and vocab < 2 ** 15 else np.int32
        self.vals = np.memmap(os.path.join(self.dir, "vals.npy"), dtype=vdt, mode="w+", shape=(self.size, 1))

    def add(self, feats: torch.Tensor, tokens: torch.Tensor):
        """feats [n, d] fp32 on the device, tokens [n]."""
        n = min(feats.shape[0], self.size - self.idx)            # eval_lm.py:227-230 (clip at dstore_size)
        if n <= 0:
            return
        f = feats[:n].contiguous()
        if self.fp16:
            f = ops.convert(f, torch.float16)
        self.keys[self.idx:self.idx + n] = f.cpu().numpy()
        self.vals[self.idx:self.idx + n] = tokens[:n].view(-1, 1).cpu().numpy().astype(self.vals.dtype)
        self.idx += n

    def close(self):
        self.keys.flush()
        self.vals.flush()
        return self.idx


@torch.no_grad()
def evaluate(model, dataset: GraphTokenBlockDataset, dstore: DeviceDatastore, scorer: SequenceScorer, *,
             knn_dstore=None, temperature: float = 1.0, max_sentences: int = 1, device="cuda", rank: int = 0,
             world_size: int = 1, process_group=None, log=None, dstore_writer: Optional[DstoreWriter] = None,
             knn_keytype: Optional[str] = None) -> dict:
    lo, hi = shard_range(len(dataset), rank, world_size)
    acc = torch.zeros(2, dtype=torch.float64, device=device)
    ntok = 0
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for ids in batches(dataset, lo, hi, max_sentences):
        batch = dataset.collater([dataset[i] for i in ids])
        sample = move_to_cuda(batch, dataset, dstore, device)
        if knn_dstore is not None and "knn_ids" in sample:
            knn_dstore.set_search_results(sample["knn_dists"], sample["knn_ids"])
        _, _, _, dec_out = scorer.score_tokens(model, sample, knn_dstore, temperature, nll_acc=acc)
        ntok += sample["ntokens"]
        if dstore_writer is not None:                           # sequence_scorer.py:180-183 + eval_lm.py:223-244
            extra = dec_out[1]
            feat = extra[knn_keytype] if knn_keytype in extra else extra["inner_states"][-1]     # [L, B, d]
            starts = sample["start_indices"].view(-1).tolist()
            for i in range(sample["target"].shape[0]):
                mask = sample["target"][i, starts[i]:].ne(scorer.pad)
                dstore_writer.add(feat[starts[i]:, i, :][mask].float(), sample["target"][i, starts[i]:][mask])
    if world_size > 1:
        torch.distributed.all_reduce(acc, group=process_group)          # the path's only collective
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    score_sum, count = acc.tolist()
    avg_nll2 = -score_sum / count / math.log(2) if count else float("nan")     # eval_lm.py:325
    res = {"score_sum": score_sum, "count": count, "loss_base2": avg_nll2, "ppl": 2 ** avg_nll2,
           "tokens_this_rank": ntok, "seconds": dt, "tokens_per_s": ntok / dt if dt > 0 else float("nan")}
    if log:
        log("Evaluated {} tokens in {:.1f}s ({:.2f} tokens/s)".format(ntok, dt, res["tokens_per_s"]))
        log("Loss (base 2): {:.4f}, Perplexity: {:.2f}".format(avg_nll2, res["ppl"]))
    return res
