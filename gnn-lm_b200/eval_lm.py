"""Evaluation loop -- mirror of fairseq_cli/eval_lm.py:208-331 for the graph LM (+ kNN-LM).

Differences that do not change results: the per-hypothesis `score_sum += pos_scores.sum().cpu()`
(one device->host sync per sample, eval_lm.py:273) becomes an fp64 accumulation on the device inside
the scoring kernel with a single read-back at the end; blocks are sharded over ranks as contiguous
ranges and the two scalars {sum log p, n_tokens} are all-reduced once (the reference's sharded
eval never combines its shards)."""
import math
import time
from typing import Optional

import torch

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


@torch.no_grad()
def evaluate(model, dataset: GraphTokenBlockDataset, dstore: DeviceDatastore, scorer: SequenceScorer, *,
             knn_dstore=None, temperature: float = 1.0, max_sentences: int = 1, device="cuda", rank: int = 0,
             world_size: int = 1, process_group=None, log=None) -> dict:
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
        scorer.score_tokens(model, sample, knn_dstore, temperature, nll_acc=acc)
        ntok += sample["ntokens"]
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
