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

from .dataset import DeviceDatastore, GraphTokenBlockDataset, move_to_cuda, sample_from_inputs
from .sequence_scorer import SequenceScorer


def shard_range(n_blocks: int, rank: int, world: int):
    """Contiguous block range of `rank`: [ceil(nb*r/R), ceil(nb*(r+1)/R))."""
    lo = -(-n_blocks * rank // world)
    hi = -(-n_blocks * (rank + 1) // world)
    return lo, hi


def batches(dataset: GraphTokenBlockDataset, lo: int, hi: int, max_sentences: int, max_tokens: Optional[int] = None,
            bucket_by_length: bool = False):
    """Batches of equal-length blocks (never padded), at most max_sentences blocks / max_tokens tokens each.
    In order (consecutive blocks; a ragged last block is its own batch), or -- bucket_by_length, for sentence-per-block
    corpora (--sample-break-mode eos) where consecutive lengths rarely match -- grouped by length over the whole shard,
    shortest first (the total score does not depend on the order)."""
    def length(i):
        return int(dataset.sizes[i]) + (0 if (dataset.context_window == 0 or i == 0) else
                                         min(dataset.context_window, dataset.slice_indices[i][0]))

    def full(cur, n):
        return len(cur) >= max_sentences or (max_tokens is not None and (len(cur) + 1) * n > max_tokens)

    order = range(lo, hi)
    if bucket_by_length:
        order = sorted(order, key=lambda i: (length(i), i))
    cur, cur_len = [], None
    for i in order:
        n = length(i)
        if cur and (n != cur_len or full(cur, n)):
            yield cur
            cur = []
        cur.append(i)
        cur_len = n
    if cur:
        yield cur


class DstoreWriter:
    """--save-knnlm-dstore (fairseq_cli/eval_lm.py:178-244): keys.npy / vals.npy raw memmaps + info.json, same names,
    dtypes and shapes; keys are the features selected by --knn-keytype (e.g. `gcn_feat`), written in token order.
    The fp32 -> fp16 cast runs on the device (gnnlm_convert) before the D2H copy.

    Shard-aware: `offset` is the global token offset of this rank's first block (sum of the sizes of the blocks before its
    contiguous range), so every rank / shard writes its own row range [offset, offset + its tokens) of the SAME files.  The
    files are sized with ftruncate (extends with zeros, never clears rows another shard already wrote) and mapped `r+`, so
    concurrent ranks and separately launched `--shard-id` runs compose into one datastore."""

    def __init__(self, dstore_mmap: str, subset: str, dstore_size: int, hidden: int, vocab: int, dstore_fp16: bool,
                 knn_keytype: Optional[str] = None, offset: int = 0, limit: Optional[int] = None, write_info: bool = True):
        suffix = "" if not knn_keytype else f"-{knn_keytype}"
        self.dir = os.path.join(dstore_mmap, f"{subset}_dstore{suffix}")
        os.makedirs(self.dir, exist_ok=True)
        self.fp16, self.size = dstore_fp16, int(dstore_size)
        self.start = self.idx = min(int(offset), self.size)
        self.end = self.size if limit is None else min(self.size, self.start + int(limit))
        info = {"dstore_size": int(dstore_size), "hidden_size": hidden, "vocab_size": vocab, "dstore_fp16": dstore_fp16,
                "val_size": 1}
        if write_info:
            json.dump(info, open(os.path.join(self.dir, "info.json"), "w"), indent=4, sort_keys=True)
        kdt = np.float16 if dstore_fp16 else np.float32
        vdt = np.int16 if dstore_fp16 and vocab < 2 ** 15 else np.int32
        self.keys = self._map(os.path.join(self.dir, "keys.npy"), kdt, (self.size, hidden))
        self.vals = self._map(os.path.join(self.dir, "vals.npy"), vdt, (self.size, 1))

    @staticmethod
    def _map(path, dtype, shape):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        with open(path, "ab") as f:
            f.truncate(nbytes)
        if nbytes == 0:
            return np.zeros(shape, dtype)
        return np.memmap(path, dtype=dtype, mode="r+", shape=shape)

    def add(self, feats: torch.Tensor, tokens: torch.Tensor):
        """feats [n, d] fp32 on the device, tokens [n]."""
        n = min(feats.shape[0], self.end - self.idx)             # eval_lm.py:227-230 (clip at dstore_size)
        if n <= 0:
            return
        f = feats[:n].contiguous()
        if self.fp16:
            f = ops.convert(f, torch.float16)
        self.keys[self.idx:self.idx + n] = f.cpu().numpy()
        self.vals[self.idx:self.idx + n] = tokens[:n].view(-1, 1).cpu().numpy().astype(self.vals.dtype)
        self.idx += n

    def close(self):
        """Rows written by THIS rank / shard."""
        if hasattr(self.keys, "flush"):
            self.keys.flush()
            self.vals.flush()
        return self.idx - self.start


class GraphedScorer:
    """One whole scoring step (graph assembly -> PQ decode -> HGT -> log-probs -> kNN mix -> NLL) captured as a CUDA
    graph per input-shape signature and replayed for every later batch of that shape.  Possible because the path has
    no host synchronisation: node / edge counts stay on the device and every buffer is sized by its host-known
    capacity.  Inputs are copied into the graph's static buffers on the current stream before each replay; outputs
    are the capture's own tensors (valid until the next replay of the same signature)."""

    def __init__(self, fn, warmup: int = 2, device=None):
        self.fn, self.warmup, self.cache, self.device = fn, warmup, {}, device

    @staticmethod
    def signature(inputs: dict):
        return tuple((k, tuple(v.shape), str(v.dtype)) for k, v in sorted(inputs.items()))

    def __call__(self, inputs: dict):
        key = self.signature(inputs)
        ent = self.cache.get(key)
        if ent is None:
            static = {k: torch.empty_like(v, device=self.device or v.device) for k, v in inputs.items()}   # inputs may be pinned host
            for k, v in inputs.items():
                static[k].copy_(v, non_blocking=True)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                       # lazy one-time setup (func attributes, workspaces) outside capture
                for _ in range(self.warmup):
                    self.fn(static, dry=True)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.fn(static, dry=False)
            ent = self.cache[key] = (graph, static, out)
        graph, static, out = ent
        for k, v in inputs.items():
            static[k].copy_(v, non_blocking=True)
        graph.replay()
        return out


def host_inputs(batch: dict) -> dict:
    """The tensors of a collated batch the device step consumes (what utils.move_to_cuda moves, fairseq/utils.py:43-67)."""
    d = {"nbr": batch["nbr"], "positions": batch["positions"], "src_tokens": batch["net_input"]["src_tokens"],
         "target": batch["target"], "start_indices": batch["start_indices"].to(torch.int32).reshape(-1)}
    for k in ("feats", "knn_dists", "knn_ids"):
        if k in batch:
            d[k] = batch[k]
    return d


def device_inputs(batch: dict, device) -> dict:
    return {k: v.to(device, non_blocking=True) for k, v in host_inputs(batch).items()}


class Stager:
    """Host -> device staging on a copy stream (pageable -> pinned -> device) into two reusable sets of device buffers, so
    the H2D copy of step i+1 runs under the kernels of step i without touching the allocator (a fresh cross-stream
    allocation per step ends in cudaMalloc, which synchronises the device).  Protocol per batch:
    `s = stage(host)` (copy stream) ... `dev = wait(s)` (compute stream waits for the copy) ... launch ... `done(s)`."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.copy = torch.cuda.Stream(self.device)
        self.slots = [None, None]            # (signature, {name: device tensor}, consumed event)
        self.n = 0

    def stage(self, host: dict):
        i = self.n & 1
        self.n += 1
        sig = tuple((k, tuple(v.shape), v.dtype) for k, v in host.items())
        if self.slots[i] is None or self.slots[i][0] != sig:
            if self.slots[i] is not None:
                self.slots[i][2].synchronize()
            self.slots[i] = (sig, {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host.items()},
                             torch.cuda.Event())
            self.slots[i][2].record(torch.cuda.current_stream(self.device))
        _, dev, consumed = self.slots[i]
        ev = torch.cuda.Event()
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(consumed)                   # the previous user of this slot has been launched and finished
            for k, v in host.items():
                dev[k].copy_(v if v.is_pinned() else v.pin_memory(), non_blocking=True)
            ev.record(self.copy)
        return i, dev, ev

    def wait(self, staged):
        _, dev, ev = staged
        torch.cuda.current_stream(self.device).wait_event(ev)
        return dev

    def done(self, staged):
        self.slots[staged[0]][2].record(torch.cuda.current_stream(self.device))


class HostBatcher:
    """Host batches of evaluate(): walks the batches of a shard and fills reusable PINNED buffer sets straight from the dataset's
    memmaps (GraphTokenBlockDataset.collate_into: one native multi-threaded copy per array, 45 MB in ~2 ms for a Wiki103-shape
    block) -- on the calling thread (workers = 0, the default), or on `workers` producer threads that fill the next batches
    under the device step of the current one and hand them over in order.  The reference does this work in DataLoader worker processes and pickles whole DGL graphs back
    (fairseq/data/iterators.py:155-167); here only the inputs of graph assembly are sliced and the consumer is the same
    process.  A buffer set returns to its pool once its H2D copy has completed (`release`); the sets are kept on the dataset, so
    a second evaluate() over the same shapes allocates nothing (one evaluate() per dataset at a time)."""

    def __init__(self, dataset, id_lists, depth: int = 4, workers: int = 0):
        import queue
        import threading
        self.dataset, self.id_lists, self.depth = dataset, list(id_lists), depth
        if not hasattr(dataset, "_host_bufs"):
            dataset._host_bufs = {}
        self.bufs = dataset._host_bufs               # shape signature -> every buffer set ever allocated (kept across calls)
        self.pools = {}                              # shape signature -> queue of this run's free buffer sets
        self._queue, self._lock, self._order, self._cv = queue, threading.Lock(), threading.Lock(), threading.Condition()
        self._next, self._done, self.error = 0, {}, None
        self.threads = [threading.Thread(target=self._run, daemon=True) for _ in range(max(0, min(workers, len(self.id_lists))))]
        for t in self.threads:
            t.start()

    def _pool(self, spec):
        key = tuple((n, tuple(sh), str(dt)) for n, sh, dt in spec)
        with self._lock:
            if key not in self.pools:                # a new run owns every set of the dataset again (the previous run is over)
                pin = torch.cuda.is_available()
                sets = self.bufs.setdefault(key, [])
                while len(sets) < self.depth + 1:
                    sets.append({n: torch.empty(sh, dtype=dt, pin_memory=pin) for n, sh, dt in spec})
                q = self._queue.Queue()
                for b in sets:
                    q.put(b)
                self.pools[key] = q
            return self.pools[key]

    def _run(self):
        try:
            while True:
                with self._order:                    # sequence number AND buffer set are taken in order: the earliest
                    seq = self._next                 # outstanding batch always owns buffers, so the in-order consumer cannot starve
                    if seq >= len(self.id_lists) or self.error is not None:
                        return
                    self._next += 1
                    ids = self.id_lists[seq]
                    pool = self._pool(self.dataset.batch_spec(ids))
                    bufs = pool.get()                # blocks while every set of this shape is in flight
                item = self.dataset.collate_into(ids, bufs)
                item["pool"] = pool
                with self._cv:
                    self._done[seq] = item
                    self._cv.notify_all()
        except BaseException as e:                   # surfaces in the consumer
            with self._cv:
                self.error = e
                self._cv.notify_all()

    def __iter__(self):
        if not self.threads:                         # workers = 0: fill on the calling thread (the copies are native and threaded)
            for ids in self.id_lists:
                pool = self._pool(self.dataset.batch_spec(ids))
                item = self.dataset.collate_into(ids, pool.get())
                item["pool"] = pool
                yield item
            return
        for seq in range(len(self.id_lists)):
            with self._cv:
                while seq not in self._done and self.error is None:
                    self._cv.wait()
                if seq not in self._done:
                    raise self.error
                item = self._done.pop(seq)
            yield item

    @staticmethod
    def release(item):
        item["pool"].put(item["host"])


def prefetch(items, to_host_inputs, device, release=None, profile: Optional[dict] = None):
    """Double-buffered input pipeline: yields (item, device tensors) while the NEXT item's tensors are already being staged,
    so the H2D copy of step i+1 overlaps the kernels of step i.  The reference stages every batch synchronously inside the
    loop (utils.move_to_cuda, fairseq_cli/eval_lm.py:217).  The yielded tensors are only valid until the next iteration."""
    st = Stager(device)
    it = iter(items)
    clock = time.perf_counter
    prof = profile if profile is not None else {}
    for k in ("wait_producer_s", "stage_s", "consumer_s", "wait_copy_s"):
        prof.setdefault(k, 0.0)
    try:
        item = next(it)
    except StopIteration:
        return
    nxt = (item, st.stage(to_host_inputs(item)))
    while nxt is not None:
        item, staged = nxt
        t0 = clock()
        try:
            n_item = next(it)
            t1 = clock()
            nxt = (n_item, st.stage(to_host_inputs(n_item)))
        except StopIteration:
            t1 = clock()
            nxt = None
        t2 = clock()
        yield item, st.wait(staged)
        t3 = clock()
        st.done(staged)
        if release is not None:             # the item's H2D copy has completed: its host buffers may be refilled
            staged[2].synchronize()
            release(item)
        t4 = clock()
        prof["wait_producer_s"] += t1 - t0
        prof["stage_s"] += t2 - t1
        prof["consumer_s"] += t3 - t2
        prof["wait_copy_s"] += t4 - t3


@torch.no_grad()
def evaluate(model, dataset: GraphTokenBlockDataset, dstore: DeviceDatastore, scorer: SequenceScorer, *,
             knn_dstore=None, temperature: float = 1.0, max_sentences: int = 1, device="cuda", rank: int = 0,
             world_size: int = 1, process_group=None, log=None, dstore_writer: Optional[DstoreWriter] = None,
             knn_keytype: Optional[str] = None, cuda_graph: bool = False, prune_unreachable: bool = True,
             max_tokens: Optional[int] = None, bucket_by_length: bool = False, host_threads: bool = True,
             host_workers: int = 0, graph_cache: Optional[dict] = None) -> dict:
    """`cuda_graph=True` replays one captured CUDA graph per batch shape instead of launching the ~100 kernels of a
    step one by one (same kernels, same results; pays off when blocks are small enough to be launch-bound).
    `prune_unreachable` drops context nodes further than graph_layer-1 hops from their centre (graph.build_token_graph
    `reach`): they cannot reach a tgt node, so scores are unchanged.
    Host side: a batch's slices go from the dataset's memmaps into reusable pinned buffers with one native multi-threaded
    copy per array (gnnlm_host_copy; ~2 ms per Wiki103 block), issued by the calling thread (`host_workers` = 0) or by that many
    producer threads (HostBatcher), then H2D on a copy stream under the previous step (Stager).
    `graph_cache`: a dict the caller keeps between calls over the SAME model / dataset settings / datastore / kNN store: the
    captured CUDA graphs (and their activation pool, ~15 GB at the Wiki103 shape) are reused instead of being re-captured.
    `host_threads=False`: slice / collate every batch synchronously on the calling thread through dataset.__getitem__ +
    collater (the reference-shaped calls) instead of the HostBatcher producer thread -- same tensors, same scores."""
    reach = model.decoder.hgt_decoder.n_layers - 1 if prune_unreachable else None
    lo, hi = shard_range(len(dataset), rank, world_size)
    cached = graph_cache.get("evaluate") if (graph_cache is not None and cuda_graph) else None
    if cached is not None:
        assert cached["key"] == (id(model), id(dataset), id(dstore), id(knn_dstore), id(scorer), temperature, reach), \
            "graph_cache belongs to another evaluate() configuration"
        acc = cached["acc"].zero_()
    else:
        acc = torch.zeros(2, dtype=torch.float64, device=device)
    ntok = 0
    host_profile = {}                   # seconds the calling thread spent waiting for the producer / staging / in the step / on H2D

    def score(inp: dict, dry: bool = False):
        sample = sample_from_inputs(inp, dataset, dstore, reach=reach)
        if knn_dstore is not None and "knn_ids" in sample:
            knn_dstore.set_search_results(sample["knn_dists"], sample["knn_ids"])
        return scorer.score_tokens(model, sample, knn_dstore, temperature, nll_acc=None if dry else acc)

    if cached is not None:
        step = cached["step"]
    else:
        step = GraphedScorer(score) if cuda_graph else score
        if graph_cache is not None and cuda_graph:
            graph_cache["evaluate"] = {"key": (id(model), id(dataset), id(dstore), id(knn_dstore), id(scorer), temperature, reach),
                                       "acc": acc, "step": step}
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    if bucket_by_length and dstore_writer is not None:
        raise ValueError("--save-knnlm-dstore writes keys in corpus order: do not bucket by length")
    id_lists = batches(dataset, lo, hi, max_sentences, max_tokens, bucket_by_length)
    if host_threads:
        # slicing + collation on a producer thread into reusable pinned buffers (HostBatcher), H2D double-buffered (Stager)
        stream = prefetch(HostBatcher(dataset, id_lists, workers=host_workers), lambda item: item["host"], device,
                          release=HostBatcher.release, profile=host_profile)
    else:
        stream = prefetch((dataset.collater([dataset[i] for i in ids]) for ids in id_lists), host_inputs, device,
                          profile=host_profile)
    for batch, inp in stream:
        _, _, _, dec_out = step(inp)
        ntok += batch["ntokens"]
        if dstore_writer is not None:                           # sequence_scorer.py:180-183 + eval_lm.py:223-244
            extra = dec_out[1]
            feat = extra[knn_keytype] if knn_keytype in extra else extra["inner_states"][-1]     # [L, B, d]
            starts = (batch["host"] if "host" in batch else batch)["start_indices"].view(-1).tolist()
            for i in range(inp["target"].shape[0]):
                mask = inp["target"][i, starts[i]:].ne(scorer.pad)
                dstore_writer.add(feat[starts[i]:, i, :][mask].float(), inp["target"][i, starts[i]:][mask])
    if world_size > 1 and torch.distributed.is_initialized():
        torch.distributed.all_reduce(acc, group=process_group)          # the path's only collective
    # (--num-shards / --shard-id without a process group: the reference's shards are never combined either)
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    score_sum, count = acc.tolist()
    avg_nll2 = -score_sum / count / math.log(2) if count else float("nan")     # eval_lm.py:325
    res = {"score_sum": score_sum, "count": count, "loss_base2": avg_nll2, "ppl": 2 ** avg_nll2,
           "tokens_this_rank": ntok, "seconds": dt, "tokens_per_s": ntok / dt if dt > 0 else float("nan"),
           "host_profile": host_profile}
    if log:
        log("Evaluated {} tokens in {:.1f}s ({:.2f} tokens/s)".format(ntok, dt, res["tokens_per_s"]))
        log("Loss (base 2): {:.4f}, Perplexity: {:.2f}".format(avg_nll2, res["ppl"]))
    return res


# ---------------------------------------------------------------------------------------------- command line
def load_model_ensemble(filenames, arg_overrides=None, task=None):
    """checkpoint_utils.load_model_ensemble (fairseq/checkpoint_utils.py:176-206): every checkpoint is {'args': Namespace,
    'model': state_dict}; the model is built from the CHECKPOINT's args (+ overrides) and loaded strictly on the hot path
    (the bypassed base-transformer keys are ignored, model.load_reference_state_dict)."""
    ensemble, args = [], None
    for filename in filenames:
        if not os.path.exists(filename):
            raise IOError("Model file not found: {}".format(filename))
        state = torch.load(filename, map_location="cpu", weights_only=False)
        args = state["args"]
        for k, v in (arg_overrides or {}).items():
            setattr(args, k, v)
        if getattr(args, "reinit_nfeat", None) is None and getattr(getattr(task, "args", None), "reinit_nfeat", False):
            args.reinit_nfeat = True
        quantizer = None
        qpath = getattr(args, "quantizer_path", "") or ""
        if qpath:
            args.quantizer_path = ""        # the codec is in the checkpoint (decoder.tgt_quantizer.*, convert_ckpt.py:40-45) ...
        if qpath and os.path.exists(qpath) and not any(k.startswith("decoder.tgt_quantizer.") for k in state["model"]):
            # ... or, for a checkpoint convert_ckpt.py was never run on, in the faiss `quantizer` file: the conversion's buffer
            # injection (convert_ckpt.py:40-45) done on the fly from the parsed file (formats.read_faiss_quantizer)
            from .formats import read_faiss_quantizer
            from .pq_codec import TorchPQCodec
            cen, A, b = read_faiss_quantizer(qpath)
            for k, v in TorchPQCodec(centroids=cen, A=A, b=b).state_dict().items():
                state["model"]["decoder.tgt_quantizer." + k] = v
        if any(k.startswith("decoder.tgt_quantizer.") for k in state["model"]):
            from .pq_codec import TorchPQCodec
            sd = state["model"]
            # A / b exist only for an OPQ index (convert_ckpt.py:40-45 writes them `if QUANTIZER.pre_torch`): a plain
            # `--index PQ64` checkpoint has the codebook alone
            A, b = sd.get("decoder.tgt_quantizer.A"), sd.get("decoder.tgt_quantizer.b")
            quantizer = TorchPQCodec(centroids=sd["decoder.tgt_quantizer.centroids_torch"].numpy(),
                                     A=None if A is None else A.numpy(), b=None if b is None else b.numpy())
        from .registry import apply_architecture
        apply_architecture(args, getattr(args, "arch", "transformer_lm"))
        if getattr(args, "max_target_positions", None) is None:
            args.max_target_positions = getattr(args, "tokens_per_sample", 1024)
        from .model import TransformerLanguageModel
        model = TransformerLanguageModel.build_model(args, task, quantizer=quantizer)
        model.load_reference_state_dict(state["model"])
        ensemble.append(model)
    return ensemble, args


def main(argv=None, device="cuda", log=print) -> dict:
    """`fairseq-eval-lm DATA --path ckpt.pt --graph --use-precompute-feat ...` (fairseq_cli/eval_lm.py:61-331) for the graph LM:
    same flags, same files, same two log lines.  kNN-LM neighbours are read from `{split}_dstore/neighbors.mmap.{--k}` (the
    precomputed-neighbour pipeline); their distances, which knn/find_knn.py:65-66 does not keep, are either
    recomputed on the device (`--knn-sim-func l2|ip`, against the PQ-decoded keys) or read from `--knn-dists-file` (raw fp32
    [N_split, k])."""
    from . import registry
    from .formats import MmapDataset
    from .dataset import neighbor_path
    from .knn_model import KNNModel
    parser = registry.eval_lm_parser()
    parser.add_argument("--knn-dists-file", default=None, help="raw fp32 [N_split, k] distances of neighbors.mmap.{k}")
    parser.add_argument("--math", default="f16x3", choices=["f16x3", "f16f8", "tf32x3", "fp32", "tf32", "bf16"])
    parser.add_argument("--cuda-graph", action="store_true")
    parser.add_argument("--share-centres", action="store_true",
                        help="run the ntgt side once per distinct centre row of a batch (identical scores; faster when "
                             "retrieved rows repeat inside a block, as they do in real kNN graphs)")
    parsed = parser.parse_args(argv)
    if parsed.path is None:
        raise ValueError("--path required for evaluation!")
    if parsed.context_window > 0:
        raise NotImplementedError("--context-window (LMContextWindowDataset) is not used by the graph LM scripts; "
                                  "use --gcn-context-window")
    task = registry.LanguageModelingTask.setup_task(parsed)
    models, args = load_model_ensemble(parsed.path.split(os.pathsep), arg_overrides=eval(parsed.model_overrides), task=task)
    if len(models) != 1:
        raise NotImplementedError("ensembles are not on the graph LM path (sequence_scorer.py:138-152)")
    for k, v in vars(parsed).items():                             # eval_lm.py:80-86
        if k not in {"self_target", "future_target", "past_target", "output_size_dictionary", "add_bos_token"}:
            setattr(args, k, v)
    task = registry.LanguageModelingTask.setup_task(args)
    dataset = task.load_dataset(args.gen_subset)
    if args.knnlm and args.save_knnlm_dstore:
        raise ValueError("Cannot use knnlm while trying to build the datastore!")
    model = models[0].eval().to(device).set_math(args.math)
    model.decoder.share_centres = bool(args.share_centres)
    dstore = task.load_datastore(device)
    knn = None
    if args.knnlm:
        n = len(dataset.tokens)
        dataset.knn_ids = MmapDataset(neighbor_path(args.data.split(os.pathsep)[0], args.gen_subset, args.k), (n, args.k),
                                      np.int64).array()
        if args.knn_dists_file:
            dataset.knn_dists = MmapDataset(args.knn_dists_file, (n, args.k), np.float32).array()
        elif args.knn_sim_func in ("l2", "ip"):
            dataset.knn_dists = np.broadcast_to(np.zeros((1, 1), np.float32), (n, args.k))       # ignored: recomputed
        else:
            raise ValueError("--knn-sim-func {} needs the neighbours' distances: pass --knn-dists-file, or recompute them with "
                             "--knn-sim-func l2|ip".format(args.knn_sim_func))
        knn = KNNModel(dstore.vals, vocab_size=len(task.target_dictionary), metric_type=args.knn_sim_func, k=args.k,
                       pq_codes=dstore.codes, quantizer=model.decoder.tgt_quantizer, index_file=args.index_file or "")
    scorer = SequenceScorer(task.target_dictionary, args.softmax_batch, args=args)
    rank, world = int(os.environ.get("RANK", args.shard_id)), int(os.environ.get("WORLD_SIZE", args.num_shards))
    use_dist = "RANK" in os.environ and world > 1
    if use_dist and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl" if str(device).startswith("cuda") else "gloo")
    writer = None
    if args.save_knnlm_dstore:
        # every rank / shard writes the rows of its own contiguous block range into the same files
        lo, hi = shard_range(len(dataset), rank, world)
        writer = DstoreWriter(args.dstore_mmap, args.gen_subset, int(dataset.sizes.sum()), args.decoder_embed_dim,
                              len(task.target_dictionary), args.dstore_fp16, args.knn_keytype,
                              offset=int(dataset.sizes[:lo].sum()), limit=int(dataset.sizes[lo:hi].sum()),
                              write_info=rank == 0 or not use_dist)
    log("{} {} {} examples".format(args.data, args.gen_subset, len(dataset)))
    log("num. model params: {}".format(sum(p.numel() for p in model.parameters())))
    res = evaluate(model, dataset, dstore, scorer, knn_dstore=knn, temperature=args.temperature,
                   max_sentences=args.max_sentences or 1, max_tokens=args.max_tokens, device=device, rank=rank,
                   world_size=world, log=log if (rank == 0 or not use_dist) else None, dstore_writer=writer,
                   knn_keytype=args.knn_keytype, cuda_graph=args.cuda_graph,
                   bucket_by_length=args.sample_break_mode not in (None, "none") and writer is None)
    if writer is not None:
        n_rows = writer.close()
        res["dstore_items_this_rank"] = n_rows
        if use_dist:
            t = torch.tensor([n_rows], dtype=torch.int64, device=device)
            torch.distributed.all_reduce(t)
            n_rows = int(t.item())
        res["dstore_items"] = n_rows
    return res


if __name__ == "__main__":
    main()
