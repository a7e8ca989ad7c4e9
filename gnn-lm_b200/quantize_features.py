"""Datastore quantisation -- the consumer side of knn/quantize_features.py (:46-152): encode every key of
`{subset}_dstore/keys.npy` with a trained product quantizer into `{subset}_dstore/quantized-keys.npy` (the uint8 [N, M] table the
graph builder gathers from, token_block_dataset.py:369-370), the step right before the hot path.

Same command line as the reference script.  What runs here is its `--pretrained_quantizer` branch (:58-72) and its encode loop
(:115-152): the faiss `quantizer` file is parsed without faiss (formats.read_faiss_quantizer), keys stream from the raw memmap
through pinned double buffers, `TorchPQCodec.encode` (OPQ pre-rotation + gnnlm_pq_encode: bit-exact nearest centroids under
||c||^2 - 2 x.c, pq_wrapper.py:131-167) runs on the GPU under the next chunk's H2D copy, and the codes go straight into an NPY memmap
(the reference holds the whole [N, M] table in RAM and np.save's it).  TRAINING a quantizer (faiss.index_factory(...).train, :74-110)
is faiss's own k-means / OPQ and stays out of scope: without --pretrained_quantizer this raises.

    python -m gnnlm_b200.quantize_features --data-dir DATA --subset train --code-size 128 --pretrained_quantizer [--norm] [--compute-error]
"""
import argparse
import json
import os
from typing import Optional

import numpy as np
import torch

from .dataset import dstore_path, feature_path, quantized_feature_path
from .formats import read_faiss_quantizer
from .pq_codec import TorchPQCodec


def quantizer_path(data_dir, suffix="", norm=False):
    """knn/path_utils.py:33-34."""
    return os.path.join(data_dir, f"quantizer{'-norm' if norm else ''}{suffix}")


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="quantize datastore keys with a trained PQ / OPQ quantizer (knn/quantize_features.py)")
    p.add_argument("--data-dir", type=str, required=True, help="path to binary dataset directory")
    p.add_argument("--prefix", type=str, default="de-en", help="prefix of binary file")
    p.add_argument("--index", type=str, default="OPQ64_512,PQ64", help="quantizer index")
    p.add_argument("--subset", type=str, default="train", help="train/valid/test")
    p.add_argument("--code-size", type=int, default=64, help="bytes of quantized feature")
    p.add_argument("--chunk-size", type=int, default=10000000, help="maximum number of features to train")
    p.add_argument("--compute-error", action="store_true", default=False, help="compute reconstruction error")
    p.add_argument("--use-gpu", action="store_true", default=False, help="use gpu (always, here)")
    p.add_argument("--norm", action="store_true", default=False, help="normalize feature vector to unit vector before quantize")
    p.add_argument("--pretrained_quantizer", action="store_true", default=False, help="use pretrained quantizer to encode features")
    p.add_argument("--batch-size", type=int, default=1 << 18, help="keys per GPU batch (the reference encodes 8192 at a time)")
    return p


@torch.no_grad()
def quantize_datastore(data_dir: str, subset: str = "train", code_size: Optional[int] = None, norm: bool = False,
                       compute_error: bool = False, batch_size: int = 1 << 18, device="cuda", quantizer: Optional[TorchPQCodec] = None,
                       log=print) -> dict:
    """Encode `{subset}_dstore/keys.npy` -> `{subset}_dstore/quantized-keys.npy`; returns {"path", "n", "M", "avg_error"}."""
    with open(os.path.join(dstore_path(data_dir, subset), "info.json")) as f:
        info = json.load(f)
    n, d = int(info["dstore_size"]), int(info["hidden_size"])
    keys = np.memmap(feature_path(data_dir, subset), dtype=np.float16 if info.get("dstore_fp16", False) else np.float32, mode="r",
                     shape=(n, d))
    if quantizer is None:
        path = quantizer_path(data_dir, norm=norm)
        log(f"load pretrained quantizer at {path}")
        cen, A, b = read_faiss_quantizer(path)
        quantizer = TorchPQCodec(centroids=cen, A=A, b=b)
    quantizer = quantizer.to(device)
    M = quantizer.M
    if code_size is not None and code_size != M:
        raise ValueError(f"--code-size {code_size} does not match the quantizer's {M} sub-quantizers")
    d_in = quantizer.A.shape[1] if quantizer.pre_torch else M * quantizer.dsub
    if d_in != d:
        raise ValueError(f"the quantizer takes {d_in}-dimensional keys, the datastore holds {d}-dimensional ones")
    out_path = quantized_feature_path(data_dir, subset)
    out = np.lib.format.open_memmap(out_path, mode="w+", dtype=np.uint8, shape=(n, M))      # a real NPY file, as np.save writes
    bufs = [torch.empty((min(batch_size, max(n, 1)), d), dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=device)
    main_stream = torch.cuda.current_stream(device)

    def stage(i: int):
        """keys[start:end] -> pinned fp32 (the reference's .astype(np.float32), :121) -> device, on the copy stream."""
        start, end = i * batch_size, min(n, (i + 1) * batch_size)
        host = bufs[i % 2][:end - start]
        np.copyto(host.numpy(), keys[start:end], casting="same_kind")
        with torch.cuda.stream(copy_stream):
            x = host.to(device, non_blocking=True)
            x.record_stream(main_stream)                        # consumed on the compute stream
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return start, end, x, ev

    n_batches = (n + batch_size - 1) // batch_size
    total_error = 0.0
    pending = stage(0) if n_batches else None
    drained = [None, None]              # per pinned buffer: the event after which it may be overwritten
    for i in range(n_batches):
        start, end, x, ev = pending
        main_stream.wait_event(ev)
        drained[i % 2] = ev
        if i + 1 < n_batches:
            if drained[(i + 1) % 2] is not None:
                drained[(i + 1) % 2].synchronize()              # the H2D copy that read this pinned buffer has finished
            pending = stage(i + 1)
        if norm:
            x = x / x.pow(2).sum(-1, keepdim=True).sqrt()       # :123-125
        codes = quantizer.encode(x)
        if compute_error:                                       # :134-139
            x2 = quantizer.decode(codes)
            err = float(((x - x2) ** 2).sum() / (x ** 2).sum())
            total_error += err * (end - start)
        out[start:end] = codes.cpu().numpy()
    out.flush()
    del out
    res = {"path": out_path, "n": n, "M": M, "avg_error": total_error / n if (compute_error and n) else None}
    if compute_error:
        log(f"Avg Reconstruction error: {res['avg_error']}")
    log(f"Save quantized feature to {out_path}")
    return res


def main(argv=None, device="cuda", log=print):
    args = build_parser().parse_args(argv)
    if not args.pretrained_quantizer:
        raise NotImplementedError("training a quantizer (faiss.index_factory(...).train, knn/quantize_features.py:74-110) is faiss's k-means / "
                                  "OPQ and out of scope here: train it with the reference script, then pass --pretrained_quantizer")
    return quantize_datastore(args.data_dir, args.subset, args.code_size, args.norm, args.compute_error, args.batch_size, device, log=log)


if __name__ == "__main__":
    main()
