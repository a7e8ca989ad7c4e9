"""kNN-LM probability -- host-side mirror of knn/knn_model.py:25-217 (KNNModel) for the
precomputed-neighbour pipeline: the faiss search (get_knns, :87-101) is out of scope, so the
retrieved (dists, ids) are inputs, exactly as find_knn.py precomputes ids for the graph
(the eval script's queries are the same precomputed features find_knn.py searched
with, hence precomputable).  Everything after the search -- similarity sign, -1 masking, softmax
over sims/T, the vote on the target or the full-vocabulary scatter -- runs in logprob_knn.cu."""
from typing import Optional

import numpy as np
import torch

from . import ops


class KNNModel(object):
    def __init__(self, vals: torch.Tensor, vocab_size: Optional[int] = None, metric_type: str = "do_not_recomp_ip",
                 k: int = 1024, dists: Optional[torch.Tensor] = None, knns: Optional[torch.Tensor] = None):
        """vals: [N_d] (or [N_d,1]) int16/int32 datastore values resident in HBM (vals.npy);
        dists/knns: optional whole-split precomputed search results [N_split, k] indexed by stream position."""
        assert metric_type in ["do_not_recomp_l2", "do_not_recomp_ip", "l2", "ip"]
        if metric_type in ("l2", "ip"):
            raise NotImplementedError("similarity recompute from keys (knn_model.py:159-177) is a 'next' row "
                                      "(SURVEY.md 8f-3); use do_not_recomp_*")
        self.vals = vals.reshape(-1)
        assert self.vals.dtype in (torch.int16, torch.int32)
        self.dstore_size = self.vals.numel()
        self.vocab_size = vocab_size
        self.metric_type = metric_type
        self.k = k
        self.dists, self.knns = dists, knns
        self._pending = None

    @property
    def sim_sign(self):
        return -1.0 if self.metric_type == "do_not_recomp_l2" else 1.0      # knn_model.py:153-157

    def set_search_results(self, dists: torch.Tensor, knns: torch.Tensor):
        """Provide the (out-of-scope) search output for the next get_knn_prob call: [num, k] each."""
        self._pending = (dists.contiguous().float(), knns.contiguous().long())

    def get_knns(self, queries, k: int = 0, positions: Optional[torch.Tensor] = None):
        if self._pending is not None:
            out, self._pending = self._pending, None
            return out
        if positions is not None and self.dists is not None:
            return self.dists[positions].contiguous(), self.knns[positions].contiguous()
        raise RuntimeError("faiss search is out of scope: call set_search_results() or pass precomputed arrays")

    @torch.no_grad()
    def get_knn_prob(self, queries, k: int = 0, output_size: int = None, return_knn: bool = False, t: float = 1.0,
                     targets: torch.Tensor = None, return_recall: bool = False, positions=None):
        """knn_model.py:103-217.  `queries` is only used for its leading dimension."""
        dists, knns = self.get_knns(queries, k=k, positions=positions)
        if targets is None:
            output_size = output_size or self.vocab_size
            if not output_size:
                raise ValueError("DataStore.info does not have vocab_size, please set output_size manually")
            probs = ops.knn_full_prob(dists, knns, self.vals, output_size, self.dstore_size, self.sim_sign, t)
            if return_knn:
                return probs, dists * self.sim_sign, knns
            return probs
        zero = torch.zeros(dists.shape[0], device=dists.device, dtype=torch.float32)
        # lambda = 0.5 is irrelevant here: only p_knn / recall are read back
        _, p, rec = ops.knn_mix_nll(zero, target=targets.reshape(-1).contiguous(), dists=dists, ids=knns, vals=self.vals,
                                    n_datastore=self.dstore_size, sim_sign=self.sim_sign, temperature=t, lmbda=0.5,
                                    want_knn=True)
        if not return_recall:
            return p
        return p, rec.long()
