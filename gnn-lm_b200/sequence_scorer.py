"""SequenceScorer -- host-side mirror of fairseq/sequence_scorer.py:16-194.

Same constructor and generate() signature / return structure.  The fast path never builds the
[bsz, len, V] tensor: the decoder returns per-target log-probs (fused adaptive-softmax epilogue) and
the kNN vote + log-space interpolation + optional NLL accumulation happen in one kernel.

Deviation documented in SURVEY.md Q4: the reference flattens kNN queries T-major but targets B-major
(sequence_scorer.py:105-117), which is only self-consistent for bsz == 1; here both are indexed
B-major consistently (identical to the reference whenever the reference is correct)."""
import sys

import torch

from . import ops


def strip_pad(tensor, pad):
    return tensor[tensor.ne(pad)]


class SequenceScorer(object):
    def __init__(self, tgt_dict, softmax_batch=None, compute_alignment=False, args=None):
        self.pad = tgt_dict.pad()
        self.eos = tgt_dict.eos()
        self.softmax_batch = softmax_batch or sys.maxsize      # kept for signature parity: nothing to chunk
        assert self.softmax_batch > 0
        self.compute_alignment = compute_alignment
        self.args = args

    @torch.no_grad()
    def score_tokens(self, model, sample, knn_dstore=None, temperature=1.0, nll_acc=None, want_knn=False):
        """[bsz, len] interpolated log-probs (+ p_knn, recall); accumulates (sum lp*w, sum w) into
        nll_acc (2 x fp64 on device) with w = 0 on pad / before start_indices."""
        net_input = sample["net_input"]
        target = sample["target"]
        model.eval()
        decoder_out = model(**net_input)
        dec = model.decoder
        lm_lp = dec.target_log_probs(decoder_out, target).reshape(-1)
        bsz, L = target.shape
        mask_kw = {}
        if nll_acc is not None:     # pad / start_indices masking happens inside the kernel
            mask_kw = dict(pad_id=self.pad, block_len=L, target=target.reshape(-1))
            if "start_indices" in sample:
                mask_kw["loss_start"] = sample["start_indices"].to(target.device, torch.int32).reshape(-1).contiguous()
        lmbda = getattr(self.args, "lmbda", 0.0) if self.args is not None else 0.0
        use_knn = knn_dstore is not None and lmbda > 0.0
        kw = {}
        if use_knn:
            dists, knns = knn_dstore.get_knns(None, positions=sample.get("positions"))
            if getattr(knn_dstore, "recompute", False):       # metric l2 / ip: queries = the --knn-keytype features (:105)
                keytype = getattr(self.args, "knn_keytype", None)
                extra = decoder_out[1]
                feat = extra[keytype] if keytype in extra else extra["inner_states"][-1]          # [L, B, d]
                queries = feat.float().permute(1, 0, 2).reshape(bsz * L, -1)                     # B-major, like the targets
                dists = knn_dstore.similarities(queries, dists, knns)
            kw = dict(target=target.reshape(-1), dists=dists, ids=knns, vals=knn_dstore.vals,
                      n_datastore=knn_dstore.dstore_size, sim_sign=knn_dstore.sim_sign, temperature=temperature,
                      lmbda=lmbda, want_knn=want_knn)
        if use_knn or nll_acc is not None:
            kw.update(mask_kw)
            lp, p, rec = ops.knn_mix_nll(lm_lp, nll_acc=nll_acc, **kw)
        else:
            lp, p, rec = lm_lp, None, None
        return lp.view(bsz, L), p, rec, decoder_out

    @torch.no_grad()
    def generate(self, models, sample, **kwargs):
        """Score a batch (sequence_scorer.py:27-194).  Returns List[List[dict]] like the reference."""
        if len(models) != 1:
            raise ValueError("Only knn *log* probs are supported.")      # :108-109; ensembles are off-path
        model = models[0]
        temperature = kwargs.get("temperature", 1.0)
        knn = kwargs.get("knn_dstore")
        probs, p_knn, recall, decoder_out = self.score_tokens(model, sample, knn, temperature, want_knn=True)
        bsz = probs.size(0)
        start_idxs = sample["start_indices"].view(-1).tolist() if "start_indices" in sample else [0] * bsz
        keytype = getattr(self.args, "knn_keytype", None) if self.args is not None else None
        extra = decoder_out[1]
        feat = extra[keytype] if keytype in extra else extra["inner_states"][-1]       # :181-182
        use_knn = recall is not None
        if use_knn:
            recall = recall.view(bsz, -1)
        hypos = []
        for i in range(bsz):
            s = start_idxs[i]
            ref = strip_pad(sample["target"][i, s:], self.pad)
            tgt_len = ref.numel()
            avg_probs_i = probs[i][s:s + tgt_len]
            mask = sample["target"][i, s:].ne(self.pad)
            hypos.append([{
                "tokens": ref,
                "score": avg_probs_i.sum() / tgt_len,
                "attention": None,
                "alignment": None,
                "positional_scores": avg_probs_i,
                "dstore_keys": feat[s:, i, :][mask],
                "knn_recall": recall[i, s:][mask] if use_knn else None,
            }])
        return hypos
