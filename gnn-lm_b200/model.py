"""`transformer_lm` with `--graph_layer N > 0`: host-side mirror of the reference model surface on the
evaluation hot path.

  TransformerLanguageModel          fairseq/models/transformer_lm.py:27-188 (flags :122-139)
  TokenGraphTransformerDecoder      fairseq/models/transformer.py:910-1085
  AdaptiveSoftmax                   fairseq/modules/adaptive_softmax.py:50-206

Same constructor arguments (an argparse-style `args`), same method names and the same state_dict
keys for everything on the path (decoder.hgt_decoder.*, decoder.tgt_quantizer.*,
decoder.adaptive_softmax.* / decoder.embed_out).  The 16-layer base transformer is out of scope: the
target configs run with --use-precompute-feat, which bypasses it (transformer.py:974-976); its keys
(decoder.layers.*, decoder.embed_tokens.*) are accepted and ignored by load_reference_state_dict().
All arithmetic goes through libgnnlm_sm100.so.
"""
import math
from argparse import Namespace
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .graph import ETYPES, TokenGraph
from .hgt import HGT, _Weight, act_dtype, as_act, as_float
from .pq_codec import TorchPQCodec


class _W(nn.Module):
    """A bare `weight` holder so that state_dict keys match the reference module tree."""

    def __init__(self, *shape):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*shape))
        nn.init.xavier_uniform_(self.weight)


class _TiedHead(nn.Module):
    """TiedHeadModule (adaptive_softmax.py:24-47): keys word_proj.weight, class_proj.weight."""

    def __init__(self, n_words, d, n_classes):
        super().__init__()
        self.word_proj = _W(n_words, d)
        self.class_proj = _W(n_classes, d)
        self.register_buffer("_float_tensor", torch.FloatTensor(1))


class AdaptiveSoftmax(nn.Module):
    """Efficient softmax approximation (adaptive_softmax.py:50-115), evaluated at the target only.

    `tied=True` reproduces the key layout of --tie-adaptive-weights checkpoints (wiki103): head.word_proj / head.class_proj;
    with `tie_proj` (--tie-adaptive-proj, default = `tied`) the tail projection is the adaptive input's TiedLinear, stored
    [d, dim_i] and applied transposed; without it the reference builds nn.Linear(d, dim_i), stored [dim_i, d]
    (adaptive_softmax.py:96-101)."""

    def __init__(self, vocab_size: int, input_dim: int, cutoff: List[int], dropout: float = 0.0, factor: float = 4.0,
                 tied: bool = False, tie_proj: Optional[bool] = None):
        super().__init__()
        self.tie_proj = bool(tied) if tie_proj is None else bool(tie_proj) and bool(tied)
        cutoff = list(cutoff)
        if vocab_size > cutoff[-1]:
            cutoff = cutoff + [vocab_size]
        else:
            assert vocab_size == cutoff[-1], "cannot specify cutoff larger than vocab size"
        self.vocab_size, self.cutoff, self.input_dim, self.factor, self.tied = vocab_size, cutoff, input_dim, factor, tied
        self.dropout = dropout                    # training only (adaptive_softmax.py:156 on the input, :101 inside the tails)
        n_tail = len(cutoff) - 1
        if tied:
            self.head = _TiedHead(cutoff[0], input_dim, n_tail)
        else:
            self.head = _W(cutoff[0] + n_tail, input_dim)
        self.tail = nn.ModuleList()
        for i in range(n_tail):
            dim = int(input_dim // factor ** (i + 1))
            proj = _W(input_dim, dim) if self.tie_proj else _W(dim, input_dim)
            self.tail.append(nn.Sequential(proj, nn.Dropout(dropout), _W(cutoff[i + 1] - cutoff[i], dim)))
        self.register_buffer("version", torch.LongTensor([1]))
        self._prep, self._prep_key = None, None

    def prepare(self, math_mode):
        key = (math_mode, self.version.device, tuple(int(p._version) for p in self.parameters()))
        if self._prep is not None and self._prep_key == key:
            return self._prep
        if self.tied:
            head = torch.cat([self.head.word_proj.weight.detach(), self.head.class_proj.weight.detach()], 0)
        else:
            head = self.head.weight.detach()
        P = {"head": _Weight(head, None, math_mode), "proj": [], "out": []}
        for seq in self.tail:
            p = seq[0].weight.detach()
            P["proj"].append(_Weight(p.t() if self.tie_proj else p, None, math_mode))
            P["out"].append(_Weight(seq[2].weight.detach(), None, math_mode))
        self._prep, self._prep_key = P, key
        return P

    def train_weights(self) -> dict:
        """fp32 matrices of the (frozen) clusters for the training loss (train._AdaptiveLoss): head [c0 + n_tail, d],
        proj[i] [dim_i, d], out[i] [size_i, dim_i]."""
        if self.tied:
            head = torch.cat([self.head.word_proj.weight.detach(), self.head.class_proj.weight.detach()], 0)
        else:
            head = self.head.weight.detach()
        W = {"head": head.float().contiguous(), "proj": [], "out": [], "plain": None}
        for seq in self.tail:
            p = seq[0].weight.detach().float()
            W["proj"].append((p.t() if self.tie_proj else p).contiguous())
            W["out"].append(seq[2].weight.detach().float().contiguous())
        return W

    @torch.no_grad()
    def target_log_prob(self, x: torch.Tensor, target: torch.Tensor, math_mode: int = L.MATH_FP32_SIMT) -> torch.Tensor:
        """log p(target) per row: the only entries of get_log_prob's [T, V] tensor that the scorer reads
        (adaptive_softmax.py:170-206 + sequence_scorer.py:48-53), computed without materialising it."""
        x = as_act(x.reshape(-1, x.shape[-1]), math_mode)
        target = target.reshape(-1).contiguous()
        P = self.prepare(math_mode)
        head_pick, tail_rows, tail_pick, tail_count = ops.adapt_target(target, self.cutoff)
        lp = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
        pm, ps, pk, nt = ops.linear_lse(x, P["head"].W, head_pick, W_lo=P["head"].lo, w_scale=P["head"].scale, math=math_mode)
        ops.lse_finish(pm, ps, pk, nt, lp)
        for i in range(len(self.tail)):
            cnt = tail_count[i:i + 1]
            xi = ops.gather_rows(x, tail_rows[i], n_dev=cnt)
            hi = ops.linear(xi, P["proj"][i].W, None, W_lo=P["proj"][i].lo, w_scale=P["proj"][i].scale, m_dev=cnt,
                            math=math_mode, out_dtype=act_dtype(math_mode))
            pm, ps, pk, nt = ops.linear_lse(hi, P["out"][i].W, tail_pick[i], W_lo=P["out"][i].lo, w_scale=P["out"][i].scale,
                                            m_dev=cnt, math=math_mode)
            ops.lse_finish(pm, ps, pk, nt, lp, row_map=tail_rows[i], accumulate=True, m_dev=cnt)
        return lp

    @torch.no_grad()
    def get_log_prob(self, input: torch.Tensor, target: Optional[torch.Tensor], math_mode: int = L.MATH_FP32_SIMT):
        """API-compatibility path: the full [bsz, len, V] tensor of adaptive_softmax.py:170-206 with
        every column populated (the reference's target=None behaviour; with a target the reference
        leaves non-target tail columns at 0, which nobody reads).  GEMMs run on the library; the
        row normalisation uses torch.log_softmax -- do not use on the hot path (it is the tensor this
        project exists to avoid)."""
        bsz, length, dim = input.shape
        x = input.reshape(-1, dim).contiguous()
        P = self.prepare(math_mode)
        c0 = self.cutoff[0]
        lin = lambda a, w: ops.linear(a, w.W, None, W_lo=w.lo, w_scale=w.scale, math=math_mode)
        head = torch.log_softmax(lin(x, P["head"]), dim=1)
        cols = [head[:, :c0]]
        for i in range(len(self.tail)):
            hi = lin(x, P["proj"][i])
            ti = lin(hi, P["out"][i])
            cols.append(torch.log_softmax(ti, dim=1) + head[:, c0 + i, None])
        return torch.cat(cols, 1).view(bsz, length, -1)


def _get(args, name, default=None):
    return getattr(args, name, default)


class AdaptiveInput(nn.Module):
    """fairseq/modules/adaptive_input.py:13-74 -- same constructor, same state_dict keys (`embeddings.{i}.0.weight` [size_i, dim_i],
    `embeddings.{i}.1.weight` [d, dim_i], `_float_tensor`).  Only used for `--reinit-nfeat` (ntgt features = embeddings of the
    neighbour tokens, transformer.py:1046-1048).  On the device the band-wise Embedding -> Linear is evaluated ONCE per
    checkpoint into a projected table [V, d] fp32 (1.1 GB at the wiki103 shape; HBM is sized for it), so a lookup is one row
    gather (gnnlm_embed_gather) instead of a mask / index / GEMM / scatter per band and batch."""

    def __init__(self, vocab_size: int, padding_idx: int, initial_dim: int, factor: float, output_dim: int, cutoff: List[int]):
        super().__init__()
        cutoff = list(cutoff)
        if vocab_size > cutoff[-1]:
            cutoff = cutoff + [vocab_size]
        else:
            assert vocab_size == cutoff[-1], "cannot specify cutoff larger than vocab size"
        self.cutoff, self.embedding_dim, self.padding_idx = cutoff, output_dim, padding_idx
        self.embeddings = nn.ModuleList()
        for i in range(len(cutoff)):
            size = cutoff[i] - (cutoff[i - 1] if i > 0 else 0)
            dim = int(initial_dim // (factor ** i))
            self.embeddings.append(nn.Sequential(nn.Embedding(size, dim, padding_idx if i == 0 else None),
                                                 nn.Linear(dim, output_dim, bias=False)))
        for m in self.modules():                                  # adaptive_input.py:48-55
            if isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, mean=0, std=m.weight.shape[1] ** -0.5)
                nn.init.constant_(m.weight[padding_idx], 0)
            elif isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
        self.register_buffer("_float_tensor", torch.FloatTensor(1))
        self._table, self._table_key = None, None

    def weights_for_band(self, band: int):
        return self.embeddings[band][0].weight, self.embeddings[band][1].weight

    @torch.no_grad()
    def table(self) -> torch.Tensor:
        """[V, d] fp32: row v = Linear_i(Embedding_i[v - cutoff[i-1]]) for the band i of v (fp32 CUDA-core GEMM)."""
        key = (self.embeddings[0][0].weight.device, tuple(int(p._version) for p in self.parameters()))
        if self._table is None or self._table_key != key:
            dev = key[0]
            t = torch.empty((self.cutoff[-1], self.embedding_dim), device=dev, dtype=torch.float32)
            for i, seq in enumerate(self.embeddings):
                lo = self.cutoff[i - 1] if i > 0 else 0
                ops.linear(seq[0].weight.detach().float().contiguous(), seq[1].weight.detach().float().contiguous(),
                           out=t[lo:self.cutoff[i]], math=L.MATH_FP32_SIMT, tag="embed_table")
            self._table, self._table_key = t, key
        return self._table

    @torch.no_grad()
    def forward(self, input: torch.Tensor) -> torch.Tensor:
        flat = input.reshape(-1).to(torch.int32)
        return ops.gather_rows(self.table(), flat).view(*input.shape, self.embedding_dim)


class Embedding(nn.Embedding):
    """Plain input embedding (transformer_lm.py:173-175) with the `table()` face of AdaptiveInput."""

    def table(self) -> torch.Tensor:
        return self.weight.detach().float().contiguous()


class TokenGraphTransformerDecoder(nn.Module):
    """transformer.py:910-1085.  `dictionary` only needs __len__ and eos()."""

    def __init__(self, args, dictionary, embed_tokens=None, no_encoder_attn=True, quantizer: Optional[TorchPQCodec] = None):
        super().__init__()
        self.args = args
        d = args.decoder_embed_dim
        self.embed_dim = d
        self.hgt_etypes = list(ETYPES)
        self.hgt_decoder = HGT(ntype2idx={"tgt": 0, "ntgt": 1}, etype2idx={"intra": 0, "inter": 1}, in_dim=d,
                               hidden_dim=_get(args, "decoder_gcn_dim", d), out_dim=d, n_layers=args.graph_layer,
                               n_heads=args.decoder_attention_heads, dropout=_get(args, "dropout", 0.0), two_stream=False,
                               attn_drop=_get(args, "attention_dropout", 0.0))
        self.num_classes = len(dictionary)
        self.eos_idx = dictionary.eos() if hasattr(dictionary, "eos") else 2
        if embed_tokens is not None:              # only read by --reinit-nfeat (the base transformer is bypassed)
            self.embed_tokens = embed_tokens
        else:
            self.embed_tokens = None
        if quantizer is not None:
            self.tgt_quantizer = quantizer
        elif _get(args, "quantizer_path"):
            from .formats import read_faiss_quantizer      # transformer.py:936-937 (faiss.read_index) without faiss
            cen, A, b = read_faiss_quantizer(args.quantizer_path)
            self.tgt_quantizer = TorchPQCodec(centroids=cen, A=A, b=b)
        else:
            self.tgt_quantizer = None
        self.short_cut = _get(args, "short_cut", False)
        self.orig_prob_ratio = _get(args, "orig_prob_ratio", 0.0)
        cut = _get(args, "adaptive_softmax_cutoff")
        if cut is not None:
            if isinstance(cut, str):
                cut = [int(c) for c in cut.split(",")]
            tied = bool(_get(args, "tie_adaptive_weights", False))
            if tied and _get(args, "decoder_input_dim", d) != _get(args, "decoder_output_dim", d):
                # TiedHeadModule then wraps word_proj in nn.Sequential(Linear, TiedLinear) (adaptive_softmax.py:32-36)
                raise NotImplementedError("--tie-adaptive-weights with decoder_input_dim != decoder_output_dim "
                                          "(Sequential head word_proj) is not used by any graph LM config")
            self.adaptive_softmax = AdaptiveSoftmax(self.num_classes, d, cut, dropout=float(_get(args, "adaptive_softmax_dropout", 0.0) or 0.0),
                                                    factor=_get(args, "adaptive_softmax_factor", 4), tied=tied,
                                                    tie_proj=_get(args, "tie_adaptive_proj", None))
            self.embed_out = None
        else:
            self.adaptive_softmax = None
            self.embed_out = nn.Parameter(torch.empty(self.num_classes, d))       # transformer.py:650-655
            nn.init.normal_(self.embed_out, mean=0, std=d ** -0.5)
        self.math_mode = L.MATH_FP32_SIMT
        self.fold_rotation = True          # MATH_F16F8: fold the OPQ rotation into HGT layer 0 (False: explicit rotation GEMM)
        # cluster-level reuse: run the ntgt side once per DISTINCT centre row of a batch (same results; pays off on real kNN graphs,
        # where centres repeat inside a block -- uniform synthetic ids have no duplicates)
        self.share_centres = bool(_get(args, "share_centres", False))
        self._out_prep, self._out_key = None, None

    def set_math(self, mode):
        self.math_mode = L.MATH_NAMES[mode] if isinstance(mode, str) else int(mode)
        self.hgt_decoder.set_math(self.math_mode)
        return self

    # ------------------------------------------------------------------ forward
    def forward(self, prev_output_tokens, encoder_out=None, incremental_state=None, features_only=False,
                alignment_layer=None, alignment_heads=None, src_lengths=None, return_all_hiddens=False,
                graph: TokenGraph = None):
        bsz, tgt_len = prev_output_tokens.shape
        if "h" not in graph.nodes["tgt"].data:
            raise NotImplementedError("the base transformer is out of scope: evaluate with --use-precompute-feat "
                                      "(transformer.py:974-976)")
        feats = graph.nodes["tgt"].data["h"]                         # [T, d] fp16 / fp32 rows of keys.npy
        x = feats.view(bsz, tgt_len, -1)
        extra = {"inner_states": [x.transpose(0, 1)]}
        orig_x = x if self.orig_prob_ratio > 0 else None
        if not self.short_cut:
            # fp16 -> fp32 (token_block_dataset.py:328), or the activation format of the math mode
            x = self.extract_graph_features(as_act(feats, self.math_mode), prev_output_tokens, graph, encoder_out,
                                            incremental_state, shape=(bsz, tgt_len))
        extra["gcn_feat"] = x.transpose(0, 1)
        if self.orig_prob_ratio > 0:
            if self.adaptive_softmax is None:
                raise NotImplementedError("orig_prob_ratio without adaptive softmax (transformer.py:1002) is unused")
            extra["orig_x"] = orig_x
            extra["orig_ratio"] = self.orig_prob_ratio
        return x, extra      # adaptive softmax: output_layer is the identity (transformer.py:843-852)

    def extract_graph_features(self, tgt_features, prev_output_tokens, graph: TokenGraph, encoder_out=None,
                               incremental_state=None, shape=None):
        """tgt_features: [bsz, len, d] tensor (reference signature) or a flat [T, d] activation (Tensor / Split)
        together with shape=(bsz, len)."""
        assert encoder_out is None and incremental_state is None, "only support lm"
        if shape is None:
            bsz, seq_len, _ = tgt_features.shape
            h_tgt = tgt_features.reshape(-1, self.embed_dim)
        else:
            bsz, seq_len = shape
            h_tgt = tgt_features
        nd = graph.nodes["ntgt"].data
        mode = self.math_mode
        NL = self.hgt_decoder.n_layers
        if "h" in nd and nd["h"].dtype != torch.uint8:               # caller supplied decoded features
            out = self.hgt_decoder.forward_tgt(graph, h_tgt, nd["h"])
        elif "h" in nd:                                               # explicit uint8 code rows (reference layout)
            h_n = self.tgt_quantizer.decode(nd["h"], math_mode=mode)
            out = self.hgt_decoder.forward_tgt(graph, h_tgt, h_n)
        else:                                                         # fused gather from the HBM-resident datastore
            codes = graph.codes_table
            q = self.tgt_quantizer
            fold, rot = False, None
            if codes is None:
                # --reinit-nfeat: the dataset carries no code rows, ntgt.h = embed_tokens(ntgt.labels) (transformer.py:1046-1048)
                if self.embed_tokens is None:
                    raise ValueError("the graph has no ntgt features (--reinit-nfeat) and the model was built without embed_tokens")
                table, labels = self.embed_tokens.table(), graph.labels_table
                if table.shape[1] != self.embed_dim:
                    raise ValueError("--reinit-nfeat needs decoder_input_dim == decoder_embed_dim")

                def decode(g, centre_only):
                    ids, cap, n_dev = (g.inter_indices, None, g.n_valid_dev) if centre_only else (None, g.node_cap, g.n_ntgt_dev)
                    x, _ = ops.embed_gather(table, labels, g.ntgt_row, row_ids=ids, n_cap=cap, n_dev=n_dev)
                    act = act_dtype(mode)
                    if act == ops.SPLIT:
                        return ops.to_split(x, n_dev)
                    return x if act == torch.float32 else ops.convert(x, act)
            else:
                # MATH_F16F8: the OPQ rotation is folded into layer 0's weights (HGTLayer.prepare `rot`): features stay un-rotated
                fold = (q.pre_torch and q.dsub == 8 and not graph.dedup and self.fold_rotation
                        and self.hgt_decoder.can_fold_rotation(q.M * q.dsub))
                rot = q.rotation_weight() if fold else None

                def decode(g, centre_only):
                    if centre_only:   # only centre nodes are ever read (single layer; every layer's inter edges when folded)
                        # (folded: their e4m3 companion only when they also feed a projection, i.e. layer 0 is the centre-only layer)
                        return q.gather_decode(codes, g.ntgt_row, row_ids=g.inter_indices, n_dev=g.n_valid_dev, math_mode=mode,
                                               rotate=not fold, q8=NL == 2)
                    return q.gather_decode(codes, g.ntgt_row, n_cap=g.node_cap, n_dev=g.n_ntgt_dev, math_mode=mode,
                                           rotate=not fold, hi_only=fold)

            # ~11 live [rows, d] fp32-sized buffers on the ntgt side; chunk over target tokens above the budget
            per_token = graph.k * graph.w * self.embed_dim * 4 * 11
            budget = float(_get(self.args, "ntgt_memory_budget_gb", 48.0)) * 1e9
            chunk = max(64, int(budget // per_token) // 64 * 64)
            if self.share_centres and not graph.dedup and graph.T <= chunk:
                # cluster-level reuse: the ntgt side once per distinct centre row (identical results, HGT.forward_tgt_shared)
                out = self.hgt_decoder.forward_tgt_shared(graph, h_tgt, decode, rot=rot)
                return as_float(out).view(bsz, seq_len, -1)
            if NL > 1 and graph.T > chunk and graph.dedup:
                raise NotImplementedError("--deprecated graphs share ntgt nodes between tokens, so the ntgt side cannot run "
                                          "in token chunks: raise ntgt_memory_budget_gb or evaluate shorter blocks")
            if NL > 1 and graph.T > chunk:
                out = self.hgt_decoder.forward_tgt_chunked(graph, h_tgt, decode, chunk, rot=rot)
            elif NL == 1:
                out = self.hgt_decoder.forward_tgt(graph, h_tgt, None, hc0=decode(graph, True), rot=rot)
            else:
                out = self.hgt_decoder.forward_tgt(graph, h_tgt, decode(graph, False), hc0=decode(graph, True) if fold else None,
                                                   rot=rot)
        return as_float(out).view(bsz, seq_len, -1)

    # ------------------------------------------------------------------ probabilities
    def _plain_out(self):
        key = (self.math_mode, self.embed_out.device, int(self.embed_out._version))
        if self._out_prep is None or self._out_key != key:
            self._out_prep, self._out_key = _Weight(self.embed_out.detach(), None, self.math_mode), key
        return self._out_prep

    @torch.no_grad()
    def target_log_probs(self, net_output, target: torch.Tensor) -> torch.Tensor:
        """Fused fast path: per-token log p(target) [bsz, len], orig/GNN mixing included
        (transformer.py:1064-1085 evaluated at the target column)."""
        x, extra = net_output
        mode = self.math_mode
        if self.adaptive_softmax is not None:
            lp = self.adaptive_softmax.target_log_prob(x, target, mode)
            if "orig_x" in extra:
                lo = self.adaptive_softmax.target_log_prob(extra["orig_x"], target, mode)
                lp, _, _ = ops.knn_mix_nll(lp, orig_lp=lo, orig_ratio=self.orig_prob_ratio)
        else:
            w = self._plain_out()
            pm, ps, pk, nt = ops.linear_lse(as_act(x.reshape(-1, x.shape[-1]), mode), w.W, target.reshape(-1).to(torch.int32),
                                            W_lo=w.lo, w_scale=w.scale, math=mode)
            lp = torch.empty(pk.shape[0], device=pk.device, dtype=torch.float32)
            ops.lse_finish(pm, ps, pk, nt, lp)
        return lp.view(target.shape)

    @torch.no_grad()
    def get_normalized_probs(self, net_output, log_probs, sample):
        """API-compatibility path returning [bsz, len, V] (transformer.py:1064-1085)."""
        x, extra = net_output
        if self.adaptive_softmax is not None:
            target = sample["target"] if sample is not None else None
            out = self.adaptive_softmax.get_log_prob(x, target, self.math_mode)
            if "orig_x" in extra:
                o = self.adaptive_softmax.get_log_prob(extra["orig_x"], target, self.math_mode)
                a = self.orig_prob_ratio
                out = torch.logsumexp(torch.stack([o + math.log(a), out + math.log(1 - a)]), 0)
        else:
            w = self._plain_out()
            logits = ops.linear(x.reshape(-1, x.shape[-1]).contiguous(), w.W, None, W_lo=w.lo, w_scale=w.scale,
                                math=self.math_mode)
            out = torch.log_softmax(logits, dim=-1).view(x.shape[0], x.shape[1], -1)
        return out if log_probs else out.exp_()

    def max_positions(self):
        return _get(self.args, "max_target_positions", 1 << 30)


class TransformerLanguageModel(nn.Module):
    """`transformer_lm` (transformer_lm.py:27); with graph_layer > 0 the decoder is the graph decoder
    (:178-182)."""

    def __init__(self, decoder: TokenGraphTransformerDecoder):
        super().__init__()
        self.decoder = decoder

    @classmethod
    def build_model(cls, args, task=None, dictionary=None, quantizer=None):
        dictionary = dictionary if dictionary is not None else task.source_dictionary
        if _get(args, "graph_layer", 0) <= 0:
            raise NotImplementedError("only the --graph_layer > 0 decoder is on the hot path")
        embed_tokens = None
        if _get(args, "reinit_nfeat", False):                    # transformer_lm.py:160-175 (only --reinit-nfeat reads it here)
            d_in = _get(args, "decoder_input_dim", args.decoder_embed_dim)
            if _get(args, "adaptive_input", False):
                cut = _get(args, "adaptive_input_cutoff")
                cut = [int(c) for c in cut.split(",")] if isinstance(cut, str) else list(cut)
                embed_tokens = AdaptiveInput(len(dictionary), dictionary.pad(), d_in, _get(args, "adaptive_input_factor", 4),
                                             args.decoder_embed_dim, cut)
            else:
                embed_tokens = Embedding(len(dictionary), d_in, dictionary.pad())
                nn.init.normal_(embed_tokens.weight, mean=0, std=d_in ** -0.5)
                nn.init.constant_(embed_tokens.weight[dictionary.pad()], 0)
        dec = TokenGraphTransformerDecoder(args, dictionary, embed_tokens, no_encoder_attn=True, quantizer=quantizer)
        if _get(args, "freeze", False):                          # transformer_lm.py:183-186: only the HGT is trained
            for name, param in dec.named_parameters():
                if "hgt" not in name:
                    param.requires_grad = False
        return cls(dec)

    def forward(self, src_tokens, **kwargs):
        kwargs.pop("src_lengths", None)
        return self.decoder(src_tokens, **kwargs)

    def get_normalized_probs(self, net_output, log_probs, sample=None):
        return self.decoder.get_normalized_probs(net_output, log_probs, sample)

    def max_positions(self):
        return self.decoder.max_positions()

    def make_generation_fast_(self, **kwargs):
        return self

    def set_math(self, mode):
        self.decoder.set_math(mode)
        return self

    def load_reference_state_dict(self, state_dict: Dict[str, torch.Tensor]):
        """Strict for every key on the hot path; the bypassed base transformer's keys are ignored."""
        own = self.state_dict()
        skip_prefix = ("decoder.layers.", "decoder.embed_tokens.", "decoder.embed_positions.", "decoder.layer_norm.",
                       "decoder.project_", "decoder.version", "decoder.xl_bias")
        picked, ignored = {}, []
        for k, v in state_dict.items():
            if k in own:
                picked[k] = v
            elif k.startswith(skip_prefix):
                ignored.append(k)
            else:
                raise KeyError(f"unexpected key on the hot path: {k}")
        missing = [k for k in own if k not in picked]
        if missing:
            raise KeyError(f"missing keys: {missing[:8]}{'...' if len(missing) > 8 else ''}")
        self.load_state_dict(picked, strict=True)
        return ignored


def default_args(**kw) -> Namespace:
    """argparse-style args with the reference flag names (transformer_lm.py:122-139, :206-263)."""
    a = dict(decoder_embed_dim=512, decoder_attention_heads=8, graph_layer=3, decoder_gcn_dim=None, quantizer_path="",
             short_cut=False, orig_prob_ratio=0.0, adaptive_softmax_cutoff=None, adaptive_softmax_factor=4,
             tie_adaptive_weights=False, dropout=0.0, attention_dropout=0.0, max_target_positions=1 << 30)
    a.update(kw)
    if a["decoder_gcn_dim"] is None:
        a["decoder_gcn_dim"] = a["decoder_embed_dim"]
    return Namespace(**a)
