"""The registration face of the drop-in boundary: command-line flags, architecture presets, the task's
`--graph` branch and the fairseq registry hooks, so that a reference command line

    fairseq-eval-lm DATA --user-dir .../gnn-lm_b200/fairseq_plugin --path ckpt.pt --graph --use-precompute-feat ...

resolves to this implementation.  Mirrors, by name and default:

  model flags            fairseq/models/transformer_lm.py:52-139            add_model_args
  task flags             fairseq/tasks/language_modeling.py:68-153          add_task_args
  eval-lm flags          fairseq/options.py:456-501 (+ :328-331)            add_eval_lm_args
  architecture presets   fairseq/models/transformer_lm.py:190-333           ARCHITECTURES / apply_architecture
  task                   fairseq/tasks/language_modeling.py:168-319         LanguageModelingTask (graph branch only)
  registry hooks         fairseq/models/__init__.py:51-119, fairseq/tasks/__init__.py, fairseq/utils.py:315-330
                                                                           register_with_fairseq

Flag spellings, types and defaults, and what every preset makes of an empty Namespace, are pinned to the reference's own
add_args / architecture functions (tests/golden/registry.json).  Host code only; nothing here touches the GPU."""
import argparse
import os
import sys
from typing import Dict, List, Optional

from .model import TransformerLanguageModel

_T, _S = "store_true", "store"
# (option strings, action, type, default[, extra kwargs]) -- transformer_lm.py:52-139
MODEL_FLAGS = [
    (["--activation-fn"], _S, None, None, {"choices": ["relu", "gelu", "gelu_fast", "gelu_accurate", "tanh", "linear"]}),
    (["--dropout"], _S, float, None), (["--attention-dropout"], _S, float, None),
    (["--activation-dropout", "--relu-dropout"], _S, float, None),
    (["--decoder-embed-dim"], _S, int, None), (["--decoder-output-dim"], _S, int, None), (["--decoder-input-dim"], _S, int, None),
    (["--decoder-ffn-embed-dim"], _S, int, None), (["--decoder-layers"], _S, int, None),
    (["--decoder-attention-heads"], _S, int, None), (["--decoder-normalize-before"], _T, None, False),
    (["--no-decoder-final-norm"], _T, None, False), (["--adaptive-softmax-cutoff"], _S, None, None),
    (["--adaptive-softmax-dropout"], _S, float, None), (["--adaptive-softmax-factor"], _S, float, None),
    (["--no-token-positional-embeddings"], _T, None, False), (["--share-decoder-input-output-embed"], _T, None, False),
    (["--character-embeddings"], _T, None, False),
    (["--character-filters"], _S, str, "[(1, 64), (2, 128), (3, 192), (4, 256), (5, 256), (6, 256), (7, 256)]"),
    (["--character-embedding-dim"], _S, int, 4), (["--char-embedder-highway-layers"], _S, int, 2),
    (["--adaptive-input"], _T, None, False), (["--adaptive-input-factor"], _S, float, None),
    (["--adaptive-input-cutoff"], _S, None, None), (["--tie-adaptive-weights"], _T, None, False),
    (["--tie-adaptive-proj"], _T, None, False), (["--decoder-learned-pos"], _T, None, False),
    (["--decoder-layerdrop"], _S, float, 0), (["--decoder-layers-to-keep"], _S, None, None),
    (["--layernorm-embedding"], _T, None, False), (["--no-scale-embedding"], _T, None, False),
    # the graph transformer (:122-139)
    (["--quantizer_path"], _S, str, ""), (["--graph_layer"], _S, int, 0), (["--decoder_gcn_dim"], _S, int, 1024),
    (["--freeze"], _T, None, False, {"default": False}), (["--short-cut"], _T, None, False, {"default": False}),
    (["--orig_prob_ratio"], _S, float, 0.0), (["--add-bias"], _T, None, False, {"default": False}),
]
# language_modeling.py:68-153 (`data` positional first)
TASK_FLAGS = [
    (["data"], _S, None, None),
    (["--sample-break-mode"], _S, None, "none", {"choices": ["none", "complete", "complete_doc", "eos"]}),
    (["--tokens-per-sample"], _S, int, 1024), (["--output-dictionary-size"], _S, int, -1), (["--self-target"], _T, None, False),
    (["--future-target"], _T, None, False), (["--past-target"], _T, None, False), (["--add-bos-token"], _T, None, False),
    (["--max-target-positions"], _S, int, None), (["--truncate-sequence"], _T, None, False),
    # kNN-LM
    (["--knn-keytype"], _S, str, None), (["--probe"], _S, int, 8), (["--k"], _S, int, 1024), (["--dstore-size"], _S, int, 103227021),
    (["--dstore-filename"], _S, str, None), (["--indexfile"], _S, str, None), (["--lmbda"], _S, float, 0.0),
    (["--knn-sim-func"], _S, str, "do_not_recomp_ip"), (["--faiss-metric-type"], _S, str, "l2"), (["--no-load-keys"], _T, None, False),
    (["--dstore-fp16"], _T, None, False), (["--move-dstore-to-mem"], _T, None, False), (["--load-neighbor"], _T, None, False),
    # graph
    (["--graph"], _T, None, False), (["--neighbor-context"], _S, None, "(2, 2)"), (["--use-precompute-feat"], _T, None, False),
    (["--invalid-neighbor-context"], _S, int, 1536), (["--gcn-k"], _S, int, 1024), (["--dstore-dir"], _S, str, None),
    (["--index-file"], _S, str, None), (["--plasma_path"], _S, str, ""), (["--gcn-context-window"], _S, int, 0),
    (["--intra-context"], _S, int, 0), (["--deprecated"], _T, None, False), (["--reinit-nfeat"], _T, None, False),
]
# options.py:456-501
EVAL_LM_FLAGS = [
    (["--path"], _S, None, None), (["--remove-bpe"], _S, None, None, {"nargs": "?", "const": "@@ "}), (["--quiet"], _T, None, False),
    (["--model-overrides"], _S, str, "{}"), (["--results-path"], _S, str, None),
    (["--output-word-probs"], _T, None, False), (["--output-word-stats"], _T, None, False), (["--context-window"], _S, int, 0),
    (["--softmax-batch"], _S, int, sys.maxsize), (["--lm-eval"], _T, None, True), (["--knnlm"], _T, None, False),
    (["--save-knnlm-dstore"], _T, None, False), (["--dstore-mmap"], _S, str, None), (["--first"], _S, int, 0),
    (["--temperature"], _S, float, 1.0), (["--output-knn-recall"], _T, None, False),
]


def _add(parser, table):
    for entry in table:
        flags, action, typ, default = entry[:4]
        kw = dict(entry[4]) if len(entry) > 4 else {}
        # a default is passed only where the reference passes one, so that a parser group created with
        # argument_default=SUPPRESS (fairseq's model-specific group) leaves the other flags absent unless given
        if action == _T:
            if default is True:
                kw.setdefault("default", True)
            parser.add_argument(*flags, action="store_true", **kw)
        elif flags[0].startswith("-"):
            if default is not None:
                kw["default"] = default
            if typ is not None:
                kw["type"] = typ
            parser.add_argument(*flags, **kw)
        else:
            parser.add_argument(*flags, **kw)             # positional
    return parser


def add_model_args(parser):
    return _add(parser, MODEL_FLAGS)


def add_task_args(parser):
    return _add(parser, TASK_FLAGS)


def add_eval_lm_args(parser):
    return _add(parser.add_argument_group("LM Evaluation"), EVAL_LM_FLAGS)


# ---- architecture presets: (defaults applied first, parent preset applied after), transformer_lm.py:190-333.  Every value is a
# `getattr(args, name, default)` default: an attribute the command line or the checkpoint already carries wins.
_GPT = dict(dropout=0.1, attention_dropout=0.1, activation_fn="gelu")
ARCHITECTURES: Dict[str, tuple] = {
    "transformer_lm": (dict(dropout=0.1, attention_dropout=0.0, decoder_embed_dim=512, decoder_ffn_embed_dim=2048, decoder_layers=6,
                            decoder_attention_heads=8, adaptive_softmax_cutoff=None, adaptive_softmax_dropout=0,
                            adaptive_softmax_factor=4, decoder_learned_pos=False, activation_fn="relu", add_bos_token=False,
                            no_token_positional_embeddings=False, share_decoder_input_output_embed=False,
                            character_embeddings=False), None),
    "transformer_lm_big": (dict(decoder_layers=12, decoder_embed_dim=1024, decoder_ffn_embed_dim=4096, decoder_attention_heads=16),
                           "transformer_lm"),
    "transformer_lm_baevski_wiki103": (dict(decoder_layers=16, decoder_attention_heads=8, dropout=0.3, adaptive_input=True,
                                            tie_adaptive_weights=True, adaptive_input_cutoff="20000,60000",
                                            adaptive_softmax_cutoff="20000,60000", adaptive_softmax_dropout=0.2,
                                            attention_dropout=0.1, activation_dropout=0.1, no_decoder_final_norm=True,
                                            tie_adaptive_proj=True), "transformer_lm_big"),
    "transformer_lm_baevski_gbw": (dict(decoder_embed_dim=512, dropout=0.1, attention_dropout=0.1, no_decoder_final_norm=True),
                                   "transformer_lm_big"),
    "transformer_lm_gpt": (dict(decoder_embed_dim=768, decoder_ffn_embed_dim=3072, decoder_layers=12, decoder_attention_heads=12,
                                **_GPT), "transformer_lm"),
    "transformer_lm_gpt2_small": (dict(decoder_embed_dim=1024, decoder_ffn_embed_dim=4096, decoder_layers=24,
                                       decoder_attention_heads=16, **_GPT), "transformer_lm"),
    "transformer_lm_gpt2_medium": (dict(decoder_embed_dim=1280, decoder_ffn_embed_dim=5120, decoder_layers=36,
                                        decoder_attention_heads=20, **_GPT), "transformer_lm"),
    "transformer_lm_gpt2_big": (dict(decoder_embed_dim=1600, decoder_ffn_embed_dim=6400, decoder_layers=48,
                                     decoder_attention_heads=25, **_GPT), "transformer_lm"),
    "transformer_lm_enwik8": (dict(decoder_embed_dim=1024, decoder_ffn_embed_dim=4096, decoder_layers=24, decoder_attention_heads=16,
                                   dropout=0.15, attention_dropout=0.15, activation_fn="gelu"), "transformer_lm"),
}
ARCHITECTURES["transformer_lm_wiki103"] = ARCHITECTURES["transformer_lm_baevski_wiki103"]
ARCHITECTURES["transformer_lm_gbw"] = ARCHITECTURES["transformer_lm_baevski_gbw"]
# tail of the base preset (:222-245): evaluated after the table above because two of them depend on decoder_embed_dim
_BASE_TAIL = dict(no_decoder_final_norm=False, adaptive_input=False, adaptive_input_factor=4, adaptive_input_cutoff=None,
                  tie_adaptive_weights=False, tie_adaptive_proj=False, no_scale_embedding=False, layernorm_embedding=False)


def apply_architecture(args, arch: str = "transformer_lm"):
    """What `ARCH_CONFIG_REGISTRY[arch](args)` does in the reference: fill in every attribute the namespace lacks."""
    if arch not in ARCHITECTURES:
        raise ValueError(f"unknown architecture: {arch}")
    defaults, parent = ARCHITECTURES[arch]
    if parent is None:                                            # base_lm_architecture, :191-245
        if hasattr(args, "no_tie_adaptive_proj"):                 # checkpoints older than --tie-adaptive-proj
            args.no_decoder_final_norm = True
            if args.no_tie_adaptive_proj is False:
                args.tie_adaptive_proj = True
        if hasattr(args, "decoder_final_norm"):
            args.no_decoder_final_norm = not args.decoder_final_norm
    for k, v in defaults.items():
        if not hasattr(args, k):
            setattr(args, k, v)
    if parent is not None:
        return apply_architecture(args, parent)
    for k in ("decoder_output_dim", "decoder_input_dim"):
        if not hasattr(args, k):
            setattr(args, k, args.decoder_embed_dim)
    args.decoder_normalize_before = True                          # unconditional in the reference (:225)
    for k, v in _BASE_TAIL.items():
        if not hasattr(args, k):
            setattr(args, k, v)
    return args


# ---- the task's --graph branch
class LanguageModelingTask:
    """fairseq/tasks/language_modeling.py:39-319 restricted to what evaluation of a `--graph` model touches: setup_task
    (dict.txt), load_dataset (the graph branch, :265-303), build_model, the dictionaries and dataset()."""

    def __init__(self, args, dictionary, output_dictionary=None, targets=None):
        self.args, self.dictionary = args, dictionary
        self.output_dictionary = output_dictionary or dictionary
        self.targets = targets if targets is not None else ["future"]
        self.graph = getattr(args, "graph", False)
        self.datasets = {}

    @staticmethod
    def add_args(parser):
        add_task_args(parser)

    @classmethod
    def setup_task(cls, args, **kwargs):
        from .formats import Dictionary
        dictionary = None
        if getattr(args, "data", None):
            paths = args.data.split(os.pathsep)
            dictionary = Dictionary.load(os.path.join(paths[0], "dict.txt"))
        if getattr(args, "output_dictionary_size", -1) >= 0:
            raise NotImplementedError("--output-dictionary-size (TruncatedDictionary) is not on the evaluation path")
        targets = [t for t in ("self", "future", "past") if getattr(args, t + "_target", False)] or ["future"]
        if targets != ["future"]:
            raise ValueError("Unsupported language modeling target: {}".format(targets))      # language_modeling.py:208-213
        return cls(args, dictionary, dictionary, targets=targets)

    @property
    def source_dictionary(self):
        return self.dictionary

    @property
    def target_dictionary(self):
        return self.output_dictionary

    def load_dataset(self, split, epoch=0, combine=False, **kwargs):
        from .formats import load_graph_lm_dataset
        a = self.args
        if not self.graph:
            raise NotImplementedError("only the --graph dataset is on the hot path (language_modeling.py:238-264 is stock fairseq)")
        if getattr(a, "truncate_sequence", False) or getattr(a, "add_bos_token", False):
            raise NotImplementedError("--truncate-sequence / --add-bos-token are not used by the graph LM scripts")
        paths = a.data.split(os.pathsep)
        ds, _ = load_graph_lm_dataset(
            paths[epoch % len(paths)], split, tokens_per_sample=a.tokens_per_sample, gcn_k=a.gcn_k,
            neighbor_context=a.neighbor_context, use_precompute_feat=a.use_precompute_feat,
            invalid_neighbor_context=a.invalid_neighbor_context, gcn_context_window=a.gcn_context_window,
            intra_context=a.intra_context, sample_break_mode=a.sample_break_mode, deprecated=a.deprecated)
        self.datasets[split] = ds
        return ds

    def dataset(self, split):
        if split not in self.datasets:
            raise KeyError("Dataset not loaded: " + split)
        return self.datasets[split]

    def build_model(self, args):
        return build_model(args, self)

    def load_datastore(self, device):
        """The HBM-resident tables the graph is assembled from (train_dstore/{quantized-keys.npy, vals.npy})."""
        from .dataset import DeviceDatastore
        return DeviceDatastore.from_dir(self.args.data.split(os.pathsep)[0], len(self.dictionary), device,
                                        reinit_nfeat=getattr(self.args, "reinit_nfeat", False))


def build_model(args, task, arch: Optional[str] = None):
    """`ARCH_MODEL_REGISTRY[args.arch].build_model(args, task)` (fairseq/models/__init__.py:46-47) for the graph LM."""
    apply_architecture(args, arch or getattr(args, "arch", "transformer_lm"))
    if getattr(args, "max_target_positions", None) is None:                                   # transformer_lm.py:151-152
        args.max_target_positions = getattr(args, "tokens_per_sample", 1024)
    if getattr(task, "args", None) is not None and getattr(task.args, "reinit_nfeat", False):
        args.reinit_nfeat = True                                  # a task flag the decoder needs (embed_tokens is built only then)
    return TransformerLanguageModel.build_model(args, task)


# ---- fairseq registry hooks
GRAPH_MODEL_NAME = "hgt_lm"       # the name north_star uses; `transformer_lm` itself is re-pointed only with override=True


def register_with_fairseq(models_module, tasks_module=None, *, override: bool = False) -> List[str]:
    """Register the graph LM with a fairseq-shaped registry (`models_module` needs register_model,
    register_model_architecture, MODEL_REGISTRY, ARCH_MODEL_REGISTRY, ARCH_CONFIG_REGISTRY, BaseFairseqModel; `tasks_module`
    register_task, TASK_REGISTRY, FairseqTask) through its OWN decorators, so its duplicate / base-class checks apply.

    Always: model `hgt_lm` with one architecture per reference preset (`hgt_lm`, `hgt_lm_big`, `hgt_lm_wiki103`, ...).
    override=True additionally re-points `transformer_lm` and its architectures (what existing checkpoints name in
    `args.arch`) and the `language_modeling` task at classes that take the `--graph_layer > 0` / `--graph` case and hand
    everything else back to the original class.  Returns the registered architecture names."""
    base = models_module.BaseFairseqModel         # (FairseqLanguageModel.__init__ insists on a FairseqDecoder instance)
    original = models_module.MODEL_REGISTRY.get("transformer_lm")

    class HGTLanguageModel(TransformerLanguageModel, base):
        @staticmethod
        def add_args(parser):
            add_model_args(parser)

        @classmethod
        def build_model(cls, args, task):
            if getattr(args, "graph_layer", 0) <= 0 and original is not None:
                return original.build_model(args, task)           # not a graph model: stock fairseq
            apply_architecture(args, "transformer_lm")            # "make sure all arguments are present in older models"
            if getattr(args, "max_target_positions", None) is None:
                args.max_target_positions = getattr(args, "tokens_per_sample", 1024)
            if getattr(getattr(task, "args", None), "reinit_nfeat", False):
                args.reinit_nfeat = True
            return TransformerLanguageModel.build_model.__func__(cls, args, task)

    models_module.register_model(GRAPH_MODEL_NAME)(HGTLanguageModel)
    names = []
    for arch in ARCHITECTURES:
        new = arch.replace("transformer_lm", GRAPH_MODEL_NAME, 1)
        models_module.register_model_architecture(GRAPH_MODEL_NAME, new)(lambda args, _a=arch: apply_architecture(args, _a))
        names.append(new)
    if override:
        models_module.MODEL_REGISTRY["transformer_lm"] = HGTLanguageModel
        for arch in ARCHITECTURES:
            models_module.ARCH_MODEL_REGISTRY[arch] = HGTLanguageModel
            models_module.ARCH_CONFIG_REGISTRY.setdefault(arch, lambda args, _a=arch: apply_architecture(args, _a))
    if tasks_module is not None:
        orig_task = tasks_module.TASK_REGISTRY.get("language_modeling")
        task_base = tasks_module.FairseqTask

        class GraphLanguageModelingTask(LanguageModelingTask, task_base):
            def __init__(self, args, dictionary, output_dictionary=None, targets=None):
                task_base.__init__(self, args)
                LanguageModelingTask.__init__(self, args, dictionary, output_dictionary, targets)

            @classmethod
            def setup_task(cls, args, **kwargs):
                if not getattr(args, "graph", False) and orig_task is not None:
                    return orig_task.setup_task(args, **kwargs)   # not a graph run: stock fairseq
                return super().setup_task(args, **kwargs)

        tasks_module.register_task("graph_language_modeling")(GraphLanguageModelingTask)
        if override:
            tasks_module.TASK_REGISTRY["language_modeling"] = GraphLanguageModelingTask
    return names


def eval_lm_parser() -> argparse.ArgumentParser:
    """A standalone parser with the reference's spellings for the flags of the path (task + model + eval-lm groups and the
    dataset / sharding flags of options.py:276-331 that eval_lm reads)."""
    p = argparse.ArgumentParser("gnnlm-eval-lm", allow_abbrev=False)
    add_task_args(p)
    # as fairseq/options.py parse_args_and_arch does: model flags without an explicit default stay ABSENT unless given, so that
    # the architecture preset (or the checkpoint's args) fills them in
    add_model_args(p.add_argument_group("Model-specific configuration", argument_default=argparse.SUPPRESS))
    add_eval_lm_args(p)
    p.add_argument("--gen-subset", default="test", metavar="SPLIT")
    p.add_argument("--num-shards", default=1, type=int, metavar="N")
    p.add_argument("--shard-id", default=0, type=int, metavar="ID")
    p.add_argument("--max-sentences", "--batch-size", type=int, metavar="N")
    p.add_argument("--max-tokens", type=int, metavar="N")
    p.add_argument("--user-dir", default=None)
    p.add_argument("--arch", default="transformer_lm")
    return p
