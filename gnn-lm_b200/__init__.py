"""gnnlm_b200: B200-native evaluation hot path of GNN-LM (graph assembly + PQ decode -> HGT ->
adaptive-softmax log-probs -> kNN-LM interpolation) behind the reference's fairseq-shaped API.
All compute runs in libgnnlm_sm100.so (hand-written sm_100a CUDA, C ABI in include/gnnlm_sm100.h)."""
from . import _lib  # noqa: F401  (does not load the .so until first use)
from ._lib import GnnlmError, MATH_NAMES  # noqa: F401
