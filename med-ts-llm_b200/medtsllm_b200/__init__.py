"""medtsllm_b200 — B200-native (sm_100a) implementation of MedTsLLM's hot path.

Layout:
  _build.py   nvcc build of libmtsb200.so (in-tree)
  _lib.py     ctypes binding of the C ABI (include/mts_b200.h)
  ops.py      tensor-level wrappers (torch owns memory/streams; kernels do the work)
  backbone.py frozen Llama / GPT-2 stacks on those kernels
  model.py    `MedTsLLM` with the reference's constructor/forward signature (models/medtsllm.py:24)
  plugin.py   registration into the reference's `models.model_lookup`
"""
from ._lib import MtsError, EXPORTED_SYMBOLS  # noqa: F401

__version__ = "0.1.0"
