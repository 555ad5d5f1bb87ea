"""medtsllm_b200 — B200-native (sm_100a) implementation of MedTsLLM's hot path.

Layout:
  _build.py   nvcc build of libmtsb200.so (in-tree)
  _lib.py     ctypes binding of the C ABI (include/mts_b200.h)
  ops.py      tensor-level wrappers (torch owns memory/streams; kernels do the work)
  backbone.py frozen Llama / GPT-2 stacks on those kernels
  model.py    `MedTsLLM` with the reference's constructor/forward signature (models/medtsllm.py:24)
  train.py    the training path: one autograd.Function with a manual backward chain; training-step CUDA graphs
  lora.py     LoRA A/B pairs on the kernel stack (forward, backward, save / load)
  graph.py    CUDA-graph replay of the inference path
  dp.py       data parallelism: gradient buckets over torch.distributed, DistributedSampler swap
  gpt4ts.py   `GPT4TS` (models/gpt4ts.py) on the same GPT-2 kernel stack, inference and training
  synthetic.py  synthetic workloads of the BASELINE configs (bench / smoke)
  plugin.py   registration into the reference's `models.model_lookup`
"""
from ._lib import MtsError, EXPORTED_SYMBOLS  # noqa: F401

__version__ = "0.1.0"
