"""ultravox_b200 - B200-native (sm_100a) implementation of the Ultravox audio->LLM hot path.

Only what the path needs lives here: ``csrc/`` (hand-written CUDA + the C ABI ``libuvx.so``),
``_lib`` (ctypes binding, fails loudly when the library is missing), ``ops`` (tensor-level wrappers over the C ABI),
``autograd`` (the ``torch.autograd.Function``s of the adapter-training path: ProjectorFn / SpliceFn / LlamaStackFn / HeadLossFn),
``engine`` (CUDA-graph prefill and decode engines), ``training`` (explicit adapter trainer on the same forward / backward pieces)
and the host-side mirrors of the reference interface (``config``, ``processing``, ``model``, ``inference``, ``data_proc``).
There is no CPU fallback and nothing here imports ``oracle/``.
"""
from .config import LossConfig, LossFunction, LossMaskType, LoraConfigSimplified, UltravoxConfig, preset  # noqa: F401

__version__ = "0.1.0"
