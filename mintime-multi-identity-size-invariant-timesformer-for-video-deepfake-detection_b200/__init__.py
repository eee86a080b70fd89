"""mintime_b200 -- B200-native (sm_100a) implementation of MINTIME's data-parallel hot path:
EfficientNet-B0 per-face feature extractor -> Size-Invariant TimeSformer (identity-masked divided
space-time attention).

Host side mirrors the reference's ``nn.Module`` surface (same class names, constructor arguments,
forward signatures and ``state_dict`` keys); all arithmetic runs in hand-written CUDA kernels
behind the C-ABI declared in ``include/mintime_b200.h`` (``libmintime_b200.so``).  There is no CPU
or PyTorch fallback: calling a forward without the built library / without a GPU raises.
"""
from . import spec, synth  # noqa: F401  (pure-python, importable without the CUDA library)

__all__ = ["spec", "synth", "EfficientNet", "Xception", "xception", "SizeInvariantTimeSformer", "lib"]


def __getattr__(name):
    # heavy / CUDA-bound parts are loaded on first use so that CPU-only tooling can import the package
    if name == "EfficientNet":
        from .efficientnet import EfficientNet
        return EfficientNet
    if name == "Xception":
        from .xception import Xception
        return Xception
    if name == "xception":
        from .xception import xception
        return xception
    if name == "SizeInvariantTimeSformer":
        from .size_invariant_timesformer import SizeInvariantTimeSformer
        return SizeInvariantTimeSformer
    if name == "lib":
        from . import _lib
        return _lib
    raise AttributeError(name)
