"""Per-frame feature transform of the reference on the GPU (jsbsim_gym/features.py:37-67)."""
import ctypes as C

import torch

from . import _lib

FEATURES_PER_FRAME = 17


def jsbsim_features(observations: torch.Tensor) -> torch.Tensor:
    """(..., 15) float32 cuda tensor -> (..., 17): what JSBSimFeatureExtractor.forward returns for each
    frame; for stacked observations (B, 10, 15) the result is (B, 10, 17), i.e. the tensor
    StackedLMAFeaturesExtractor builds before its LMA blocks (jsbsim_gym/LMA_features.py:757-771)."""
    if observations.device.type != "cuda":
        raise _lib.F16Error("jsbsim_features only runs on CUDA tensors (no CPU fallback)")
    x = observations.to(torch.float32).contiguous()
    assert x.shape[-1] == 15
    out = torch.empty(x.shape[:-1] + (FEATURES_PER_FRAME,), dtype=torch.float32, device=x.device)
    n = x.numel() // 15
    stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().f16_features17(n, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), stream), "f16_features17")
    return out
