"""GPU image pre-processing: the reference's demo.py:191-196 (BGR->RGB, cv2.resize, imagenet_normalize, HWC->CHW)
as one kernel over a batch of uint8 frames already on the device.  Bit-exact against cv2 4.13 + numpy."""
import ctypes as C

import torch

from . import _native as nv


def preprocess(images, size, out=None):
    """images: uint8 CUDA tensor [N, h, w, 3] (or [h, w, 3]) in OpenCV's B,G,R order, rows may be strided;
    size: (width, height) of the network input, as cv2.resize takes it (demo.py:104 ``net_input_size``).
    Returns float32 [N, 3, height, width]: the ``img`` tensor demo.py:196 feeds to ``hydranet(img)``."""
    if images.dim() == 3:
        images = images.unsqueeze(0)
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[3] != 3:
        raise TypeError("preprocess expects a uint8 tensor [N, h, w, 3] (BGR), got %s %s" % (images.dtype, tuple(images.shape)))
    if not images.is_cuda:
        raise RuntimeError("hydranet_b200 runs on CUDA only (no CPU fallback); got a CPU tensor")
    if images.stride(3) != 1 or images.stride(2) != 3:
        images = images.contiguous()
    n, h, w, _ = images.shape
    width, height = int(size[0]), int(size[1])
    if out is None:
        out = torch.empty((n, 3, height, width), dtype=torch.float32, device=images.device)
    elif out.shape != (n, 3, height, width) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != images.device:
        raise ValueError("preprocess: `out` must be a contiguous float32 [%d, 3, %d, %d] tensor on %s" % (n, height, width, images.device))
    if n == 0:
        return out
    d = nv.PreprocessDesc(images.data_ptr(), n, h, w, images.stride(1), images.stride(0) if n > 1 else max(images.stride(0), images.stride(1) * h),
                          out.data_ptr(), height, width)
    with torch.cuda.device(images.device):
        nv.check(nv.lib.hn_preprocess_fwd(C.byref(d), torch.cuda.current_stream(images.device).cuda_stream))
    return out
