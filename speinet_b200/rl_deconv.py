"""The reference's Richardson-Lucy "edge prior" (/root/reference/model/rcl.py:18-51) on the B200 library.

Same two functions, same argument meaning as the reference:

    create_blur_kernel(kernel_size=5)                                        rcl.py:18-20
    r_l_per_channel(image_tensor, blur_kernel, num_iterations=1,
                    regularization_strength=0.01)                            rcl.py:22-51

`SPEINet` calls it on the blurry input frames (speinet.py:81 one iteration per neighbour frame, :129 / :141 five
iterations on the middle frame) right before the encoders that feed SearchTransfer.  The reference runs about eight
tiny ATen launches per channel and iteration; here every iteration of every channel is one launch of
`spei_rl_deconv`.  CUDA tensors only, no fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .search_transfer import _check_inputs, _ptr


def create_blur_kernel(kernel_size: int = 5) -> torch.Tensor:
    """[1, 1, ks, ks] box filter, ones / ks^2 (rcl.py:18-20)."""
    return torch.ones((1, 1, kernel_size, kernel_size)).to(torch.float32) / (kernel_size ** 2)


def r_l_per_channel(image_tensor: torch.Tensor, blur_kernel: torch.Tensor, num_iterations: int = 1,
                    regularization_strength: float = 0.01) -> torch.Tensor:
    """image_tensor [N, C, H, W]; blur_kernel [1, 1, ks, ks] (or [ks, ks]), ks in {3, 5, 7} -> [N, C, H, W]."""
    lib = _lib.load()
    out_dtype = image_tensor.dtype
    x = image_tensor.float().contiguous()
    k = blur_kernel.detach().to(device=x.device, dtype=torch.float32).contiguous()
    _check_inputs((x, k))
    if x.dim() != 4:
        raise RuntimeError(f"r_l_per_channel: expected [N, C, H, W], got {tuple(x.shape)}")
    ks = int(k.shape[-1])
    if k.numel() != ks * ks:
        raise RuntimeError(f"r_l_per_channel: blur_kernel must be one {ks}x{ks} filter shared by all channels "
                           f"(rcl.py:33), got {tuple(blur_kernel.shape)}")
    if num_iterations < 1:   # the reference returns the clone of the input
        return x.clone() if out_dtype == torch.float32 else x.to(out_dtype)
    n, c, h, w = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty_like(x)
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        rc = lib.spei_rl_deconv(n, c, h, w, ks, int(num_iterations), float(regularization_strength), _ptr(x), _ptr(k), _ptr(out), stream)
        _lib.check(rc, "spei_rl_deconv")
    return out if out_dtype == torch.float32 else out.to(out_dtype)
