"""numpy restatement of the reference's Richardson-Lucy "edge prior" (test infrastructure only).

  /root/reference/model/rcl.py:18-20   create_blur_kernel: 5x5 box filter, ones / 25
  /root/reference/model/rcl.py:22-51   r_l_per_channel(image, blur_kernel, num_iterations, regularization_strength)
  call sites: /root/reference/model/speinet.py:81 (1 iteration, neighbour frames), :129 / :141 (5 iterations, mid frame)

Per channel (channels are independent, rcl.py:27-28) and per iteration (rcl.py:32-46):
    blurred = conv2d(d, blur_kernel, padding=ks//2)            zero padding, cross-correlation (F.conv2d)
    cf      = x / blurred ; cf[cf != cf] = 0 ; cf[cf < 0] = 0     (0/0 -> NaN -> 0; negative -> 0; +inf stays)
    reg     = d + lambda * conv2d(d, [[0,-1,0],[-1,4,-1],[0,-1,0]], padding=1)
    d       = cf * reg
with d = x before the first iteration.  All arithmetic fp32.  The 25 taps of the blur are added in ascending
(ky, kx) order with one multiply per tap, the Laplacian as 4*d - up - left - right - down in ascending tap order:
torch's CPU convolution may associate differently, so the comparison against the reference is by tolerance
(1e-5 relative on the fixtures), not bit-exact.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def create_blur_kernel(kernel_size: int = 5) -> np.ndarray:
    """rcl.py:18-20."""
    return (np.ones((kernel_size, kernel_size), dtype=F32) / F32(kernel_size ** 2)).astype(F32)


def _corr2d_zero_pad(img: np.ndarray, k: np.ndarray) -> np.ndarray:
    """F.conv2d(img[None,None], k[None,None], padding=ks//2) for one [H, W] plane, fp32 accumulation in tap order."""
    ks = k.shape[0]
    p = ks // 2
    H, W = img.shape
    pad = np.zeros((H + 2 * p, W + 2 * p), dtype=F32)
    pad[p:p + H, p:p + W] = img
    acc = np.zeros((H, W), dtype=F32)
    for ky in range(ks):
        for kx in range(ks):
            if k[ky, kx] != 0:
                acc = (acc + pad[ky:ky + H, kx:kx + W] * F32(k[ky, kx])).astype(F32)
    return acc


_LAPLACIAN = np.array([[0, -1, 0], [-1, 4, -1], [0, -1, 0]], dtype=F32)   # rcl.py:30


def r_l_per_channel(image: np.ndarray, blur_kernel: np.ndarray, num_iterations: int = 1,
                    regularization_strength: float = 0.01) -> np.ndarray:
    """image [N, C, H, W] fp32; blur_kernel [ks, ks] (or [1,1,ks,ks]) fp32 -> [N, C, H, W] fp32 (rcl.py:22-51)."""
    x = np.asarray(image, dtype=F32)
    k = np.asarray(blur_kernel, dtype=F32).reshape(blur_kernel.shape[-2], blur_kernel.shape[-1])
    lam = F32(regularization_strength)
    out = np.empty_like(x)
    for n in range(x.shape[0]):
        for c in range(x.shape[1]):
            xc = x[n, c]
            d = xc.copy()
            for _ in range(num_iterations):
                blurred = _corr2d_zero_pad(d, k)
                with np.errstate(divide="ignore", invalid="ignore"):
                    cf = (xc / blurred).astype(F32)
                cf[cf != cf] = 0          # rcl.py:39
                cf[cf < 0] = 0            # rcl.py:40
                reg = (d + lam * _corr2d_zero_pad(d, _LAPLACIAN)).astype(F32)
                with np.errstate(invalid="ignore"):
                    d = (cf * reg).astype(F32)
            out[n, c] = d
    return out
