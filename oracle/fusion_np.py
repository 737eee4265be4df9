"""numpy restatement of the three fusion lines of SPEINet._decode (test infrastructure only).

  /root/reference/model/speinet.py:93-94   f_lv3 = f + conv_lv3(cat(f, T3)) * S
  /root/reference/model/speinet.py:96-97   f_lv2 = d2 + conv_lv2(cat(d2, T2)) * bicubic_x2(S)
  /root/reference/model/speinet.py:108-109 f_lv1 = d1 + conv_lv1(cat(d1, T1)) * bicubic_x4(S)

`F.interpolate(mode='bicubic')` semantics follow torch/include/ATen/native/UpSample.h:
source index (dst+0.5)/scale-0.5 un-clamped for cubic (:289-300), A=-0.75
coefficients (:400-423), border-clamped taps.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
_A = -0.75


def _cc1(x):  # cubic_convolution1, UpSample.h:400-403
    return ((_A + 2) * x - (_A + 3)) * x * x + 1


def _cc2(x):  # cubic_convolution2, UpSample.h:405-408
    return ((_A * x - 5 * _A) * x + 8 * _A) * x - 4 * _A


def cubic_coeffs(t):
    """get_cubic_upsample_coefficients, UpSample.h:410-423."""
    t = np.asarray(t, dtype=F32)
    return np.stack([_cc2(t + 1), _cc1(t), _cc1(1 - t), _cc2(2 - t)], axis=-1).astype(F32)


def _taps(n_in: int, scale: int):
    dst = np.arange(n_in * scale, dtype=F32)
    src = (dst + F32(0.5)) * F32(1.0 / scale) - F32(0.5)     # area_pixel_compute_source_index, cubic=True
    i0 = np.floor(src)
    t = (src - i0).astype(F32)
    idx = np.clip(i0.astype(np.int64)[:, None] + np.arange(-1, 3)[None, :], 0, n_in - 1)  # bounded access
    return idx, cubic_coeffs(t)


def bicubic_upsample(x: np.ndarray, scale: int) -> np.ndarray:
    """`F.interpolate(x, scale_factor=scale, mode='bicubic')` (align_corners=False)."""
    x = np.asarray(x, dtype=F32)
    if scale == 1:
        return x.copy()
    n, c, h, w = x.shape
    iy, wy = _taps(h, scale)
    ix, wx = _taps(w, scale)
    # interpolate along x inside each of the four source rows, then along y (cubic_interp1d order)
    rows = x[:, :, iy, :]                                   # [N,C,Ho,4,W]
    g = rows[..., ix]                                       # [N,C,Ho,4,Wo,4]
    alongx = np.einsum("nchkwj,wj->nchkw", g, wx).astype(F32)
    return np.einsum("nchkw,hk->nchw", alongx, wy).astype(F32)


def conv1x1(x: np.ndarray, weight: np.ndarray, bias: np.ndarray) -> np.ndarray:
    """nn.Conv2d(kernel_size=1): out[n,o,y,x] = sum_i W[o,i] x[n,i,y,x] + b[o]  (speinet.py:55-57)."""
    wt = np.asarray(weight, dtype=F32).reshape(weight.shape[0], -1)
    out = np.einsum("oi,nihw->nohw", wt, np.asarray(x, dtype=F32)).astype(F32)
    return out + np.asarray(bias, dtype=F32)[None, :, None, None]


def fuse_level(dec, t, s, weight, bias, scale: int) -> np.ndarray:
    """dec + conv1x1(cat(dec, T)) * bicubic_up(S, scale)  -- speinet.py:93-94 / 96-97 / 108-109."""
    dec = np.asarray(dec, dtype=F32)
    cat = np.concatenate([dec, np.asarray(t, dtype=F32)], axis=1)
    return (dec + conv1x1(cat, weight, bias) * bicubic_upsample(s, scale)).astype(F32)
