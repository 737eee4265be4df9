#!/usr/bin/env python
"""Launch the gather/fold kernels a few times on a smooth match field (for ncu captures)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from speinet_b200 import _lib  # noqa: E402
import _util as U  # noqa: E402

lib = _lib.load()
h, w = 180, 320
shape = U.make_shape(1, h, w, h, w)
st = U.cur_stream()
ws, wsp, nbytes = U.alloc_workspace(shape)
ident = torch.arange(h * w, device="cuda", dtype=torch.int64)
yy, xx = ident // w, ident % w
torch.manual_seed(1)
jit = lambda m: torch.randint(-m, m + 1, (h * w,), device="cuda")
field = sys.argv[1] if len(sys.argv) > 1 else "smooth"
if field == "smooth":
    arg = ((yy + jit(2)).clamp(0, h - 1) * w + (xx + jit(2)).clamp(0, w - 1)).to(torch.int32)[None].contiguous()
else:
    arg = torch.randint(0, h * w, (1, h * w), device="cuda", dtype=torch.int32)
for lvl, c, s in ((3, 128, 1), (2, 64, 2), (1, 32, 4)):
    ref = torch.randn(1, 1, c, s * h, s * w, device="cuda")
    out = torch.empty(1, c, s * h, s * w, device="cuda")
    for _ in range(3):
        _lib.check(lib.spei_gather_fold(ctypes.byref(shape), lvl, U.vp(arg), U.vp(ref), U.vp(out), ctypes.c_void_p(0),
                                        ctypes.c_void_p(wsp), nbytes, st), "gf")
torch.cuda.synchronize()
print("ok")
