#!/usr/bin/env python
"""Times the three gather/fold levels at 720p on random / smooth / identity match fields (GPU box)."""
import ctypes, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from speinet_b200 import _lib
import _util as U
lib = _lib.load()
h, w = 180, 320
shape = U.make_shape(1, h, w, h, w)
st = U.cur_stream()
ws, wsp, nbytes = U.alloc_workspace(shape)
ident = torch.arange(h * w, device="cuda", dtype=torch.int64)
yy, xx = ident // w, ident % w
torch.manual_seed(1)
jit = lambda m: torch.randint(-m, m + 1, (h * w,), device="cuda")
fields = {
    "random": torch.randint(0, h * w, (1, h * w), device="cuda", dtype=torch.int32),
    "smooth2": ((yy + jit(2)).clamp(0, h - 1) * w + (xx + jit(2)).clamp(0, w - 1)).to(torch.int32)[None].contiguous(),
    "identity": ident.to(torch.int32)[None].contiguous(),
}
res = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, arg in fields.items():
    for lvl, c, s in ((3, 128, 1), (2, 64, 2), (1, 32, 4)):
        ref = torch.randn(1, 1, c, s * h, s * w, device="cuda")
        out = torch.empty(1, c, s * h, s * w, device="cuda")
        fn = lambda: _lib.check(lib.spei_gather_fold(ctypes.byref(shape), lvl, U.vp(arg), U.vp(ref), U.vp(out), ctypes.c_void_p(0),
                                                     ctypes.c_void_p(wsp), nbytes, st), "gf")
        for _ in range(2): fn()
        # cold L2 per call, host launch latency hidden: 10 x (flush + call) against 10 x flush, each in one timed region
        def region(with_fn):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                flush.zero_()
                if with_fn: fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 100.0   # us per iteration
        ts = sorted(region(True) - region(False) for _ in range(3))
        res[f"{name}_lv{lvl}_us"] = round(ts[1], 1)
print(json.dumps(res))
