#!/usr/bin/env python
"""Times spei_stage_norm and spei_rescore alone at 720p (kernel experiments, GPU box)."""
import ctypes, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from speinet_b200 import _lib
import _util as U
lib = _lib.load()
torch.manual_seed(0)
h, w = 180, 320
q = torch.randn(1, 128, h, w, device="cuda") * 0.2
k = (torch.randn(1, 1, 128, h, w, device="cuda") * 0.04).contiguous()
shape = U.make_shape(1, h, w, h, w)
ws, ptr, nbytes = U.alloc_workspace(shape)
st = U.cur_stream(); wsp = ctypes.c_void_p(ptr)
S = torch.empty(1, 1, h, w, device="cuda"); arg32 = torch.empty(1, h * w, dtype=torch.int32, device="cuda")
stats = torch.zeros(8, dtype=torch.int32, device="cuda")
stage = lambda: _lib.check(lib.spei_stage_norm(ctypes.byref(shape), U.vp(q), U.vp(k), wsp, nbytes, st), "stage")
stage()
_lib.check(lib.spei_relevance_candidates(ctypes.byref(shape), wsp, nbytes, st), "cand")
resc = lambda: _lib.check(lib.spei_rescore(ctypes.byref(shape), U.vp(S), U.vp(arg32), ctypes.c_void_p(0), U.vp(stats), wsp, nbytes, st), "rescore")
def timed(fn, n=20):
    for _ in range(3): fn()
    g = torch.cuda.CUDAGraph()           # graph replay: GPU time without per-call host latency
    s = torch.cuda.Stream()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize(); ev[0].record()
    for _ in range(n): fn()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n * 1e3
print(json.dumps({"rescore_us": round(timed(resc), 1), "stage_us": round(timed(stage), 1), "stats": stats.cpu().tolist()}))
