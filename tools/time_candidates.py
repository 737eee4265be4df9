#!/usr/bin/env python
"""Times spei_relevance_candidates alone at 720p (kernel experiments; test infrastructure, GPU box)."""
import ctypes, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from speinet_b200 import _lib
import _util as U

def main():
    lib = _lib.load()
    torch.manual_seed(0)
    h, w = 180, 320
    search = {"tcs": _lib.SEARCH_TCS, "tc": _lib.SEARCH_TC}[sys.argv[1] if len(sys.argv) > 1 else "tcs"]
    q = torch.randn(1, 128, h, w, device="cuda") * 0.2
    k = (torch.randn(1, 1, 128, h, w, device="cuda") * 0.04).contiguous()
    shape = U.make_shape(1, h, w, h, w, search=search)
    ws, ptr, nbytes = U.alloc_workspace(shape)
    st = U.cur_stream(); wsp = ctypes.c_void_p(ptr)
    _lib.check(lib.spei_stage_norm(ctypes.byref(shape), U.vp(q), U.vp(k), wsp, nbytes, st), "stage")
    fn = lambda: _lib.check(lib.spei_relevance_candidates(ctypes.byref(shape), wsp, nbytes, st), "cand")
    for _ in range(3): fn()
    torch.cuda.synchronize()
    n = 15
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    flag = ctypes.c_int32(0)
    lib.spei_debug_error_flag(ctypes.byref(shape), wsp, nbytes, st, ctypes.byref(flag))
    print(json.dumps({"lib": os.environ.get("SPEINET_B200_LIB", "default"), "median_ms": ts[n // 2], "min_ms": ts[0], "max_ms": ts[-1], "error_flag": flag.value}))

main()
