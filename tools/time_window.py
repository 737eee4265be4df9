#!/usr/bin/env python
"""Cost of the certified candidate window against the fixed 2e-3 window of round 1, 720p, on the bench's randn features and on
image-like (smooth) features: event-timed tcgen05 pass and exactness layer, counters, and agreement with the exhaustive
fp32 search.  Writes one JSON object to stdout (profiles/r02_window_cost.json)."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from speinet_b200 import _lib  # noqa: E402
import _util as U  # noqa: E402
import test_gpu_fullsize as FS  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
out = {}
for kind in ("randn", "image_like", "image_like_x16_lownoise"):
    if kind == "image_like_x16_lownoise":
        g = torch.Generator(device="cuda").manual_seed(3)
        low = torch.randn(1, 128, 13, 21, device="cuda", generator=g)
        base = torch.nn.functional.interpolate(low, scale_factor=16, mode="bicubic")[:, :, :bench.H, :bench.W].contiguous()
        q = (base + 0.002 * torch.randn(base.shape, device="cuda", generator=g)) * 0.2
        k = (base + 0.002 * torch.randn(base.shape, device="cuda", generator=g)) * 0.04
    else:
        q, lv3, _, _ = FS.make_features(kind, 1, bench.H, bench.W, 1, seed=7)
        k = lv3[0]
    k5 = k.unsqueeze(1).contiguous()
    S0, a0, _, _ = U.run_search(q, k5, search=_lib.SEARCH_EXACT)
    res = {}
    for name, eps in (("fixed_2e-3", 2e-3), ("certified", 0.0)):
        P = bench.DevicePath(dev, "tcs", eps=eps)
        P.stage(q, k5, stream)

        def timed(fn, iters=8):
            for _ in range(2):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / iters
        t_c = timed(lambda: P.candidates(stream))
        t_r = timed(lambda: P.rescore(stream))
        diff = (P.arg32.view(1, -1) != a0.view(1, -1))
        worst = float((P.S.view(1, -1)[diff] - S0.view(1, -1)[diff]).abs().max()) if diff.any() else 0.0
        res[name] = {"candidates_ms": t_c, "exactness_ms": t_r, "stats": dict(zip(_lib.STATS_NAMES, P.stats.cpu().tolist())),
                     "indices_differing_from_exhaustive": int(diff.sum()), "worst_S_gap_on_differing": worst,
                     "S_max_abs_diff": float((P.S - S0).abs().max()), "S_mean": float(S0.mean())}
        del P
    res["certified_over_fixed"] = (res["certified"]["candidates_ms"] + res["certified"]["exactness_ms"]) / (
        res["fixed_2e-3"]["candidates_ms"] + res["fixed_2e-3"]["exactness_ms"])
    out[kind] = res
print(json.dumps(out, indent=1))
