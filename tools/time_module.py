#!/usr/bin/env python
"""Latency of one `speinet_b200.SearchTransfer` call (the drop-in module at speinet.py:135), eager vs cuda_graph=True, at the
256x256 (64x64 grid) and 1280x720 (180x320 grid) shapes: host time per call (how long the Python call blocks the CPU) and
device time per call (CUDA events around 50 back-to-back calls).  One JSON object on stdout (profiles/r02_module_latency.json)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import speinet_b200  # noqa: E402

torch.cuda.set_device(0)
out = {}
for name, (h, w) in (("256x256", (64, 64)), ("1280x720", (180, 320))):
    g = torch.Generator(device="cuda").manual_seed(1)
    rn = lambda *s, std: torch.randn(*s, device="cuda", generator=g) * std
    q, lv3 = rn(1, 128, h, w, std=0.2), rn(1, 128, h, w, std=0.04)
    lv2, lv1 = rn(1, 64, 2 * h, 2 * w, std=0.04), rn(1, 32, 4 * h, 4 * w, std=0.04)
    res = {}
    ref = None
    for mode in ("eager", "cuda_graph"):
        m = speinet_b200.SearchTransfer(cuda_graph=(mode == "cuda_graph")).cuda()
        with torch.no_grad():
            for _ in range(5):
                o = m(q, lv3, lv1, lv2, lv3)
            torch.cuda.synchronize()
            n = 50
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            for _ in range(n):
                o = m(q, lv3, lv1, lv2, lv3)
            b.record()
            t_host = time.perf_counter() - t0
            torch.cuda.synchronize()
            t_wall = time.perf_counter() - t0
        res[mode] = {"host_us_per_call": t_host / n * 1e6, "device_us_per_call": a.elapsed_time(b) / n * 1e3, "wall_us_per_call": t_wall / n * 1e6}
        if ref is None:
            ref = [t.clone() for t in o]
        else:
            res["graph_equals_eager"] = all(bool(torch.equal(x, y)) for x, y in zip(ref, o))
    out[name] = res
print(json.dumps(out, indent=1))
