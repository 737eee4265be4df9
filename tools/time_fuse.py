#!/usr/bin/env python
"""Runs the three fusion levels at 720p a few times (kernel experiments, GPU box)."""
import sys, os, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import speinet_b200
h, w = 180, 320
torch.manual_seed(0)
S = torch.rand(1, 1, h, w, device="cuda") * 0.2
res = {}
only = [int(a) for a in sys.argv[1:]] or [3, 2, 1]
for lvl, c, sc in ((3, 128, 1), (2, 64, 2), (1, 32, 4)):
    if lvl not in only: continue
    dec = torch.randn(1, c, sc * h, sc * w, device="cuda"); tt = torch.randn(1, c, sc * h, sc * w, device="cuda")
    wgt = torch.randn(c, 2 * c, 1, 1, device="cuda") * 0.05; b = torch.randn(c, device="cuda")
    n = int(os.environ.get("FUSE_ITERS", "5"))
    for _ in range(2): speinet_b200.fuse_level(dec, tt, S, wgt, b, sc)
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): speinet_b200.fuse_level(dec, tt, S, wgt, b, sc)
    e.record(); torch.cuda.synchronize()
    res[f"lv{lvl}_us"] = round(a.elapsed_time(e) / n * 1e3, 1)
print(json.dumps(res))
