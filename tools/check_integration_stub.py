#!/usr/bin/env python
"""Runs the minimal ctypes stub printed in INTEGRATION.md section 3 verbatim (extracted from the markdown) on the GPU box and
compares it with the drop-in module: the documented binding must keep working."""
import os, re, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT); sys.path.insert(0, ROOT)
md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
code = [b for b in re.findall(r"```python\n(.*?)```", md, flags=re.S) if "def search_transfer(q, k, ref1, ref2, ref3)" in b][0]
ns = {}
exec(code, ns)
torch.manual_seed(0)
n, h, w = 1, 20, 30
q = torch.randn(n, 128, h, w, device="cuda") * 0.2
k = (torch.randn(n, 1, 128, h, w, device="cuda") * 0.04).contiguous()
r2 = (torch.randn(n, 1, 64, 2 * h, 2 * w, device="cuda") * 0.04).contiguous()
r1 = (torch.randn(n, 1, 32, 4 * h, 4 * w, device="cuda") * 0.04).contiguous()
S, T3, T2, T1 = ns["search_transfer"](q, k, r1, r2, k)
torch.cuda.synchronize()
import speinet_b200
mS, m3, m2, m1 = speinet_b200.SearchTransfer().cuda()(q, k[:, 0], r1[:, 0], r2[:, 0], k[:, 0])
assert torch.equal(S, mS) and torch.equal(T3, m3) and torch.equal(T2, m2) and torch.equal(T1, m1)
print("INTEGRATION.md stub ok")
