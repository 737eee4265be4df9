#!/usr/bin/env python
"""Where does fuse_level differ from the golden model_forward fixture?  (test infrastructure, GPU box)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import speinet_b200
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "model_forward.npz")))
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for lvl, scale in ((3, 1), (2, 2), (1, 4)):
    f = speinet_b200.fuse_level(cu(g[f"dec{lvl}"]), cu(g[f"t{lvl}"]) if f"t{lvl}" in g else cu(g[f"T_lv{lvl}"]), cu(g["S"]), cu(g[f"w{lvl}"]), cu(g[f"b{lvl}"]), scale).cpu().numpy()
    want = g[f"f{lvl}"]
    err = np.abs(f - want)[0]
    C, H, W = err.shape
    e2 = err.reshape(C, -1)
    ntile = (H * W + 127) // 128
    print("level", lvl, err.shape, "max", err.max())
    print(" per pixel tile:", [float(f"{e2[:, i*128:(i+1)*128].max():.2e}") for i in range(min(ntile, 12))])
    print(" per 16-channel block:", [float(f"{e2[i*16:(i+1)*16].max():.2e}") for i in range(C // 16)])
    print(" per lane%32 (pixel within tile, first tile):", [float(f"{e2[:, j:128:32].max():.1e}") for j in range(0, 32, 4)])
