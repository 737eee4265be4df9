"""torch-CPU restatement of the reference hot path, used ONLY as the timed CPU baseline
(`bench.py` cpu_baseline / `--impl reference`) and as a second checker in tests.

It issues the same ATen operator sequence as /root/reference/model/SearchTransfer.py:24-51
(unfold -> normalize -> bmm -> max -> unfold x3 -> gather x3 -> fold x3 -> /9) so that its
wall time on the GPU box's host cores is a faithful stand-in for the reference module,
which cannot travel to the GPU box.  Test infrastructure: see oracle/__init__.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_PYRAMID = {3: dict(kernel_size=3, padding=1, stride=1),      # SearchTransfer.py:36,44
            2: dict(kernel_size=6, padding=2, stride=2),      # :37,45
            1: dict(kernel_size=12, padding=4, stride=4)}     # :38,46


def _aslist(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


@torch.no_grad()
def search_transfer_torch(lrsr_lv3, refsr_lv3, ref_lv1, ref_lv2, ref_lv3, query_slice=None):
    """Returns (S, T_lv3, T_lv2, T_lv1, arg).

    `query_slice=(lo, hi)` restricts the relevance search (bmm + max, the O(L*Lk) part) to
    query columns lo:hi -- the bounded sample used by bench.py; the other query columns get
    index 0 / S 0 so the remaining (linear-cost) stages still run at full size.
    Lists of reference tensors = several sharp frames, key sets concatenated.
    """
    n, _, h, w = lrsr_lv3.shape
    keys = torch.cat([F.unfold(r, 3, padding=1) for r in _aslist(refsr_lv3)], dim=2)   # :27
    keys = F.normalize(keys.permute(0, 2, 1), dim=2)                                   # :28,:30
    qcols = F.normalize(F.unfold(lrsr_lv3, 3, padding=1), dim=1)                       # :26,:31
    if query_slice is None:
        r_star, r_arg = torch.max(torch.bmm(keys, qcols), dim=1)                       # :33-34
    else:
        lo, hi = query_slice
        r_star = torch.zeros(n, h * w, dtype=qcols.dtype)
        r_arg = torch.zeros(n, h * w, dtype=torch.int64)
        r_star[:, lo:hi], r_arg[:, lo:hi] = torch.max(torch.bmm(keys, qcols[:, :, lo:hi]), dim=1)
    del keys, qcols
    gather_index = r_arg[:, None, :]
    outs = {}
    for lvl, refs in ((3, ref_lv3), (2, ref_lv2), (1, ref_lv1)):
        p = _PYRAMID[lvl]
        cols = torch.cat([F.unfold(r, **p) for r in _aslist(refs)], dim=2)             # :36-38
        picked = torch.gather(cols, 2, gather_index.expand(-1, cols.size(1), -1))      # :40-42 (bis)
        s = p["stride"]
        outs[lvl] = F.fold(picked, output_size=(h * s, w * s), **p) / (3. * 3.)        # :44-46
    return r_star.view(n, 1, h, w), outs[3], outs[2], outs[1], r_arg                   # :49,:51


@torch.no_grad()
def fuse_level_torch(dec, t, s, weight, bias, scale):
    """speinet.py:93-94 / 96-97 / 108-109."""
    up = s if scale == 1 else F.interpolate(s, scale_factor=scale, mode="bicubic")
    return dec + F.conv2d(torch.cat((dec, t), dim=1), weight, bias) * up
