"""torch-CPU restatement of the reference hot path, used ONLY as the timed CPU baseline
(`bench.py` cpu_baseline / `--impl reference`) and as a second checker in tests.

It issues the same ATen operator sequence as /root/reference/model/SearchTransfer.py:24-51
(unfold -> normalize -> bmm -> max -> unfold x3 -> gather x3 -> fold x3 -> /9) so that its
wall time on the GPU box's host cores is a faithful stand-in for the reference module,
which cannot travel to the GPU box.  The sequence is split into its key-side, query-side and
fold parts so the benchmark can time a bounded sample (a slice of the query columns) and
extrapolate.  Test infrastructure: see oracle/__init__.py.
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
def key_side(refsr_lv3, ref_lv1, ref_lv2, ref_lv3):
    """Everything that does not depend on the queries: normalised key rows [N, Lk, 9C] (:27-28,:30)
    and the three unfolded reference pyramids (:36-38).  Lists = several sharp frames."""
    keys = torch.cat([F.unfold(r, 3, padding=1) for r in _aslist(refsr_lv3)], dim=2)
    keys = F.normalize(keys.permute(0, 2, 1), dim=2)
    cols = {lvl: torch.cat([F.unfold(r, **_PYRAMID[lvl]) for r in _aslist(refs)], dim=2)
            for lvl, refs in ((3, ref_lv3), (2, ref_lv2), (1, ref_lv1))}
    return keys, cols


@torch.no_grad()
def query_side(lrsr_lv3):
    """Normalised query columns [N, 9C, L] (:26,:31)."""
    return F.normalize(F.unfold(lrsr_lv3, 3, padding=1), dim=1)


@torch.no_grad()
def search_slice(keys, qcols, lo, hi):
    """bmm + max over keys for query columns lo:hi (:33-34)."""
    return torch.max(torch.bmm(keys, qcols[:, :, lo:hi]), dim=1)


@torch.no_grad()
def gather_slice(cols, r_arg):
    """The three `bis` gathers (:12-22, :40-42) for the given query columns."""
    idx = r_arg[:, None, :]
    return {lvl: torch.gather(c, 2, idx.expand(-1, c.size(1), -1)) for lvl, c in cols.items()}


@torch.no_grad()
def fold_all(picked, h, w):
    """The three overlap-adds and the constant /9 (:44-46)."""
    out = {}
    for lvl, t in picked.items():
        s = _PYRAMID[lvl]["stride"]
        out[lvl] = F.fold(t, output_size=(h * s, w * s), **_PYRAMID[lvl]) / (3. * 3.)
    return out


@torch.no_grad()
def search_transfer_torch(lrsr_lv3, refsr_lv3, ref_lv1, ref_lv2, ref_lv3):
    """Whole SearchTransfer.forward.  Returns (S, T_lv3, T_lv2, T_lv1, arg)."""
    n, _, h, w = lrsr_lv3.shape
    keys, cols = key_side(refsr_lv3, ref_lv1, ref_lv2, ref_lv3)
    qcols = query_side(lrsr_lv3)
    r_star, r_arg = search_slice(keys, qcols, 0, h * w)
    del keys, qcols
    outs = fold_all(gather_slice(cols, r_arg), h, w)
    return r_star.view(n, 1, h, w), outs[3], outs[2], outs[1], r_arg                   # :49,:51


@torch.no_grad()
def fuse_level_torch(dec, t, s, weight, bias, scale):
    """speinet.py:93-94 / 96-97 / 108-109."""
    up = s if scale == 1 else F.interpolate(s, scale_factor=scale, mode="bicubic")
    return dec + F.conv2d(torch.cat((dec, t), dim=1), weight, bias) * up
