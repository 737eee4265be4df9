"""numpy restatement of the reference SearchTransfer (test infrastructure only).

Every function cites the line(s) of /root/reference/model/SearchTransfer.py it
follows.  fp32 throughout unless stated.  See oracle/__init__.py for the rules
on who may import this file.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- #
# F.unfold  (SearchTransfer.py:26-27, 36-38)
# --------------------------------------------------------------------------- #
def unfold(x: np.ndarray, k: int, pad: int, stride: int) -> np.ndarray:
    """`F.unfold(x, (k,k), padding=pad, stride=stride)`.

    out[n, c*k*k + ki*k + kj, hc*Wc + wc] = xpad[n, c, hc*stride + ki, wc*stride + kj]
    (zero padding), i.e. the row/column convention of SURVEY.md appendix B.
    """
    n, c, h, w = x.shape
    hc = (h + 2 * pad - k) // stride + 1
    wc = (w + 2 * pad - k) // stride + 1
    xp = np.zeros((n, c, h + 2 * pad, w + 2 * pad), dtype=x.dtype)
    xp[:, :, pad:pad + h, pad:pad + w] = x
    out = np.empty((n, c, k, k, hc, wc), dtype=x.dtype)
    for ki in range(k):
        for kj in range(k):
            out[:, :, ki, kj] = xp[:, :, ki:ki + stride * hc:stride, kj:kj + stride * wc:stride]
    return out.reshape(n, c * k * k, hc * wc)


# --------------------------------------------------------------------------- #
# F.fold + constant divide  (SearchTransfer.py:44-46)
# --------------------------------------------------------------------------- #
def fold(cols: np.ndarray, out_hw, k: int, pad: int, stride: int, order: str = "cpu") -> np.ndarray:
    """`F.fold(cols, output_size=out_hw, kernel_size=k, padding=pad, stride=stride)`.

    fp32 overlap-add.  `order` selects the summation order of the <=9 overlapping
    contributions per output pixel, which matters for bit-exactness:
      "cpu"  : ascending (ki,kj)   -- ATen/native/im2col.h:131-146 (CPU col2im)
      "cuda" : ascending patch origin (h_col,w_col) == descending (ki,kj)
               -- ATen/native/cuda/im2col.cuh:139-154 (CUDA col2im_device)
    """
    n, ckk, l = cols.shape
    c = ckk // (k * k)
    h, w = out_hw
    hc = (h + 2 * pad - k) // stride + 1
    wc = (w + 2 * pad - k) // stride + 1
    assert hc * wc == l, (hc, wc, l)
    cols = cols.reshape(n, c, k, k, hc, wc)
    acc = np.zeros((n, c, h + 2 * pad, w + 2 * pad), dtype=F32)
    offsets = [(ki, kj) for ki in range(k) for kj in range(k)]
    if order == "cuda":
        offsets = offsets[::-1]
    elif order != "cpu":
        raise ValueError(order)
    for ki, kj in offsets:
        acc[:, :, ki:ki + stride * hc:stride, kj:kj + stride * wc:stride] += cols[:, :, ki, kj]
    return np.ascontiguousarray(acc[:, :, pad:pad + h, pad:pad + w])


def divide9(x: np.ndarray, mode: str = "cpu") -> np.ndarray:
    """The `/ (3.*3.)` of SearchTransfer.py:44-46.

    "cpu"  : true fp32 division (what torch CPU does).
    "cuda" : x * fp32(1/9)      (torch CUDA `div` by a Python scalar multiplies
             by the reciprocal; SURVEY.md section 7 hard part 4).
    """
    if mode == "cpu":
        return (x / F32(9.0)).astype(F32)
    if mode == "cuda":
        return (x * F32(1.0 / 9.0)).astype(F32)
    raise ValueError(mode)


# --------------------------------------------------------------------------- #
# F.normalize  (SearchTransfer.py:30-31)
# --------------------------------------------------------------------------- #
def l2_normalize(v: np.ndarray, axis: int, eps: float = 1e-12) -> np.ndarray:
    """`F.normalize(v, dim=axis)` = v / max(||v||_2, eps)."""
    nrm = np.sqrt(np.sum(v.astype(F32) ** 2, axis=axis, keepdims=True, dtype=F32))
    return (v / np.maximum(nrm, F32(eps))).astype(F32)


# --------------------------------------------------------------------------- #
# bmm + max  (SearchTransfer.py:33-34)
# --------------------------------------------------------------------------- #
def relevance(q_unf_n: np.ndarray, k_unf_n: np.ndarray, chunk: int = 4096):
    """R[n, j, i] = sum_k K^[n, k, j] * Q^[n, k, i]; returns (max_j R, first argmax_j).

    q_unf_n: [N, K, L] normalised query columns; k_unf_n: [N, K, Lk] normalised keys.
    The [Lk, L] matrix is produced in query chunks so the oracle stays in memory;
    torch.max(dim=1) returns the FIRST maximal index, as does np.argmax.
    """
    n, kk, l = q_unf_n.shape
    r_star = np.empty((n, l), dtype=F32)
    r_arg = np.empty((n, l), dtype=np.int64)
    for b in range(n):
        kt = np.ascontiguousarray(k_unf_n[b].T)  # [Lk, K]
        for s in range(0, l, chunk):
            r = kt @ q_unf_n[b][:, s:s + chunk]  # [Lk, chunk]
            r_arg[b, s:s + chunk] = np.argmax(r, axis=0)
            r_star[b, s:s + chunk] = np.max(r, axis=0)
    return r_star, r_arg


def bis(inp: np.ndarray, index: np.ndarray) -> np.ndarray:
    """`bis(input, 2, index)` (SearchTransfer.py:12-22): out[n,k,i] = input[n,k,index[n,i]]."""
    return np.take_along_axis(inp, index[:, None, :], axis=2)


# --------------------------------------------------------------------------- #
# SearchTransfer.forward  (SearchTransfer.py:24-51)
# --------------------------------------------------------------------------- #
_LEVELS = ((3, 1, 1), (6, 2, 2), (12, 4, 4))  # (kernel, pad, stride) lv3, lv2, lv1 -- :36-38,:44-46


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def search_transfer(lrsr_lv3, refsr_lv3, ref_lv1, ref_lv2, ref_lv3,
                    fold_order: str = "cpu", div_mode: str = "cpu", index=None):
    """Restatement of SearchTransfer.forward.  Returns (S, T_lv3, T_lv2, T_lv1, arg, R_star).

    The four reference arguments may each be a list of Rf arrays (one per sharp
    reference frame); their unfolded key sets are concatenated along the key axis
    (SURVEY.md F2 / section 8(c)), key index j = f*Hr*Wr + hr*Wr + wr.  With single arrays
    this is exactly the reference.  `index` overrides the argmax (used to test the
    gather/fold stage in isolation with the reference's own indices).
    """
    q = np.asarray(lrsr_lv3, dtype=F32)
    refsr = [np.asarray(a, dtype=F32) for a in _as_list(refsr_lv3)]
    refs = {3: [np.asarray(a, dtype=F32) for a in _as_list(ref_lv3)],
            2: [np.asarray(a, dtype=F32) for a in _as_list(ref_lv2)],
            1: [np.asarray(a, dtype=F32) for a in _as_list(ref_lv1)]}
    n, c, h, w = q.shape

    # :26-28  unfold query and keys
    q_unf = unfold(q, 3, 1, 1)                                            # [N, 9C, L]
    k_unf = np.concatenate([unfold(r, 3, 1, 1) for r in refsr], axis=2)   # [N, 9C, Lk]
    # :30-31  per-patch L2 normalisation
    k_unf = l2_normalize(k_unf, axis=1)
    q_unf = l2_normalize(q_unf, axis=1)
    # :33-34  relevance and hard-attention index
    if index is None:
        r_star, r_arg = relevance(q_unf, k_unf)
    else:
        r_arg = np.asarray(index, dtype=np.int64)
        r_star = np.einsum("nki,nki->ni", np.take_along_axis(k_unf, r_arg[:, None, :], axis=2), q_unf,
                           dtype=F32).astype(F32)
    outs = []
    for lvl, (kk, pad, st) in zip((3, 2, 1), _LEVELS):
        # :36-38 unfold the reference pyramid; :40-42 gather; :44-46 fold and /9
        ref_unf = np.concatenate([unfold(r, kk, pad, st) for r in refs[lvl]], axis=2)
        t_unf = bis(ref_unf, r_arg)
        scale = {3: 1, 2: 2, 1: 4}[lvl]
        t = fold(t_unf, (h * scale, w * scale), kk, pad, st, order=fold_order)
        outs.append(divide9(t, div_mode))
    s = r_star.reshape(n, 1, h, w)                                        # :49
    return s, outs[0], outs[1], outs[2], r_arg, r_star


def self_transfer_S(lrsr_lv3):
    """The search half of SelfTransfer.forward (SearchTransfer.py:59-72): keys are the
    query transposed and flipped (`x.transpose(2,3).flip(2)`), only S is used."""
    q = np.asarray(lrsr_lv3, dtype=F32)
    ref = np.ascontiguousarray(np.flip(np.swapaxes(q, 2, 3), axis=2))
    q_unf = l2_normalize(unfold(q, 3, 1, 1), axis=1)
    k_unf = l2_normalize(unfold(ref, 3, 1, 1), axis=1)
    r_star, r_arg = relevance(q_unf, k_unf)
    n, c, h, w = q.shape
    return r_star.reshape(n, 1, h, w), r_arg


# --------------------------------------------------------------------------- #
# Closed form of unfold -> gather -> fold -> /9   
# --------------------------------------------------------------------------- #
def closed_form_transfer(arg: np.ndarray, ref, scale: int, h: int, w: int,
                         fold_order: str = "cuda", div_mode: str = "cuda") -> np.ndarray:
    """T_s[n,c,y,x] = (sum over the <=9 query cells (Y+dy, X+dx) in the H x W grid of
    ref_s[n,c, y + (hr(q)-(Y+dy))*s, x + (wr(q)-(X+dx))*s]) / 9, Y=y//s, X=x//s,
    (hr,wr) = divmod(arg[n,q] % (Hr*Wr), Wr), frame f = arg // (Hr*Wr); out-of-image
    reads are 0.  This is the formula the CUDA gather/fold kernel evaluates; the
    test-suite checks it bit-for-bit against fold(bis(unfold(.))) above.
    """
    refs = [np.asarray(a, dtype=F32) for a in _as_list(ref)]
    n, c, hs, ws = refs[0].shape
    hr, wr = hs // scale, ws // scale
    arg = np.asarray(arg).reshape(n, h, w)
    out = np.zeros((n, c, h * scale, w * scale), dtype=F32)
    ys, xs = np.meshgrid(np.arange(h * scale), np.arange(w * scale), indexing="ij")
    cy, cx = ys // scale, xs // scale
    deltas = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]  # ascending (h_col,w_col) == "cuda"
    if fold_order == "cpu":
        deltas = deltas[::-1]                                       # ascending (ki,kj)
    refs_stacked = np.stack(refs, axis=1)  # [N, Rf, C, hs, ws]
    for b in range(n):
        for dy, dx in deltas:
            qy, qx = cy + dy, cx + dx
            inside = (qy >= 0) & (qy < h) & (qx >= 0) & (qx < w)
            j = arg[b][np.clip(qy, 0, h - 1), np.clip(qx, 0, w - 1)]
            f, rem = j // (hr * wr), j % (hr * wr)
            sy = ys + (rem // wr - qy) * scale
            sx = xs + (rem % wr - qx) * scale
            ok = inside & (sy >= 0) & (sy < hs) & (sx >= 0) & (sx < ws)
            vals = refs_stacked[b][f, :, np.clip(sy, 0, hs - 1), np.clip(sx, 0, ws - 1)]  # [H,W,C]
            vals = np.where(ok[..., None], vals, F32(0)).astype(F32)
            out[b] += np.moveaxis(vals, -1, 0)
    return divide9(out, div_mode)


# --------------------------------------------------------------------------- #
# Index comparison rule  (BASELINE.json north_star)
# --------------------------------------------------------------------------- #
def near_tie_agreement(lrsr_lv3, refsr_lv3, arg_a, arg_b, tol: float = 1e-5):
    """Per-query agreement mask: indices equal, or the relevances of the two candidate
    keys -- both recomputed here from the normalised patches (fp32 normalisation as
    the reference does it, fp64 dot) -- differ by less than `tol`.
    Returns (agree[N,L] bool, n_exact_equal, n_near_tie)."""
    q = np.asarray(lrsr_lv3, dtype=F32)
    refsr = [np.asarray(a, dtype=F32) for a in _as_list(refsr_lv3)]
    q_unf = l2_normalize(unfold(q, 3, 1, 1), axis=1).astype(np.float64)
    k_unf = l2_normalize(np.concatenate([unfold(r, 3, 1, 1) for r in refsr], axis=2), axis=1).astype(np.float64)
    arg_a = np.asarray(arg_a, dtype=np.int64).reshape(q.shape[0], -1)
    arg_b = np.asarray(arg_b, dtype=np.int64).reshape(q.shape[0], -1)
    ra = np.einsum("nki,nki->ni", np.take_along_axis(k_unf, arg_a[:, None, :], axis=2), q_unf)
    rb = np.einsum("nki,nki->ni", np.take_along_axis(k_unf, arg_b[:, None, :], axis=2), q_unf)
    equal = arg_a == arg_b
    near = (~equal) & (np.abs(ra - rb) < tol)
    return equal | near, int(equal.sum()), int(near.sum())
