"""Shared helpers for the GPU parity tests and tests/diag/gpu_diag.py (test infrastructure)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from speinet_b200 import _lib


def make_shape(n, h, w, hr, wr, rf=1, fold_mode=_lib.FOLD_CUDA, search=_lib.SEARCH_TCS, eps=0.0):
    return _lib.SpeiShape(n=n, h=h, w=w, hr=hr, wr=wr, rf=rf, c3=128, c2=64, c1=32, fold_mode=fold_mode,
                          search=search, eps=eps)


def plan_info(shape):
    out = (ctypes.c_int32 * 16)()
    _lib.check(_lib.load().spei_plan_info(ctypes.byref(shape), out), "spei_plan_info")
    keys = ["q_orient", "q_tu", "q_tv", "q_Upad", "q_Vpad", "k_orient", "k_tu", "k_tv", "k_Ny", "k_Upad", "k_Vpad",
            "QT", "KT", "G", "maxseg", "num_sms"]
    return dict(zip(keys, [int(v) for v in out]))


def alloc_workspace(shape, device="cuda"):
    n = ctypes.c_size_t(0)
    _lib.check(_lib.load().spei_workspace_bytes(ctypes.byref(shape), ctypes.byref(n)), "spei_workspace_bytes")
    ws = torch.zeros(int(n.value) + 256, dtype=torch.uint8, device=device)
    ptr = (ws.data_ptr() + 255) // 256 * 256
    return ws, ptr, int(n.value)


def cur_stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def vp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def bf16_round(a: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float32).numpy()


def expected_debug_tile(q, k, info):
    """Raw (un-normalised) bf16-operand dot products of query tile 0 x key tile 0 of item 0, frame 0,
    as the tcgen05 kernel lays them out: row m = (v=m//8, u=m%8) of the query tile, column c likewise of
    the key tile, (u,v) per operand orientation.  fp64 accumulation of exactly-representable products."""
    qb = bf16_round(q[0]).astype(np.float64)          # [C,H,W]
    kb = bf16_round(k[0]).astype(np.float64)
    qp = np.pad(qb, ((0, 0), (1, 1), (1, 1)))
    kp = np.pad(kb, ((0, 0), (1, 1), (1, 1)))

    def patches(img_p, H, W, orient, rows):
        out = np.zeros((rows * 8, img_p.shape[0] * 9))
        for m in range(rows * 8):
            u, v = m % 8, m // 8
            x, y = (u, v) if orient == 0 else (v, u)
            if x < W and y < H:
                out[m] = img_p[:, y:y + 3, x:x + 3].reshape(-1)
        return out

    A = patches(qp, q.shape[2], q.shape[3], info["q_orient"], 16)
    B = patches(kp, k.shape[2], k.shape[3], info["k_orient"], info["k_Ny"])
    return A @ B.T  # [128, 8*Ny]


def expected_debug_tile_tcs(q, k, info):
    """Tap-sharing kernel (relevance_tcs.cu): raw accumulator D of query tile 0 x key tile 0 = the bf16-operand dot
    over the 3 taps along v and the 128 channels.  Row m = (v = m // 32, u = m % 32 - 1), column c likewise for the
    key tile (u = -1 and u = 30 are halo positions; positions outside the image contribute zeros)."""
    qb = bf16_round(q[0]).astype(np.float64)
    kb = bf16_round(k[0]).astype(np.float64)

    def vcols(img, orient, rows):
        C, H, W = img.shape
        img_uv = img if orient == 0 else img.transpose(0, 2, 1)   # [C, V, U]
        V, Uu = img_uv.shape[1:]
        pad = np.zeros((C, rows + 2, 32))
        for vv in range(-1, rows + 1):
            for uu in range(-1, 31):
                if 0 <= vv < V and 0 <= uu < Uu:
                    pad[:, vv + 1, uu + 1] = img_uv[:, vv, uu]
        out = np.zeros((rows * 32, C * 3))
        for m in range(rows * 32):
            v, ub = m // 32, m % 32
            out[m] = pad[:, v:v + 3, ub].reshape(-1)
        return out

    A = vcols(qb, info["q_orient"], 4)
    B = vcols(kb, info["k_orient"], info["k_Ny"])
    return A @ B.T  # [128, 32*Ny]


def run_search(q, k, search=_lib.SEARCH_TCS, eps=0.0):
    """Stage + relevance only, through the C-ABI.  q [N,128,H,W], k [N,Rf,128,Hr,Wr] CUDA fp32.
    Returns (S [N,1,H,W], arg32 [N,L], stats[8], error_flag)."""
    lib = _lib.load()
    n, _, h, w = q.shape
    _, rf, _, hr, wr = k.shape
    shape = make_shape(n, h, w, hr, wr, rf, search=search, eps=eps)
    ws, ptr, nbytes = alloc_workspace(shape)
    S = torch.empty((n, 1, h, w), device="cuda")
    arg32 = torch.empty((n, h * w), dtype=torch.int32, device="cuda")
    stats = torch.zeros(_lib.STATS_WORDS, dtype=torch.int32, device="cuda")
    st = cur_stream()
    _lib.check(lib.spei_stage_norm(ctypes.byref(shape), vp(q), vp(k), ctypes.c_void_p(ptr), nbytes, st), "stage_norm")
    _lib.check(lib.spei_relevance_argmax(ctypes.byref(shape), vp(S), vp(arg32), ctypes.c_void_p(0), vp(stats),
                                         ctypes.c_void_p(ptr), nbytes, st), "relevance_argmax")
    flag = ctypes.c_int32(0)
    _lib.check(lib.spei_debug_error_flag(ctypes.byref(shape), ctypes.c_void_p(ptr), nbytes, st, ctypes.byref(flag)), "error_flag")
    torch.cuda.synchronize()
    return S, arg32, stats, int(flag.value)


def run_debug_tile(q, k, search=_lib.SEARCH_TC):
    lib = _lib.load()
    n, _, h, w = q.shape
    _, rf, _, hr, wr = k.shape
    shape = make_shape(n, h, w, hr, wr, rf, search=search)
    ws, ptr, nbytes = alloc_workspace(shape)
    acc = torch.full((128, 256), float("nan"), device="cuda")
    st = cur_stream()
    _lib.check(lib.spei_stage_norm(ctypes.byref(shape), vp(q), vp(k), ctypes.c_void_p(ptr), nbytes, st), "stage_norm")
    _lib.check(lib.spei_debug_relevance_tile(ctypes.byref(shape), vp(acc), ctypes.c_void_p(ptr), nbytes, st), "debug_tile")
    flag = ctypes.c_int32(0)
    _lib.check(lib.spei_debug_error_flag(ctypes.byref(shape), ctypes.c_void_p(ptr), nbytes, st, ctypes.byref(flag)), "error_flag")
    torch.cuda.synchronize()
    return acc.cpu().numpy(), plan_info(shape), int(flag.value)


def gather_fold(arg32, ref, level, n, h, w, hr, wr, rf, fold_mode):
    lib = _lib.load()
    shape = make_shape(n, h, w, hr, wr, rf, fold_mode=fold_mode)
    s = {3: 1, 2: 2, 1: 4}[level]
    c = ref.shape[2]
    out = torch.empty((n, c, s * h, s * w), device="cuda")
    ws, ptr, nbytes = alloc_workspace(shape)
    _lib.check(lib.spei_gather_fold(ctypes.byref(shape), level, vp(arg32), vp(ref), vp(out), ctypes.c_void_p(0),
                                    ctypes.c_void_p(ptr), nbytes, cur_stream()), "gather_fold")
    torch.cuda.synchronize()
    return out
