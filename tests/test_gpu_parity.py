"""Parity of the CUDA hot path against the oracle, through the C-ABI / the drop-in module.
Run on the B200 box: python -m pytest tests -m gpu.

Bars (BASELINE.json north_star): argmax indices bit-exact except near-ties with
|delta relevance| < 1e-5 (fp32 normalisation, fp64 dot); given identical indices gather/fold is
bit-exact; R_star (S) and fused features within 1e-4 relative in fp32, 1e-2 for bf16 inputs.
"""
import ctypes

import numpy as np
import pytest
import torch

import oracle
from oracle.torch_port import search_transfer_torch, fuse_level_torch
from speinet_b200 import _lib
import speinet_b200
import _util as U

pytestmark = pytest.mark.gpu

GOLDEN_CASES = ["st_same_grid", "st_ragged", "st_edge"]
RTOL_S = 1e-4


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def assert_indices_agree(q, k_list, got, want):
    agree, n_eq, n_tie = oracle.near_tie_agreement(q, k_list, got, want)
    assert agree.all(), f"{int((~agree).sum())} of {agree.size} argmax indices differ beyond the 1e-5 near-tie rule"
    return n_tie


def test_library_is_the_native_one():
    lib = speinet_b200.load_library()
    assert lib.spei_version() == _lib.VERSION
    info = U.plan_info(U.make_shape(1, 180, 320, 180, 320))
    assert info["num_sms"] >= 100 and info["G"] == info["num_sms"]


def test_torch_cuda_divides_by_multiplying_reciprocal():
    """SURVEY section 7 hard part 4: pins the SPEI_FOLD_CUDA default."""
    x = torch.randn(1 << 16, device="cuda")
    assert torch.equal(x / (3. * 3.), x * torch.tensor(1.0 / 9.0, device="cuda"))


# ------------------------------------------------------------------ (a)+(b) search ---------------
SEARCH_MODES = {"tcs": _lib.SEARCH_TCS, "tc": _lib.SEARCH_TC, "exact": _lib.SEARCH_EXACT}
TC_MODES = ["tcs", "tc"]  # the two tcgen05 candidate passes: tap-sharing (default) and dense


@pytest.mark.parametrize("search", ["tcs", "tc", "exact"])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_search_matches_reference_golden(golden, name, search):
    g = golden(name)
    q, k = g["q"], g["ref_lv3"]
    S, arg32, stats, flag = U.run_search(cu(q), cu(k).unsqueeze(1).contiguous(), search=SEARCH_MODES[search])
    assert flag == 0
    assert_indices_agree(q, k, arg32.cpu().numpy(), g["arg"])
    np.testing.assert_allclose(S.cpu().numpy(), g["S"], rtol=RTOL_S, atol=1e-6)


@pytest.mark.parametrize("search", TC_MODES)
def test_edge_semantics_zero_patch_and_duplicate_keys(golden, search):
    g = golden("st_edge")
    S, arg32, _, _ = U.run_search(cu(g["q"]), cu(g["ref_lv3"]).unsqueeze(1).contiguous(), search=SEARCH_MODES[search])
    arg = arg32.cpu().numpy().reshape(8, 16)
    S = S.cpu().numpy()[0, 0]
    assert (arg[0:2, 0:2] == 0).all() and (S[0:2, 0:2] == 0).all()      # zero query patch -> index 0, S = 0
    assert np.array_equal(arg[4:, 9:15], arg[4:, 1:7]) and (arg[:, 9:15] % 16 < 8).all()  # first duplicate wins
    assert np.array_equal(arg, g["arg"].reshape(8, 16))


# grids around the tile sizes of both kernels (8 / 16 dense, 30 / 4 / 8 tap-sharing), ragged, tiny, two frames
@pytest.mark.parametrize("search", TC_MODES)
@pytest.mark.parametrize("shape", [(1, 37, 50, 29, 44, 1), (2, 24, 40, 24, 40, 1), (1, 33, 21, 40, 35, 2), (1, 64, 64, 64, 64, 1),
                                   (1, 30, 30, 31, 61, 1), (1, 4, 60, 9, 29, 1), (1, 1, 1, 1, 1, 1), (1, 2, 3, 1, 7, 1),
                                   (1, 5, 91, 17, 32, 2)])
def test_tc_search_matches_oracle_random(shape, search):
    n, h, w, hr, wr, rf = shape
    rng = np.random.default_rng(hash(shape) % (1 << 31))
    q = (rng.standard_normal((n, 128, h, w)) * 0.2).astype(np.float32)
    ks = [(rng.standard_normal((n, 128, hr, wr)) * 0.04).astype(np.float32) for _ in range(rf)]
    qu = oracle.l2_normalize(oracle.unfold(q, 3, 1, 1), axis=1)
    ku = oracle.l2_normalize(np.concatenate([oracle.unfold(k, 3, 1, 1) for k in ks], axis=2), axis=1)
    want_S, want_arg = oracle.relevance(qu, ku)
    S, arg32, stats, flag = U.run_search(cu(q), torch.stack([cu(k) for k in ks], dim=1).contiguous(), search=SEARCH_MODES[search])
    assert flag == 0
    assert_indices_agree(q, ks, arg32.cpu().numpy(), want_arg)
    np.testing.assert_allclose(S.cpu().numpy().reshape(n, -1), want_S, rtol=RTOL_S, atol=1e-6)


@pytest.mark.parametrize("search", TC_MODES)
def test_tc_search_random_shape_sweep_vs_exhaustive_fp32(search):
    """24 seeded random problem shapes (batch 1-3, 1-3 reference frames, grids from 1x1 to ~100x150, query and reference
    grids unrelated) through both tcgen05 engines against the exhaustive fp32 search on the same GPU: every persistent-CTA
    split, tile-padding and barrier-phase pattern the planner can produce at these sizes."""
    rng = np.random.default_rng(2024)
    for case in range(24):
        n, rf = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        h, w, hr, wr = (int(rng.integers(1, hi)) for hi in ((20, 40, 20, 40) if case % 3 else (100, 150, 100, 150)))
        q = torch.from_numpy((rng.standard_normal((n, 128, h, w)) * 0.2).astype(np.float32)).cuda()
        k = torch.from_numpy((rng.standard_normal((n, rf, 128, hr, wr)) * 0.04).astype(np.float32)).cuda()
        S0, a0, _, f0 = U.run_search(q, k, search=_lib.SEARCH_EXACT)
        S1, a1, st1, f1 = U.run_search(q, k, search=SEARCH_MODES[search])
        assert f0 == 0 and f1 == 0, (case, n, rf, h, w, hr, wr)
        diff = (a0 != a1)
        if diff.any():   # differing indices must be near-ties: compare the exact scores both engines report
            assert float((S0.reshape(n, -1)[diff] - S1.reshape(n, -1)[diff]).abs().max()) < 1e-5, (case, n, rf, h, w, hr, wr)
        torch.testing.assert_close(S1, S0, rtol=RTOL_S, atol=1e-6, msg=lambda m: f"case {case} {(n, rf, h, w, hr, wr)}: {m}")


@pytest.mark.parametrize("search", TC_MODES)
def test_tc_search_smooth_features_many_near_candidates(search):
    """Image-like (spatially smooth) features put many keys inside the candidate window; the
    saturated-list -> exhaustive fp32 fallback must keep the result exact."""
    rng = np.random.default_rng(9)
    base = rng.standard_normal((1, 128, 10, 12)).astype(np.float32)
    up = torch.nn.functional.interpolate(torch.from_numpy(base), scale_factor=4, mode="bicubic").numpy()
    q = (up + 0.01 * rng.standard_normal(up.shape)).astype(np.float32)
    k = (up + 0.01 * rng.standard_normal(up.shape)).astype(np.float32)
    qu = oracle.l2_normalize(oracle.unfold(q, 3, 1, 1), axis=1)
    ku = oracle.l2_normalize(oracle.unfold(k, 3, 1, 1), axis=1)
    want_S, want_arg = oracle.relevance(qu, ku)
    S, arg32, stats, flag = U.run_search(cu(q), cu(k).unsqueeze(1).contiguous(), search=SEARCH_MODES[search])
    assert flag == 0
    assert_indices_agree(q, k, arg32.cpu().numpy(), want_arg)
    np.testing.assert_allclose(S.cpu().numpy().reshape(1, -1), want_S, rtol=RTOL_S, atol=1e-6)


@pytest.mark.parametrize("search", TC_MODES)
def test_debug_tile_accumulator_matches_bf16_dot(search):
    """The raw tcgen05 accumulator of (query tile 0, key tile 0) equals the bf16-operand dot products
    (all nine taps for the dense kernel, the three v taps for the tap-sharing kernel)."""
    rng = np.random.default_rng(1)
    q = rng.standard_normal((1, 128, 20, 24)).astype(np.float32)
    k = rng.standard_normal((1, 128, 20, 24)).astype(np.float32)
    acc, info, flag = U.run_debug_tile(cu(q), cu(k).unsqueeze(1).contiguous(), search=SEARCH_MODES[search])
    want = (U.expected_debug_tile_tcs if search == "tcs" else U.expected_debug_tile)(q, k, info)
    assert flag == 0
    np.testing.assert_allclose(acc[:, :want.shape[1]], want, rtol=0, atol=2e-3)


# ------------------------------------------------------------------ (c) gather / fold -----------
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_gather_fold_bit_exact_given_reference_indices(golden, name):
    g = golden(name)
    n, _, h, w = g["q"].shape
    hr, wr = g["ref_lv3"].shape[2:]
    arg32 = cu(g["arg"].astype(np.int32))
    idx64 = cu(g["arg"].astype(np.int64))
    # torch CUDA runs the same ATen ops as the reference: the bit-exact comparator for the default mode
    tq = cu(g["q"])
    for lvl, key, gk in ((3, "ref_lv3", "T_lv3"), (2, "ref_lv2", "T_lv2"), (1, "ref_lv1", "T_lv1")):
        ref = cu(g[key]).unsqueeze(1).contiguous()
        got_cpu_mode = U.gather_fold(arg32, ref, lvl, n, h, w, hr, wr, 1, _lib.FOLD_CPU)
        assert np.array_equal(got_cpu_mode.cpu().numpy(), g[gk]), f"lv{lvl}: CPU-order fold differs from the reference golden"
        p = {3: dict(kernel_size=3, padding=1, stride=1), 2: dict(kernel_size=6, padding=2, stride=2),
             1: dict(kernel_size=12, padding=4, stride=4)}[lvl]
        cols = torch.nn.functional.unfold(cu(g[key]), **p)
        picked = torch.gather(cols, 2, idx64[:, None, :].expand(-1, cols.size(1), -1))
        s = p["stride"]
        want = torch.nn.functional.fold(picked, output_size=(h * s, w * s), **p) / (3. * 3.)
        got = U.gather_fold(arg32, ref, lvl, n, h, w, hr, wr, 1, _lib.FOLD_CUDA)
        assert torch.equal(got, want), f"lv{lvl}: CUDA-order fold differs from torch CUDA F.fold"


@pytest.mark.parametrize("dims", [(2, 11, 13, 9, 17, 2), (1, 5, 70, 6, 45, 1), (1, 3, 33, 40, 3, 3)])
def test_gather_fold_multi_frame_and_oracle_closed_form(dims):
    """Ragged grids (query grid != reference grid, widths across the 32-cell block boundary), several
    reference frames, arbitrary (not argmax) indices: every level against the oracle's closed form."""
    rng = np.random.default_rng(4)
    n, h, w, hr, wr, rf = dims
    arg = rng.integers(0, rf * hr * wr, size=(n, h * w)).astype(np.int32)
    for lvl, c, s in ((3, 128, 1), (2, 64, 2), (1, 32, 4)):
        refs = [rng.standard_normal((n, c, s * hr, s * wr)).astype(np.float32) for _ in range(rf)]
        want = oracle.closed_form_transfer(arg, refs, s, h, w, fold_order="cuda", div_mode="cuda")
        ref = torch.stack([cu(r) for r in refs], dim=1).contiguous()
        got = U.gather_fold(cu(arg), ref, lvl, n, h, w, hr, wr, rf, _lib.FOLD_CUDA)
        assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("layout", ["auto", "planar", "cells"])
@pytest.mark.parametrize("field", ["identity", "jitter2", "random"])
@pytest.mark.parametrize("dims", [(1, 45, 80, 45, 80, 1), (2, 9, 37, 11, 41, 2)])
def test_gather_fold_lv1_both_source_layouts(field, dims, layout):
    """The finest level gathers from the planar input or from a re-tiled cell-major copy (gather_fold.cu; chosen on the device
    from the match field for large levels, pinned here through the SPEI_FOLD_LV1_* bits): every combination must give the
    oracle's closed form bit for bit."""
    rng = np.random.default_rng(11)
    n, h, w, hr, wr, rf = dims
    yy, xx = np.divmod(np.arange(h * w), w)
    if field == "random":
        arg = rng.integers(0, rf * hr * wr, size=(n, h * w))
    else:
        m = 0 if field == "identity" else 2
        cy = np.clip(yy[None] + rng.integers(-m, m + 1, size=(n, h * w)), 0, hr - 1)
        cx = np.clip(xx[None] + rng.integers(-m, m + 1, size=(n, h * w)), 0, wr - 1)
        arg = rng.integers(0, rf, size=(n, 1)) * hr * wr + cy * wr + cx
    arg = arg.astype(np.int32)
    refs = [rng.standard_normal((n, 32, 4 * hr, 4 * wr)).astype(np.float32) for _ in range(rf)]
    want = oracle.closed_form_transfer(arg, refs, 4, h, w, fold_order="cuda", div_mode="cuda")
    ref = torch.stack([cu(r) for r in refs], dim=1).contiguous()
    mode = _lib.FOLD_CUDA | {"auto": 0, "planar": _lib.FOLD_LV1_PLANAR, "cells": _lib.FOLD_LV1_CELLS}[layout]
    got = U.gather_fold(cu(arg), ref, 1, n, h, w, hr, wr, rf, mode)
    assert np.array_equal(got.cpu().numpy(), want)


# ------------------------------------------------------------------ whole module ---------------
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_module_matches_reference_golden(golden, name):
    g = golden(name)
    st = speinet_b200.SearchTransfer(fold_mode="cpu").cuda()
    k = cu(g["ref_lv3"])
    with torch.no_grad():
        S, T3, T2, T1, arg = st(cu(g["q"]), k, cu(g["ref_lv1"]), cu(g["ref_lv2"]), k, return_index=True)
    n_tie = assert_indices_agree(g["q"], g["ref_lv3"], arg.cpu().numpy(), g["arg"])
    np.testing.assert_allclose(S.cpu().numpy(), g["S"], rtol=RTOL_S, atol=1e-6)
    assert S.shape == g["S"].shape and S.dtype == torch.float32 and arg.dtype == torch.int64
    if n_tie == 0:
        assert np.array_equal(T3.cpu().numpy(), g["T_lv3"])
        assert np.array_equal(T2.cpu().numpy(), g["T_lv2"])
        assert np.array_equal(T1.cpu().numpy(), g["T_lv1"])


def test_module_vs_torch_cuda_reference_ops_medium():
    """Against the torch-CUDA restatement (same ATen ops as the reference) at a mid size."""
    torch.manual_seed(3)
    n, h, w = 2, 45, 80
    q = torch.randn(n, 128, h, w, device="cuda") * 0.2
    lv3 = torch.randn(n, 128, h, w, device="cuda") * 0.04
    lv2 = torch.randn(n, 64, 2 * h, 2 * w, device="cuda") * 0.04
    lv1 = torch.randn(n, 32, 4 * h, 4 * w, device="cuda") * 0.04
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        wS, w3, w2, w1, warg = search_transfer_torch_cuda(q, lv3, lv1, lv2, lv3)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    st = speinet_b200.SearchTransfer().cuda()
    with torch.no_grad():
        S, T3, T2, T1, arg = st(q, lv3, lv1, lv2, lv3, return_index=True)
    n_tie = assert_indices_agree(q.cpu().numpy(), lv3.cpu().numpy(), arg.cpu().numpy(), warg.cpu().numpy())
    torch.testing.assert_close(S, wS, rtol=RTOL_S, atol=1e-6)
    same = (arg == warg)
    if bool(same.all()):
        assert torch.equal(T3, w3) and torch.equal(T2, w2) and torch.equal(T1, w1)


def test_module_random_shape_sweep_vs_torch_cuda_ops():
    """12 seeded random shapes (query and reference grids unrelated, batch 1-3) through the drop-in module against the
    reference's ATen sequence on the same GPU: S within 1e-4, and wherever the argmax indices are identical the three
    transferred pyramids must be BIT-identical (torch-CUDA col2im order, x * (1/9f))."""
    import torch.nn.functional as F
    rng = np.random.default_rng(77)
    st = speinet_b200.SearchTransfer().cuda()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for case in range(12):
            n = int(rng.integers(1, 4))
            h, w, hr, wr = (int(rng.integers(1, hi)) for hi in (24, 36, 24, 36))
            mk = lambda *shape, std: torch.from_numpy((rng.standard_normal(shape) * std).astype(np.float32)).cuda()
            q, lv3 = mk(n, 128, h, w, std=0.2), mk(n, 128, hr, wr, std=0.04)
            lv2, lv1 = mk(n, 64, 2 * hr, 2 * wr, std=0.04), mk(n, 32, 4 * hr, 4 * wr, std=0.04)
            # reference ops with a reference grid that differs from the query grid (fold sizes follow the QUERY, :44-46)
            keys = F.normalize(F.unfold(lv3, 3, padding=1).permute(0, 2, 1), dim=2)
            r_star, r_arg = torch.max(torch.bmm(keys, F.normalize(F.unfold(q, 3, padding=1), dim=1)), dim=1)
            want = {}
            for lvl, ref, p in ((3, lv3, dict(kernel_size=3, padding=1, stride=1)), (2, lv2, dict(kernel_size=6, padding=2, stride=2)),
                                (1, lv1, dict(kernel_size=12, padding=4, stride=4))):
                cols = F.unfold(ref, **p)
                picked = torch.gather(cols, 2, r_arg[:, None, :].expand(-1, cols.size(1), -1))
                want[lvl] = F.fold(picked, output_size=(h * p["stride"], w * p["stride"]), **p) / (3. * 3.)
            with torch.no_grad():
                S, T3, T2, T1, arg = st(q, lv3, lv1, lv2, lv3, return_index=True)
            tag = f"case {case} {(n, h, w, hr, wr)}"
            torch.testing.assert_close(S.reshape(n, -1), r_star, rtol=RTOL_S, atol=1e-6, msg=lambda m: f"{tag}: {m}")
            assert_indices_agree(q.cpu().numpy(), lv3.cpu().numpy(), arg.cpu().numpy(), r_arg.cpu().numpy())
            if bool((arg == r_arg).all()):
                assert torch.equal(T3, want[3]) and torch.equal(T2, want[2]) and torch.equal(T1, want[1]), tag
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def search_transfer_torch_cuda(q, lv3, lv1, lv2, ref3):
    import torch.nn.functional as F
    n, _, h, w = q.shape
    keys = F.normalize(F.unfold(lv3, 3, padding=1).permute(0, 2, 1), dim=2)
    qc = F.normalize(F.unfold(q, 3, padding=1), dim=1)
    r_star, r_arg = torch.max(torch.bmm(keys, qc), dim=1)
    outs = {}
    for lvl, ref, p in ((3, ref3, dict(kernel_size=3, padding=1, stride=1)), (2, lv2, dict(kernel_size=6, padding=2, stride=2)),
                        (1, lv1, dict(kernel_size=12, padding=4, stride=4))):
        cols = F.unfold(ref, **p)
        picked = torch.gather(cols, 2, r_arg[:, None, :].expand(-1, cols.size(1), -1))
        s = p["stride"]
        outs[lvl] = F.fold(picked, output_size=(h * s, w * s), **p) / (3. * 3.)
    return r_star.view(n, 1, h, w), outs[3], outs[2], outs[1], r_arg


def test_two_reference_frames_extension():
    rng = np.random.default_rng(12)
    n, h, w = 1, 14, 18
    q = (rng.standard_normal((n, 128, h, w)) * 0.2).astype(np.float32)
    pyr = lambda: ((rng.standard_normal((n, 32, 4 * h, 4 * w)) * 0.04).astype(np.float32),
                   (rng.standard_normal((n, 64, 2 * h, 2 * w)) * 0.04).astype(np.float32),
                   (rng.standard_normal((n, 128, h, w)) * 0.04).astype(np.float32))
    a1, a2, a3 = pyr()
    b1, b2, b3 = pyr()
    wS, w3, w2, w1, warg, _ = oracle.search_transfer(q, [a3, b3], [a1, b1], [a2, b2], [a3, b3], fold_order="cuda", div_mode="cuda")
    st = speinet_b200.SearchTransfer().cuda()
    with torch.no_grad():
        S, T3, T2, T1, arg = st(cu(q), [cu(a3), cu(b3)], [cu(a1), cu(b1)], [cu(a2), cu(b2)], [cu(a3), cu(b3)], return_index=True)
    n_tie = assert_indices_agree(q, [a3, b3], arg.cpu().numpy(), warg)
    np.testing.assert_allclose(S.cpu().numpy(), wS, rtol=RTOL_S, atol=1e-6)
    assert (arg >= h * w).any() and (arg < h * w).any()  # both frames win somewhere
    if n_tie == 0:
        assert np.array_equal(T1.cpu().numpy(), w1) and np.array_equal(T2.cpu().numpy(), w2) and np.array_equal(T3.cpu().numpy(), w3)


def test_real_model_forward_features(golden):
    """The boundary tensors of a real SPEINet forward (random-init weights; tests/golden/make_golden_model.py):
    what the reference computed at speinet.py:135 and :93-109 vs. the CUDA path."""
    g = golden("model_forward")
    st = speinet_b200.SearchTransfer(fold_mode="cpu").cuda()
    k = cu(g["ref_lv3"])
    with torch.no_grad():
        S, T3, T2, T1, arg = st(cu(g["q"]), k, cu(g["ref_lv1"]), cu(g["ref_lv2"]), k, return_index=True)
        n_tie = assert_indices_agree(g["q"], g["ref_lv3"], arg.cpu().numpy(), g["arg"])
        np.testing.assert_allclose(S.cpu().numpy(), g["S"], rtol=RTOL_S, atol=1e-6)
        if n_tie == 0:
            assert np.array_equal(T3.cpu().numpy(), g["T_lv3"])
            assert np.array_equal(T2.cpu().numpy(), g["T_lv2"])
            assert np.array_equal(T1.cpu().numpy(), g["T_lv1"])
        for lvl, scale, T in ((3, 1, T3), (2, 2, T2), (1, 4, T1)):
            f = speinet_b200.fuse_level(cu(g[f"dec{lvl}"]), T, S, cu(g[f"w{lvl}"]), cu(g[f"b{lvl}"]), scale)
            np.testing.assert_allclose(f.cpu().numpy(), g[f"f{lvl}"], rtol=1e-4, atol=1e-6)


def test_bf16_inputs_within_1e2():
    rng = np.random.default_rng(21)
    h, w = 16, 24
    q = cu((rng.standard_normal((1, 128, h, w)) * 0.2).astype(np.float32)).bfloat16()
    lv3 = cu((rng.standard_normal((1, 128, h, w)) * 0.04).astype(np.float32)).bfloat16()
    lv2 = cu((rng.standard_normal((1, 64, 2 * h, 2 * w)) * 0.04).astype(np.float32)).bfloat16()
    lv1 = cu((rng.standard_normal((1, 32, 4 * h, 4 * w)) * 0.04).astype(np.float32)).bfloat16()
    st = speinet_b200.SearchTransfer().cuda()
    with torch.no_grad():
        S, T3, T2, T1 = st(q, lv3, lv1, lv2, lv3)
    assert S.dtype == torch.bfloat16 and T1.dtype == torch.bfloat16
    # oracle for bf16 = fp32 reference on the bf16 inputs upcast to fp32
    f = lambda t: t.float().cpu().numpy()
    wS, w3, w2, w1, _, _ = oracle.search_transfer(f(q), f(lv3), f(lv1), f(lv2), f(lv3), fold_order="cuda", div_mode="cuda")
    np.testing.assert_allclose(f(S), wS, rtol=1e-2, atol=1e-3)
    np.testing.assert_allclose(f(T1), w1, rtol=1e-2, atol=1e-3)
    np.testing.assert_allclose(f(T3), w3, rtol=1e-2, atol=1e-3)


def test_self_transfer_matches_reference_golden(golden):
    """The other branch of SPEINet.forward (speinet.py:147): S from the search, T_lv3 = the input, T_lv2 / T_lv1 from the
    module's own search1 / search2 chains -- all four outputs against what the reference SelfTransfer produced."""
    g = golden("self_transfer")
    m = speinet_b200.SelfTransfer().cuda()
    m.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd_")}, strict=True)
    q = cu(g["q"])
    with torch.no_grad():
        S, T3, T2, T1 = m(q)
    np.testing.assert_allclose(S.cpu().numpy(), g["S"], rtol=RTOL_S, atol=1e-6)
    assert T3 is q
    np.testing.assert_allclose(T2.cpu().numpy(), g["T_lv2"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(T1.cpu().numpy(), g["T_lv1"], rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------ (d) fusion ------------------
def test_fusion_matches_reference_golden(golden):
    g = golden("fusion")
    for lvl, scale in ((3, 1), (2, 2), (1, 4)):
        f = speinet_b200.fuse_level(cu(g[f"dec{lvl}"]), cu(g[f"t{lvl}"]), cu(g["S"]), cu(g[f"w{lvl}"]), cu(g[f"b{lvl}"]), scale)
        np.testing.assert_allclose(f.cpu().numpy(), g[f"f{lvl}"], rtol=1e-4, atol=1e-6)


def test_fusion_odd_sizes_vs_oracle():
    rng = np.random.default_rng(8)
    h, w = 7, 9  # plane not a multiple of 4 or of the 128-pixel tile
    S = (rng.random((2, 1, h, w)) * 0.2).astype(np.float32)
    for c, scale in ((128, 1), (64, 2), (32, 4)):
        dec = rng.standard_normal((2, c, h * scale, w * scale)).astype(np.float32)
        t = rng.standard_normal((2, c, h * scale, w * scale)).astype(np.float32)
        wt = (rng.standard_normal((c, 2 * c, 1, 1)) * 0.05).astype(np.float32)
        b = rng.standard_normal(c).astype(np.float32)
        got = speinet_b200.fuse_level(cu(dec), cu(t), cu(S), cu(wt), cu(b), scale).cpu().numpy()
        want = oracle.fuse_level(dec, t, S, wt, b, scale)
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("hw", [(4, 4), (8, 8), (12, 11), (16, 40)])
def test_fusion_small_and_partial_tiles_vs_oracle(hw):
    """Planes smaller than / not a multiple of the 128-pixel tile on the TMA path (plane % 4 == 0): the tensor maps
    clip the partial boxes of loads and stores.  (12, 11) at scale 1 has plane % 4 == 0; at scale 2 / 4 too."""
    rng = np.random.default_rng(11)
    h, w = hw
    S = (rng.random((3, 1, h, w)) * 0.2).astype(np.float32)
    for c, scale in ((128, 1), (64, 2), (32, 4)):
        dec = rng.standard_normal((3, c, h * scale, w * scale)).astype(np.float32)
        t = rng.standard_normal((3, c, h * scale, w * scale)).astype(np.float32)
        wt = (rng.standard_normal((c, 2 * c, 1, 1)) * 0.05).astype(np.float32)
        b = rng.standard_normal(c).astype(np.float32)
        got = speinet_b200.fuse_level(cu(dec), cu(t), cu(S), cu(wt), cu(b), scale).cpu().numpy()
        want = oracle.fuse_level(dec, t, S, wt, b, scale)
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)


def test_fixed_window_mode_and_its_deviation_counter():
    """eps > 0 selects the fixed (uncertified) window of round 1: with a deliberately tiny eps the measured bf16 deviation
    (stats[2]) exceeds half the window -- the situation the certified default exists for -- while the default call on the
    same inputs agrees with the exhaustive fp32 search and reports no bound violation."""
    rng = np.random.default_rng(31)
    q = (rng.standard_normal((1, 128, 20, 28)) * 0.2).astype(np.float32)
    k = (rng.standard_normal((1, 128, 20, 28)) * 0.04).astype(np.float32)
    exact = speinet_b200.search_transfer(cu(q), cu(k), search="exact")
    tiny = speinet_b200.search_transfer(cu(q), cu(k), eps=2e-5)
    assert float(tiny[5][2].item()) * 1e-9 > 0.5 * 2e-5            # the window really was too tight for bf16
    safe = speinet_b200.search_transfer(cu(q), cu(k))
    assert_indices_agree(q, [k], safe[4].cpu().numpy(), exact[4].cpu().numpy())
    np.testing.assert_allclose(safe[0].cpu().numpy(), exact[0].cpu().numpy(), rtol=RTOL_S, atol=1e-6)
    st = safe[5].cpu().tolist()
    assert st[6] == 0 and st[3] == 0
    assert st[1] >= int(tiny[5][1].item())                          # the certified window rescores at least as many candidates


@pytest.mark.parametrize("search", TC_MODES)
def test_certified_bound_holds_and_second_pass_handles_saturated_lists(search):
    """A constant feature vector + 20 % noise: ~240 keys per query sit within the certified window, every candidate list saturates,
    and the second tcgen05 pass (relevance_flagged.cu) must enumerate them.  The result still equals the exhaustive fp32
    search; stats: saturated queries > 0, pairs emitted > 0, no capacity fallback, zero bound violations."""
    rng = np.random.default_rng(5)
    base = rng.standard_normal((1, 128, 1, 1)).astype(np.float32)
    # two items with different queue lengths (the packed query tiles of item 1 start where item 0's end), two frames each
    q = (base + 0.2 * rng.standard_normal((2, 128, 33, 47))).astype(np.float32)
    k = (base + 0.2 * rng.standard_normal((2, 2, 128, 29, 41))).astype(np.float32)
    q[1, :, 20:] = rng.standard_normal((128, 13, 47)).astype(np.float32)     # item 1: a third of the queries are ordinary
    S0, a0, _, f0 = U.run_search(cu(q), cu(k), search=_lib.SEARCH_EXACT)
    S1, a1, st1, f1 = U.run_search(cu(q), cu(k), search=SEARCH_MODES[search])
    st = st1.cpu().tolist()
    assert f0 == 0 and f1 == 0
    assert st[0] > 2000 and st[5] >= 8 * st[0] and st[3] == 0 and st[6] == 0, st
    diff = a0 != a1
    if diff.any():
        assert float((S0.reshape(2, -1)[diff] - S1.reshape(2, -1)[diff]).abs().max()) < 1e-5
    torch.testing.assert_close(S1, S0, rtol=RTOL_S, atol=1e-6)


def test_second_pass_capacity_overflow_falls_back_to_exhaustive_search():
    """A constant image: every key is inside every query's window, the second pass would have to emit L x Lk pairs.  The
    emission buffer overflows, the device flag routes the queued queries to the exhaustive fp32 search, and the reference
    semantics survive: every relevance equal -> first index (torch.max), S = 1."""
    q = torch.full((1, 128, 40, 64), 0.25, device="cuda")
    k = torch.full((1, 1, 128, 24, 50), 0.5, device="cuda")
    q[0, :, 0, 0] += 0.001   # keep the borders from being the unique best
    S, a, st, flag = U.run_search(q, k)
    S0, a0, _, _ = U.run_search(q, k, search=_lib.SEARCH_EXACT)
    st = st.cpu().tolist()
    assert flag == 0 and st[3] != 0 and st[6] == 0, st
    diff = a != a0
    if diff.any():
        assert float((S.reshape(1, -1)[diff] - S0.reshape(1, -1)[diff]).abs().max()) < 1e-5
    torch.testing.assert_close(S, S0, rtol=RTOL_S, atol=1e-6)


# ------------------------------------------------------------------ (f-1 / f-3) resize + 1x1 conv + ReLU chains
@pytest.mark.parametrize("dims", [(2, 128, 64, 9, 13), (1, 64, 32, 32, 48), (1, 128, 64, 5, 3)])
def test_up2_conv1x1_relu_vs_reference_ops_and_oracle(dims):
    """relu(conv1x1(F.interpolate(x, 2, bicubic))) (SearchTransfer.py:70-76, speinet.py:99-100,111-112): against the
    numpy oracle in the reference's op order and against torch's own ops on the GPU (TF32 off)."""
    import torch.nn.functional as F
    n, cin, cout, h, w = dims
    rng = np.random.default_rng(21)
    x = rng.standard_normal((n, cin, h, w)).astype(np.float32)
    wt = (rng.standard_normal((cout, cin, 1, 1)) * 0.1).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    got = speinet_b200.up2_conv1x1_act(cu(x), cu(wt), cu(b)).cpu().numpy()
    up = oracle.bicubic_upsample(x, 2)
    want = np.maximum(oracle.conv1x1(up, wt, b), 0)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = F.relu(F.conv2d(F.interpolate(cu(x), scale_factor=2, mode="bicubic"), cu(wt), cu(b))).cpu().numpy()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------ (f-3) Richardson-Lucy edge prior
@pytest.mark.parametrize("name", ["uni", "img"])
@pytest.mark.parametrize("iters", [1, 5])
def test_rl_deconv_matches_reference_golden(golden, name, iters):
    g = golden("rl_deconv")
    got = speinet_b200.r_l_per_channel(cu(g[name]), cu(g["blur_kernel"]), iters, 0.01).cpu().numpy()
    want = g[f"{name}_it{iters}"]
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)   # reference r_l_per_channel (rcl.py:22-51), CPU fp32


def test_rl_deconv_vs_oracle_ragged_tiles_and_kernel_sizes():
    rng = np.random.default_rng(5)
    x = rng.random((2, 3, 45, 131)).astype(np.float32)           # not a multiple of the 64 x 32 tile
    x[0, 0, 5:20, 60:90] = 0.0
    for ks, iters in ((5, 5), (5, 2), (3, 4), (7, 2)):
        k = rng.random((1, 1, ks, ks)).astype(np.float32)
        k /= k.sum()
        got = speinet_b200.r_l_per_channel(cu(x), cu(k), iters, 0.02).cpu().numpy()
        want = oracle.r_l_per_channel(x, k, iters, 0.02)
        assert np.array_equal(np.isfinite(got), np.isfinite(want))
        fin = np.isfinite(want)
        np.testing.assert_allclose(got[fin], want[fin], rtol=1e-4, atol=1e-5)


def test_rl_deconv_720p_matches_torch_ops_and_is_one_launch():
    """Full 1280x720 frame, 5 iterations (speinet.py:129): against the reference's op sequence run with torch on the same
    GPU (TF32 off so the convolutions are fp32), and channel / batch independence as a size-independent property."""
    import torch.nn.functional as F
    torch.manual_seed(3)
    x = torch.rand(1, 3, 720, 1280, device="cuda")
    k = speinet_b200.create_blur_kernel().cuda()
    got = speinet_b200.r_l_per_channel(x, k, 5, 0.01)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        lap = torch.tensor([[0, -1, 0], [-1, 4, -1], [0, -1, 0]], dtype=torch.float32, device="cuda")[None, None]
        chans = []
        for c in range(3):
            xc = x[:, c:c + 1]
            d = xc.clone()
            for _ in range(5):
                cf = xc / F.conv2d(d, k, padding=2)
                cf[cf != cf] = 0.0
                cf[cf < 0] = 0.0
                d = cf * (d + 0.01 * F.conv2d(d, lap, padding=1))
            chans.append(d)
        want = torch.cat(chans, dim=1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    rel = ((got - want).abs() / want.abs().clamp_min(1e-3)).max().item()
    assert rel < 1e-4, rel
    again = speinet_b200.r_l_per_channel(x[:, 1:2].contiguous(), k, 5, 0.01)
    assert torch.equal(again[:, 0], got[:, 1])


# ------------------------------------------------------------------ full size properties --------
def test_full_size_720p_properties():
    """At BASELINE.json's full size the oracle is too slow; use size-independent properties
    (SURVEY section 8(c)): query == key => identity match, S ~ 1, interior T_lv3 == key, borders
    attenuated by the constant /9; linearity of the transfer in the reference pyramid."""
    torch.manual_seed(1)
    h, w = 180, 320
    k = torch.randn(1, 128, h, w, device="cuda") * 0.04
    lv2 = torch.randn(1, 64, 2 * h, 2 * w, device="cuda") * 0.04
    lv1 = torch.randn(1, 32, 4 * h, 4 * w, device="cuda") * 0.04
    st = speinet_b200.SearchTransfer().cuda()
    with torch.no_grad():
        S, T3, T2, T1, arg = st(k.clone(), k, lv1, lv2, k, return_index=True)
        ident = torch.arange(h * w, device="cuda")[None]
        assert torch.equal(arg, ident)
        assert float((S - 1).abs().max()) < 1e-5
        torch.testing.assert_close(T3[:, :, 1:-1, 1:-1], k[:, :, 1:-1, 1:-1], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(T1[:, :, 4:-4, 4:-4], lv1[:, :, 4:-4, 4:-4], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(T3[:, :, 0, 0], k[:, :, 0, 0] * (4.0 / 9.0), rtol=1e-5, atol=1e-7)
        # linearity: T(2*ref) == 2*T(ref) exactly (power-of-two scaling commutes with every rounding)
        _, _, _, T1b = st(k.clone(), k, lv1 * 2, lv2, k)
        assert torch.equal(T1b, T1 * 2)
    assert st.last_stats.cpu().tolist()[0] == 0


@pytest.mark.parametrize("search", TC_MODES)
def test_full_size_720p_random_subset_vs_exhaustive_fp32(search):
    """TC paths vs the exhaustive fp32 CUDA-core search on the same staged operands, 720p, random data."""
    torch.manual_seed(2)
    h, w = 180, 320
    q = torch.randn(1, 128, h, w, device="cuda") * 0.2
    k = (torch.randn(1, 1, 128, h, w, device="cuda") * 0.04).contiguous()
    S_tc, a_tc, stats, flag = U.run_search(q, k, search=SEARCH_MODES[search])
    S_ex, a_ex, _, _ = U.run_search(q, k, search=_lib.SEARCH_EXACT)
    assert flag == 0
    diff = (a_tc != a_ex)
    # any disagreement must be a near tie in relevance
    assert float((S_tc - S_ex).abs().max()) < 1e-5
    assert int(diff.sum()) <= 0.001 * h * w


# ------------------------------------------------------------------ boundary behaviour ----------
def test_errors_are_loud():
    st = speinet_b200.SearchTransfer().cuda()
    q = torch.randn(1, 128, 8, 8, device="cuda")
    with pytest.raises(RuntimeError, match="forward-only"):
        st(q.clone().requires_grad_(True), q, torch.randn(1, 32, 32, 32, device="cuda"), torch.randn(1, 64, 16, 16, device="cuda"), q)
    with torch.no_grad(), pytest.raises(RuntimeError, match="expected"):
        st(q, q, torch.randn(1, 32, 30, 32, device="cuda"), torch.randn(1, 64, 16, 16, device="cuda"), q)
    with torch.no_grad(), pytest.raises(RuntimeError, match="channels"):
        st(torch.randn(1, 64, 8, 8, device="cuda"), torch.randn(1, 64, 8, 8, device="cuda"), None, None, None)
    lib = speinet_b200.load_library()
    shape = U.make_shape(1, 8, 8, 8, 8)
    assert lib.spei_stage_norm(ctypes.byref(shape), U.vp(q), U.vp(q), ctypes.c_void_p(0), 0, U.cur_stream()) == -1
    ws, ptr, nbytes = U.alloc_workspace(shape)
    assert lib.spei_stage_norm(ctypes.byref(shape), U.vp(q), U.vp(q), ctypes.c_void_p(ptr), 16, U.cur_stream()) == -4


def test_runs_on_non_default_stream_and_batch_mask():
    """speinet.py:163 feeds boolean-mask sub-batches; inference may use side streams."""
    torch.manual_seed(5)
    q = torch.randn(3, 128, 12, 16, device="cuda") * 0.2
    lv3 = torch.randn(3, 128, 12, 16, device="cuda") * 0.04
    lv2 = torch.randn(3, 64, 24, 32, device="cuda") * 0.04
    lv1 = torch.randn(3, 32, 48, 64, device="cuda") * 0.04
    st = speinet_b200.SearchTransfer().cuda()
    mask = torch.tensor([True, False, True], device="cuda")
    with torch.no_grad():
        full = st(q, lv3, lv1, lv2, lv3)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            sub = st(q[mask], lv3[mask], lv1[mask], lv2[mask], lv3[mask])
        torch.cuda.current_stream().wait_stream(side)
    for a, b in zip(full, sub):
        assert torch.equal(a[mask], b)


def test_module_cuda_graph_mode_replays_and_matches_eager():
    """SearchTransfer(cuda_graph=True): first call eager, second call with the same input tensors captures, later calls
    replay; results are bit-identical to the eager module, new input VALUES in the same tensors are picked up, and tensors
    at new addresses fall back to a fresh eager call."""
    torch.manual_seed(41)
    h, w = 20, 28
    q = torch.randn(1, 128, h, w, device="cuda") * 0.2
    lv3 = torch.randn(1, 128, h, w, device="cuda") * 0.04
    lv2 = torch.randn(1, 64, 2 * h, 2 * w, device="cuda") * 0.04
    lv1 = torch.randn(1, 32, 4 * h, 4 * w, device="cuda") * 0.04
    eager = speinet_b200.SearchTransfer().cuda()
    graphed = speinet_b200.SearchTransfer(cuda_graph=True).cuda()
    with torch.no_grad():
        want = [t.clone() for t in eager(q, lv3, lv1, lv2, lv3)]
        for call in range(4):
            got = graphed(q, lv3, lv1, lv2, lv3)
            assert all(torch.equal(a, b) for a, b in zip(got, want)), f"call {call}"
        assert len(graphed._graphs) == 1 and next(iter(graphed._graphs.values()))["graph"] is not None
        q.mul_(-1.0).add_(0.01)                       # same tensor, new content: the replay must see it
        want2 = [t.clone() for t in eager(q, lv3, lv1, lv2, lv3)]
        got2 = graphed(q, lv3, lv1, lv2, lv3)
        assert all(torch.equal(a, b) for a, b in zip(got2, want2))
        q3 = q.clone()                                # new address: eager path again, then its own graph
        got3 = graphed(q3, lv3, lv1, lv2, lv3)
        assert all(torch.equal(a, b) for a, b in zip(got3, want2)) and len(graphed._graphs) == 2


# ------------------------------------------------------------------ host pipeline / install ------
def test_host_pipeline_matches_direct_calls():
    torch.manual_seed(13)
    h, w = 12, 16
    mk = lambda *s, std: (torch.randn(*s) * std).pin_memory()
    clips = [{"q": mk(1, 128, h, w, std=0.2), "lv3": mk(1, 128, h, w, std=0.04), "lv2": mk(1, 64, 2 * h, 2 * w, std=0.04),
              "lv1": mk(1, 32, 4 * h, 4 * w, std=0.04), "dec3": mk(1, 128, h, w, std=0.3),
              "dec2": mk(1, 64, 2 * h, 2 * w, std=0.3), "dec1": mk(1, 32, 4 * h, 4 * w, std=0.3)} for _ in range(5)]
    convs = {3: torch.nn.Conv2d(256, 128, 1).cuda(), 2: torch.nn.Conv2d(128, 64, 1).cuda(), 1: torch.nn.Conv2d(64, 32, 1).cuda()}
    wb = {l: (c.weight.detach(), c.bias.detach()) for l, c in convs.items()}
    outs = [{"S": torch.empty(1, 1, h, w).pin_memory(), "f3": torch.empty(1, 128, h, w).pin_memory(),
             "f2": torch.empty(1, 64, 2 * h, 2 * w).pin_memory(), "f1": torch.empty(1, 32, 4 * h, 4 * w).pin_memory()} for _ in range(5)]
    pipe = speinet_b200.HostPipeline(wb, "cuda")
    pipe.run(clips, outs)
    torch.cuda.synchronize()
    st = speinet_b200.SearchTransfer().cuda()
    with torch.no_grad():
        for clip, out in zip(clips, outs):
            d = {k: v.cuda() for k, v in clip.items()}
            S, T3, T2, T1 = st(d["q"], d["lv3"], d["lv1"], d["lv2"], d["lv3"])
            f1 = speinet_b200.fuse_level(d["dec1"], T1, S, *wb[1], 4)
            f3 = speinet_b200.fuse_level(d["dec3"], T3, S, *wb[3], 1)
            assert torch.equal(out["S"], S.cpu()) and torch.equal(out["f1"], f1.cpu()) and torch.equal(out["f3"], f3.cpu())


class _TinyRecons(torch.nn.Module):
    """Stand-in with the attribute names decode_fused touches (recons_video_ori.py decoders / outBlock)."""
    def __init__(self):
        super().__init__()
        self.decoder_second = torch.nn.ConvTranspose2d(128, 64, 3, stride=2, padding=1, output_padding=1)
        self.decoder_first = torch.nn.ConvTranspose2d(64, 32, 3, stride=2, padding=1, output_padding=1)
        self.outBlock = torch.nn.Conv2d(32, 3, 3, padding=1)


class _TinySPEINet(torch.nn.Module):
    """The parameters of SPEINet that _decode uses (speinet.py:53-66), reference-style PyTorch _decode."""
    def __init__(self):
        super().__init__()
        import torch.nn as nn
        n_feat = 32
        self.recons_net = _TinyRecons()
        self.SearchTransfer = speinet_b200.SearchTransfer()
        self.SelfTransfer = speinet_b200.SelfTransfer()
        self.conv_lv1 = nn.Conv2d(n_feat * 2, n_feat, 1)
        self.conv_lv2 = nn.Conv2d(n_feat * 4, n_feat * 2, 1)
        self.conv_lv3 = nn.Conv2d(n_feat * 8, n_feat * 4, 1)
        self.search3 = nn.Conv2d(n_feat * 2, n_feat * 2, 3, padding=1)
        self.search2 = nn.Conv2d(n_feat * 4, n_feat * 2, 1)
        self.search1 = nn.Conv2d(n_feat * 4, n_feat * 2, 1)
        self.search43 = nn.Conv2d(n_feat, n_feat, 3, padding=1)
        self.search33 = nn.Conv2d(n_feat * 2, n_feat, 3, padding=1)
        self.search13 = nn.Conv2d(n_feat * 2, n_feat, 1)

    def _decode(self, f, S, T3, T2, T1):  # torch restatement of speinet.py:92-120 (the checker for decode_fused)
        import torch.nn.functional as F
        up = lambda x, s=2: F.interpolate(x, scale_factor=s, mode="bicubic")
        f_lv3 = fuse_level_torch(f, T3, S, self.conv_lv3.weight, self.conv_lv3.bias, 1)
        d2 = self.recons_net.decoder_second(f_lv3)
        f_lv2 = fuse_level_torch(d2, T2, S, self.conv_lv2.weight, self.conv_lv2.bias, 2)
        s1 = F.relu(self.search1(up(f_lv3)))
        s2 = F.relu(self.search3(f_lv2))
        f_v3 = d2 + F.relu(self.search2(torch.cat((d2, s1), 1)))
        f_lv2 = f_lv2 + F.relu(self.search2(torch.cat((f_lv2, s2), 1)))
        d1 = self.recons_net.decoder_first(f_lv2)
        f_lv1 = fuse_level_torch(d1, T1, S, self.conv_lv1.weight, self.conv_lv1.bias, 4)
        s13 = F.relu(self.search13(up(f_v3)))
        s23 = F.relu(self.search33(up(f_lv2)))
        s33 = F.relu(self.search43(f_lv1))
        pair = lambda a, b: F.relu(self.search33(torch.cat((a, b), 1)))
        return self.recons_net.outBlock(f_lv1 + pair(s13, s23) + pair(s13, s33) + pair(s23, s33))


def test_install_and_decode_fused_match_torch_decode():
    torch.manual_seed(17)
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        net = _TinySPEINet().cuda().eval()
        h, w = 10, 14
        f = torch.randn(2, 128, h, w, device="cuda") * 0.2
        lv3 = torch.randn(2, 128, h, w, device="cuda") * 0.04
        lv2 = torch.randn(2, 64, 2 * h, 2 * w, device="cuda") * 0.04
        lv1 = torch.randn(2, 32, 4 * h, 4 * w, device="cuda") * 0.04
        with torch.no_grad():
            S, T3, T2, T1 = net.SearchTransfer(f, lv3, lv1, lv2, lv3)
            want = net._decode(f, S, T3, T2, T1)
            sd_before = {k: v.clone() for k, v in net.state_dict().items()}
            speinet_b200.install(net)
            assert sorted(net.state_dict()) == sorted(sd_before)          # checkpoint keys survive install()
            got = net._decode(f, S, T3, T2, T1)
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_frame_equals_unsharded(world):
    """Single large frame split by query rows: the bands of all ranks, computed
    one after the other on this GPU, stitch to exactly the unsharded result."""
    torch.manual_seed(23)
    h, w = 26, 24
    q = torch.randn(1, 128, h, w, device="cuda") * 0.2
    lv3 = torch.randn(1, 128, 20, 28, device="cuda") * 0.04
    lv2 = torch.randn(1, 64, 40, 56, device="cuda") * 0.04
    lv1 = torch.randn(1, 32, 80, 112, device="cuda") * 0.04
    st = speinet_b200.SearchTransfer().cuda()
    with torch.no_grad():
        full = st(q, lv3, lv1, lv2, lv3)
        bands = [speinet_b200.search_transfer_rows(st, q, lv3, lv1, lv2, lv3, r, world) for r in range(world)]
    for i in range(4):
        stitched = torch.cat([b[i] for b in bands], dim=2)
        assert torch.equal(stitched, full[i]), f"output {i} differs between row-sharded and unsharded"


def test_integration_md_ctypes_stub_matches_module():
    """The minimal ctypes binding printed in INTEGRATION.md section 3, executed verbatim, against the drop-in module."""
    import os
    import runpy
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runpy.run_path(os.path.join(root, "tools", "check_integration_stub.py"), run_name="__main__")


# ------------------------------------------------------------------ multi-GPU (needs >= 2 devices) ------
def _peer_worker(rank, world, port, ret):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NCCL_DEBUG="WARN")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ok = True
        for num_clips in (world, 2 * world):              # one clip per rank per exchange, then two
            mine = speinet_b200.shard_clips(num_clips, rank, world)
            pg = speinet_b200.PeerGather((3, 8, 16), num_clips, rank, world, device=torch.device("cuda", rank))
            for rnd in range(4):                            # reuse of both slots, result() before the next push()
                local = torch.stack([torch.full((3, 8, 16), float(100 * rnd + c), device="cuda") for c in mine])
                slot = pg.push(local)
                got = pg.result(slot)
                want = speinet_b200.gather_outputs(local, num_clips, rank, world)      # NCCL all-gather of the same shards
                ok = ok and bool(torch.equal(got, want)) and got[:, 0, 0, 0].tolist() == [float(100 * rnd + c) for c in range(num_clips)]
        # row bands of one frame: stitched result equals the unsharded module (NCCL gather_rows)
        torch.manual_seed(5)
        q = torch.randn(1, 128, 22, 20, device="cuda") * 0.2
        lv3 = torch.randn(1, 128, 22, 20, device="cuda") * 0.04
        lv2 = torch.randn(1, 64, 44, 40, device="cuda") * 0.04
        lv1 = torch.randn(1, 32, 88, 80, device="cuda") * 0.04
        st = speinet_b200.SearchTransfer().cuda()
        with torch.no_grad():
            parts = speinet_b200.gather_rows(speinet_b200.search_transfer_rows(st, q, lv3, lv1, lv2, lv3, rank, world), 22, rank, world)
            full = st(q, lv3, lv1, lv2, lv3)
        ok = ok and all(bool(torch.equal(a, b)) for a, b in zip(parts, full))
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_peer_memory_gather_and_row_bands_on_two_gpus():
    """PeerGather (NVLink peer memory, copy engines) against NCCL's all-gather, and the NCCL row-band gather of one frame,
    one process per GPU.  Skipped on single-GPU boxes (bench.py --gpus N runs the same checks at N = 2 .. 8)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_peer_worker, args=(2, port, ret), nprocs=2, join=True)
        assert ret[0] and ret[1]


def test_misaligned_views_take_a_working_path():
    """Contiguous views whose storage offset is not a multiple of 16 bytes (e.g. x[1:] of an odd-sized plane): the search
    wrapper clones them, the fusion entry point routes them to its LDG-fed kernel -- no RuntimeError, same results
    (round-1 ADVICE: such views used to be rejected)."""
    torch.manual_seed(19)
    h, w = 9, 13                                           # odd planes: every slice below starts at an odd float offset
    big_q = torch.randn(2, 128, h, w, device="cuda") * 0.2
    big_k = torch.randn(2, 128, h, w, device="cuda") * 0.04
    big2 = torch.randn(2, 64, 2 * h, 2 * w, device="cuda") * 0.04
    big1 = torch.randn(2, 32, 4 * h, 4 * w, device="cuda") * 0.04
    flat = torch.randn(128 * h * w + 1, device="cuda") * 0.2
    q_odd = flat[1:].view(1, 128, h, w)                    # data_ptr % 16 == 4
    assert q_odd.data_ptr() % 16 != 0 and q_odd.is_contiguous()
    st = speinet_b200.SearchTransfer().cuda()
    with torch.no_grad():
        got = st(q_odd, big_k[1:], big1[1:], big2[1:], big_k[1:])
        want = st(q_odd.clone(), big_k[1:].clone(), big1[1:].clone(), big2[1:].clone(), big_k[1:].clone())
    assert all(torch.equal(a, b) for a, b in zip(got, want))
    S, T3 = got[0], got[1]
    wt = torch.randn(128, 256, 1, 1, device="cuda") * 0.05
    b = torch.randn(128, device="cuda")
    dflat = torch.randn(128 * h * w + 3, device="cuda")
    dec_odd = dflat[3:].view(1, 128, h, w)
    assert dec_odd.data_ptr() % 16 != 0
    f_odd = speinet_b200.fuse_level(dec_odd, T3, S, wt, b, 1)
    f_ref = speinet_b200.fuse_level(dec_odd.clone(), T3.clone(), S, wt, b, 1)
    torch.testing.assert_close(f_odd, f_ref, rtol=1e-5, atol=1e-6)   # (LDG-fed vs TMA-fed kernel when the plane allows TMA)
