"""The oracle (numpy restatement) against the golden vectors produced by executing the
reference module (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import oracle
from oracle import search_transfer_np as st_np

CASES = ["st_same_grid", "st_ragged", "st_edge"]


@pytest.mark.parametrize("name", CASES)
def test_search_transfer_matches_reference(golden, name):
    g = golden(name)
    S, T3, T2, T1, arg, _ = oracle.search_transfer(g["q"], g["ref_lv3"], g["ref_lv1"], g["ref_lv2"], g["ref_lv3"],
                                                   fold_order="cpu", div_mode="cpu")
    agree, n_eq, n_tie = oracle.near_tie_agreement(g["q"], g["ref_lv3"], arg, g["arg"])
    assert agree.all(), f"{(~agree).sum()} index mismatches beyond the 1e-5 near-tie rule"
    # R_star: within 1e-4 relative (north_star); here it is ~1e-7
    np.testing.assert_allclose(S, g["S"], rtol=1e-4, atol=1e-6)
    if n_tie == 0:
        # identical indices => gather/fold is bit exact (CPU col2im order, true division)
        assert np.array_equal(T3, g["T_lv3"])
        assert np.array_equal(T2, g["T_lv2"])
        assert np.array_equal(T1, g["T_lv1"])


@pytest.mark.parametrize("name", CASES)
def test_fold_bit_exact_given_reference_indices(golden, name):
    g = golden(name)
    _, T3, T2, T1, _, _ = oracle.search_transfer(g["q"], g["ref_lv3"], g["ref_lv1"], g["ref_lv2"], g["ref_lv3"],
                                                 fold_order="cpu", div_mode="cpu", index=g["arg"].astype(np.int64))
    assert np.array_equal(T3, g["T_lv3"])
    assert np.array_equal(T2, g["T_lv2"])
    assert np.array_equal(T1, g["T_lv1"])


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("order,div", [("cpu", "cpu"), ("cuda", "cuda")])
def test_closed_form_equals_fold_path(golden, name, order, div):
    """the 9-neighbour displaced sum is bit-identical to
    fold(bis(unfold(.))) for either summation order."""
    g = golden(name)
    idx = g["arg"].astype(np.int64)
    _, T3, T2, T1, _, _ = oracle.search_transfer(g["q"], g["ref_lv3"], g["ref_lv1"], g["ref_lv2"], g["ref_lv3"],
                                                 fold_order=order, div_mode=div, index=idx)
    n, _, h, w = g["q"].shape
    for ref, scale, want in ((g["ref_lv3"], 1, T3), (g["ref_lv2"], 2, T2), (g["ref_lv1"], 4, T1)):
        got = oracle.closed_form_transfer(idx, ref, scale, h, w, fold_order=order, div_mode=div)
        assert np.array_equal(got, want)
    if order == "cpu":
        assert np.array_equal(T1, g["T_lv1"])


def test_edge_case_semantics(golden):
    g = golden("st_edge")
    S, arg = g["S"][0, 0], g["arg"][0].reshape(8, 16)
    # all-zero query patch -> relevance 0 for every key -> first index, S = 0
    assert S[1, 1] == 0.0 and arg[1, 1] == 0
    # duplicated keys: key patches at columns 9..14 are exact copies of those at columns 1..6
    # (columns 0/7/8/15 differ through their neighbours); the first (left) copy always wins
    assert np.array_equal(arg[4:, 9:15], arg[4:, 1:7])
    assert (arg[:, 9:15] % 16 < 8).all()
    assert (arg[0:2, 0:2] == 0).all() and (S[0:2, 0:2] == 0).all()
    # query == key away from the zeroed corner: S ~ 1 and the match is the identity
    assert np.allclose(S[5:, :8], 1.0, atol=1e-5)
    ident = (np.arange(8)[:, None] * 16 + np.arange(16)[None, :])
    assert np.array_equal(arg[5:, :8], ident[5:, :8])


def test_constant_divisor_not_fold_of_ones(golden):
    """the reference divides by 9 everywhere, so borders are attenuated."""
    g = golden("st_edge")
    idx = np.arange(8 * 16, dtype=np.int64)[None]
    _, T3, _, _, _, _ = oracle.search_transfer(g["ref_lv3"], g["ref_lv3"], g["ref_lv1"], g["ref_lv2"], g["ref_lv3"],
                                               index=idx)
    k = g["ref_lv3"]
    np.testing.assert_allclose(T3[:, :, 1:-1, 1:-1], k[:, :, 1:-1, 1:-1], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(T3[:, :, 0, 0], k[:, :, 0, 0] * 4 / 9, rtol=1e-5, atol=1e-7)


def test_self_transfer_S(golden):
    g = golden("self_transfer")
    S, _ = oracle.self_transfer_S(g["q"])
    np.testing.assert_allclose(S, g["S"], rtol=1e-4, atol=1e-6)


def test_self_transfer_chains_vs_reference_golden(golden):
    """relu(conv1x1(bicubic_x2(.))) twice with the reference module's own search1 / search2 (SearchTransfer.py:70-76)."""
    g = golden("self_transfer")
    t2 = np.maximum(oracle.conv1x1(oracle.bicubic_upsample(g["q"], 2), g["sd_search1.weight"], g["sd_search1.bias"]), 0)
    t1 = np.maximum(oracle.conv1x1(oracle.bicubic_upsample(t2, 2), g["sd_search2.weight"], g["sd_search2.bias"]), 0)
    np.testing.assert_allclose(t2, g["T_lv2"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(t1, g["T_lv1"], rtol=1e-4, atol=1e-6)


def test_multi_reference_extension_reduces_to_single():
    """Rf=2 with the two key sets concatenated: the winner is the better of the two
    single-frame searches."""
    rng = np.random.default_rng(3)
    q = rng.standard_normal((1, 128, 6, 7)).astype(np.float32) * 0.2
    pyr = lambda: (rng.standard_normal((1, 32, 20, 24)).astype(np.float32),
                   rng.standard_normal((1, 64, 10, 12)).astype(np.float32),
                   rng.standard_normal((1, 128, 5, 6)).astype(np.float32))
    a1, a2, a3 = pyr()
    b1, b2, b3 = pyr()
    S, T3, T2, T1, arg, _ = oracle.search_transfer(q, [a3, b3], [a1, b1], [a2, b2], [a3, b3])
    Sa, *_, arga, _ = oracle.search_transfer(q, a3, a1, a2, a3)
    Sb, *_, argb, _ = oracle.search_transfer(q, b3, b1, b2, b3)
    np.testing.assert_allclose(S, np.maximum(Sa, Sb), rtol=1e-6)
    pick_b = (Sb > Sa).reshape(1, -1)
    assert np.array_equal(arg, np.where(pick_b, argb + 30, arga))
    got = oracle.closed_form_transfer(arg, [a1, b1], 4, 6, 7, fold_order="cpu", div_mode="cpu")
    assert np.array_equal(got, T1)


def test_fusion_matches_reference(golden):
    g = golden("fusion")
    np.testing.assert_allclose(oracle.bicubic_upsample(g["S"], 2), g["S_up2"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(oracle.bicubic_upsample(g["S"], 4), g["S_up4"], rtol=1e-5, atol=1e-7)
    for lvl, scale in ((3, 1), (2, 2), (1, 4)):
        f = oracle.fuse_level(g[f"dec{lvl}"], g[f"t{lvl}"], g["S"], g[f"w{lvl}"], g[f"b{lvl}"], scale)
        np.testing.assert_allclose(f, g[f"f{lvl}"], rtol=1e-4, atol=1e-6)


def test_oracle_on_real_model_forward(golden):
    """Features captured from a real SPEINet forward (tests/golden/make_golden_model.py): the oracle
    reproduces what crossed the hot-path boundary at speinet.py:135 and the fused features of :94/:97/:109."""
    g = golden("model_forward")
    S, T3, T2, T1, arg, _ = oracle.search_transfer(g["q"], g["ref_lv3"], g["ref_lv1"], g["ref_lv2"], g["ref_lv3"])
    agree, _, n_tie = oracle.near_tie_agreement(g["q"], g["ref_lv3"], arg, g["arg"])
    assert agree.all()
    np.testing.assert_allclose(S, g["S"], rtol=1e-4, atol=1e-6)
    if n_tie == 0:
        assert np.array_equal(T3, g["T_lv3"]) and np.array_equal(T2, g["T_lv2"]) and np.array_equal(T1, g["T_lv1"])
    assert np.array_equal(g["dec3"], g["q"])                    # speinet.py:93: the lv3 decoder feature is f_fusion itself
    for lvl, scale in ((3, 1), (2, 2), (1, 4)):
        f = oracle.fuse_level(g[f"dec{lvl}"], g[f"T_lv{lvl}"], g["S"], g[f"w{lvl}"], g[f"b{lvl}"], scale)
        np.testing.assert_allclose(f, g[f"f{lvl}"], rtol=1e-4, atol=1e-6)


def test_torch_port_matches_golden(golden):
    import torch
    from oracle.torch_port import search_transfer_torch, fuse_level_torch
    g = golden("st_ragged")
    t = lambda a: torch.from_numpy(a)
    S, T3, T2, T1, arg = search_transfer_torch(t(g["q"]), t(g["ref_lv3"]), t(g["ref_lv1"]), t(g["ref_lv2"]), t(g["ref_lv3"]))
    agree, _, n_tie = oracle.near_tie_agreement(g["q"], g["ref_lv3"], arg.numpy(), g["arg"])
    assert agree.all()
    np.testing.assert_allclose(S.numpy(), g["S"], rtol=1e-4, atol=1e-6)
    if n_tie == 0:
        assert np.array_equal(T1.numpy(), g["T_lv1"]) and np.array_equal(T2.numpy(), g["T_lv2"])
    gf = golden("fusion")
    f = fuse_level_torch(t(gf["dec2"]), t(gf["t2"]), t(gf["S"]), t(gf["w2"]), t(gf["b2"]), 2)
    np.testing.assert_allclose(f.numpy(), gf["f2"], rtol=1e-5, atol=1e-6)


# ---- Richardson-Lucy edge prior against the reference's own r_l_per_channel ----
@pytest.mark.parametrize("name", ["uni", "img"])
@pytest.mark.parametrize("iters", [1, 5])
def test_rl_deconv_matches_reference(golden, name, iters):
    g = golden("rl_deconv")
    got = oracle.r_l_per_channel(g[name], g["blur_kernel"], iters, 0.01)
    want = g[f"{name}_it{iters}"]
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    # fp32, but torch's CPU convolution associates the 25 taps differently and 5 iterations amplify it: 1e-4 relative
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)


def test_rl_deconv_semantics_zero_and_negative():
    """0/0 -> NaN -> 0 (rcl.py:39), negative correction -> 0 (rcl.py:40), channels independent (rcl.py:27)."""
    x = np.zeros((1, 2, 9, 11), dtype=np.float32)
    x[0, 1] = np.linspace(0.1, 1.0, 99, dtype=np.float32).reshape(9, 11)
    out = oracle.r_l_per_channel(x, oracle.create_blur_kernel(), 3, 0.01)
    assert np.array_equal(out[0, 0], np.zeros((9, 11), dtype=np.float32))          # all-zero channel stays zero
    assert np.isfinite(out).all() and (out[0, 1] > 0).all()
    single = oracle.r_l_per_channel(x[:, 1:2], oracle.create_blur_kernel(), 3, 0.01)
    assert np.array_equal(single[0, 0], out[0, 1])
    neg = -x
    out_neg = oracle.r_l_per_channel(neg, oracle.create_blur_kernel(), 1, 0.01)  # x/blurred > 0 for an all-negative frame
    assert np.isfinite(out_neg).all()


def test_conv1x1_commutes_with_bicubic_resize():
    """The identity `up2_conv1x1_act` relies on: relu(conv1x1(up2(x)) + b) == relu(up2(W . x) + b) (both maps are linear,
    one mixes channels per pixel, the other pixels per channel; the border-clamped bicubic taps sum to 1)."""
    rng = np.random.default_rng(4)
    x = rng.standard_normal((2, 16, 7, 9)).astype(np.float32)
    w = (rng.standard_normal((8, 16, 1, 1)) * 0.2).astype(np.float32)
    b = rng.standard_normal(8).astype(np.float32)
    ref_order = np.maximum(oracle.conv1x1(oracle.bicubic_upsample(x, 2), w, b), 0)          # SearchTransfer.py:70-72 order
    low_res_first = np.maximum(oracle.bicubic_upsample(oracle.conv1x1(x, w, np.zeros(8, np.float32)), 2) + b[None, :, None, None], 0)
    np.testing.assert_allclose(low_res_first, ref_order, rtol=1e-4, atol=1e-5)
