"""C-ABI boundary checks that need no GPU: the library loads, exports every symbol that
include/speinet_b200.h declares, and the ctypes table mirrors the header."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "speinet_b200.h")


@pytest.fixture(scope="module")
def lib():
    from speinet_b200 import build, _lib
    build.build()          # no-op when up to date; nvcc cross-compiles sm_100a without a GPU
    return _lib.load()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spei_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("spei_search_transfer", "spei_stage_norm", "spei_relevance_argmax", "spei_gather_fold",
                 "spei_fuse_level", "spei_rl_deconv", "spei_workspace_bytes", "spei_last_error", "spei_version"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/speinet_b200.h but not exported"


def test_ctypes_table_matches_header(lib):
    from speinet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert ctypes.sizeof(_lib.SpeiShape) == 13 * 4


def test_version_and_error_string(lib):
    text = open(HEADER).read()
    assert lib.spei_version() == int(re.search(r"#define SPEI_VERSION (\d+)", text).group(1))
    assert isinstance(lib.spei_last_error(), bytes)


def test_argument_errors_do_not_need_a_gpu(lib):
    from speinet_b200 import _lib
    bad = _lib.SpeiShape(n=0, h=4, w=4, hr=4, wr=4, rf=1, c3=128, c2=64, c1=32)
    n = ctypes.c_size_t(0)
    assert lib.spei_workspace_bytes(ctypes.byref(bad), ctypes.byref(n)) == -1
    assert b"dimension" in lib.spei_last_error()
    bad = _lib.SpeiShape(n=1, h=4, w=4, hr=4, wr=4, rf=1, c3=64, c2=32, c1=16)
    assert lib.spei_workspace_bytes(ctypes.byref(bad), ctypes.byref(n)) == -1
    assert b"channels" in lib.spei_last_error()


def test_no_silent_fallback_without_gpu(lib):
    """Without a CUDA device every compute entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from speinet_b200 import _lib
    ok = _lib.SpeiShape(n=1, h=8, w=8, hr=8, wr=8, rf=1, c3=128, c2=64, c1=32)
    n = ctypes.c_size_t(0)
    assert lib.spei_workspace_bytes(ctypes.byref(ok), ctypes.byref(n)) < 0
    # the stand-alone entry points check the device before anything else: no GPU -> error code + message, never a result
    null = ctypes.c_void_p(0)
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.spei_fuse_level(1, 32, 4, 4, 1, p, p, p, p, p, p, null) < 0 and lib.spei_last_error()
    assert lib.spei_rl_deconv(1, 1, 4, 4, 5, 1, ctypes.c_float(0.01), p, p, p, null) < 0 and lib.spei_last_error()
    assert lib.spei_upsample2_bias_act(1, 1, 4, 4, p, p, 1, p, null) < 0 and lib.spei_last_error()


def test_python_modules_refuse_cpu_tensors():
    """The public wrappers of the rows next to the path (edge prior, resize + conv chains) have no CPU path either."""
    import torch
    import speinet_b200
    x = torch.rand(1, 3, 8, 8)
    with pytest.raises(RuntimeError):
        speinet_b200.r_l_per_channel(x, speinet_b200.create_blur_kernel(), 1, 0.01)
    with pytest.raises(RuntimeError):
        speinet_b200.up2_conv1x1_act(torch.rand(1, 8, 4, 4), torch.rand(4, 8, 1, 1), torch.rand(4))
    assert tuple(speinet_b200.create_blur_kernel().shape) == (1, 1, 5, 5)
    assert abs(float(speinet_b200.create_blur_kernel().sum()) - 1.0) < 1e-6     # rcl.py:18-20
