"""Host-side logic of the drop-in (no GPU): module surface, state-dict keys, error behaviour,
clip sharding and the world-size-2 gather over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import oracle
import speinet_b200
from speinet_b200 import sharding


def test_state_dict_keys_and_shapes_match_reference():
    # SearchTransfer.py:10-11 (n_feat=32): search1 Conv2d(128->64,1x1), search2 Conv2d(64->32,1x1)
    sd = speinet_b200.SearchTransfer().state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "search1.weight": (64, 128, 1, 1), "search1.bias": (64,),
        "search2.weight": (32, 64, 1, 1), "search2.bias": (32,)}
    sd2 = speinet_b200.SelfTransfer().state_dict()
    assert sorted(sd2) == sorted(sd)


def test_reference_checkpoint_loads_strict():
    import importlib.util
    ref_path = "/root/reference/model/SearchTransfer.py"
    if not os.path.exists(ref_path):
        pytest.skip("reference not mounted")
    spec = importlib.util.spec_from_file_location("_ref_st", ref_path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ours = speinet_b200.SearchTransfer()
    ours.load_state_dict(mod.SearchTransfer().state_dict(), strict=True)


def test_bis_helper_matches_oracle():
    st = speinet_b200.SearchTransfer()
    x = torch.randn(2, 5, 7)
    idx = torch.randint(0, 7, (2, 9))
    got = st.bis(x, 2, idx).numpy()
    assert np.array_equal(got, oracle.search_transfer_np.bis(x.numpy(), idx.numpy()))


def test_cpu_tensors_raise_no_fallback():
    st = speinet_b200.SearchTransfer()
    q = torch.randn(1, 128, 8, 8)
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA|cuda|libspeinet|missing"):
        st(q, q, torch.randn(1, 32, 32, 32), torch.randn(1, 64, 16, 16), q)


def test_shard_clips_round_robin():
    assert sharding.shard_clips(64, 3, 8) == list(range(3, 64, 8))
    all_ids = sorted(i for r in range(3) for i in sharding.shard_clips(10, r, 3))
    assert all_ids == list(range(10))
    with pytest.raises(ValueError):
        sharding.shard_clips(4, 4, 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank, world, port, num_clips, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = sharding.shard_clips(num_clips, rank, world)
        local = torch.stack([torch.full((3, 4, 5), float(i)) for i in ids]) if ids else torch.empty(0, 3, 4, 5)
        out = sharding.gather_outputs(local, num_clips, rank, world)
        pending = sharding.gather_outputs(local, num_clips, rank, world, async_op=True)   # overlappable form: same result
        assert torch.equal(pending.result(), out)
        ret[rank] = out[:, 0, 0, 0].tolist()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_clips", [4, 5])
def test_gather_outputs_world2_gloo(num_clips):
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gather_worker, args=(world, port, num_clips, ret), nprocs=world, join=True)
        for r in range(world):
            assert ret[r] == [float(i) for i in range(num_clips)]


@pytest.mark.parametrize("h,world", [(180, 8), (120, 3), (7, 2), (5, 1)])
def test_row_bands_tile_the_grid_with_two_row_halo(h, world):
    covered = []
    for r in range(world):
        y0, y1, p0, p1 = sharding.row_band(h, r, world)
        covered += list(range(y0, y1))
        assert p0 == max(0, y0 - 2) and p1 == min(h, y1 + 2)
    assert covered == list(range(h))


def _rows_worker(rank, world, port, h, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        y0, y1, _, _ = sharding.row_band(h, rank, world)
        rows = torch.arange(h, dtype=torch.float32)
        band = lambda s: rows.repeat_interleave(s)[y0 * s:y1 * s].view(1, 1, -1, 1).expand(1, 2, -1, 3).contiguous()
        full = sharding.gather_rows((band(1), band(1), band(2), band(4)), h, rank, world)
        ret[rank] = [t[0, 0, :, 0].tolist() for t in full]
    finally:
        dist.destroy_process_group()


def test_gather_rows_world2_gloo():
    world, port, h = 2, _free_port(), 7   # ragged: bands of 3 and 4 rows
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_rows_worker, args=(world, port, h, ret), nprocs=world, join=True)
        rows = torch.arange(h, dtype=torch.float32)
        for r in range(world):
            for got, s in zip(ret[r], (1, 1, 2, 4)):
                assert got == rows.repeat_interleave(s).tolist()


def test_install_on_the_real_reference_speinet_swaps_modules_and_keeps_checkpoint_keys():
    """`install()` against the reference's own `SPEINet` class (speinet.py:53-54, 81, 129), imported behind the shims of
    SURVEY.md section 8(c): the two transfer modules are replaced by this package's, `_decode` is rebound, the module
    global `r_l_per_channel` of model/speinet.py points at the CUDA edge prior, and the state-dict keys / values are
    untouched (a strict checkpoint load before or after keeps working).  Mechanics only: no forward pass without a GPU."""
    import importlib.util
    import sys
    import types
    ref = os.environ.get("SPEINET_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "model")):
        pytest.skip("reference not mounted")
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("_mk_golden_model", os.path.join(here, "golden", "make_golden_model.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    mk.install_shims()
    env_before = os.environ.get("CUDA_VISIBLE_DEVICES")
    saved_modules = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.")}
    sys.path.insert(0, ref)
    try:
        from model import speinet as ref_speinet  # sets CUDA_VISIBLE_DEVICES at import (rcl.py:16)
        args = types.SimpleNamespace(patch_size=64, window_size=4, rgb_range=1, depths=[2] * 2, embed_dim=64, num_heads=[4] * 2,
                                     mlp_ratio=2, resi_connection="1conv", n_colors=3, n_sequence=3, n_resblock=1, n_feat=32, cpu=True)
        net = ref_speinet.SPEINet(in_channels=3, n_sequence=3, out_channels=3, n_resblock=1, n_feat=32, device="cpu", args=args).eval()
        sd_before = {k: v.clone() for k, v in net.state_dict().items()}
        ref_st_cls, ref_rl = type(net.SearchTransfer), ref_speinet.r_l_per_channel
        assert ref_st_cls.__module__.startswith("model.")
        speinet_b200.install(net)
        assert isinstance(net.SearchTransfer, speinet_b200.SearchTransfer) and isinstance(net.SelfTransfer, speinet_b200.SelfTransfer)
        assert net._decode.__func__ is speinet_b200.decode_fused
        assert ref_speinet.r_l_per_channel is speinet_b200.r_l_per_channel and ref_rl is not speinet_b200.r_l_per_channel
        sd_after = net.state_dict()
        assert list(sd_after) == list(sd_before)
        assert all(torch.equal(sd_after[k], sd_before[k]) for k in sd_before)
        net.load_state_dict(sd_before, strict=True)                       # inference_SPEINet.py:232 loads strict
        # the call site speinet.py:135 passes five positional tensors and unpacks four results
        import inspect
        params = list(inspect.signature(net.SearchTransfer.forward).parameters)
        assert params[:5] == ["lrsr_lv3", "refsr_lv3", "ref_lv1", "ref_lv2", "ref_lv3"]
        ref_speinet.r_l_per_channel = ref_rl
    finally:
        sys.path.remove(ref)
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
        sys.modules.update(saved_modules)
        if env_before is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = env_before


def test_forward_sync_free_equals_reference_forward_on_a_mixed_batch():
    """`forward_sync_free` against the reference's own `SPEINet.forward` (CPU, import shims, tiny Swin): a batch whose rows
    take different branches (frame 3 zero / non-zero), with the mask given by the caller and derived on the device."""
    import importlib.util
    import sys
    import types
    ref = os.environ.get("SPEINET_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "model")):
        pytest.skip("reference not mounted")
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("_mk_golden_model2", os.path.join(here, "golden", "make_golden_model.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    mk.install_shims()
    env_before = os.environ.get("CUDA_VISIBLE_DEVICES")
    saved_modules = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.")}
    saved_cuda = torch.Tensor.cuda
    sys.path.insert(0, ref)
    try:
        from model import speinet as ref_speinet
        torch.Tensor.cuda = lambda self, *a, **k: self       # rcl.py:29-30 calls .cuda() unconditionally
        torch.manual_seed(0)
        args = types.SimpleNamespace(patch_size=32, window_size=4, rgb_range=1, depths=[1], embed_dim=32, num_heads=[2],
                                     mlp_ratio=2, resi_connection="1conv", n_colors=3, n_sequence=3, n_resblock=1, n_feat=32, cpu=True)
        net = ref_speinet.SPEINet(in_channels=3, n_sequence=3, out_channels=3, n_resblock=1, n_feat=32, device="cpu", args=args).eval()
        x = torch.rand(3, 5, 3, 32, 32)
        x[1, 3] = 0                                             # row 1 has no sharp frame -> SelfTransfer branch
        with torch.no_grad():
            want = net(x)
            got_given = speinet_b200.forward_sync_free(net, x, has_sharp=[True, False, True])
            got_derived = speinet_b200.forward_sync_free(net, x)
            all_sharp = speinet_b200.forward_sync_free(net, x[[0, 2]], has_sharp=[True, True])
        assert torch.equal(got_given, want) and torch.equal(got_derived, want)
        assert torch.equal(all_sharp, want[[0, 2]])
    finally:
        torch.Tensor.cuda = saved_cuda
        sys.path.remove(ref)
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
        sys.modules.update(saved_modules)
        if env_before is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = env_before


def test_fastdiv_header_matches_integer_division(tmp_path):
    """speinet_b200/csrc/fastdiv.h (multiply-high division used by the rescoring kernels) against `/`: every divisor class the
    kernels use (1, powers of two, image widths, grid sizes, 2^31 - 1), dividends up to 2^31 - 1 incl. all multiples' borders."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "fastdiv_check.cpp")
    exe = str(tmp_path / "fastdiv_check")
    subprocess.run([gxx, "-O2", "-o", exe, src], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip() == "bad 0", res.stdout + res.stderr
