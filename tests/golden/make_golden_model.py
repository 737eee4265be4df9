#!/usr/bin/env python
"""Golden fixture from a REAL forward pass of the reference SPEINet (random-init weights, CPU).

    python tests/golden/make_golden_model.py        # build container only (/root/reference needed)

Imports /root/reference/model/speinet.py behind the import shims of SURVEY.md section 8(c)
(fake `timm.models.layers`, `pypardiso`, `scipy.signal.gaussian`, `scipy.ndimage.filters`, `.cuda()`
neutralised, CUDA_VISIBLE_DEVICES restored), runs `SPEINet.forward` on one seeded 5-frame 96x96 clip
(window_size=4; the shipped window_size=5 does not divide this size) and records what
crosses the hot-path boundary:
  * the arguments and results of `self.SearchTransfer(...)` at speinet.py:135
  * the inputs / outputs of conv_lv3/2/1 in `_decode` (speinet.py:93, 96, 108), from which the three
    fused features of lines 94, 97, 109 follow with the reference's own expressions.
Unlike the synthetic fixtures these features come from the real encoders + Swin fusion.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SPEINET_REFERENCE", "/root/reference")


def install_shims():
    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")

    class DropPath(nn.Module):  # eval-mode identity; drop_path_rate only matters in training
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            return x

    layers.DropPath = DropPath
    layers.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
    layers.trunc_normal_ = nn.init.trunc_normal_
    sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": layers})

    import scipy.signal
    import scipy.signal.windows
    import scipy.sparse.linalg
    import scipy.ndimage
    if not hasattr(scipy.signal, "gaussian"):
        scipy.signal.gaussian = scipy.signal.windows.gaussian
    if "scipy.ndimage.filters" not in sys.modules:
        filt = types.ModuleType("scipy.ndimage.filters")
        filt.convolve = scipy.ndimage.convolve
        sys.modules["scipy.ndimage.filters"] = filt
    pyp = types.ModuleType("pypardiso")
    pyp.spsolve = scipy.sparse.linalg.spsolve
    sys.modules["pypardiso"] = pyp
    try:
        import cv2  # noqa: F401
    except ImportError:
        sys.modules["cv2"] = types.ModuleType("cv2")


def main():
    install_shims()
    env_before = os.environ.get("CUDA_VISIBLE_DEVICES")
    sys.path.insert(0, REF)
    from model import speinet  # noqa: E402  (sets CUDA_VISIBLE_DEVICES as an import side effect)
    if env_before is None:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = env_before
    torch.Tensor.cuda = lambda self, *a, **k: self       # rcl.py:29-30 calls .cuda() unconditionally

    torch.manual_seed(0)
    torch.set_num_threads(1)
    size = 96
    args = types.SimpleNamespace(patch_size=size, window_size=4, rgb_range=1, depths=[6] * 6, embed_dim=256, num_heads=[8] * 6,
                                 mlp_ratio=2, resi_connection="1conv", n_colors=3, n_sequence=3, n_resblock=3, n_feat=32, cpu=True)
    net = speinet.SPEINet(in_channels=3, n_sequence=3, out_channels=3, n_resblock=3, n_feat=32, device="cpu", args=args).eval()

    rec = {}
    orig_st = net.SearchTransfer.forward

    def st_spy(lrsr_lv3, refsr_lv3, ref_lv1, ref_lv2, ref_lv3):
        out = orig_st(lrsr_lv3, refsr_lv3, ref_lv1, ref_lv2, ref_lv3)
        assert refsr_lv3 is ref_lv3
        rec.update(q=lrsr_lv3, ref_lv1=ref_lv1, ref_lv2=ref_lv2, ref_lv3=ref_lv3, S=out[0], T_lv3=out[1], T_lv2=out[2], T_lv1=out[3])
        return out

    net.SearchTransfer.forward = st_spy
    for lvl in (3, 2, 1):
        conv = getattr(net, f"conv_lv{lvl}")
        conv.register_forward_hook(lambda m, inp, out, lvl=lvl: rec.update({f"cat{lvl}": inp[0], f"conv{lvl}": out}))

    gen = torch.Generator().manual_seed(42)
    low = torch.rand(1, 5, 3, size // 8, size // 8, generator=gen)
    x = F.interpolate(low.view(5, 3, size // 8, size // 8), size=(size, size), mode="bicubic").clamp(0, 1).view(1, 5, 3, size, size)
    x = (x + 0.02 * torch.rand(x.shape, generator=gen)).clamp(0, 1)          # image-like frames, frame 3 non-zero
    with torch.no_grad():
        y = net(x)
        qu = F.normalize(F.unfold(rec["q"], 3, padding=1), dim=1)
        ku = F.normalize(F.unfold(rec["ref_lv3"], 3, padding=1).permute(0, 2, 1), dim=2)
        _, r_arg = torch.max(torch.bmm(ku, qu), dim=1)
        S = rec["S"]
        out = {k: v.numpy() for k, v in rec.items() if not k.startswith(("cat", "conv"))}
        out["arg"] = r_arg.numpy().astype(np.int32)
        out["frame"] = y.numpy()
        for lvl, scale in ((3, 1), (2, 2), (1, 4)):
            cat, conv_out = rec[f"cat{lvl}"], rec[f"conv{lvl}"]
            c = cat.shape[1] // 2
            dec = cat[:, :c]
            conv = getattr(net, f"conv_lv{lvl}")
            assert torch.equal(cat[:, c:], rec[f"T_lv{lvl}"])                  # cat(dec, T): speinet.py:93/96/108
            up = S if scale == 1 else F.interpolate(S, scale_factor=scale, mode="bicubic")
            out[f"dec{lvl}"] = dec.numpy()
            out[f"w{lvl}"] = conv.weight.detach().numpy()
            out[f"b{lvl}"] = conv.bias.detach().numpy()
            out[f"f{lvl}"] = (dec + conv_out * up).numpy()                     # speinet.py:93-94 / 96-97 / 108-109
    np.savez_compressed(os.path.join(HERE, "model_forward.npz"), **out)
    print("model_forward", {k: v.shape for k, v in out.items()})
    s = out["S"]
    print("S range", float(s.min()), float(s.max()), "q std", float(out["q"].std()), "ref3 std", float(out["ref_lv3"].std()))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"reference not found at {REF}")
    main()
