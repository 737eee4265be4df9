#!/usr/bin/env python
"""Golden fixture for the Richardson-Lucy edge prior.

    python tests/golden/make_golden_rl.py        # build container only (/root/reference needed)

Imports /root/reference/model/rcl.py behind the import shims of make_golden_model.py (pypardiso,
scipy.signal.gaussian, scipy.ndimage.filters, `.cuda()` neutralised, CUDA_VISIBLE_DEVICES restored) and runs the
reference's own `r_l_per_channel` (rcl.py:22-51) with `create_blur_kernel()` (rcl.py:18-20) on CPU, fp32, for the two
ways speinet.py calls it: 1 iteration (:81) and 5 iterations (:129).  Inputs: a seeded uniform [0,1] frame, and an
image-like frame with an all-zero block (0/0 -> NaN -> 0 path of rcl.py:39) and a saturated block.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_model import install_shims, REF  # noqa: E402


def main():
    install_shims()
    env_before = os.environ.get("CUDA_VISIBLE_DEVICES")
    sys.path.insert(0, REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_rcl", os.path.join(REF, "model", "rcl.py"))
    rcl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rcl)   # sets CUDA_VISIBLE_DEVICES as an import side effect (rcl.py:16)
    if env_before is None:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = env_before
    torch.Tensor.cuda = lambda self, *a, **k: self       # rcl.py:29-30 call .cuda() unconditionally
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(77)
    uni = torch.rand(2, 3, 37, 53, generator=g)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, 48), torch.linspace(0, 1, 64), indexing="ij")
    img = torch.stack([0.5 + 0.4 * torch.sin(9 * xx + 3 * yy), yy * xx, (xx - 0.5).abs()])[None].clone()
    img += 0.02 * torch.rand(img.shape, generator=g)
    img = img.clamp(0, 1)
    img[:, :, 10:22, 30:45] = 0.0      # black block: blurred == 0 inside -> 0/0
    img[:, :, 30:40, 5:20] = 1.0       # saturated block
    k = rcl.create_blur_kernel()
    out = {"blur_kernel": k.numpy(), "uni": uni.numpy(), "img": img.numpy()}
    with torch.no_grad():
        for name, x in (("uni", uni), ("img", img)):
            for it in (1, 5):
                out[f"{name}_it{it}"] = rcl.r_l_per_channel(x, k, it, 0.01).numpy()
    np.savez_compressed(os.path.join(HERE, "rl_deconv.npz"), **out)
    print({k_: v.shape for k_, v in out.items()})


if __name__ == "__main__":
    main()
