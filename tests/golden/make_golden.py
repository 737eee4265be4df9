#!/usr/bin/env python
"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It loads /root/reference/model/SearchTransfer.py by file path (avoids model/__init__.py
and the import side effects of model/rcl.py), feeds it seeded inputs and
stores inputs + outputs as .npz.  The fusion fixtures evaluate the exact torch expressions
of /root/reference/model/speinet.py:93, 96 and 108 with Conv2d layers shaped as
speinet.py:55-57.  The reference ships no golden vectors of its own, so
these files are what pins the oracle; the GPU box never sees /root/reference.
"""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SPEINET_REFERENCE", "/root/reference")


def load_reference():
    spec = importlib.util.spec_from_file_location("_ref_search_transfer", os.path.join(REF, "model", "SearchTransfer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def pyramid(gen, n, hr, wr, std=0.04):
    lv3 = torch.randn(n, 128, hr, wr, generator=gen) * std
    lv2 = torch.randn(n, 64, 2 * hr, 2 * wr, generator=gen) * std
    lv1 = torch.randn(n, 32, 4 * hr, 4 * wr, generator=gen) * std
    return lv1, lv2, lv3


def run_case(ref_mod, name, q, lv1, lv2, lv3):
    torch.set_num_threads(1)  # deterministic summation order inside MKL
    st = ref_mod.SearchTransfer()
    with torch.no_grad():
        # call convention of speinet.py:135: (f_fusion, sharp_lv3, sharp_lv1, sharp_lv2, sharp_lv3)
        S, T3, T2, T1 = st(q, lv3, lv1, lv2, lv3)
        # the module does not return the index: recompute it with the same ops (:26-34)
        qu = F.normalize(F.unfold(q, 3, padding=1), dim=1)
        ku = F.normalize(F.unfold(lv3, 3, padding=1).permute(0, 2, 1), dim=2)
        r_star, r_arg = torch.max(torch.bmm(ku, qu), dim=1)
    assert torch.equal(r_star.view_as(S), S)
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        q=q.numpy(), ref_lv1=lv1.numpy(), ref_lv2=lv2.numpy(), ref_lv3=lv3.numpy(),
                        S=S.numpy(), T_lv3=T3.numpy(), T_lv2=T2.numpy(), T_lv1=T1.numpy(),
                        arg=r_arg.numpy().astype(np.int32))
    print(name, tuple(S.shape), tuple(T3.shape), tuple(T2.shape), tuple(T1.shape))


def main():
    ref_mod = load_reference()
    gen = torch.Generator().manual_seed(20261018)

    # 1. same-size query / reference grid, N=1 (the shape relation of speinet.py:135)
    q = torch.randn(1, 128, 10, 12, generator=gen) * 0.2
    run_case(ref_mod, "st_same_grid", q, *pyramid(gen, 1, 10, 12))

    # 2. ragged: N=2, reference grid differs from the query grid, odd sizes
    q = torch.randn(2, 128, 9, 11, generator=gen) * 0.2
    run_case(ref_mod, "st_ragged", q, *pyramid(gen, 2, 7, 13))

    # 3. edge cases: an all-zero query neighbourhood (S=0, arg=0), duplicated key patches
    #    (first index wins), and query == key (S~1 in the interior)
    lv1, lv2, lv3 = pyramid(gen, 1, 8, 16)
    lv3[:, :, :, 8:] = lv3[:, :, :, :8]           # right half duplicates the left half
    q = lv3.clone()
    q[:, :, 0:3, 0:3] = 0                          # query patch at (1,1) is entirely zero
    run_case(ref_mod, "st_edge", q, lv1, lv2, lv3)

    # 4. SelfTransfer (SearchTransfer.py:53-79): S from the search (:59-72), T_lv3 = the input itself, T_lv2 / T_lv1 =
    #    relu(conv1x1(bicubic_x2(.))) with the module's own search1 / search2 (:70-76)
    torch.manual_seed(7)
    selft = ref_mod.SelfTransfer()
    q = torch.randn(1, 128, 8, 12, generator=gen) * 0.2
    with torch.no_grad():
        S, T3, T2, T1 = selft(q)
    assert T3 is q
    np.savez_compressed(os.path.join(HERE, "self_transfer.npz"), q=q.numpy(), S=S.numpy(), T_lv2=T2.numpy(), T_lv1=T1.numpy(),
                        **{"sd_" + k: v.numpy() for k, v in selft.state_dict().items()})
    print("self_transfer", tuple(S.shape), tuple(T2.shape), tuple(T1.shape))

    # 5. fusion lines speinet.py:93-94, 96-97, 108-109 with conv_lv3/2/1 of speinet.py:55-57
    torch.manual_seed(11)
    n_feat, h, w = 32, 6, 10
    convs = {3: nn.Conv2d(n_feat * 8, n_feat * 4, 1), 2: nn.Conv2d(n_feat * 4, n_feat * 2, 1),
             1: nn.Conv2d(n_feat * 2, n_feat, 1)}
    S = torch.rand(2, 1, h, w, generator=gen) * 0.2
    out = {"S": S.numpy()}
    with torch.no_grad():
        for lvl, scale, ch in ((3, 1, 128), (2, 2, 64), (1, 4, 32)):
            dec = torch.randn(2, ch, h * scale, w * scale, generator=gen) * 0.3
            t = torch.randn(2, ch, h * scale, w * scale, generator=gen) * 0.05
            if scale == 1:
                f = dec + convs[lvl](torch.cat((dec, t), dim=1)) * S                                   # :93-94
            else:
                f = dec + convs[lvl](torch.cat((dec, t), dim=1)) * F.interpolate(S, scale_factor=scale, mode='bicubic')  # :96-97,:108-109
            out.update({f"dec{lvl}": dec.numpy(), f"t{lvl}": t.numpy(), f"f{lvl}": f.numpy(),
                        f"w{lvl}": convs[lvl].weight.numpy(), f"b{lvl}": convs[lvl].bias.numpy()})
        out["S_up2"] = F.interpolate(S, scale_factor=2, mode='bicubic').numpy()
        out["S_up4"] = F.interpolate(S, scale_factor=4, mode='bicubic').numpy()
    np.savez_compressed(os.path.join(HERE, "fusion.npz"), **out)
    print("fusion ok")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"reference not found at {REF}")
    main()
