#!/usr/bin/env python
"""Candidate-window statistics of the bf16 pass on REAL SPEINet features (build container, CPU only).

Runs the shimmed reference SPEINet (random-init weights, image-like synthetic clip) at a size the CPU
can handle, captures (f_fusion, sharp_lv3) at speinet.py:135 and emulates the tcgen05 candidate pass in
numpy: operands rounded to bf16, fp32 patch norms, exact accumulation.  Reports, for the default and the
rigorous window, how many keys fall inside the window of the best bf16 score (= candidates rescored),
how many queries would saturate a kTopK=8 list (= exhaustive fallback) and the largest bf16 scoring
error -- the evidence behind `eps` in DESIGN.md section 4(b').   Usage: python tests/diag/analyze_real_features.py [H W]
"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_model as shim  # noqa: E402
import oracle  # noqa: E402


def bf16(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float32).numpy()


def main():
    hh = int(sys.argv[1]) if len(sys.argv) > 1 else 192
    ww = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    shim.install_shims()
    sys.path.insert(0, shim.REF)
    from model import speinet
    os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.manual_seed(0)
    args = types.SimpleNamespace(patch_size=hh, window_size=4, rgb_range=1, depths=[6] * 6, embed_dim=256, num_heads=[8] * 6,
                                 mlp_ratio=2, resi_connection="1conv", n_colors=3, n_sequence=3, n_resblock=3, n_feat=32, cpu=True)
    net = speinet.SPEINet(in_channels=3, n_sequence=3, out_channels=3, n_resblock=3, n_feat=32, device="cpu", args=args).eval()
    rec = {}
    orig = net.SearchTransfer.forward

    def spy(a, b, c, d, e):
        rec.update(q=a.numpy(), k=b.numpy())
        n, _, h, w = a.shape
        return (torch.zeros(n, 1, h, w), torch.zeros_like(e), torch.zeros_like(d), torch.zeros_like(c))  # skip the slow CPU search

    net.SearchTransfer.forward = spy
    gen = torch.Generator().manual_seed(7)
    low = torch.rand(5, 3, hh // 16, ww // 16, generator=gen)
    x = F.interpolate(low, size=(hh, ww), mode="bicubic").clamp(0, 1)
    # sharp frames = the middle frame plus texture; blurry frames = box-blurred versions (crude motion blur)
    tex = 0.1 * torch.rand(1, 3, hh, ww, generator=gen)
    sharp = (x[1:2] + tex).clamp(0, 1)
    blur = F.avg_pool2d(F.pad(sharp, (4, 4, 0, 0), mode="replicate"), (1, 9), stride=1)
    clip = torch.stack([blur[0], blur[0].roll(2, -1), blur[0].roll(-2, -1), sharp[0].roll(3, -1), sharp[0].roll(-3, -1)])[None]
    with torch.no_grad():
        net(clip)
    q, k = rec["q"], rec["k"]
    n, c, h, w = q.shape
    L = h * w
    qu = oracle.l2_normalize(oracle.unfold(q, 3, 1, 1), axis=1)[0]         # [1152, L] fp32-normalised
    ku = oracle.l2_normalize(oracle.unfold(k, 3, 1, 1), axis=1)[0]
    R = (ku.T.astype(np.float64) @ qu.astype(np.float64)).astype(np.float32)   # exact relevance [Lk, L]
    # bf16 pass: raw operands rounded to bf16, fp32 reciprocal norms applied afterwards
    qn = np.maximum(np.linalg.norm(oracle.unfold(q, 3, 1, 1)[0], axis=0), 1e-12)
    kn = np.maximum(np.linalg.norm(oracle.unfold(k, 3, 1, 1)[0], axis=0), 1e-12)
    qb, kb = oracle.unfold(bf16(q), 3, 1, 1)[0], oracle.unfold(bf16(k), 3, 1, 1)[0]
    Rb = ((kb.T.astype(np.float64) @ qb.astype(np.float64)) / kn[:, None] / qn[None, :]).astype(np.float32)
    best_b = Rb.max(axis=0)
    top = np.argmax(R, axis=0)
    out = {"grid": [h, w], "queries": L, "S_mean": float(R.max(axis=0).mean()), "S_min": float(R.max(axis=0).min()),
           "max_abs_bf16_error_all_pairs": float(np.abs(Rb - R).max()),
           "max_abs_bf16_error_at_true_best": float(np.abs(Rb[top, np.arange(L)] - R[top, np.arange(L)]).max()),
           "argmax_changed_by_bf16_pct": float((np.argmax(Rb, axis=0) != top).mean() * 100)}
    for eps in (1e-3, 2e-3, 4e-3, 8e-3):
        inside = (Rb >= best_b[None, :] - eps).sum(axis=0)
        true_kept = Rb[top, np.arange(L)] >= best_b - eps
        out[f"eps_{eps:g}"] = {"candidates_mean": float(inside.mean()), "candidates_max": int(inside.max()),
                               "queries_with_8_or_more_pct": float((inside >= 8).mean() * 100),
                               "true_argmax_inside_window_pct": float(true_kept.mean() * 100)}
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"r01_real_feature_window_stats_{h}x{w}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
