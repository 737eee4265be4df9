#!/bin/bash
# GPU check of the Richardson-Lucy kernel: parity tests, then timing against the reference op sequence in torch
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "rl_deconv or install" 2>&1 | tail -3
python - <<PY
import torch, torch.nn.functional as F, speinet_b200, json
x = torch.rand(1, 3, 720, 1280, device="cuda"); k = speinet_b200.create_blur_kernel().cuda()
lap = torch.tensor([[0, -1, 0], [-1, 4, -1], [0, -1, 0]], dtype=torch.float32, device="cuda")[None, None]
def ref(it):
    out = []
    for c in range(3):
        xc = x[:, c:c + 1]; d = xc.clone()
        for _ in range(it):
            cf = xc / F.conv2d(d, k, padding=2); cf[cf != cf] = 0.0; cf[cf < 0] = 0.0
            d = cf * (d + 0.01 * F.conv2d(d, lap, padding=1))
        out.append(d)
    return torch.cat(out, 1)
def timed(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e3
res = {}
for it in (1, 5):
    res[f"ours_it{it}_us"] = round(timed(lambda: speinet_b200.r_l_per_channel(x, k, it, 0.01)), 1)
    res[f"torch_ops_it{it}_us"] = round(timed(lambda: ref(it)), 1)
res["bytes_read_write"] = 2 * x.numel() * 4
res["ours_it5_GBs"] = round(res["bytes_read_write"] / res["ours_it5_us"] / 1e3, 1)
print(json.dumps(res)); open("gpurun_out/rl_deconv_timing.json", "w").write(json.dumps(res))
PY
