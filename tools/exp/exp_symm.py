"""Experiment: does torch symmetric memory (NVLink peer buffers) work on the GPU box, and what does a copy-engine
all-gather of one [3,720,1280] fp32 frame per rank cost next to NCCL's?"""
import os, time, json, sys
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
out = {"world": world}
try:
    import torch.distributed._symmetric_memory as symm
    shape = (2, world, 3, 720, 1280)
    buf = symm.empty(shape, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    peers = [hdl.get_buffer(r, shape, torch.float32) for r in range(world)]
    local = torch.full((3, 720, 1280), float(rank + 1), device=dev)
    side = torch.cuda.Stream(dev)
    def p2p(slot):
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for r in range(world):
                peers[r][slot, rank].copy_(local, non_blocking=True)
            hdl.barrier(channel=slot)
    p2p(0); torch.cuda.synchronize(); dist.barrier()
    ok = all(float(buf[0, r].mean()) == r + 1 for r in range(world))
    out["p2p_correct"] = ok
    for name, fn in (("p2p_copy_engine", lambda i: p2p(i & 1)),):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for i in range(20): fn(i)
        side.synchronize(); torch.cuda.synchronize()
        out[name + "_us"] = (time.perf_counter() - t0) / 20 * 1e6
    recv = torch.empty(world, 3, 720, 1280, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for i in range(20): dist.all_gather_into_tensor(recv, local.unsqueeze(0))
    torch.cuda.synchronize()
    out["nccl_all_gather_us"] = (time.perf_counter() - t0) / 20 * 1e6
except Exception as e:
    import traceback
    out["error"] = repr(e)[:300]; out["tb"] = traceback.format_exc()[-800:]
if rank == 0: print(json.dumps(out))
dist.destroy_process_group()
