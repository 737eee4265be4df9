"""Multi-GPU partitioning of the hot path: independent clips, one process per GPU.

Every clip (batch row) is independent end to end (speinet.py:150-168 only masks rows), so the path
shards with NO data-path collective: rank r owns the clips with `clip_id % world == r`
(SURVEY.md section 8(e)).  The only exchange is one all-gather of the per-clip outputs at the end,
replacing the reference's single-process `nn.DataParallel` gather (inference_SPEINet.py:235,569).
Works with NCCL (GPU tensors) and gloo (CPU tensors, used by the CPU tests).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_clips(num_clips: int, rank: int, world: int) -> List[int]:
    """Clip ids owned by `rank` (round-robin, so ragged totals differ by at most one clip)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, num_clips, world))


class _PendingGather:
    """Handle of an asynchronous `gather_outputs`: `result()` waits for the collective (stream-side) and returns the
    [num_clips, ...] tensor in clip order."""

    def __init__(self, work, recv, send, num_clips, world, per, tail):
        self.work, self.recv, self.send = work, recv, send
        self.num_clips, self.world, self.per, self.tail = num_clips, world, per, tail

    def result(self) -> torch.Tensor:
        if self.work is not None:
            self.work.wait()
        recv = self.recv.view((self.world, self.per) + tuple(self.tail))
        out = recv.new_empty((self.num_clips,) + tuple(self.tail))
        for r in range(self.world):
            ids = shard_clips(self.num_clips, r, self.world)
            if ids:
                out[ids] = recv[r, :len(ids)]
        return out


def gather_outputs(local: torch.Tensor, num_clips: int, rank: int, world: int, group=None, async_op: bool = False):
    """All-gather per-clip outputs back into clip order.

    `local` is [len(shard_clips(num_clips, rank, world)), ...]; the result is [num_clips, ...] on every
    rank.  Ragged shards are padded to the longest shard for the collective and trimmed afterwards.
    With `async_op` the collective is only enqueued (NCCL's stream waits for the producer of `local`, the caller's
    stream does not wait for NCCL) and a handle is returned whose `result()` completes it -- the gather of clip i then
    overlaps the kernels of clip i+1."""
    mine = shard_clips(num_clips, rank, world)
    if local.shape[0] != len(mine):
        raise ValueError(f"rank {rank} holds {local.shape[0]} clips, expected {len(mine)}")
    tail = local.shape[1:]
    if world == 1:
        return _PendingGather(None, local, local, num_clips, 1, num_clips, tail) if async_op else local
    per = (num_clips + world - 1) // world
    if len(mine) == per:
        send = local.contiguous()
    else:
        send = local.new_zeros((per,) + tuple(tail))
        send[:len(mine)] = local
    recv = local.new_empty((world * per,) + tuple(tail))
    work = dist.all_gather_into_tensor(recv, send, group=group, async_op=async_op)
    pending = _PendingGather(work if async_op else None, recv, send, num_clips, world, per, tail)
    return pending if async_op else pending.result()


class PeerGather:
    """`gather_outputs` over NVLink PEER MEMORY instead of NCCL kernels: every rank copies its shard straight into the peers'
    symmetric buffers (`torch.distributed._symmetric_memory`; cudaMemcpyPeer = the GPUs' copy engines through NVSwitch) and a
    device-side signal barrier closes the exchange.  No SM runs collective code, so the power-capped persistent tcgen05 search
    kernel of the next clip keeps the whole chip -- measured at 8 GPUs the NCCL all-gather of the frames cost 0.5 ms per
    3.2 ms step even when launched asynchronously (bench.py collective_ab).

    One instance serves a fixed per-clip shape; `slots` exchanges can be in flight.  Usage per step:
        slot = pg.push(local)          # enqueue (side stream waits for the producer of `local` on the current stream)
        ...                            # next clip's kernels
        full = pg.result(slot)         # [num_clips, ...] in clip order; the current stream waits for the exchange"""

    def __init__(self, tail_shape, num_clips: int, rank: int, world: int, dtype=torch.float32, device=None, group=None,
                 slots: int = 2):
        import torch.distributed._symmetric_memory as symm
        self.rank, self.world, self.num_clips, self.slots = rank, world, num_clips, slots
        self.per = (num_clips + world - 1) // world
        self.mine = shard_clips(num_clips, rank, world)
        self.tail = tuple(tail_shape)
        self.dev = torch.device(device if device is not None else torch.cuda.current_device())
        shape = (slots, world, self.per) + self.tail
        self.buf = symm.empty(shape, dtype=dtype, device=self.dev)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.peers = [self.hdl.get_buffer(r, shape, dtype) for r in range(world)]
        self.side = torch.cuda.Stream(self.dev)
        self.done = [torch.cuda.Event() for _ in range(slots)]
        self.consumed = [None] * slots
        self.pending = [False] * slots
        self.next_slot = 0

    def push(self, local: torch.Tensor) -> int:
        if local.shape[0] != len(self.mine) or tuple(local.shape[1:]) != self.tail:
            raise ValueError(f"rank {self.rank}: expected [{len(self.mine)}, {self.tail}], got {tuple(local.shape)}")
        slot = self.next_slot
        self.next_slot = (slot + 1) % self.slots
        if any(self.pending):
            # one exchange in flight at a time: result(i) before push(i+1) is what orders this rank's reads of a slot
            # before the peers' next writes to it (the barrier of exchange i+1 is entered only after those reads)
            raise RuntimeError("PeerGather: take result() of the previous exchange before the next push()")
        cur = torch.cuda.current_stream(self.dev)
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            # every earlier result of ANY slot on this rank has been consumed before this rank arrives at the barrier
            # below; a peer that has passed the barrier may therefore overwrite those slots
            for ev in self.consumed:
                if ev is not None:
                    self.side.wait_event(ev)
            n = len(self.mine)
            if n:
                for r in range(self.world):
                    self.peers[r][slot, self.rank, :n].copy_(local, non_blocking=True)
            self.hdl.barrier(channel=slot)
            self.done[slot].record(self.side)
        local.record_stream(self.side)
        self.pending[slot] = True
        return slot

    def result(self, slot: int, copy: bool = True, out: torch.Tensor = None) -> torch.Tensor:
        """[num_clips, ...] in clip order (written into `out` when given).  With copy=False (only when every rank holds
        exactly one clip) a view of the exchange buffer is returned: valid until this rank's next push()."""
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.done[slot])
        self.pending[slot] = False
        recv = self.buf[slot]                                        # [world, per, ...]; clip id = i * world + r
        if self.per == 1 and self.num_clips == self.world and not copy and out is None:
            res = recv[:, 0]
        else:
            src = recv.transpose(0, 1).reshape((self.world * self.per,) + self.tail)[:self.num_clips] if self.per > 1 \
                else recv[:self.num_clips, 0]
            res = src.clone() if out is None else out.copy_(src)
        ev = self.consumed[slot] or torch.cuda.Event()
        ev.record(cur)
        self.consumed[slot] = ev
        return res


# --------------------------------------------------------------------------------------------------
# Single very large frame: shard the QUERY ROWS of the lv3 grid (second row)
# --------------------------------------------------------------------------------------------------
def row_band(h: int, rank: int, world: int):
    """Rows [y0, y1) of the lv3 query grid owned by `rank`, and the padded band [p0, p1) it must search.

    T at row y needs the argmax of rows y-1..y+1 (3x3 fold neighbourhood) and the patch of row y needs
    pixel rows y-1..y+1, so the searched band carries a 2-row halo; recomputing the halo instead of
    exchanging indices keeps the path free of any data-path collective.  Unlike the reference's
    `forward_chop` quadrants (inference_SPEINet.py:545-607) every band still searches ALL keys, so
    the result is identical to the unsharded one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    y0, y1 = rank * h // world, (rank + 1) * h // world
    return y0, y1, max(0, y0 - 2), min(h, y1 + 2)


def search_transfer_rows(module, lrsr_lv3, refsr_lv3, ref_lv1, ref_lv2, ref_lv3, rank: int, world: int):
    """This rank's row band of SearchTransfer.forward for one large frame: returns
    (S, T_lv3, T_lv2, T_lv1) restricted to lv3 rows [y0, y1) (x1 / x2 / x4 for the pyramid levels)."""
    h = lrsr_lv3.shape[2]
    y0, y1, p0, p1 = row_band(h, rank, world)
    if y1 <= y0:
        raise ValueError(f"rank {rank} of {world} owns no rows of a {h}-row grid")
    S, T3, T2, T1 = module(lrsr_lv3[:, :, p0:p1].contiguous(), refsr_lv3, ref_lv1, ref_lv2, ref_lv3)
    a, b = y0 - p0, y1 - p0
    return (S[:, :, a:b], T3[:, :, a:b], T2[:, :, 2 * a:2 * b], T1[:, :, 4 * a:4 * b])


def gather_rows(parts, h: int, rank: int, world: int, group=None):
    """All-gather the row bands of `search_transfer_rows` into full-height tensors (NCCL or gloo).
    `parts` = (S, T_lv3, T_lv2, T_lv1) bands of this rank; returns the four full tensors."""
    if world == 1:
        return tuple(parts)
    out = []
    for t, s in zip(parts, (1, 1, 2, 4)):
        full = t.new_empty(t.shape[:2] + (h * s,) + t.shape[3:])
        per = max(((r + 1) * h // world - r * h // world) for r in range(world)) * s
        send = t.new_zeros(t.shape[:2] + (per,) + t.shape[3:])
        send[:, :, :t.shape[2]] = t
        recv = t.new_empty((world * send.shape[0],) + tuple(send.shape[1:]))  # concatenated along dim 0 (gloo and NCCL)
        dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
        recv = recv.view((world,) + tuple(send.shape))
        for r in range(world):
            y0, y1 = r * h // world, (r + 1) * h // world
            full[:, :, y0 * s:y1 * s] = recv[r][:, :, :(y1 - y0) * s]
        out.append(full)
    return tuple(out)
