"""Host-buffer front end of the hot path: clips whose features live in pinned host memory are
streamed through SearchTransfer + the three fusion kernels with copy/compute overlap.

Three CUDA streams (host->device, compute, device->host) and two device buffer SLOTS: while clip i
computes, clip i+1's inputs are uploading and clip i-1's outputs are downloading, so the PCIe time
(~620 MB per 720p clip) hides behind the kernels instead of adding to it.  Every device buffer of a slot
(inputs, S / T / arg, fused features, the library workspace) is allocated once and reused, and -- because the
pointers of a slot never change -- the ~20 kernel launches of a clip are captured ONCE per slot into a CUDA graph
and replayed (`cuda_graph=True`): the per-launch host work (tensor-map encoding, attribute calls, ctypes) leaves the
steady state.  This is the inference-driver hygiene `SURVEY.md` section 8(f) row 4 asks for (the reference's
`.to(cuda)` / `.cpu()` per frame: inference_SPEINet.py:390,398), applied to the path's own boundary; it uses only the
public API (`search_transfer`, `fuse_level`).

`dec3` is optional: at speinet.py:93 the lv3 decoder feature IS the query (`f_fusion`), so a clip without a "dec3"
entry fuses into `q` and saves the 29.5 MB upload.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from .fusion import fuse_level
from .search_transfer import SearchTransfer, search_transfer

IN_KEYS = ("q", "lv3", "lv2", "lv1", "dec3", "dec2", "dec1")
OUT_KEYS = ("S", "f3", "f2", "f1")


class _Slot:
    def __init__(self):
        self.inp: Dict[str, torch.Tensor] = {}
        self.out: Dict[str, torch.Tensor] = {}      # persistent result buffers: S, T3, T2, T1, arg, stats, f3, f2, f1 (bf16 T / f for all-bf16 clips)
        self.fused: Dict[str, torch.Tensor] = {}    # what goes back to the host (S, f3, f2, f1 in the clip's dtype)
        self.graph = None
        self.sig = None
        self.warm = False


class HostPipeline:
    """`run(clips, outs)`: clips = list of dicts of pinned CPU tensors with keys IN_KEYS ("dec3" optional: defaults to
    "q"); outs = list of dicts of pinned CPU tensors with keys OUT_KEYS to receive S and the fused features.
    `convs` maps level -> (weight, bias) of conv_lv3/2/1 (speinet.py:55-57), already on the device.
    When `run` returns, every `outs[i]` is complete on the host."""

    def __init__(self, convs: Dict[int, Sequence[torch.Tensor]], device, search_module: SearchTransfer = None,
                 cuda_graph: bool = False):
        self.dev = torch.device(device)
        self.convs = convs
        self.st = search_module if search_module is not None else SearchTransfer().to(self.dev)
        self.cuda_graph = cuda_graph
        self.s_h2d = torch.cuda.Stream(self.dev)
        self.s_cmp = torch.cuda.Stream(self.dev)
        self.s_d2h = torch.cuda.Stream(self.dev)
        self._slots = [_Slot(), _Slot()]

    # ------------------------------------------------------------------ buffers
    def _prepare(self, slot: _Slot, clip: Dict[str, torch.Tensor]):
        keys = [k for k in IN_KEYS if k in clip]
        sig = tuple((k, tuple(clip[k].shape), clip[k].dtype) for k in keys)
        if sig != slot.sig:
            slot.inp = {k: torch.empty(clip[k].shape, dtype=clip[k].dtype, device=self.dev) for k in keys}
            n, c3, h, w = clip["q"].shape
            # all-bf16 clips run the native bf16 kernels (T / fused features stay bf16 on the device); anything else is fp32
            bf16 = all(clip[k].dtype == torch.bfloat16 for k in keys)
            f32 = dict(dtype=torch.float32, device=self.dev)
            io = dict(dtype=torch.bfloat16 if bf16 else torch.float32, device=self.dev)
            fo = io if (h * w) % 8 == 0 else f32          # the bf16 fusion kernel wants planes that are a multiple of 8 pixels
            slot.out = {"S": torch.empty((n, 1, h, w), **f32), "T3": torch.empty((n, c3, h, w), **io),
                        "T2": torch.empty((n, c3 // 2, 2 * h, 2 * w), **io), "T1": torch.empty((n, c3 // 4, 4 * h, 4 * w), **io),
                        "arg": torch.empty((n, h * w), dtype=torch.int64, device=self.dev),
                        "stats": torch.empty(8, dtype=torch.int32, device=self.dev),
                        "f3": torch.empty((n, c3, h, w), **fo), "f2": torch.empty((n, c3 // 2, 2 * h, 2 * w), **fo),
                        "f1": torch.empty((n, c3 // 4, 4 * h, 4 * w), **fo)}
            slot.fused = {}
            slot.graph, slot.sig, slot.warm = None, sig, False
        return slot.inp

    def _compute(self, slot: _Slot):
        d, o = slot.inp, slot.out
        st = self.st
        S, T3, T2, T1, arg, stats = search_transfer(d["q"], d["lv3"], d["lv1"], d["lv2"], d["lv3"], fold_mode=st.fold_mode,
                                                     search=st.search, eps=st.eps, out=o)
        st.last_index, st.last_stats = arg, stats
        dec3 = d.get("dec3", d["q"])
        fo = slot.fused
        fo["S"] = S
        for name, dec, T, lvl, sc in (("f3", dec3, T3, 3, 1), ("f2", d["dec2"], T2, 2, 2), ("f1", d["dec1"], T1, 1, 4)):
            fo[name] = fuse_level(dec, T, S, self.convs[lvl][0], self.convs[lvl][1], sc, out=o[name])

    def _launch(self, slot: _Slot):
        """Enqueue one clip's kernels on the compute stream (current stream): eagerly the first time (allocations,
        driver entry points, attribute calls happen here), then -- with cuda_graph -- captured once and replayed."""
        if not self.cuda_graph:
            self._compute(slot)
            return
        if slot.graph is not None:
            slot.graph.replay()
            return
        if not slot.warm:
            self._compute(slot)
            slot.warm = True
            return
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.s_cmp):
            self._compute(slot)
        slot.graph = g
        g.replay()          # capture does not execute: run the clip that triggered it

    # ------------------------------------------------------------------ run
    @torch.no_grad()
    def run(self, clips: List[Dict[str, torch.Tensor]], outs: List[Dict[str, torch.Tensor]]) -> None:
        n = len(clips)
        ev_h2d = [torch.cuda.Event() for _ in range(n)]
        ev_cmp = [torch.cuda.Event() for _ in range(n)]
        ev_d2h = [torch.cuda.Event() for _ in range(n)]
        start = torch.cuda.Event()
        start.record(torch.cuda.current_stream(self.dev))
        for s in (self.s_h2d, self.s_cmp, self.s_d2h):
            s.wait_event(start)
        for i, clip in enumerate(clips):
            slot = self._slots[i & 1]
            with torch.cuda.stream(self.s_h2d):
                if i >= 2:
                    self.s_h2d.wait_event(ev_cmp[i - 2])      # the slot's input buffers are free again
                d = self._prepare(slot, clip)
                for k, buf in d.items():
                    buf.copy_(clip[k], non_blocking=True)
                ev_h2d[i].record(self.s_h2d)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(ev_h2d[i])
                if i >= 2:
                    self.s_cmp.wait_event(ev_d2h[i - 2])      # the slot's output buffers have been read back
                self._launch(slot)
                ev_cmp[i].record(self.s_cmp)
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(ev_cmp[i])
                for name in OUT_KEYS:
                    outs[i][name].copy_(slot.fused[name], non_blocking=True)
                ev_d2h[i].record(self.s_d2h)
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_stream(self.s_d2h)
        cur.wait_stream(self.s_cmp)
        if n:
            ev_d2h[-1].synchronize()    # the pinned `outs` are complete when run() returns (s_d2h is in order)
