"""Fork/join of independent launch sequences on CUDA streams ("lanes").

The routed stack is a tree of independent chains -- the two branch stacks of a batch, and inside every routing
layer the K cells that read the same input (DynamicInteraction.py:44-48, 83-100).  Each chain is a sequence of
dependent kernels, and on a B200 most of them leave SMs idle: the 12800-row GEMMs run 2.03 waves of tiles, the
[B, D] global branches launch 6-24 CTAs.  Issuing the chains on separate streams lets the hardware fill those
holes with another chain's CTAs.  The pattern (side streams fork from and join into the caller's stream) is
capturable in a CUDA graph, which is how bench.py replays it.

Memory rule (torch's caching allocator is stream-ordered): a tensor allocated while lane i is current is only
ever reused by lane i, every fork waits on the caller's stream and every join happens before the caller
continues.  Callers must keep tensors that were allocated on the caller's stream and are read by a side lane
alive until ``join()``.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Tuple

import torch


class Lanes:
    """``n`` lanes forked from the current stream: lane 0 IS the current stream, lanes 1..n-1 are cached side
    streams private to (device, parent stream, fork site), so nested forks (a stack inside run_pair) never share streams."""

    _side: Dict[Tuple[int, int, str, int, int], "torch.cuda.Stream"] = {}

    def __init__(self, device: torch.device, n: int, tag: str = "", priorities=None):
        """``priorities[i]`` (CUDA stream priority, -1 = high, 0 = default) of side lane i; default: the
        priority of the caller's stream, so nested forks inherit it."""
        self.n = n
        self.cur = None
        self.streams: List = [None]
        if n <= 1:
            return
        self.cur = torch.cuda.current_stream(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        for i in range(1, n):
            prio = self.cur.priority if priorities is None else priorities[i]
            key = (idx, self.cur.cuda_stream, tag, i, prio)
            st = Lanes._side.get(key)
            if st is None:
                st = Lanes._side[key] = torch.cuda.Stream(device=device, priority=prio)
            self.streams.append(st)
        # fork NOW, before lane 0 puts any work on the caller's stream: a later wait would order the side
        # streams behind that work and serialise everything
        for st in self.streams[1:]:
            st.wait_stream(self.cur)

    def lane(self, i: int):
        """Context manager under which lane i's launches are issued (i is taken modulo the number of lanes)."""
        i %= max(self.n, 1)
        if i == 0:
            return contextlib.nullcontext()
        return torch.cuda.stream(self.streams[i])

    def catch_up(self, i: int) -> None:
        """Lane i additionally waits for everything issued on the caller's stream so far."""
        i %= max(self.n, 1)
        if i:
            self.streams[i].wait_stream(self.cur)

    def mark(self, i: int):
        """Event at the current tail of lane i (None for lane 0): lets the caller's stream wait for the work issued
        on that lane SO FAR without waiting for what is issued on it later (see wait_mark)."""
        i %= max(self.n, 1)
        if i == 0:
            return None
        ev = torch.cuda.Event()
        ev.record(self.streams[i])
        return ev

    def wait_mark(self, ev) -> None:
        if ev is not None:
            self.cur.wait_event(ev)

    def join(self) -> None:
        for st in self.streams[1:]:
            self.cur.wait_stream(st)


class NoLanes:
    """Same interface, everything on the current stream (single-stream mode and the CPU emulation in tests/)."""

    def __init__(self, device=None, n: int = 1):
        self.n = 1

    def lane(self, i: int):
        return contextlib.nullcontext()

    def catch_up(self, i: int) -> None:
        pass

    def mark(self, i: int):
        return None

    def wait_mark(self, ev) -> None:
        pass

    def join(self) -> None:
        pass


_AUX: Dict[Tuple[int, int], "torch.cuda.Stream"] = {}


def aux_stream(parent) -> "torch.cuda.Stream":
    """Cached helper stream of ``parent`` for small side computations (bias-gradient column sums) that nothing
    on the parent's chain waits for."""
    key = (parent.device.index, parent.cuda_stream)
    st = _AUX.get(key)
    if st is None:
        st = _AUX[key] = torch.cuda.Stream(device=parent.device, priority=parent.priority)
    return st


# Concurrency switches (read at every call): ENABLED=False puts every launch on the caller's stream.
ENABLED = True
CELL_LANES = 5      # lanes per routing layer: [K/V + GLAC local | IMRC | CMRC | CRCMC | routers + GESC + GLAC global]
PRIORITIZE_FIRST_BLOCK = False  # run_pair: high stream priority for the first (text, heavier) stack -- measured
                                # neutral (23.2 vs 23.1 ms), off
AUX_BIAS = True     # bias-gradient column sums on a helper stream beside the wgrad / dgrad GEMMs
AUX_WGRAD = False    # (experiment) weight-gradient GEMMs on the helper stream as well
FWD_LANES = True    # (bring-up switches: cell lanes in the forward / backward pass)
BWD_LANES = True


def fork(device: torch.device, n: int, tag: str = "", priorities=None):
    """``tag`` names the fork site: forks at different sites under the same parent stream get different side
    streams (run_pair's second stack must not share a stream with a cell lane of the first)."""
    if not ENABLED or n <= 1 or device.type != "cuda":
        return NoLanes()
    return Lanes(device, n, tag, priorities)
