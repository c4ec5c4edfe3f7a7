"""Host side of the peer-memory ClipLoss exchange (csrc/p2p.cu): symmetric buffers, pointer tables, epochs.

PyTorch is plumbing here: `torch.distributed._symmetric_memory` allocates one buffer per rank and maps every peer's copy
into this process; the kernels only ever see the raw, pre-offset pointers collected in the small device tables below.

Buffer layout per rank (fp32 words, S = n * 2D = one rank's packed img | txt rows, W = world * S, R = (world + 2) * S):
    [ gather ring: RING x W ][ receive ring: RING x R ]
A receive buffer has one slot per rank plus two local slots (the two K halves of the local-row gradient terms).
Sync pad per rank (uint32 words): [0, 16) all-gather flags (word q = epoch last published by rank q), [16, 32) reduce-scatter
flags, [32, 48) block counters of the local all-gather kernel, [48, 48 + RING) busy words of the ring slots (the epoch whose
gathered rows a pending backward of THIS rank still needs, 0 = free; peers wait on it before they overwrite the slot).

Waits inside the kernels are bounded in wall time (B200CLIP_P2P_TIMEOUT_S, default 600 s — rank skew of many seconds is normal
when rank 0 evaluates / checkpoints between epochs); an expired wait raises a pinned-host error word that `check_error`
turns into a B200ClipError at the next exchange instead of trapping the CUDA context.
"""
from __future__ import annotations

import os
import warnings

import torch

from .. import _lib as L

RING = 4          # forward passes whose gathered features may be alive (saved for a backward) at the same time
MAX_WORLD = 16


def enabled() -> bool:
    return os.environ.get("B200CLIP_P2P", "1") != "0"


class _SlotToken:
    """Held by an autograd ctx: the ring slot is released when the ctx dies (after its backward, or when the graph is freed)."""

    def __init__(self, state: "PeerExchange", slot: int):
        self.state, self.slot = state, slot
        state.inflight.add(slot)

    def __del__(self):
        self.state.inflight.discard(self.slot)


_err_host = None


def _error_word() -> torch.Tensor:
    """Process-wide host-visible error word of the bounded waits (pinned: the kernels store to it, the host reads it without
    a synchronisation); registered once together with the wait bound."""
    global _err_host
    if _err_host is None:
        _err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        L.check(L.load().b200clip_p2p_configure(float(os.environ.get("B200CLIP_P2P_TIMEOUT_S", "600")), _err_host.data_ptr()),
                "b200clip_p2p_configure")
    return _err_host


class PeerExchange:
    """Two-phase set-up so that the ranks can agree after each phase (get_exchange): `__init__` only allocates (local, may
    fail on one rank alone: out of memory, no P2P), `connect` runs the collective rendezvous and builds the pointer tables."""

    def __init__(self, n: int, D: int, rank: int, world: int, device: torch.device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        group = group if group is not None else dist.group.WORLD
        self.n, self.D, self.rank, self.world, self.device = n, D, rank, world, device
        self.S = n * 2 * D
        self.W = world * self.S
        self.R = (world + 2) * self.S
        self.data = symm_mem.empty(RING * (self.W + self.R), dtype=torch.float32, device=device)
        self.sync = symm_mem.empty(64, dtype=torch.int32, device=device)
        self.sync.zero_()
        self.err_host = _error_word()
        torch.cuda.synchronize(device)
        self.group = group

    def connect(self) -> None:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        n, rank, world, device, group = self.n, self.rank, self.world, self.device, self.group
        hd = symm_mem.rendezvous(self.data, group.group_name)
        hs = symm_mem.rendezvous(self.sync, group.group_name)
        dptr, sptr = [int(p) for p in hd.buffer_ptrs], [int(p) for p in hs.buffer_ptrs]

        def table(vals):
            return torch.tensor(vals, dtype=torch.int64, device=device)

        # where THIS rank's rows go on every peer p (slot `rank` of p's gather buffer), per ring slot
        self.ag_dst = [table([dptr[p] + 4 * (s * self.W + rank * self.S) for p in range(world)]) for s in range(RING)]
        # where the gradient block of rank j's rows goes (slot `rank` of j's receive buffer), per ring slot
        # (+ this rank's two local slots)
        self.rs_dst = [table([dptr[j] + 4 * (RING * self.W + s * self.R + rank * self.S) for j in range(world)] +
                             [dptr[rank] + 4 * (RING * self.W + s * self.R + (world + h) * self.S) for h in range(2)]) for s in range(RING)]
        self.ag_flag = table([sptr[p] + 4 * rank for p in range(world)])
        self.rs_flag = table([sptr[p] + 4 * (16 + rank) for p in range(world)])
        base = self.sync.data_ptr()
        self.my_ag_flags, self.my_rs_flags, self.counters = base, base + 64, base + 128
        # busy word of ring slot s: mine, and every peer's (checked before a peer's slot is overwritten)
        self.my_busy = [base + 4 * (48 + s) for s in range(RING)]
        self.peer_busy = [table([sptr[p] + 4 * (48 + s) for p in range(world)]) for s in range(RING)]
        # loss workspace (raw logits, row statistics) per ring slot: lives exactly as long as the slot is held
        self.ws = [torch.empty((2 * n * world * n + 8 * n + 8,), dtype=torch.float32, device=device) for _ in range(RING)]
        self.ag_epoch = 0
        self.rs_epoch = 0
        self.inflight: set[int] = set()
        dist.barrier(group=group)            # every rank's sync pad is zeroed and mapped before anyone signals
        torch.cuda.synchronize(device)

    # ------------------------------------------------------------------------------------------------------------
    def gathered_view(self, slot: int) -> torch.Tensor:
        return self.data[slot * self.W:(slot + 1) * self.W].view(self.world * self.n, 2 * self.D)

    def recv_view(self, slot: int) -> torch.Tensor:
        off = RING * self.W + slot * self.R
        return self.data[off:off + self.R]

    def check_error(self) -> None:
        """Raise if a bounded wait of an earlier exchange expired (its results were undefined)."""
        code = int(self.err_host[0])
        if code != 0:
            what = {1: "a peer's flag never arrived (dead or badly skewed rank)",
                    2: "a peer never released the ring slot (its forward is still waiting for a backward)"}.get(code, f"code {code}")
            raise L.B200ClipError(f"ClipLoss peer exchange: a wait timed out — {what}; results of that step are invalid "
                                  "(B200CLIP_P2P_TIMEOUT_S sets the bound, B200CLIP_P2P=0 selects the NCCL path)")

    def all_gather(self, img: torch.Tensor, txt: torch.Tensor, hold: bool = True) -> tuple[torch.Tensor, int]:
        """img, txt [n, D] (fp32 / bf16 / fp16, contiguous) -> (gathered [N, 2D] fp32 view of the local ring slot, slot).
        `hold`: a backward will read the gathered rows (the slot stays protected from the peers until reduce_scatter_finish)."""
        self.check_error()
        self.ag_epoch += 1
        # epoch 0 means "free" in the busy words: skip it when the 32-bit counter wraps
        if self.ag_epoch & 0xFFFFFFFF == 0:
            self.ag_epoch += 1
        slot = self.ag_epoch % RING
        if slot in self.inflight:
            raise L.B200ClipError(f"ClipLoss peer exchange: more than {RING - 1} forward passes are waiting for their backward")
        rc = L.load().b200clip_p2p_allgather(L.dtype_code(img.dtype), img.data_ptr(), txt.data_ptr(), self.n, self.D,
                                             self.ag_dst[slot].data_ptr(), self.ag_flag.data_ptr(), self.my_ag_flags, self.counters,
                                             self.world, self.ag_epoch & 0xFFFFFFFF, self.peer_busy[slot].data_ptr(),
                                             self.my_busy[slot], int(hold), L.stream_ptr())
        L.check(rc, "b200clip_p2p_allgather")
        return self.gathered_view(slot), slot

    def reduce_scatter_finish(self, slot: int) -> torch.Tensor:
        """After the slot-addressed backward GEMM of every rank: -> [2, n, D] = (d_img, d_txt) of the local rows, each a dense
        tensor (autograd's AccumulateGrad takes a dense gradient as is; a column slice of [n, 2D] would be copied)."""
        self.check_error()
        self.rs_epoch += 1
        out = torch.empty((2, self.n, self.D), dtype=torch.float32, device=self.device)
        rc = L.load().b200clip_p2p_reduce_finish(self.recv_view(slot).data_ptr(), out.data_ptr(), self.S, self.rs_flag.data_ptr(),
                                                 self.my_rs_flags, self.world, self.world + 2, self.rs_epoch & 0xFFFFFFFF, self.my_busy[slot],
                                                 2 * self.D, L.stream_ptr())
        L.check(rc, "b200clip_p2p_reduce_finish")
        return out


_states: dict = {}
_failed = False


def get_exchange(n: int, D: int, rank: int, world: int, device: torch.device, group=None):
    """One PeerExchange per (shape, device, group); None when peer memory is unavailable (the caller then uses NCCL)."""
    global _failed
    if _failed or not enabled():
        return None
    key = (n, D, rank, world, device.index, id(group))
    st = _states.get(key)
    if st is None:
        import torch.distributed as dist
        if world > MAX_WORLD:
            raise L.B200ClipError(f"peer exchange supports up to {MAX_WORLD} ranks, got {world}")
        err = None

        def all_ok(flag: bool) -> bool:
            # The choice between the peer path and NCCL must be the SAME on every rank (a rank in NCCL's all-gather and a
            # rank spinning on peer flags would wait for each other forever): MIN over the ranks' success flags.
            ok = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            return int(ok) == 1

        try:
            st = PeerExchange(n, D, rank, world, device, group)     # local allocation only
        except Exception as e:  # symmetric memory not supported here (no P2P, out of memory, old driver ...)
            st, err = None, e
        ok = all_ok(st is not None)
        if ok:
            try:
                st.connect()                                        # collective rendezvous: every rank is in it
            except Exception as e:
                err = e
            ok = all_ok(err is None)
        if not ok:
            _failed = True
            st = None           # drops this rank's symmetric buffers if it had got that far
            why = f"{type(err).__name__}: {err}" if err is not None else "another rank could not set it up"
            warnings.warn(f"b200clip: peer-memory ClipLoss exchange unavailable ({why}); every rank uses NCCL collectives")
            return None
        _states[key] = st
    return st
