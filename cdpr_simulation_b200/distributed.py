"""Multi-GPU plumbing: one process per GPU, instances sharded by contiguous range, no collective inside a step.

NCCL (torch.distributed) is used only where the path has a real exchange:
  * the decimated trajectory gather (snapshots written by the step kernel straight into the send buffer,
    all-gathered on a side stream while the next pass runs), and
  * the all-reduce of the per-sequence rollout cost vector.
Everything here also runs on CPU tensors with the gloo backend (tests/test_distributed_gloo.py)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .workloads import shard_range


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def gather_trajectory(local: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """local: [n_snap, 13, n_local] of this rank; returns [world, n_snap, 13, n_local] ordered by rank, i.e. by
    global instance id when every rank owns the same number of instances."""
    rank, ws = world()
    if out is None:
        out = torch.empty((ws,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    if ws == 1:
        out[0].copy_(local)
        return out
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1))
    return out


def allreduce_cost(cost_seq: torch.Tensor) -> torch.Tensor:
    """Sum over ranks of the per-sequence rollout cost (each rank holds the sum over ITS robots)."""
    _, ws = world()
    if ws > 1:
        dist.all_reduce(cost_seq, op=dist.ReduceOp.SUM)
    return cost_seq


def global_trajectory_to_instance_major(gathered: torch.Tensor) -> torch.Tensor:
    """[world, n_snap, 13, n_local] -> [n_snap, 13, world * n_local] (global instance id fastest)."""
    ws, n_snap, f, n_local = gathered.shape
    return gathered.permute(1, 2, 0, 3).reshape(n_snap, f, ws * n_local)


class TrajectoryGather:
    """Double-buffered snapshot send buffers + an all-gather per pass on a side stream.

    before_pass(): point the step kernel at the send buffer of this pass (after the gather that last read it).
    after_pass():  enqueue the all-gather of that buffer behind the step kernel, on the communication stream.
    finish():      make the compute stream wait for every outstanding gather."""

    def __init__(self, batch, every: int, steps_per_pass: int, stream: torch.cuda.Stream):
        self.batch, self.every, self.stream = batch, int(every), stream
        self.n_snap = steps_per_pass // self.every
        _, ws = world()
        dev = torch.device("cuda", batch.device)
        self.send = [torch.empty((self.n_snap, 13, batch.n), dtype=torch.float64, device=dev) for _ in range(2)]
        self.recv = torch.empty((ws, self.n_snap, 13, batch.n), dtype=torch.float64, device=dev)
        self.comm = torch.cuda.Stream(device=dev)
        self.free = [None, None]
        self.i = 0

    def before_pass(self):
        slot = self.i % 2
        if self.free[slot] is not None:
            self.stream.wait_event(self.free[slot])
        self.batch.set_snapshots(self.every, self.send[slot].data_ptr(), self.n_snap)

    def after_pass(self):
        slot = self.i % 2
        done = torch.cuda.Event()
        done.record(self.stream)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(done)
            gather_trajectory(self.send[slot], self.recv)
            ev = torch.cuda.Event()
            ev.record(self.comm)
            self.free[slot] = ev
        self.i += 1

    def finish(self):
        self.stream.wait_stream(self.comm)


class FusedTrajectoryGather:
    """The trajectory all-gather fused into the step kernel over NVLink peer memory (no collective call).

    Every rank allocates the FULL gather buffer [n_snap][13][N_total] in symmetric memory (double-buffered) and maps
    all peers' buffers (torch.distributed._symmetric_memory rendezvous). The step kernel stores each snapshot of its
    own instances into column range [rank*n, (rank+1)*n) of EVERY rank's buffer with plain stores -- the transfer
    rides along with the compute, snapshot by snapshot. after_pass() enqueues a cross-rank barrier on the compute
    stream: once it passes, every peer's kernel has finished, so this rank's buffer holds the whole trajectory in
    global instance order."""

    def __init__(self, batch, every: int, steps_per_pass: int, stream: torch.cuda.Stream, multicast: bool = True):
        import torch.distributed._symmetric_memory as symm
        self.batch, self.every, self.stream = batch, int(every), stream
        self.n_snap = steps_per_pass // self.every
        self.rank, self.ws = world()
        if self.ws > 8:
            raise RuntimeError("cdpr_set_snapshot_peers takes at most 8 peers")
        dev = torch.device("cuda", batch.device)
        self.total = batch.n * self.ws
        self.bufs, self.hdls = [], []
        for _ in range(2):
            t = symm.empty((self.n_snap, 13, self.total), dtype=torch.float64, device=dev)
            self.bufs.append(t)
            self.hdls.append(symm.rendezvous(t, dist.group.WORLD))
        # NVLS: one multimem.st per value instead of one store per rank, when the switch supports it
        self.multicast = bool(multicast) and all(getattr(h, "has_multicast_support", False) and int(h.multicast_ptr) != 0 for h in self.hdls)
        self.i = 0

    def before_pass(self):
        slot = self.i % 2
        if self.multicast:
            self.batch.set_snapshot_multicast(self.every, int(self.hdls[slot].multicast_ptr), self.rank * self.batch.n, self.total, self.n_snap)
            return
        self.batch.set_snapshot_peers(self.every, [int(p) for p in self.hdls[slot].buffer_ptrs], self.rank * self.batch.n, self.total, self.n_snap)

    def after_pass(self):
        slot = self.i % 2
        with torch.cuda.stream(self.stream):
            self.hdls[slot].barrier(channel=slot)
        self.i += 1

    def finish(self):
        pass

    def latest(self) -> torch.Tensor:
        """[n_snap][13][N_total] of the last completed pass (valid after the stream reached its barrier)."""
        return self.bufs[(self.i - 1) % 2]


def make_trajectory_gather(batch, every, steps_per_pass, stream, prefer_fused: bool = True):
    """Fused peer-memory gather when symmetric memory is available on EVERY rank, else NCCL all-gather on a side stream.
    The rendezvous is collective, so the ranks agree on the path (MIN over a success flag) before any of them returns:
    a rank that failed locally must not wait in all_gather while the others wait in the fused barrier."""
    if prefer_fused:
        g, reason = None, ""
        try:
            g = FusedTrajectoryGather(batch, every, steps_per_pass, stream)
        except Exception as e:  # no symmetric memory / no P2P on this box
            reason = f"{type(e).__name__}: {e}"
        ok = torch.tensor([1 if g is not None else 0], dtype=torch.int32, device=torch.device("cuda", batch.device))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            how = "one NVLS multimem.st per value" if g.multicast else "one NVLink peer store per rank and value"
            return g, f"fused: step kernel stores snapshots into every rank's symmetric-memory buffer ({how})"
        if g is not None:  # some other rank failed: drop this rank's symmetric buffers and fall back with everybody
            g.bufs.clear(); g.hdls.clear()
            del g
            reason = "fused path unavailable on another rank"
        return TrajectoryGather(batch, every, steps_per_pass, stream), f"NCCL all-gather on a side stream (fused path unavailable: {reason[:120]})"
    return TrajectoryGather(batch, every, steps_per_pass, stream), "NCCL all-gather on a side stream"


__all__ = ["FusedTrajectoryGather", "make_trajectory_gather", "shard_range", "gather_trajectory", "allreduce_cost", "global_trajectory_to_instance_major", "TrajectoryGather", "world"]
