"""CPU-only, world_size 2, gloo: the host-side multi-GPU logic (contiguous sharding, trajectory gather ordering,
rollout-cost all-reduce) without a GPU.  The per-rank "device" work is stood in by the CPU oracle."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cdpr_simulation_b200 import workloads as wl
    from cdpr_simulation_b200 import distributed as D
    from oracle import binding as ob
    cfg = ob.default_config(4)
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n_total, seed=4)
    lo, hi = D.shard_range(n_total, rank, world)
    b = ob.Batch(cfg, hi - lo, pose7[lo:hi], twist6[lo:hi], amp[lo:hi], freq[lo:hi], phase[lo:hi])
    snaps = []
    for _ in range(3):
        b.step(20)
        pose, twist = b.platform_state()
        # device snapshot layout [13][n]: px py pz qw qx qy qz v w
        snaps.append(np.concatenate([pose[:, :3], pose[:, 6:7], pose[:, 3:6], twist], axis=1).T)
    local = torch.from_numpy(np.ascontiguousarray(np.stack(snaps)))
    gathered = D.gather_trajectory(local)
    traj = D.global_trajectory_to_instance_major(gathered)
    cost = torch.from_numpy(np.arange(5, dtype=np.float64) * (rank + 1))
    D.allreduce_cost(cost)
    if rank == 0:
        np.save(os.path.join(out_dir, "traj.npy"), traj.numpy())
        np.save(os.path.join(out_dir, "cost.npy"), cost.numpy())
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    from cdpr_simulation_b200 import workloads as wl
    from oracle import binding as ob
    n_total, world = 64, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    traj = np.load(tmp_path / "traj.npy")
    cost = np.load(tmp_path / "cost.npy")
    cfg = ob.default_config(4)
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n_total, seed=4)
    b = ob.Batch(cfg, n_total, pose7, twist6, amp, freq, phase)
    for s in range(3):
        b.step(20)
        pose, twist = b.platform_state()
        ref = np.concatenate([pose[:, :3], pose[:, 6:7], pose[:, 3:6], twist], axis=1).T
        assert np.array_equal(traj[s], ref)            # sharding is invisible: bitwise equal, global instance order
    assert np.array_equal(cost, np.arange(5) * 3.0)


def test_shard_ranges_partition_the_instances():
    from cdpr_simulation_b200.workloads import shard_range
    for n, w in [(10, 3), (8, 8), (1 << 23, 8), (7, 2), (5, 8)]:
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
