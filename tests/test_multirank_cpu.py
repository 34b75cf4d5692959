"""
The N > 1 path on the CPU: two gloo ranks shard a network by drainage basin, each "routes" its shard with the
CPU emulation of the device schedule, and the gathered result must equal the single-process oracle bit for bit
(no collective is needed inside the time loop; the only exchange is the gather of outputs at the end).
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import river_route_b200 as rr
from river_route_b200 import synth
from river_route_b200.sharding import shard_by_basin
from oracle import oracle
from tests.emulator import emulate
from tests.helpers import network_arrays

N, T, K, WORLD = 900, 6, 2, 2


def _inputs():
    down = synth.forest(N, 7, seed=21, depth_bias=0.6)
    k, x = synth.muskingum_params(N, 21)
    a = network_arrays(down, k, x, 1800, 3600)
    ql = synth.lateral_volumes(T, N, 21)
    q0 = np.random.default_rng(21).uniform(0, 30, N)
    return down, a, ql, q0


def _worker(rank, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=WORLD)
    try:
        down, a, ql, q0 = _inputs()
        idx, local = shard_by_basin(down, WORLD, rank)
        plan = rr.Plan(local, renumber='always' if rank else 'never')
        out, q, _ = emulate(plan, rr.MODE_RAPID, a['c1'][idx], a['c2'][idx], a['c3'][idx], a['c4_dt'][idx], q0[idx],
                            np.ascontiguousarray(ql[:, idx]), T, K, 8, 1)
        # gather (index, outputs) on every rank -- the only communication of the whole run
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(WORLD)]
        dist.all_gather(sizes, torch.tensor([idx.shape[0]]))
        m = int(max(s.item() for s in sizes))
        pad = lambda t: torch.nn.functional.pad(t, (0, m - t.shape[-1]))
        g_idx = [torch.zeros(m, dtype=torch.int64) for _ in range(WORLD)]
        g_out = [torch.zeros((T, m), dtype=torch.float64) for _ in range(WORLD)]
        g_q = [torch.zeros(m, dtype=torch.float64) for _ in range(WORLD)]
        dist.all_gather(g_idx, pad(torch.from_numpy(idx)))
        dist.all_gather(g_out, pad(torch.from_numpy(out)))
        dist.all_gather(g_q, pad(torch.from_numpy(q)))
        full_out, full_q = np.full((T, N), np.nan), np.full(N, np.nan)
        for r in range(WORLD):
            n_r = int(sizes[r].item())
            ii = g_idx[r][:n_r].numpy()
            full_out[:, ii] = g_out[r][:, :n_r].numpy()
            full_q[ii] = g_q[r][:n_r].numpy()
        q_ref, ref = q0.copy(), np.zeros((T, N))
        oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q_ref, ql, ref, K)
        ret[rank] = bool(np.array_equal(full_out, ref) and np.array_equal(full_q, q_ref))
    finally:
        dist.destroy_process_group()


def test_two_rank_basin_sharding_matches_single_process_oracle():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(port, ret), nprocs=WORLD, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shards_partition_the_network_without_cutting_basins():
    down, *_ = _inputs()
    basin, nb, _ = rr.label_basins(down)
    seen = np.zeros(N, dtype=int)
    for parts in (2, 4, 8):
        seen[:] = 0
        sizes = []
        for r in range(parts):
            idx, local = shard_by_basin(down, parts, r)
            seen[idx] += 1
            sizes.append(idx.shape[0])
            assert np.all(np.diff(idx) > 0)                                  # original relative order kept
            assert np.all(local[local >= 0] > np.flatnonzero(local >= 0))     # still topologically sorted
            assert int((local < 0).sum()) == len(set(basin[idx]))             # whole basins only
        assert np.all(seen == 1)
        assert max(sizes) - min(sizes) <= np.bincount(basin).max()


def test_numa_binding_is_best_effort():
    """Without NVML / a GPU the helper declines instead of raising, and never widens the affinity mask."""
    from river_route_b200.sharding import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa(0) in (True, False)
    assert os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)
