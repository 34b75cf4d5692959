#!/usr/bin/env python
"""
Concurrent H2D + D2H ceiling of the box at N = 1, 2, 4, 8 ranks (one process per GPU, torchrun), with pinned host
buffers: the bound the end-to-end (`e2e`) bench numbers should be read against.  Every rank copies `--mb` MB in each
direction, H2D and D2H on separate streams at the same time, `--reps` times; the ranks start together (barrier) and
rank 0 prints one JSON line with per-rank and aggregate GB/s, H2D-only, D2H-only and both at once, plus which NUMA
node / CPUs each rank runs on.

    python tools/pcie_ceiling_probe.py                                   # N = 1
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_ceiling_probe.py
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mb', type=int, default=2048)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--bind', action='store_true', help='pin each rank to the CPUs NVML reports local to its GPU')
    args = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    bound = False
    if args.bind:
        from river_route_b200.sharding import bind_to_gpu_numa
        bound = bind_to_gpu_numa(local)
    nbytes = args.mb << 20
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.ones(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def timed(h2d, d2h):
        barrier()
        t = time.perf_counter()
        for _ in range(args.reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        x = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
        return float(x.item())
    timed(True, True)
    res = {}
    for name, (a, b) in {'h2d_only': (True, False), 'd2h_only': (False, True), 'both': (True, True)}.items():
        dt = timed(a, b)
        per_dir = nbytes * args.reps / dt / 1e9
        res[name] = {'seconds_max_over_ranks': dt, 'GBps_per_rank_per_direction': per_dir,
                     'GBps_aggregate_per_direction': per_dir * world}
    cpus = sorted(os.sched_getaffinity(0))
    info = torch.tensor([float(len(cpus)), float(cpus[0]), float(cpus[-1])], dtype=torch.float64, device=dev)
    infos = [info]
    if dist is not None:
        infos = [torch.empty_like(info) for _ in range(world)]
        dist.all_gather(infos, info)
    if rank == 0:
        print(json.dumps({'probe': 'pinned H2D / D2H ceiling, all ranks at once', 'n_gpus': world, 'mb_per_direction': args.mb,
                          'reps': args.reps, 'numa_bound': bound, 'host_cpus': os.cpu_count(),
                          'rank_cpu_affinity[count, first, last]': [i.cpu().tolist() for i in infos], **res,
                          'e2e_ceiling_reach_steps_per_s': {
                              'f64_in_f32_out': res['both']['GBps_aggregate_per_direction'] * 1e9 / 8.0,
                              'f32_in_f32_out': res['both']['GBps_aggregate_per_direction'] * 1e9 / 4.0,
                              'note': 'H2D-bound: bytes per reach-timestep entering the GPU (8 or 4) at the both-directions rate'}}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
