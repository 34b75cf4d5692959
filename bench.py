#!/usr/bin/env python
"""
bench.py -- reach-timesteps/sec of the RapidMuskingum routing hot path (BASELINE.json metric).

Workload (config.workload = "C4"): the GEOGLOWS-scale configuration of BASELINE.json configs[3] --
a synthetic 7M-reach forest in 5000 independent basins (SURVEY.md 8d generator, seed 4), hourly
lateral inflow volumes, dt_routing = dt_runoff = 3600 s, fp64.  One "step" routes one resident chunk
of `--rows` hourly time steps over the whole network, chained in time through the channel state
(a 1-year run is 8760/rows such steps).  The 7M-reach network fits one B200, so it is the N=1
workload; with N GPUs the basins of the same network are bin-packed over the ranks (no collective in the
time loop) and every step routes N x `--rows` time steps, so the bytes each GPU streams per step stay fixed
("scaling": "weak"; `--scaling strong` keeps the rows, and so the total work, fixed instead).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU algorithm (oracle port)

Prints ONE JSON line on rank 0.  `value` is timed with CUDA events around device-resident launches;
`e2e` is the same metric through the host-array API the router classes call (pinned fp64 lateral inflows in,
float32 discharge out as the reference hands it to its writer; H2D + D2H inside the timed region); `e2e.variants`
adds the kernel-level fp64-out call and the gridded-runoff -> discharge residency.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_ALG = 56.0  # algorithmic bytes per reach-timestep: 8 lateral + 8 discharge + 32 coefficients + 8 topology
DT = 3600


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--reaches', type=int, default=7_000_000)
    ap.add_argument('--basins', type=int, default=5000)
    ap.add_argument('--rows', type=int, default=240, help='hourly time steps resident per step')
    ap.add_argument('--e2e-rows', type=int, default=96, help='time steps per end-to-end (host array) step, per GPU')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: N x rows time steps per step on N GPUs (bytes per GPU fixed); strong: rows fixed')
    ap.add_argument('--no-e2e-variants', action='store_true')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--ref-rows', type=int, default=24, help='time steps per step of the CPU reference arm')
    ap.add_argument('--cpu-sample-reaches', type=int, default=1_000_000)
    ap.add_argument('--depth-bias', type=float, default=0.5)
    ap.add_argument('--time-tile', type=int, default=0)
    ap.add_argument('--tile-stride', type=int, default=0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--staging', default='auto', choices=['auto', 'registers', 'registers-tiled', 'tma', 'out-reach-major', 'lateral-grouped', 'direct'])
    ap.add_argument('--order', default='growth', choices=['growth', 'level'],
                    help='reach order of the synthetic params file: generator order or sorted by topological level')
    return ap.parse_args()


def network(args):
    """C4 network + Muskingum parameters, identical on every rank."""
    from river_route_b200 import synth
    down = synth.forest(args.reaches, args.basins, seed=4, depth_bias=args.depth_bias)
    k, x = synth.muskingum_params(args.reaches, 4)
    if args.order == 'level':
        down = synth.relabel(down, synth.level_sorted_order(down))
    return down, k, x


def coefficients(k, x, dt_routing, dt_runoff):
    """The reference's exact expressions (routers/Muskingum.py:174-179, TransformMuskingum.py:104)."""
    dt_div_k = dt_routing / k
    denominator = dt_div_k + (2 * (1 - x))
    _2x = 2 * x
    c1 = (dt_div_k - _2x) / denominator
    c2 = (dt_div_k + _2x) / denominator
    c3 = ((2 * (1 - x)) - dt_div_k) / denominator
    return c1, c2, c3, (c1 + c2) / dt_runoff


def shard(down, n_parts, part_id):
    """Reaches of this rank's basins, in their original relative order, and the local downstream index."""
    from river_route_b200.sharding import shard_by_basin
    return shard_by_basin(down, n_parts, part_id)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled every few
    milliseconds from a thread (the timed region is only a few hundred ms); nvidia-smi as a fallback."""
    SMI_Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread, self.proc = index, [], False, None, None
        self.max_mhz = None

    def _visible_index(self):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis:
            try:
                return int(vis.split(',')[self.index])
            except Exception:
                pass
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons = {'hw_slowdown': pynvml.nvmlClocksThrottleReasonHwSlowdown,
                       'hw_thermal_slowdown': pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                       'sw_thermal_slowdown': pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                       'sw_power_cap': pynvml.nvmlClocksThrottleReasonSwPowerCap}

            def poll():
                while not self.stop_flag:
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.samples.append((mhz, [k for k, bit in reasons.items() if mask & bit]))
                    except Exception:
                        pass
                    time.sleep(0.004)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.SMI_Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self._visible_index()), '-lms', '50'], stdout=subprocess.PIPE, text=True)

            def read():
                names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
                for line in self.proc.stdout:
                    f = [x.strip() for x in line.split(',')]
                    try:
                        self.max_mhz = float(f[2])
                        self.samples.append((float(f[1]), [nm for nm, v in zip(names, f[5:9]) if v.lower().startswith('active')]))
                    except Exception:
                        continue
            threading.Thread(target=read, daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        if self.proc:
            time.sleep(0.1)
            self.proc.terminate()
        sm = [x[0] for x in self.samples]
        reasons = sorted({r for x in self.samples for r in x[1]})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': self.max_mhz, 'reasons': reasons,
                'samples': len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic():
    """DRAM bytes per launch of the routing kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f)
    except Exception:
        return None


def cpu_baseline(down, k, x, target_reaches, rows, threads=1):
    """The oracle port of rapid_route on the host cores, on a bounded sample of the same workload:
    the first whole basins of the network (basins are contiguous index ranges) x `rows` hourly steps."""
    from oracle import oracle
    from river_route_b200 import synth
    outlets = np.flatnonzero(down < 0)
    m = int(outlets[np.searchsorted(outlets, min(target_reaches, down.shape[0]) - 1)]) + 1
    sub = down[:m]
    c1, c2, c3, c4 = coefficients(k[:m], x[:m], DT, DT)
    indptr, indices = oracle.csc_from_down(sub)
    lhs = oracle.lhs_off_data(c1, indices)
    ql = synth.lateral_volumes(rows, m, 99)
    out = np.zeros((rows, m))
    q = np.zeros(m)
    oracle.rapid_route(indptr, indices, lhs, c2, c3, c4, q, ql[:2], out[:2], 1)   # warm caches / page in
    t = time.perf_counter()
    oracle.rapid_route(indptr, indices, lhs, c2, c3, c4, q, ql, out, 1)
    dt = time.perf_counter() - t
    return {'value': m * rows / dt, 'unit': 'reach-timesteps/s', 'cores': threads, 'kind': 'port',
            'sample': f'first {m} reaches (whole basins) x {rows} hourly steps, oracle/rr_oracle.c (-O3, scalar), '
                      f'{dt:.2f} s; host has {os.cpu_count()} cores'}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm with every host core, parallel over independent basins as
    the reference documents (docs/references/parallelism.md:67-112).  The reference is Python+numba and cannot
    travel to the GPU box, so this is the oracle port (kind = "port")."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import river_route_b200 as rr
    from river_route_b200 import synth
    from oracle import oracle
    down, k, x = network(args)
    cores = os.cpu_count() or 1
    _, _, part = rr.label_basins(down, cores)
    shards = []
    rows = args.ref_rows
    for c in range(cores):
        idx = np.flatnonzero(part == c)
        new_of_old = np.full(down.shape[0], -1, dtype=np.int64)
        new_of_old[idx] = np.arange(idx.shape[0])
        d = down[idx]
        local = np.where(d >= 0, new_of_old[np.where(d >= 0, d, 0)], -1)
        c1, c2, c3, c4 = coefficients(k[idx], x[idx], DT, DT)
        indptr, indices = oracle.csc_from_down(local)
        shards.append(dict(indptr=indptr, indices=indices, lhs_off=oracle.lhs_off_data(c1, indices), c2=c2, c3=c3,
                           c4_dt=c4, q_t=np.zeros(idx.shape[0]), ql=synth.lateral_volumes(rows, idx.shape[0], 100 + c),
                           out=np.zeros((rows, idx.shape[0]))))
    for _ in range(args.warmup):
        oracle.rapid_route_sharded(shards, 1, cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        oracle.rapid_route_sharded(shards, 1, cores)
    dt = time.perf_counter() - t
    value = args.reaches * rows * args.steps / dt
    sample = (f'{args.reaches} reaches x {rows} hourly steps per step, basins bin-packed over {cores} threads, '
              f'oracle/rr_oracle.c rapid_route')
    line = {
        'impl': 'reference', 'metric': 'reach-timesteps/sec (RapidMuskingum, fp64)', 'value': value,
        'unit': 'reach-timesteps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, rows),
        'cpu_baseline': {'value': value, 'unit': 'reach-timesteps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'reach-timesteps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, rows):
    return {'workload': 'C4', 'router': 'RapidMuskingum', 'reaches': args.reaches, 'basins': args.basins,
            'scaling_rule': (f'weak: rows_per_step = {args.rows} x n_gpus on the same {args.reaches}-reach network'
                             if args.scaling == 'weak' else 'strong: rows_per_step fixed'),
            'network': f'synthetic forest seed 4 depth_bias {args.depth_bias} (SURVEY.md 8d), reach order: {args.order}',
            'rows_per_step': rows, 'dt_runoff_s': DT, 'dt_routing_s': DT, 'substeps': 1,
            'parallelism': f'basin-sharded x{args.gpus}, no collective in the time loop',
            'l2': 'inputs larger than L2: every step streams rows x reaches x 16 B'}


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import river_route_b200 as rr
    from river_route_b200 import synth  # noqa: F401

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    if not rr.cuda_available():
        raise SystemExit('bench.py needs a CUDA device: river_route_b200 has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    numa_bound = False
    if world > 1:                                   # each rank streams over its own PCIe link from its own NUMA node
        from river_route_b200.sharding import bind_to_gpu_numa
        numa_bound = bind_to_gpu_numa(local_rank)
    dist = None
    real_stdout = None
    if world > 1:
        # stdout carries exactly one JSON line: anything libraries print there meanwhile (NCCL's version banner at
        # communicator creation) goes to stderr instead
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    # ---- network, shard, plan ----
    down, k, x = network(args)
    idx, local_down = shard(down, world, rank)
    n = int(idx.shape[0])
    plan = rr.Plan(local_down, time_tile=args.time_tile, tile_stride=args.tile_stride, device=local_rank,
                   staging=args.staging)
    c1, c2, c3, c4 = coefficients(k[idx], x[idx], DT, DT)
    plan.set_coefficients(c1, c2, c3, c4)
    info = plan.info

    # ---- device-resident inputs (synthetic lateral volumes: gamma(0.3, 5e4 m3), half zeros) ----
    rows = args.rows * (world if args.scaling == 'weak' else 1)
    ld = ((n + 31) // 32) * 32
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    d_lat = torch.empty((rows, ld), dtype=torch.float64, device=dev)
    conc = torch.full((ld,), 0.3, dtype=torch.float64, device=dev)
    for t in range(rows):
        g = torch._standard_gamma(conc, generator=gen) * 5.0e4
        g[torch.rand(ld, device=dev, generator=gen) < 0.5] = 0.0
        d_lat[t] = g
    d_out = torch.empty((rows, ld), dtype=torch.float64, device=dev)
    d_q = torch.zeros(n, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        plan.route_dev(rr.MODE_RAPID, d_q.data_ptr(), d_lat.data_ptr(), ld, d_out.data_ptr(), ld, rows, 1, stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    rr.launch_count(reset=True)
    from river_route_b200.plan import timing_enable, timing_read
    timing_enable(True)
    timing_read(reset=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for s in range(args.steps):
        step()
        ev[s + 1].record()
    barrier()
    launches = rr.launch_count()
    ktimes = timing_read(reset=True)
    prof = plan.read_profile()
    timing_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(args.steps)]
    t_max = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    total_ms_max = float(t_max.item())
    value = args.reaches * rows * args.steps / (total_ms_max * 1e-3)

    # ---- summary outputs gathered over NVLink (outside the time loop): outlet discharge of the last step ----
    outlets = torch.from_numpy(np.flatnonzero(local_down < 0)).to(dev)
    summary = torch.stack([d_out[rows - 1, outlets].sum(), d_q.sum(), torch.tensor(float(n), device=dev, dtype=torch.float64)])
    if dist is not None:
        parts = [torch.empty_like(summary) for _ in range(world)]
        dist.all_gather(parts, summary)
        summary_all = torch.stack(parts).cpu().numpy()
    else:
        summary_all = summary.cpu().numpy()[None]
    finite = bool(torch.isfinite(d_out[:, :n]).all().item())

    # ---- end to end through the host-array API: pinned H2D + route + D2H every step ----
    er = min(args.e2e_rows * (world if args.scaling == 'weak' else 1), rows)
    h_lat = rr.pinned_empty((er, n))
    h_lat[:] = d_lat[:er, :n].cpu().numpy()
    h_q = rr.pinned_empty((n,))

    def timed_host(call):
        """max over ranks of the wall time of `e2e_steps` host-array calls (state chained between them)."""
        h_q[:] = 0.0
        call()                                                          # warm-up (allocates the staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            call()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return args.reaches * er * args.e2e_steps / float(t.item())

    # headline: what RapidMuskingum.route() does between reading a qlateral file and handing the float32 array to
    # the writer (routers.TransformMuskingum._route_lateral -> rr_route_host_ex)
    h_out32 = rr.pinned_empty((er, n), dtype=np.float32)
    e2e_value = timed_host(lambda: plan.route_host(rr.MODE_RAPID, h_q, h_lat, h_out32, 1))
    # the host path must agree with the device path on the same inputs (zero state)
    d_chk_q = torch.zeros(n, dtype=torch.float64, device=dev)
    d_chk = torch.empty((er, ld), dtype=torch.float64, device=dev)
    h_q[:] = 0.0
    plan.route_host(rr.MODE_RAPID, h_q, h_lat, h_out32, 1)
    plan.route_dev(rr.MODE_RAPID, d_chk_q.data_ptr(), d_lat.data_ptr(), ld, d_chk.data_ptr(), ld, er, 1, stream)
    torch.cuda.synchronize()
    host_equals_dev = bool(np.array_equal(d_chk[:, :n].cpu().numpy().astype(np.float32), h_out32))
    del d_chk
    variants = {}
    if not args.no_e2e_variants and world == 1:                        # extra pinned host memory: one GPU only
        # (a) the reference kernel's own contract: fp64 discharge array back to the host (rapid_route drop-in)
        h_out64 = rr.pinned_empty((er, n))
        variants['kernel_level_f64_out'] = {
            'value': timed_host(lambda: plan.route_host(rr.MODE_RAPID, h_q, h_lat, h_out64, 1)),
            'h2d_bytes_per_step': int(n * er * 8 + n * 8), 'd2h_bytes_per_step': int(n * er * 8 + n * 8),
            'api': 'Plan.route_host -> rr_route_host (kernels.rapid_route signature)'}
        del h_out64
        # (b) gridded runoff in (float32, ERA5 0.25 degree grid), float32 discharge out: weight table -> route in
        #     one device residency (routers.TransformMuskingum._route_runoff -> rr_runoff_route_host)
        from river_route_b200.transforms import Transform
        n_cells = 721 * 1440
        rng = np.random.default_rng(77 + rank)
        per = rng.integers(4, 9, n)
        indptr = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(per, out=indptr[1:])
        first = rng.integers(0, n_cells - 8, n)
        indices = (np.repeat(first, per) + (np.arange(indptr[-1]) - np.repeat(indptr[:-1], per))).astype(np.int32)
        w = rng.random(indptr[-1])
        w /= np.repeat(np.add.reduceat(w, indptr[:-1]), per)
        tf = Transform(indptr, indices, w, n_cells, area=rng.uniform(1e5, 5e8, n), device=local_rank)
        h_grid = rr.pinned_empty((er, n_cells), dtype=np.float32)
        h_grid[:] = (rng.gamma(0.3, 2e-3, (er, n_cells)) * (rng.random((er, n_cells)) < 0.4)).astype(np.float32)
        variants['grid_runoff_to_discharge_f32'] = {
            'value': timed_host(lambda: plan.runoff_route_host(tf, rr.MODE_RAPID, h_q, h_grid, h_out32, 1, as_volumes=True)),
            'h2d_bytes_per_step': int(n_cells * er * 4 + n * 8), 'd2h_bytes_per_step': int(n * er * 4 + n * 8),
            'api': 'Plan.runoff_route_host -> rr_runoff_route_host (weights SpMM + route + float32 cast on the device)',
            'weight_table': f'{int(indptr[-1])} entries, 4-8 cells per river, {n_cells} grid cells'}
        # (c) the same, copying back only the basin outlets (rr_plan_set_output_subset: the device-side form of the
        #     reference's subset writer, docs/tutorial/advanced.md:147-170): neither PCIe direction is the bound any more
        outlets_idx = np.flatnonzero(local_down < 0).astype(np.int32)
        plan.set_output_subset(outlets_idx)
        h_out_sub = rr.pinned_empty((er, outlets_idx.shape[0]), dtype=np.float32)
        variants['grid_runoff_to_outlet_discharge_f32'] = {
            'value': timed_host(lambda: plan.runoff_route_host(tf, rr.MODE_RAPID, h_q, h_grid, h_out_sub, 1, as_volumes=True)),
            'h2d_bytes_per_step': int(n_cells * er * 4 + n * 8),
            'd2h_bytes_per_step': int(outlets_idx.shape[0] * er * 4 + n * 8), 'outlets': int(outlets_idx.shape[0]),
            'api': 'Plan.set_output_subset + Plan.runoff_route_host (all reaches routed, outlet columns copied back)'}
        plan.set_output_subset(None)
        tf.close()

    if rank == 0:
        peak, peak_src = measured_peak()
        kernel_ms = ktimes['route']['ms'] / max(ktimes['route']['launches'], 1)   # the routing kernel alone (CUDA events)
        achieved = B_ALG * n * rows / (kernel_ms * 1e-3) / 1e9
        traffic = ncu_traffic()
        if traffic and args.staging != 'registers-tiled':
            # the committed capture is of the ring-exchange kernel; the default (direct exchange) moves fewer bytes
            traffic = {'source': 'none for this kernel variant: ' + str(traffic.get('applies_to', traffic.get('source')))}
        line = {
            'metric': 'reach-timesteps/sec (RapidMuskingum, fp64)', 'value': value, 'unit': 'reach-timesteps/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms_max / args.steps,
            'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': dict(workload_config(args, rows), reaches_rank0=n, plan={k_: int(v) for k_, v in info.items()}),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': ((traffic or {}).get('dram_bytes_per_reach_step') or 0) * n * rows or None,
                         'kernel': 'rr_wavefront_kernel<RAPID>', 'kernel_ms': kernel_ms,
                         'step_ms_by_kernel': {k_: v['ms'] / args.steps for k_, v in ktimes.items()},
                         'algorithmic_bytes_per_reach_step': B_ALG, 'peak_source': peak_src,
                         'traffic_source': (traffic or {}).get('source')},
            'e2e': {'value': e2e_value, 'unit': 'reach-timesteps/s', 'h2d_bytes_per_step': int(n * er * 8 + n * 8),
                    'd2h_bytes_per_step': int(n * er * 4 + n * 8), 'rows_per_step': er, 'steps': args.e2e_steps,
                    'api': 'Plan.route_host with a float32 output array -> rr_route_host_ex: pinned fp64 lateral inflows '
                           'in, route, float32 cast on the device (TransformMuskingum.py:146), float32 discharge out; '
                           'chunked cudaMemcpyAsync on three streams',
                    'host_equals_device_path': host_equals_dev, 'numa_bound': numa_bound, 'variants': variants},
            'gpu_launches': int(launches), 'kernel_phase_cycles': prof,
            'clocks': clocks,
            'checks': {'finite': finite, 'summary_per_rank[outlet_q_last_step, state_sum, reaches]': summary_all.tolist()},
        }
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(down, k, x, args.cpu_sample_reaches, rows)
        if real_stdout is not None:
            sys.stdout.flush()
            os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
