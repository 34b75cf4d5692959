#!/usr/bin/env python
"""Per-group event timeline of the direct wavefront kernel on a small deep network (C1: 50k reaches, depth 189).

Needs the instrumented build (make -C river_route_b200/csrc trace -> tools/librr_trace.so): lane 0 of every warp stamps
{globaltimer, clock64} at six events per (block, 16-row group):
  0 item start (ticket decoded)        5 before the wait for this group's dependencies
  1 dependencies seen                  2 first row of the group computed (upstream data arrived)
  3 last row computed, stores issued   4 after the release of the group's flag
Writes gpurun_out/trace_<net>_T<rows>.npz and prints medians of the hops along the level and the time direction."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['RR_B200_LIB'] = os.path.join(ROOT, 'tools', 'librr_trace.so')
import torch  # noqa: E402
import river_route_b200 as rr  # noqa: E402
from river_route_b200 import synth, _lib  # noqa: E402
from tests.helpers import network_arrays  # noqa: E402

NEV = 6
dev = torch.device('cuda:0')
which = sys.argv[1] if len(sys.argv) > 1 else 'c1'
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
K = int(sys.argv[3]) if len(sys.argv) > 3 else 1
TIME_TILE = int(sys.argv[4]) if len(sys.argv) > 4 else 0
STRIDE = int(sys.argv[5]) if len(sys.argv) > 5 else 0
SAVE = (sys.argv[6] == 'save') if len(sys.argv) > 6 else False
if which == 'c1':
    n, down = 50_000, synth.forest(50_000, 1, seed=0, depth_bias=0.9)
else:
    n, down = 500_000, synth.forest(500_000, 2, seed=1, depth_bias=0.5, main_stem=3000)
k, x = synth.muskingum_params(n, 0)
a = network_arrays(down, k, x, 3600 // K, 3600)
d_lat = torch.from_numpy(synth.lateral_volumes(64, n, 0)).to(dev).repeat((T + 63) // 64, 1)[:T].contiguous()
d_out = torch.empty((T, n), dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
plan = rr.Plan(down, staging='direct-nohw' if K > 1 else 'auto', time_tile=TIME_TILE, tile_stride=STRIDE)
plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
arr = plan.arrays()
nb = plan.info['n_blocks']
rows = plan.tile_rows(T, K)
gpt = (rows * K + 15) // 16
n_tiles = (T + rows - 1) // rows
G = n_tiles * gpt
lib = _lib.lib
lib.rr_trace_set.argtypes = [C.c_void_p]
lib.rr_trace_set.restype = None
trace = torch.zeros((nb, G, NEV, 2), dtype=torch.int64, device=dev)
for rep in range(3):                                    # the last (warm) run is the one kept
    trace.zero_()
    d_q = torch.zeros(n, dtype=torch.float64, device=dev)
    lib.rr_trace_set(C.c_void_p(trace.data_ptr()))
    plan.route_dev(rr.MODE_RAPID, d_q.data_ptr(), d_lat.data_ptr(), n, d_out.data_ptr(), n, T, K, stream)
    torch.cuda.synchronize()
tr = trace.cpu().numpy()
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
if SAVE:
    np.savez_compressed(os.path.join(ROOT, 'gpurun_out', f'trace_{which}_T{T}_K{K}_tt{TIME_TILE}_s{STRIDE}.npz'), trace=tr,
                        blk_level=arr['blk_level'], dep_ptr=arr['dep_ptr'], dep_idx=arr['dep_idx'], rows=rows, gpt=gpt, K=K)

gt, ck = tr[..., 0].astype(np.float64), tr[..., 1].astype(np.float64)
lvl = arr['blk_level']
t0 = gt[gt > 0].min()
gt = np.where(gt > 0, gt - t0, np.nan) / 1e3            # us
d = np.diff(np.unique(tr[..., 0][tr[..., 0] > 0]))
res = {'network': which, 'T': T, 'K': K, 'time_tile': TIME_TILE, 'tile_stride': STRIDE, 'blocks': int(nb), 'levels': int(lvl.max()) + 1, 'tile_rows': int(rows), 'gpt': int(gpt),
       'groups': int(G), 'globaltimer_step_ns': float(d[d > 0].min()), 'span_us': float(np.nanmax(gt))}


def med(v):
    v = np.asarray(v, dtype=np.float64)
    v = v[np.isfinite(v)]
    return round(float(np.median(v)), 3) if v.size else None


# the deepest dependency chain: from the deepest block follow the upstream block of the highest level
chain = [int(np.argmax(lvl))]
while True:
    b = chain[-1]
    deps = arr['dep_idx'][arr['dep_ptr'][b]:arr['dep_ptr'][b + 1]]
    if deps.size == 0:
        break
    chain.append(int(deps[np.argmax(lvl[deps])]))
chain = chain[::-1]                                       # level 0 first
c = np.array(chain)
done = gt[c][:, :, 3]                                     # [level on chain][group] stores issued
seen = gt[c][:, :, 1]
first = gt[c][:, :, 2]
rel = gt[c][:, :, 4]
pre = gt[c][:, :, 5]
res['chain_len'] = len(chain)
res['hop_level_us: upstream stores issued -> dependencies seen'] = med(seen[1:] - done[:-1])
res['hop_level_us: upstream release returned -> dependencies seen'] = med(seen[1:] - rel[:-1])
res['seen -> first row (data load + 1 row)'] = med(first - seen)
res['first row -> stores issued (15 rows + stores)'] = med(done - first)
res['stores issued -> release returned'] = med(rel - done)
res['release returned -> next group wait begins'] = med(pre[:, 1:] - rel[:, :-1])
res['wait for dependencies (pre -> seen)'] = med(seen - pre)
sp = ck[c][:, :, 0]
ingrp = (np.arange(G) % gpt) != 0
res['polls per group inside a tile, chain (median, p90)'] = [med(sp[:, ingrp]), round(float(np.nanpercentile(sp[:, ingrp], 90)), 1)]
res['us per poll'] = med(((seen - pre)[:, ingrp] / np.maximum(sp[:, ingrp], 1))[sp[:, ingrp] > 0])
res['per level: done(l,g) - done(l-1,g)'] = med(done[1:] - done[:-1])
res['per group: done(l,g) - done(l,g-1), chain'] = med(done[:, 1:] - done[:, :-1])
in_tile = (np.arange(1, G) % gpt) != 0
res['per group inside a tile, chain'] = med((done[:, 1:] - done[:, :-1])[:, in_tile])
res['per group across a tile boundary, chain'] = med((done[:, 1:] - done[:, :-1])[:, ~in_tile]) if (~in_tile).any() else None
l0 = np.flatnonzero(lvl == 0)
d0 = gt[l0][:, :, 3]
res['per group: level-0 blocks'] = med(d0[:, 1:] - d0[:, :-1])
res['per group inside a tile: level-0 blocks'] = med((d0[:, 1:] - d0[:, :-1])[:, in_tile])
res['level-0 item start spread of tile 0 (p5, p50, p95 us)'] = [round(float(q), 2) for q in np.nanpercentile(gt[l0][:, 0, 0], [5, 50, 95])]
# clock-based durations inside one warp (cycles): same item only
ckc = ck[c]
res['cycles seen -> first row'] = med(ckc[:, :, 2] - ckc[:, :, 1])
res['cycles first row -> stores issued'] = med(ckc[:, :, 3] - ckc[:, :, 2])
res['cycles stores issued -> release returned'] = med((ckc[:, :, 4] - ckc[:, :, 3])[:, (np.arange(G) % gpt) != gpt - 1])
# finish time of group g at the chain's end vs the start: slope per group and per level
res['done(last level, g) us'] = [round(float(v), 1) for v in done[-1, :: max(1, G // 8)]]
res['done(l, group 0) us'] = [round(float(v), 1) for v in done[:: max(1, len(chain) // 8), 0]]
res['done(level 0, g) us'] = [round(float(v), 1) for v in done[0, :: max(1, G // 8)]]
print(json.dumps(res), flush=True)
