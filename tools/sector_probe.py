#!/usr/bin/env python
"""What HBM delivers for the routing kernel's access shapes (rr_probe_sector_bandwidth): coalesced 256-bit copy vs
one 640-byte series per lane (rows in order / scattered).  One JSON line; the ceiling the wavefront kernel's
4.3-4.9 TB/s should be read against."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

LIB = os.path.join(ROOT, 'tools', 'librr_probe.so')   # make -C river_route_b200/csrc probe
lib = C.CDLL(LIB)
lib.rr_probe_sector_bandwidth.restype = C.c_int
lib.rr_probe_sector_bandwidth.argtypes = [C.c_int, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                          C.POINTER(C.c_double)]
lib.rr_probe_last_error.restype = C.c_char_p


def check(rc):
    if rc:
        raise RuntimeError(lib.rr_probe_last_error().decode())

dev = torch.device('cuda:0')
res = {'probe': 'sector bandwidth, read + write bytes / kernel time, best of 5'}
for row_doubles, label in ((80, 'exchange row 640 B'), (64, 'lateral tile row 512 B')):
    rows = (6 << 30) // (row_doubles * 8) // 4096 * 4096   # 6 GiB per array: far larger than L2
    src = torch.rand(rows * row_doubles, dtype=torch.float64, device=dev)
    dst = torch.empty_like(src)
    perm = torch.randperm(rows, device=dev, dtype=torch.int32)
    local = (torch.arange(rows, device=dev, dtype=torch.int64) // 4096 * 4096
             + torch.randperm(4096, device=dev).repeat(rows // 4096 + 1)[:rows]).clamp_(max=rows - 1).to(torch.int32)
    for mode, name, p in ((0, 'coalesced', perm), (1, 'series_in_order', perm), (2, 'series_scattered', perm),
                          (3, 'series_in_order_L2_128B_hint', perm), (4, 'series_scattered_L2_128B_hint', perm),
                          (5, 'coalesced_reads_series_writes', perm), (6, 'series_reads_coalesced_writes', perm),
                          (2, 'series_scattered_within_4096_rows', local)):
        ms = C.c_double(0)
        check(lib.rr_probe_sector_bandwidth(mode, rows, row_doubles, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                                            C.c_void_p(p.data_ptr()), 5, C.byref(ms)))
        res[f'{label}: {name}'] = {'ms': ms.value, 'TBps': 2 * rows * row_doubles * 8 / (ms.value * 1e-3) / 1e12}
    if mode == 2:
        ok = bool(torch.equal(dst.view(rows, row_doubles)[:1000], src.view(rows, row_doubles)[local[:1000].long()]))
        res[f'{label}: check'] = ok
    del src, dst, perm, local
print(json.dumps(res))
