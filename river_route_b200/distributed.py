"""
Basin-sharded multi-GPU runs of the router classes: one process per GPU, launched with torchrun.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m river_route_b200.distributed \\
        RapidMuskingum config.yaml [--dt_routing 900 ...]

Reaches only connect inside a drainage basin (one downstream per reach, outlets have downstream_river_id < 0:
river_route/tools.py:98-99) and the reference names whole watersheds as its unit of parallelism
(docs/references/parallelism.md:67-75).  Every rank reads the same config, params file and input files, keeps the
river segments of the basins LPT-packed to it (``shard_by_basin``: original relative order, so every confluence still
sums its inflows in the reference's order and results are bit-identical to a single-GPU run), builds its own plan and
routes its shard on its own GPU.  There is no collective in the time loop.  ``torch.distributed`` (NCCL over NVLink on
GPUs, gloo in the CPU tests) only gathers the discharge columns and the final state to rank 0, which puts them back in
params-file order and writes ONE discharge file per input file and ONE state file, exactly what a single-GPU run writes.
"""
from __future__ import annotations

import os
import sys

import numpy as np

__all__ = ['Shard', 'main']


class Shard:
    """This rank's share of a basin-sharded run and the gathers that put results back in params-file order."""

    def __init__(self, rank: int, world: int, group=None, device=None, gather_rows: int = 64):
        if world < 1 or not (0 <= rank < world):
            raise ValueError('rank must be in [0, world)')
        self.rank, self.world, self.group = int(rank), int(world), group
        self.device = device                 # torch device the collectives run on (cuda:<local rank> for NCCL, None: CPU)
        self.gather_rows = int(gather_rows)
        self.idx = None                      # params-file indices of this rank's segments
        self._all_idx = None

    @classmethod
    def from_env(cls):
        """Inside torchrun: RANK / WORLD_SIZE / LOCAL_RANK; initialises the default process group if needed."""
        import torch
        import torch.distributed as dist
        rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
        local = int(os.environ.get('LOCAL_RANK', '0'))
        device = None
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            device = torch.device('cuda', local)
        if world > 1 and not dist.is_initialized():
            if device is not None:
                dist.init_process_group('nccl', device_id=device)
            else:
                dist.init_process_group('gloo')
        return cls(rank, world, device=device)

    @property
    def is_root(self) -> bool:
        return self.rank == 0

    @property
    def device_index(self) -> int:
        return -1 if self.device is None else int(self.device.index)

    def split(self, down: np.ndarray):
        """(idx, local_down) of this rank; remembers every rank's index set for the gathers."""
        from .plan import label_basins
        down = np.ascontiguousarray(down, dtype=np.int32)
        self.n_full = int(down.shape[0])
        if self.world == 1:
            self.idx = np.arange(self.n_full)
            self._all_idx = [self.idx]
            return self.idx, down
        _, _, part = label_basins(down, self.world)
        self._all_idx = [np.flatnonzero(part == r) for r in range(self.world)]
        self.idx = self._all_idx[self.rank]
        new_of_old = np.full(self.n_full, -1, dtype=np.int64)
        new_of_old[self.idx] = np.arange(self.idx.shape[0])
        d = down[self.idx]
        local = np.where(d >= 0, new_of_old[np.where(d >= 0, d, 0)], -1).astype(np.int32)
        if np.any((d >= 0) & (local < 0)):
            raise RuntimeError('a basin was cut across ranks')
        return self.idx, local

    def take(self, full: np.ndarray, axis: int = -1) -> np.ndarray:
        """This rank's segments of an array laid out over all segments along ``axis``."""
        if self.world == 1:
            return full
        return np.ascontiguousarray(np.take(full, self.idx, axis=axis))

    # ---- gathers: the only communication of a run, after the time loop of each file ----
    def gather_columns(self, local: np.ndarray):
        """(T, n_local) per rank -> (T, n_full) in params-file order on rank 0 (None elsewhere)."""
        if self.world == 1:
            return local
        import torch
        import torch.distributed as dist
        T = int(local.shape[0])
        sizes = [int(i.shape[0]) for i in self._all_idx]
        n_max = max(sizes)
        dt = torch.from_numpy(np.empty(0, dtype=local.dtype)).dtype
        full = np.empty((T, self.n_full), dtype=local.dtype) if self.is_root else None
        cols = [torch.from_numpy(i).to(self.device) if self.device is not None else torch.from_numpy(i)
                for i in self._all_idx] if self.is_root else None
        for t0 in range(0, T, self.gather_rows):
            t1 = min(T, t0 + self.gather_rows)
            mine = torch.zeros((t1 - t0, n_max), dtype=dt, device=self.device)
            mine[:, :sizes[self.rank]] = torch.from_numpy(np.ascontiguousarray(local[t0:t1])).to(mine.device)
            parts = [torch.empty_like(mine) for _ in range(self.world)] if self.is_root else None
            dist.gather(mine, parts, dst=0, group=self.group)
            if self.is_root:
                chunk = torch.empty((t1 - t0, self.n_full), dtype=dt, device=self.device)
                for r in range(self.world):
                    chunk[:, cols[r]] = parts[r][:, :sizes[r]]
                full[t0:t1] = chunk.cpu().numpy()
        return full

    def gather_vector(self, local: np.ndarray):
        out = self.gather_columns(np.ascontiguousarray(local)[np.newaxis, :])
        return None if out is None else out[0]

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)


def main(argv=None):
    """``python -m river_route_b200.distributed <Router> <config.yaml|json> [--key value ...]`` under torchrun."""
    import argparse
    import river_route_b200 as rr
    ap = argparse.ArgumentParser(prog='python -m river_route_b200.distributed',
                                 description='basin-sharded multi-GPU run of a river_route_b200 router (launch with torchrun)')
    ap.add_argument('router', choices=['Muskingum', 'RapidMuskingum', 'UnitMuskingum'])
    ap.add_argument('config', nargs='?', default=None)
    args, extra = ap.parse_known_args(argv)
    if len(extra) % 2:
        ap.error('overrides come in "--key value" pairs')
    overrides = {}
    for key, val in zip(extra[::2], extra[1::2]):
        if not key.startswith('--'):
            ap.error(f'bad override {key!r}')
        try:
            import yaml
            val = yaml.safe_load(val)
        except Exception:
            pass
        overrides[key[2:]] = val
    shard = Shard.from_env()
    try:
        getattr(rr, args.router)(args.config, _shard=shard, **overrides).route()
        shard.barrier()
    finally:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == '__main__':
    sys.exit(main())
