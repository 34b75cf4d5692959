"""
river_route_b200 -- B200 (sm_100a) implementation of river-route's routing hot path.

Host code is Python; every computation goes through the C ABI of ``librr_b200.so``
(include/rr_b200.h).  There is no CPU fallback: importing this package fails if the shared
library has not been built, and compute calls raise ``RuntimeError`` without a CUDA device.
"""
from . import _lib  # noqa: F401  (fails loudly when librr_b200.so is missing)
from ._lib import cuda_available, pinned_empty
from .kernels import muskingum_route, rapid_route, unit_route
from .routers import Configs, Muskingum, RapidMuskingum, UnitMuskingum
from .uhkernels import UnitHydrograph
from .runoff import runoff_to_qlateral, weights_to_qlateral
from .plan import MODE_MUSKINGUM, MODE_RAPID, MODE_UNIT, Plan, downstream_index, label_basins, launch_count

__version__ = '0.1.0'

__all__ = [
    'Configs', 'Muskingum', 'RapidMuskingum', 'UnitMuskingum', 'UnitHydrograph', 'runoff_to_qlateral',
    'weights_to_qlateral',
    'Plan', 'MODE_MUSKINGUM', 'MODE_RAPID', 'MODE_UNIT',
    'muskingum_route', 'rapid_route', 'unit_route',
    'downstream_index', 'label_basins', 'launch_count',
    'cuda_available', 'pinned_empty',
]
