"""stateline_b200 -- B200-native parallel-tempering MCMC engine behind stateline's config and
likelihood contract.  The product is the C-ABI library (include/slgpu.h, libslgpu.so) plus the kernel
header plugins instantiate (include/slgpu_device.cuh); this package is the Python host-side mirror."""
from .engine import Engine, EngineGroup, SlgpuError, make_config, format_row, rhat_finish, measure_fp64_peak, lib  # noqa: F401
from . import build, configs  # noqa: F401
