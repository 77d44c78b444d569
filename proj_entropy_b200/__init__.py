"""proj_entropy_b200 -- B200 (sm_100a) replacement for the all-pairs gravity
step of sthaid/proj_entropy (sim_gravity.c:1183-1255).

The product is two in-tree native libraries:

  libgravity_cuda.so   hand-written sm_100a kernels behind the C ABI of
                       include/gravity_cuda.h (what the reference's C host binds)
  libgravity_host.so   host C: scenario reader, Plummer model, day sampling
                       (include/gravity_host.h)

This Python package is only the ctypes mirror of those two headers, used by the
tests and by bench.py.  It never computes anything itself and has no CPU
fallback: without the built CUDA library every call raises.
"""
from .gravity import (GravityCuda, GravityError, KERNEL_AUTO, KERNEL_TILED, KERNEL_STRICT,
                      device_count, fp64_peak, nccl_unique_id, load_cuda_library, plan_probe, path_samples_in)
from .host import (Scenario, ScenarioError, load_scenario, plummer, steps_until_day_change, load_host_library,
                   snapshot_read, snapshot_write)

__all__ = ["GravityCuda", "GravityError", "KERNEL_AUTO", "KERNEL_TILED", "KERNEL_STRICT",
           "device_count", "fp64_peak", "nccl_unique_id", "load_cuda_library", "plan_probe", "path_samples_in",
           "Scenario", "ScenarioError", "load_scenario", "plummer", "steps_until_day_change",
           "load_host_library", "snapshot_read", "snapshot_write"]
