#!/usr/bin/env python3
"""Sweep-kernel tuning probe: walker-steps/s of tdvmc_gpu_sweep alone for the headline system.
Tuning knobs are read by the library from the environment (TDVMC_SWEEP_UNROLL, TDVMC_SWEEP_WPB)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tdvmc_b200 import capi, systems  # noqa: E402

spec, uR, uI, R = bench.golden_spec()
probe = capi.Handle(spec, 1)
per_sm, sms = probe.resident_walkers()
probe.close()
W = per_sm * sms
h = capi.Handle(spec, W, seed=1, mc_step=bench.MC_STEP)
rng = np.random.default_rng(1)
h.set_params(uR, uI)
h.set_positions(systems.jittered_lattice(bench.N, bench.LBOX, rng)[None] + rng.uniform(-0.02, 0.02, (W, bench.N, 3)))
h.sweep(20 * bench.N)
h.synchronize()
n = 5000
best = 0.0
for _ in range(3):
    h.timer_start()
    h.sweep(n)
    ms = h.timer_stop()
    best = max(best, W * n / ms / 1e3)
print(f"unroll={os.environ.get('TDVMC_SWEEP_UNROLL', '2')} wpb={os.environ.get('TDVMC_SWEEP_WPB', 'auto')} "
      f"resident/SM={per_sm} W={W}: {best:.1f} M walker-steps/s ({ms / n * 1e3:.2f} us per step of all walkers)")
