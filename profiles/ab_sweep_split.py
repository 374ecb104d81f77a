#!/usr/bin/env python3
"""Sweep rate (walker-steps/s, sweep kernel only) against the ensemble size, with several warps per walker
(sweep_split_kernel, chosen by sweep_split_warps) and with one warp per walker (TDVMC_SWEEP_SPLIT=1).
    python profiles/ab_sweep_split.py; TDVMC_SWEEP_SPLIT=1 python profiles/ab_sweep_split.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tdvmc_b200 import capi, systems  # noqa: E402

out = {"split": os.environ.get("TDVMC_SWEEP_SPLIT", "auto")}
for name, Ws, steps in (("bosonsbulk_n343_equil", (128, 256, 512, 1024, 2960), 5000), ("nubosonsbulkpb_n1728_equil", (128, 512), 2000)):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    spec = systems.from_golden(g)
    for W in Ws:
        h = capi.Handle(spec, W, seed=1, mc_step=0.5, max_samples=1)
        h.set_params(g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), float(g["time"]))
        rng = np.random.default_rng(1)
        h.set_positions(g["R"][None] + rng.uniform(-0.01, 0.01, (W, spec.n_particles, 3)))
        h.sweep(steps)
        h.synchronize()
        h.profile(True, True)
        for _ in range(3):
            h.sweep(steps)
        n, ms = h.kernel_stats()["sweep"]
        h.profile(False, False)
        e = h.evaluate_fixed(h.get_positions()[:1])
        out[f"{name}_W{W}"] = {"walker_steps_per_s": W * steps * n / (ms * 1e-3), "ms_per_launch": ms / n, "e_r_walker0": float(e["e_r"][0])}
        h.close()
print(json.dumps(out))
