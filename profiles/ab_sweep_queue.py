#!/usr/bin/env python3
"""Sweep rate for ensembles that are not a whole number of waves: time-shared launch (sweep_queue_kernel) vs plain launch
(TDVMC_SWEEP_QUEUE=0).  python profiles/ab_sweep_queue.py; TDVMC_SWEEP_QUEUE=0 python profiles/ab_sweep_queue.py"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tdvmc_b200 import capi, systems
out = {"queue": os.environ.get("TDVMC_SWEEP_QUEUE", "auto")}
g = np.load(os.path.join(ROOT, "tests", "golden", "bosonsbulk_n343_equil.npz"))
spec = systems.from_golden(g)
for W in (2960, 3500, 4096, 5000, 5920, 8000):
    h = capi.Handle(spec, W, seed=1, mc_step=0.5, max_samples=1)
    h.set_params(g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), float(g["time"]))
    rng = np.random.default_rng(1)
    h.set_positions(g["R"][None] + rng.uniform(-0.01, 0.01, (W, 343, 3)))
    h.sweep(5000); h.synchronize()
    h.profile(True, True)
    for _ in range(3):
        h.sweep(5000)
    n, ms = h.kernel_stats()["sweep"]
    h.profile(False, False)
    R = h.get_positions()
    e = h.evaluate_fixed(R[[0, W - 1]])
    h.sample_and_accumulate(1, 0, 0)
    f = h.allreduce_and_fetch()
    out[f"W{W}"] = {"walker_steps_per_s": W * 5000 * n / (ms * 1e-3), "ms_per_launch": ms / n, "e_r_first_last": [float(x) for x in e["e_r"]],
                    "acceptance": f["n_acceptances"] / f["n_trials"]}
    h.close()
print(json.dumps(out))
