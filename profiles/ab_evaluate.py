#!/usr/bin/env python3
"""A/B timing of the evaluation kernel alone: samples/s at the headline size (BosonsBulk N=343, P=201, one full wave of
2960 configurations per launch) and at config 4's size (NUBosonsBulkPB N=1728, P=200, 512 configurations per launch).
    python profiles/ab_evaluate.py            # current kernel;  TDVMC_EVAL_V1=1 python ... : the r01 kernel"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tdvmc_b200 import capi, systems  # noqa: E402

out = {}
for name, W, S, n_therm in (("bosonsbulk_n343_equil", 2960, 4, 343), ("nubosonsbulkpb_n1728_equil", 512, 8, 200)):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    spec = systems.from_golden(g)
    h = capi.Handle(spec, W, seed=1, mc_step=0.5, max_samples=S)
    h.set_params(g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), float(g["time"]))
    rng = np.random.default_rng(1)
    h.set_positions(g["R"][None] + rng.uniform(-0.01, 0.01, (W, spec.n_particles, 3)))
    h.sweep(10 * spec.n_particles)
    h.sample_and_accumulate(S, n_therm, 0)
    e = h.allreduce_and_fetch()
    h.profile(True, True)
    for _ in range(3):
        h.sample_and_accumulate(S, n_therm, 0)
    st = h.kernel_stats()
    h.profile(False, False)
    n, ms = st["evaluate"]
    out[name] = {"launches": n, "ms_per_launch": ms / n, "configs_per_launch": 3 * S * W // n, "samples_per_s": 3 * S * W / (ms * 1e-3),
                 "ms_per_4096_samples": ms * 4096 / (3 * S * W), "e_r": float(e["e_r"][0])}
    h.close()
print(json.dumps(out))
