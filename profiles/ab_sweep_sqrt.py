#!/usr/bin/env python3
"""A/B of the sampler's square root (sweep_math.cuh: one Newton step, three FP64 instructions, against the third-order step of
round 1, five).  Run once per library build (TDVMC_LIB selects a measurement build); writes the rate of the headline sweep
(BosonsBulk N = 343, 4096 walkers x 5000 proposals per launch) and, for the comparison of the chains, the positions of 64
walkers after 20 000 proposals and the acceptance count: identical accept decisions give bit-identical positions.
    python profiles/ab_sweep_sqrt.py <tag>    ->  gpurun_out/ab_sweep_sqrt_<tag>.{json,npy}"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tdvmc_b200 import capi, systems
if os.environ.get("TDVMC_LIB"):
    capi.LIB_PATH = os.environ["TDVMC_LIB"]
tag = sys.argv[1]
g = np.load(os.path.join(ROOT, "tests", "golden", "bosonsbulk_n343_equil.npz"))
spec = systems.from_golden(g)
W = 4096
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
R = h.get_positions(0, 64)
h.sample_and_accumulate(1, 0, 0)
f = h.allreduce_and_fetch()
out = {"tag": tag, "walker_steps_per_s": W * 5000 * n / (ms * 1e-3), "ms_per_launch": ms / n,
       "n_acceptances": int(f["n_acceptances"]), "n_trials": int(f["n_trials"]), "e_r": float(f["e_r"][0])}
h.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "ab_sweep_sqrt_%s.npy" % tag), R)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_sweep_sqrt_%s.json" % tag), "w"))
print(json.dumps(out))
