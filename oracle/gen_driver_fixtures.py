#!/usr/bin/env python3
"""Golden trajectories of the UNMODIFIED reference program (oracle/_ref/TDVMC_ref, the real main() with its config
file, output directory and .dat files) for the driver-level parity tests of tests/test_gpu_driver.py.

    python oracle/gen_driver_fixtures.py cfg3 [n_seeds]      # -> tests/golden/driver_cfg3_reference.npz

cfg3: BASELINE configs[2] at its own size (BosonsBulk, N = 343, LBOX = 7, N_PARAM = 201), imaginary-time Euler steps
with the Cholesky solve.  The reference is ONE Markov chain per process, so a time step with M samples costs
M x (343 proposals + one evaluation) = M x 39 ms; M = 4096 decorrelated samples are needed for a stable evolution of
201 parameters (profiles/r02_driver_stability.txt), i.e. 160 s per time step and core.  Hence 10 steps of 1e-5 per seed,
seeds in parallel processes (RNG state files, tdvmc_b200/driver.py::write_rng_state).
"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tdvmc_b200 import driver, systems  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref", "TDVMC_ref")


def cfg3_config(**over):
    uR, uI = systems.smooth_params(201, 3.5)
    c = driver.headline_config(uR, uI, MC_NSTEPS=4096, MC_NTHERMSTEPS=343, MC_NINITIALIZATIONSTEPS=1000,
                               MC_VERY_FIRST_NINITIALIZATIONSTEPS=34300, TIMESTEP=1e-5, TOTALTIME=1e-5 * 9.5,
                               LINEAR_EQUATION_SOLVER_TYPE=0, USE_PRECONDITIONING=1)
    c.update(over)
    return c


def gen_cfg3(n_seeds):
    g = np.load(os.path.join(GOLDEN, "bosonsbulk_n343_equil.npz"))
    cfg = cfg3_config()

    def one(seed):
        r = driver.run_driver(REF, cfg, f"/tmp/driver_fixture_cfg3_{seed}", R0=g["R"], seed=seed, timeout=6 * 3600)
        return dict(e_r=r.local_energy_r, e_i=r.local_energy_i, p_r=r.parameters_r, p_i=r.parameters_i, o=r.local_operators,
                    step_ms=r.step_ms, acc=r.acceptance)

    seeds = list(range(11, 11 + n_seeds))
    with ThreadPoolExecutor(max_workers=n_seeds) as ex:
        runs = list(ex.map(one, seeds))
    out = {k: np.stack([r[k] for r in runs]) for k in runs[0]}
    np.savez_compressed(os.path.join(GOLDEN, "driver_cfg3_reference.npz"), seeds=np.array(seeds), config=np.array(repr(cfg)),
                        MC_NSTEPS=cfg["MC_NSTEPS"], MC_NTHERMSTEPS=cfg["MC_NTHERMSTEPS"], TIMESTEP=cfg["TIMESTEP"],
                        source=np.array("bosonsbulk_n343_equil"), **out)
    print("E_R mean per step", out["e_r"].mean(axis=0), "+-", out["e_r"].std(axis=0, ddof=1))
    print("step seconds", out["step_ms"].mean() / 1e3)


if __name__ == "__main__":
    what = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    if what == "cfg3":
        gen_cfg3(n)
    else:
        sys.exit("usage: gen_driver_fixtures.py cfg3 [n_seeds]")
