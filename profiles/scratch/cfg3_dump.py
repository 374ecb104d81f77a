import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tdvmc_b200 import driver, systems
f = np.load("tests/golden/driver_cfg3_reference.npz")
g = np.load("tests/golden/bosonsbulk_n343_equil.npz")
uR, uI = systems.smooth_params(201, 3.5)
n_steps = f["e_r"].shape[1]
cfg = driver.headline_config(uR, uI, MC_NSTEPS=2, MC_NTHERMSTEPS=int(f["MC_NTHERMSTEPS"]), MC_NINITIALIZATIONSTEPS=1000,
                             MC_VERY_FIRST_NINITIALIZATIONSTEPS=34300, TIMESTEP=float(f["TIMESTEP"]),
                             TOTALTIME=float(f["TIMESTEP"]) * (n_steps - 0.5), LINEAR_EQUATION_SOLVER_TYPE=0, USE_PRECONDITIONING=1,
                             GPU_WALKERS=int(f["MC_NSTEPS"]) // 2)
out = {}
for sd in range(1, 9):
    r = driver.run_driver(driver.TDVMC_GPU, dict(cfg, GPU_SEED=sd), "/tmp/cfg3dump", R0=g["R"], seed=sd)
    out[f"e_r_{sd}"] = r.local_energy_r; out[f"p_r_{sd}"] = r.parameters_r; out[f"p_i_{sd}"] = r.parameters_i; out[f"o_{sd}"] = r.local_operators
np.savez_compressed("gpurun_out/cfg3_gpu_runs.npz", **out)
