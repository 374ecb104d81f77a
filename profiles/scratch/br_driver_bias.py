"""One-off: E_R of the BoxAndRadial driver test with 24 seeds per arm (is the 3.4 sigma of the 8-seed run a bias?)."""
import sys, os, tempfile, pathlib
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
from tdvmc_b200 import driver
import test_gpu_driver as T
g = np.load(os.path.join(T.ROOT, "tests/golden/boxradial_n27_equil.npz"), allow_pickle=True)
cfg = driver.base_config(SYSTEM_TYPE="NUBosonsBulkPBBoxAndRadial", N=27, LBOX=float(g["LBOX"]), DIM=3, N_PARAM=100, MC_STEP=0.5,
                         MC_NSTEPS=4000, MC_NTHERMSTEPS=125, MC_NINITIALIZATIONSTEPS=250, MC_VERY_FIRST_NINITIALIZATIONSTEPS=10000,
                         TIMESTEP=1e-6, TOTALTIME=1e-6 * 5.5, IMAGINARY_TIME=1, ODE_SOLVER_TYPE=0, LINEAR_EQUATION_SOLVER_TYPE=0,
                         USE_PRECONDITIONING=1, USE_PARAMETER_ACCEPTANCE_CHECK=1, PARAMETER_ACCEPTANCE_CHECK_TYPE=5, USE_NORMALIZE_WF=1,
                         GR_BIN_COUNT=400, USE_NURBS=1, NURBS_GRID=[float(x) for x in g["NURBS_GRID"]],
                         SYSTEM_PARAMS=[float(x) for x in g["SYSTEM_PARAMS"]], PARAMS_REAL=[float(x) for x in g["uR"]],
                         PARAMS_IMAGINARY=[float(x) for x in g["uI"]], PARAM_PHIR=float(g["phiR"]))
tmp = pathlib.Path(tempfile.mkdtemp())
seeds = list(range(1, 25))
ref = T.run_seeds(T.TDVMC_REF, cfg, "ref", g["R"], seeds, tmp)
for mcinit in (250, 5000):
    dev = T.run_seeds(driver.TDVMC_GPU, dict(cfg, GPU_WALKERS=2000, MC_NSTEPS=2, MC_NINITIALIZATIONSTEPS=mcinit), f"gpu{mcinit}", g["R"], seeds, tmp, gpu_seed=True)
    a = np.stack([r.local_energy_r for r in ref]); b = np.stack([r.local_energy_r for r in dev])
    z = (b.mean(0) - a.mean(0)) / np.sqrt(a.var(0, ddof=1) / len(ref) + b.var(0, ddof=1) / len(dev))
    print("init", mcinit, "ref", a.mean(0), "dev", b.mean(0), "z", z, "sd", a.std(0, ddof=1), b.std(0, ddof=1))
