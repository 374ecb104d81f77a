import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tdvmc_b200 import driver, systems
g = np.load("tests/golden/bosonsbulk_n343_equil.npz")
uR, uI = systems.smooth_params(201, 3.5)
for walkers, nsteps, ntherm, dt, solver, ui_on in [(4096, 2, 343, 1e-5, 0, 1), (4096, 8, 343, 1e-5, 0, 0), (4096, 8, 343, 1e-5, 0, 1)]:
    cfg = driver.headline_config(uR, uI if ui_on else 0 * uI, MC_NSTEPS=nsteps, MC_NTHERMSTEPS=ntherm, MC_NINITIALIZATIONSTEPS=1000,
                                 MC_VERY_FIRST_NINITIALIZATIONSTEPS=34300, TIMESTEP=dt, TOTALTIME=dt * 3.5,
                                 LINEAR_EQUATION_SOLVER_TYPE=solver, USE_PRECONDITIONING=1 if solver == 0 else 0, GPU_WALKERS=walkers)
    a = driver.run_driver(driver.TDVMC_GPU, cfg, "/tmp/drv_gpu", R0=g["R"])
    pr = a.parameters_r
    print(f"walkers={walkers} nsteps={nsteps} ntherm={ntherm} dt={dt} solver={solver} uI={ui_on}")
    print(pr[:, :4], pr[:, 100:103], pr[:, -1])
    print(open(a.out_dir + "/ParametersR.dat").read()[:300])
    print(a.log[:6000])
