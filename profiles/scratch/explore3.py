import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tdvmc_b200 import driver, systems
np.set_printoptions(linewidth=220)
g = np.load("tests/golden/bosonsbulk_n343_equil.npz")
uR, uI = systems.smooth_params(201, 3.5)
z = np.zeros(201)
cases = [("smooth", uR, uI, 1024, 2, 343, 1e-5, 0), ("smooth", uR, uI, 2048, 2, 343, 1e-5, 0), ("smooth", uR, uI, 4096, 2, 343, 1e-5, 0),
         ("smooth", uR, z, 2048, 2, 343, 1e-5, 0),
         ("zero", z, z, 1024, 2, 343, 2e-4, 1), ("zero", z, z, 1024, 2, 343, 2e-4, 0), ("zero", z, z, 4096, 2, 343, 2e-4, 1),
         ("zero", z, z, 1024, 2, 343, 5e-5, 0), ("zero", z, z, 256, 2, 343, 5e-5, 0)]
for name, a_r, a_i, walkers, nsteps, ntherm, dt, solver in cases:
    cfg = driver.headline_config(a_r, a_i, MC_NSTEPS=nsteps, MC_NTHERMSTEPS=ntherm, MC_NINITIALIZATIONSTEPS=1000,
                                 MC_VERY_FIRST_NINITIALIZATIONSTEPS=34300, TIMESTEP=dt, TOTALTIME=dt * 49.5,
                                 LINEAR_EQUATION_SOLVER_TYPE=solver, USE_PRECONDITIONING=1 if solver == 0 else 0, GPU_WALKERS=walkers)
    t = time.time()
    try:
        a = driver.run_driver(driver.TDVMC_GPU, cfg, "/tmp/drv_gpu", R0=g["R"])
    except Exception as ex:
        print("FAILED", name, walkers, nsteps, dt, solver, str(ex)[-800:])
        continue
    pr = a.parameters_r
    du = np.abs(pr[1:, :201] - pr[:-1, :201]).max(axis=1)
    print(f"{name} walkers={walkers} nsteps={nsteps} ntherm={ntherm} dt={dt} solver={solver}: wall {time.time()-t:.1f}s step_ms={np.median(a.step_ms):.1f} steps={len(a.local_energy_r)}")
    print("  E_R:", " ".join(f"{x:.4g}" for x in a.local_energy_r[::3]))
    print("  max|du|:", " ".join(f"{x:.2g}" for x in du[::3]))
    print("  max|u| end:", np.abs(pr[-1, :201]).max(), "acc", a.acceptance[[0, -1]])
