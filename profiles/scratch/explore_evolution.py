"""Scratch: stability of the Euler evolution at config-3 size through TDVMC_gpu for several sample counts / time steps."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tdvmc_b200 import driver, systems
g = np.load("tests/golden/bosonsbulk_n343_equil.npz")
uR, uI = systems.smooth_params(201, 3.5)
for walkers, nsteps, ntherm, dt, solver, ui_on in [(512, 2, 100, 1e-5, 0, 1), (4096, 2, 343, 1e-5, 0, 1), (4096, 8, 343, 1e-5, 0, 1),
                                                   (4096, 8, 343, 1e-4, 0, 1), (4096, 8, 343, 1e-5, 1, 1), (4096, 8, 343, 1e-5, 0, 0),
                                                   (4096, 8, 343, 1e-4, 0, 0)]:
    cfg = driver.headline_config(uR, uI if ui_on else 0 * uI, MC_NSTEPS=nsteps, MC_NTHERMSTEPS=ntherm, MC_NINITIALIZATIONSTEPS=1000,
                                 MC_VERY_FIRST_NINITIALIZATIONSTEPS=34300, TIMESTEP=dt, TOTALTIME=dt * 19.5,
                                 LINEAR_EQUATION_SOLVER_TYPE=solver, USE_PRECONDITIONING=1 if solver == 0 else 0, GPU_WALKERS=walkers)
    t = time.time()
    try:
        a = driver.run_driver(driver.TDVMC_GPU, cfg, "/tmp/drv_gpu", R0=g["R"])
    except Exception as ex:
        print("FAILED", walkers, nsteps, dt, solver, str(ex)[-1500:])
        continue
    pr = a.parameters_r
    print(f"walkers={walkers} nsteps={nsteps} ntherm={ntherm} dt={dt} solver={solver} uI={ui_on}: wall {time.time()-t:.1f}s step_ms={np.median(a.step_ms):.1f}")
    print("  E_R:", np.array2string(a.local_energy_r, precision=2, max_line_width=200))
    print("  E_I:", np.array2string(a.local_energy_i, precision=2, max_line_width=200))
    print("  max|du| per step:", np.array2string(np.abs(pr[1:, :201] - pr[:-1, :201]).max(axis=1), precision=2, max_line_width=200))
    print("  acceptance:", a.acceptance[:3])
print(a.log[:3000])
