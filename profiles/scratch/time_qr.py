"""solve_qr kernels at P = 201: cluster (default) vs one-CTA global-memory kernel (force_global)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tdvmc_b200 import capi, systems
g = np.load(os.path.join(ROOT, "tests/golden/bosonsbulk_n343_equil.npz"))
spec = systems.from_golden(g)
W = 592
h = capi.Handle(spec, W, seed=11, mc_step=0.5, max_samples=2)
h.set_params(g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), float(g["time"]))
h.set_positions(np.broadcast_to(g["R"], (W, 343, 3)).copy())
h.sample_and_accumulate(2, 343, 343)
out = {}
for fg in (False, True):
    for pre in (False, True):
        d = h.solve_parameters_dot(imaginary_time=1, solver_type=1, use_preconditioning=pre, force_global=fg)
        h.profile(True, True)
        for _ in range(5):
            d = h.solve_parameters_dot(imaginary_time=1, solver_type=1, use_preconditioning=pre, force_global=fg)
        n, ms = h.kernel_stats()["solve"]
        h.profile(False, False)
        out[(fg, pre)] = (ms / n, d["u_dot_r"].copy(), d["phi_dot_r"])
        print("force_global", fg, "precond", pre, "ms per solve", ms / n, "u_dot_r[:3]", d["u_dot_r"][:3])
for pre in (False, True):
    print("bit-identical (precond %s):" % pre, np.array_equal(out[(False, pre)][1], out[(True, pre)][1]), out[(False, pre)][2] == out[(True, pre)][2])
h.close()
