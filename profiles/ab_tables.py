#!/usr/bin/env python3
"""K3 / K4 exhibit timing: tables_kernel and contract_kernel on 1024 resident configurations of the headline system."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tdvmc_b200 import capi
spec, uR, uI, R = bench.golden_spec()
W = 1024
h = capi.Handle(spec, W, seed=1, mc_step=0.5, max_samples=1)
h.set_params(uR, uI, 0.0, 0.0, 0.0)
rng = np.random.default_rng(1)
h.set_positions(R[None] + rng.uniform(-0.02, 0.02, (W, 343, 3)))
h.sweep(3430)
h.tables_resident(W); h.contract_resident(W, fetch=False); h.flush_l2()
h.profile(True, True)
for _ in range(3):
    h.tables_resident(W); h.flush_l2(); h.contract_resident(W, fetch=False); h.flush_l2()
st = h.kernel_stats(); h.profile(False, False)
t_tab = st["tables"][1] / st["tables"][0] * 1e-3; t_con = st["contract"][1] / st["contract"][0] * 1e-3
e = h.contract_resident(8)
f = h.evaluate_fixed(h.get_positions()[:8])
print(json.dumps({"tables_ms": t_tab * 1e3, "tables_GBps": W * bench.BYTES_PER_TABLE / t_tab / 1e9, "frac_of_6491": W * bench.BYTES_PER_TABLE / t_tab / 1e9 / 6491.2,
                  "contract_ms": t_con * 1e3, "contract_GBps": W * 8 * 203 * 343 * 4 / t_con / 1e9,
                  "max_rel_diff_vs_fused": float(np.max(np.abs(e[0] - f["e_r"]) / np.abs(f["e_r"])))}))
