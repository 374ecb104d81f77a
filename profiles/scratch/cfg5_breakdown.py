"""one-off: where the time of a config-5 pass goes (kernel classes + wall)"""
import sys, json; sys.path.insert(0, "."); sys.path.insert(0, "profiles")
import numpy as np
import bench_configs as B
from tdvmc_b200 import capi, systems
path, fixture, system, mc_step, n_therm, n_init, W = B.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 5]
g = np.load(B.GOLDEN + "/" + fixture + ".npz"); spec = systems.from_golden(g)
h = capi.Handle(spec, W, seed=1, mc_step=mc_step, max_samples=B.SAMPLES)
h.set_params(g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), float(g["time"]))
h.set_positions(np.broadcast_to(g["R"], (W, spec.n_particles, 3)).copy()); h.sweep(100)
h.sample_and_accumulate(B.SAMPLES, n_therm, n_init); h.allreduce_and_fetch()
h.profile(True, True); h.synchronize(); h.timer_start()
h.sample_and_accumulate(B.SAMPLES, n_therm, n_init); ms1 = h.timer_stop()
h.timer_start(); out = h.allreduce_and_fetch(); ms2 = h.timer_stop()
print(json.dumps({"sample_and_accumulate_ms": ms1, "allreduce_and_fetch_ms": ms2, "stats": h.kernel_stats(), "W": W}))
