# one-off: evaluate_he_tile_kernel with __launch_bounds__(256, 1..4) on config 2 (HeBulk N = 64)
for mb in 1 12 16; do
  echo "minblocks=$mb"; python - <<PY
import sys, json; sys.path.insert(0,'.'); sys.path.insert(0,'profiles')
from tdvmc_b200 import capi
capi.LIB_PATH = "tdvmc_b200/libtdvmc_b200_he$mb.so"
import bench_configs as B
CFG = int(__import__("os").environ.get("CFG", "2"))
B.reference_host = lambda *a, **k: None
l = B.run_config(CFG, 3)
print(json.dumps({k: l[k] for k in ("ms_per_pass", "sweep_ms", "evaluate_ms", "samples_per_s", "local_energy_r")}))
PY
done
