set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "small_ensemble_forced or time_shared" 2>&1 | tail -3
TDVMC_SWEEP_QUEUE=2 TDVMC_SWEEP_SPLIT=1 python - <<'PY'
import numpy as np, os, sys
sys.path.insert(0,'.')
from tdvmc_b200 import capi, systems
g=np.load('tests/golden/bosonsbulk_n64_equil.npz'); spec=systems.from_golden(g)
h=capi.Handle(spec,333,seed=1,mc_step=0.5,max_samples=1)
h.set_params(g["uR"], g["uI"], 0.0,0.0,0.0); h.set_positions(np.broadcast_to(g["R"],(333,64,3)).copy())
h.profile(True,True); h.sweep(640); print(h.kernel_stats()); h.close()
PY
bash profiles/run_sanitizer.sh 2>&1 | tail -16
