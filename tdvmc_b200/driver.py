"""Running the reference's own C++ driver (src/TDVMC.cpp) - GPU-bound (`tdvmc_b200/host/build/TDVMC_gpu`, built by
`tdvmc_b200/host/driver/Makefile`) or unmodified - from Python: config files in the reference's JSON format (version
0.24, all 49 registered items of src/TDVMC.cpp:297-349 plus the three GPU items of TDVMC_gpu_hooks.h), the k-vector
files its InitSystem() insists on, the start coordinates, and a reader for the `.dat` files and the per-step timing
lines it writes (src/Utils.cpp:788-1018, src/TDVMC.cpp:3919-3923).

Used by the time-evolution parity tests and by bench.py's time-steps/s measurement.  Nothing here touches oracle/.
"""
import json
import math
import os
import re
import shutil
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TDVMC_GPU = os.path.join(ROOT, "tdvmc_b200", "host", "build", "TDVMC_gpu")


def write_kvectors(out):
    """kVectors{1,2,3}D.json / kNorm{1,2,3}D.csv / kVectors.json in the format ReadKValuesFromJsonFile reads
    (src/Utils.cpp:1020-1049; BosonsBulk.cpp:124-137): integer wave vectors grouped by shells of equal norm.  They feed
    the S(k) observable only, but no bulk system initialises without them.  Generated from first principles."""
    os.makedirs(out, exist_ok=True)
    m = 21
    shells = {}
    for x in range(m + 1):
        for y in range(m + 1):
            for z in range(m + 1):
                n2 = x * x + y * y + z * z
                if 0 < n2 <= m * m:
                    shells.setdefault(n2, []).append([x, y, z])
    keys = sorted(shells)[:400]
    for name in ("kVectors3D.json", "kVectors.json"):       # the second: the name HeBulk / HeDrop read (HeBulk.cpp:139)
        with open(os.path.join(out, name), "w") as f:
            json.dump({"data": [sorted(shells[k]) for k in keys]}, f)
    with open(os.path.join(out, "kNorm3D.csv"), "w") as f:
        f.write("\n".join(repr(math.sqrt(k)) for k in keys) + "\n")
    with open(os.path.join(out, "kVectors1D.json"), "w") as f:   # InhContactBosons.cpp:170-171
        json.dump({"data": [[[k]] for k in range(1, 401)]}, f)
    with open(os.path.join(out, "kNorm1D.csv"), "w") as f:
        f.write("\n".join(repr(float(k)) for k in range(1, 401)) + "\n")
    shells2 = {}
    for x in range(m + 1):
        for y in range(m + 1):
            n2 = x * x + y * y
            if 0 < n2 <= m * m:
                shells2.setdefault(n2, []).append([x, y])
    keys2 = sorted(shells2)[:120]
    with open(os.path.join(out, "kVectors2D.json"), "w") as f:
        json.dump({"data": [sorted(shells2[k]) for k in keys2]}, f)
    with open(os.path.join(out, "kNorm2D.csv"), "w") as f:
        f.write("\n".join(repr(math.sqrt(k)) for k in keys2) + "\n")


def base_config(**over):
    """Every item RegisterAllConfigItems knows (src/TDVMC.cpp:297-349) with neutral values; absent keys would be left at
    zero by the reference (ConfigItem.cpp:45-48), some of which divide (WRITE_EVERY_NTH_STEP_TO_FILE, :1047)."""
    c = {
        "CONFIG_VERSION": "0.24", "SYSTEM_TYPE": "BosonsBulk", "OUTPUT_DIRECTORY": "", "OUT_DIR_SUFFIX": "", "OUT_DIR_NAME": "out",
        "N": 64, "LBOX": 4.0, "DIM": 3, "N_PARAM": 33, "USE_PARAM_START": 0, "USE_PARAM_END": 0, "RHO": 1.0, "RC": 0.0,
        "MC_STEP": 0.5, "MC_STEP_OFFSET": 0.0, "MC_NSTEPS": 2, "MC_NTHERMSTEPS": 100, "MC_NINITIALIZATIONSTEPS": 100,
        "MC_VERY_FIRST_NINITIALIZATIONSTEPS": 1000, "MC_NADDITIONALSTEPS": 0, "MC_NADDITIONALTHERMSTEPS": 0,
        "MC_NADDITIONALINITIALIZATIONSTEPS": 0, "MC_NFINALSTEPS_MULTIPLICATOR": 1, "TIMESTEP": 1e-4, "TOTALTIME": 1e-3,
        "IMAGINARY_TIME": 1, "ODE_SOLVER_TYPE": 0, "LINEAR_EQUATION_SOLVER_TYPE": 0, "USE_PRECONDITIONING": 1,
        "USE_PARAMETER_ACCEPTANCE_CHECK": 0, "PARAMETER_ACCEPTANCE_CHECK_TYPE": 0, "WRITE_EVERY_NTH_STEP_TO_FILE": 1,
        "CALCULATE_ADDITIONAL_DATA_EVERY_NTH_STEP": 0, "WRITE_SINGLE_FILES": 0, "MC_NSTEP_MULTIPLICATION_FACTOR_FOR_WRITE_DATA": 1,
        "USE_MEAN_FOR_FINAL_PARAMETERS": 0, "USE_NORMALIZE_WF": 1, "USE_ADJUST_PARAMETERS": 0, "UPDATE_SAMPLES_EVERY_NTH_STEP": 0,
        "UPDATE_SAMPLES_PERCENT": 100.0, "GR_BIN_COUNT": 100, "RHO_BIN_COUNT": 100, "USE_NURBS": 0, "NURBS_GRID": [0.0],
        "PARTICLE_TYPES": [0], "SYSTEM_PARAMS": [1.0, 1.0], "PARAMS_REAL": [0.0], "PARAMS_IMAGINARY": [0.0],
        "PARAM_PHIR": 0.0, "PARAM_PHII": 0.0,
        "GPU_WALKERS": 0, "GPU_SEED": 1, "GPU_DEVICE_SOLVE": 0,
    }
    unknown = set(over) - set(c)
    if unknown:
        raise KeyError(f"not a config item of the reference driver: {sorted(unknown)}")
    c.update(over)
    return c


def headline_config(uR, uI, **over):
    """BASELINE configs[2]: config/BosonsBulk3D.config scaled to N = 343, LBOX = 7, N_PARAM = 201 with its own Monte-Carlo
    counts (MC_NSTEPS = 2, MC_NTHERMSTEPS = 5000, MC_NINITIALIZATIONSTEPS = 1000), Euler integrator."""
    c = base_config(SYSTEM_TYPE="BosonsBulk", N=343, LBOX=7.0, DIM=3, N_PARAM=201, MC_STEP=0.5, MC_NSTEPS=2, MC_NTHERMSTEPS=5000,
                    MC_NINITIALIZATIONSTEPS=1000, MC_VERY_FIRST_NINITIALIZATIONSTEPS=100000, TIMESTEP=2e-4, IMAGINARY_TIME=1,
                    ODE_SOLVER_TYPE=0, LINEAR_EQUATION_SOLVER_TYPE=1, USE_PRECONDITIONING=0, SYSTEM_PARAMS=[1.0, 1.0],
                    GR_BIN_COUNT=500, PARAMS_REAL=[float(x) for x in uR], PARAMS_IMAGINARY=[float(x) for x in uI])
    c.update(over)
    return c


def nubosons_config(nurbs_grid, uR, uI, **over):
    """BASELINE configs[3]: config/NUBosonsBulkPB3D.config with its own values (N = 1728, LBOX = 12, N_PARAM = 200 on the
    config's 201-point NURBS grid, MC_NSTEPS = 50, MC_NTHERMSTEPS = 200, real time, Euler, Cholesky, sample reuse on:
    UPDATE_SAMPLES_EVERY_NTH_STEP = 1, UPDATE_SAMPLES_PERCENT = 100)."""
    c = base_config(SYSTEM_TYPE="NUBosonsBulkPB", N=1728, LBOX=12.0, DIM=3, N_PARAM=200, MC_STEP=0.5, MC_NSTEPS=50,
                    MC_NTHERMSTEPS=200, MC_NINITIALIZATIONSTEPS=400, MC_VERY_FIRST_NINITIALIZATIONSTEPS=500000, TIMESTEP=1e-5,
                    IMAGINARY_TIME=0, ODE_SOLVER_TYPE=0, LINEAR_EQUATION_SOLVER_TYPE=0, USE_PRECONDITIONING=1, USE_NURBS=1,
                    NURBS_GRID=[float(x) for x in nurbs_grid], SYSTEM_PARAMS=[0.0, 0.0, 0.0, 0.1, 50.0], GR_BIN_COUNT=400,
                    UPDATE_SAMPLES_EVERY_NTH_STEP=1, UPDATE_SAMPLES_PERCENT=100.0,
                    PARAMS_REAL=[float(x) for x in uR], PARAMS_IMAGINARY=[float(x) for x in uI])
    c.update(over)
    return c


def read_dat(path):
    """One of the driver's `.dat` files (header line, then rows of setw(24) scientific numbers, src/Utils.h:25-28)."""
    rows = []
    with open(path) as f:
        f.readline()
        for line in f:
            vals = line.split()
            if vals:
                rows.append([float(v) for v in vals])
    if not rows:
        return np.zeros((0,))
    width = max(len(r) for r in rows)
    a = np.array([r + [np.nan] * (width - len(r)) for r in rows])
    return a[:, 0] if width == 1 else a


class DriverRun:
    """Result of one driver run: the per-step series the root rank appended (src/TDVMC.cpp:3833-3839) and the log."""

    def __init__(self, out_dir, log, wall_s):
        self.out_dir, self.log, self.wall_s = out_dir, log, wall_s
        self.step_ms = np.array([float(m) for m in re.findall(r"duration for full timestep: ([0-9.eE+-]+) ms", log)])
        self.acceptance = np.array([float(m) for m in re.findall(r"Acceptance AVG: ([0-9.eE+-]+)%", log)])

    def series(self, name):
        return read_dat(os.path.join(self.out_dir, name + ".dat"))

    @property
    def local_energy_r(self):
        return self.series("LocalEnergyR")

    @property
    def local_energy_i(self):
        return self.series("LocalEnergyI")

    @property
    def parameters_r(self):
        return self.series("ParametersR")        # [step][N_PARAM + 1]: uR..., phiR

    @property
    def parameters_i(self):
        return self.series("ParametersI")

    @property
    def local_operators(self):
        return self.series("LocalOperators")

    @property
    def times(self):
        return self.series("timesSystem")


def write_rng_state(cfg_dir, seed, n_particles, rank=0):
    """random/state_{generator,uniform,normal,particleIndex}_<rank>.dat as Init() restores them (src/TDVMC.cpp:516-519,
    src/Utils.cpp:48-63): the text form of std::mt19937_64(seed) and of the three distributions - the only way to give a
    single-rank process of the reference another stream than mt19937_64(rank + 1) (:524)."""
    M = (1 << 64) - 1
    x = [seed & M]
    for i in range(1, 312):
        x.append((6364136223846793005 * (x[-1] ^ (x[-1] >> 62)) + i) & M)
    d = os.path.join(cfg_dir, "random")
    os.makedirs(d, exist_ok=True)
    for name, text in (("generator", " ".join(map(str, x)) + " 312"), ("uniform", "0 1"), ("normal", "0 1 0"),
                       ("particleIndex", f"0 {n_particles - 1}")):
        with open(os.path.join(d, f"state_{name}_{rank}.dat"), "w") as f:
            f.write(text + "\n")


def run_driver(binary, config, workdir, R0=None, timeout=3600, env=None, seed=None, prefix=None):
    """Writes `config` (dict from base_config & co.) and the auxiliary files into `workdir`, runs `binary <config>` there
    and returns a DriverRun.  R0 ([N][DIM]) becomes coords/particleconfiguration_<N>_<D>D_0.csv, the restart format
    InitCoordinateConfiguration reads (src/TDVMC.cpp:677, :640-670).  `seed` (host RNG of the reference's CPU path and of the
    start jitter) goes in through the RNG state files; the device ensemble's stream is the config item GPU_SEED.  `prefix`:
    command prefix such as ["taskset", "-c", "3"]."""
    os.makedirs(workdir, exist_ok=True)
    cfg_dir = os.path.join(workdir, "cfg")
    out_root = os.path.join(workdir, "output")
    for d in (cfg_dir, out_root):
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
    write_kvectors(cfg_dir)
    c = dict(config)
    c["OUTPUT_DIRECTORY"] = out_root + "/"
    if R0 is not None:
        R0 = np.asarray(R0, float)[:, :c["DIM"]]
        os.makedirs(os.path.join(cfg_dir, "coords"), exist_ok=True)
        with open(os.path.join(cfg_dir, "coords", f"particleconfiguration_{c['N']}_{c['DIM']}D_0.csv"), "w") as f:
            f.write(",".join(repr(float(x)) for x in R0.ravel()) + "\n")
    if seed is not None:
        write_rng_state(cfg_dir, int(seed), c["N"])
    path = os.path.join(cfg_dir, "run.config")
    with open(path, "w") as f:
        json.dump(c, f, indent=1)
    import time
    t0 = time.perf_counter()
    p = subprocess.run(list(prefix or []) + [os.path.abspath(binary), path], cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout,
                       env=env)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(f"{os.path.basename(binary)} exited with {p.returncode}:\n{p.stdout[-3000:]}")
    return DriverRun(os.path.join(out_root, c["OUT_DIR_NAME"]), p.stdout, wall)
