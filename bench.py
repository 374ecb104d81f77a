#!/usr/bin/env python3
"""Headline benchmark: walker-steps/s of the walker-parallel sampling + evaluation path.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores

Workload (BASELINE.json configs[2]): config/BosonsBulk3D.config scaled to N=343, LBOX=7, N_PARAM=201.
One "step" is one ParallelUpdateExpectationValues pass (src/TDVMC.cpp:1152-1188) with the config's own
sample counts: MC_NINITIALIZATIONSTEPS=1000 + MC_NSTEPS=2 x MC_NTHERMSTEPS=5000 Metropolis proposals
per walker, 2 evaluations per walker, the S/F accumulation, the packed all-reduce and the fetch of
the seven estimator arrays, for W = 4096 walkers per GPU.  value = proposals of all walkers on all GPUs /
time (walker-steps/s).

Prints ONE JSON line (rank 0).  Under torchrun (N > 1) every rank drives one GPU; walkers are
sharded (fixed count per GPU -> "weak"), the only collective is the packed NCCL all-reduce.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tdvmc_b200 import systems  # noqa: E402

# config/BosonsBulk3D.config, scaled as BASELINE.json configs[2] says
N, LBOX, N_PARAM = 343, 7.0, 201
MC_STEP, MC_NSTEPS, MC_NTHERMSTEPS, MC_NINIT = 0.5, 2, 5000, 1000
SYSTEM_PARAMS = [1.0, 1.0]
STEPS_PER_WALKER = MC_NINIT + MC_NSTEPS * MC_NTHERMSTEPS
# algorithmic work per unit
FLOP_PER_WALKER_STEP = 2 * (N - 1) * (22 + 1 + 6)          # 19 836
FLOP_PER_EVALUATION = 5.0e6
BYTES_PER_TABLE = 8 * 203 * N * 4 + 24 * N                  # 2 236 360 (K3 table kernel)
METRIC = "walker-steps/s"
SURVEY_WALKERS_PER_GPU = 4096   # W = 4096 x G walkers
# dram__bytes_read.sum + dram__bytes_write.sum of one sweep launch (4096 walkers, 5000 steps): read from the committed ncu
# --set full summary of THIS build's kernel (profiles/run_r02.sh); algorithmic traffic is 2 x 8.2 KB per walker per launch =
# 67 MB; the time-shared kernel moves a walker through L2 once per chunk (10 chunks), which stays in the 126 MB L2
SWEEP_NCU_SUMMARY = os.path.join("profiles", "r02f_sweep_queue_ncu.txt")


def sweep_traffic_from_ncu():
    """(bytes per launch, source) from the committed summary, or (None, reason)."""
    path = os.path.join(ROOT, SWEEP_NCU_SUMMARY)
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        tot, seen = 0.0, 0
        for line in open(path):
            t = line.split()
            if len(t) >= 3 and t[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(t[1].replace(",", "")) * unit.get(t[2], 1.0)
                seen += 1
        if seen == 2:
            return tot, SWEEP_NCU_SUMMARY + " (ncu --set full, sweep_queue_kernel<1,0>, 4096 walkers x 5000 steps per launch, final r02 build)"
    except OSError:
        pass
    return None, "no ncu summary found"


def golden_spec():
    g = np.load(os.path.join(ROOT, "tests", "golden", "bosonsbulk_n343_equil.npz"))
    spec = systems.bosons_bulk(N, LBOX, N_PARAM, SYSTEM_PARAMS, weights=g["spline_weights"])
    assert np.array_equal(spec.knots, g["knots"])
    uR, uI = systems.smooth_params(N_PARAM, LBOX / 2)
    return spec, uR, uI, g["R"]


def workload_config(extra=None):
    c = {"workload": "BosonsBulk3D.config scaled to N=343 (LBOX=7, N_PARAM=201): ParallelUpdateExpectationValues pass",
         "N": N, "LBOX": LBOX, "N_PARAM": N_PARAM, "MC_STEP": MC_STEP, "MC_NSTEPS": MC_NSTEPS, "walkers_per_gpu_survey": SURVEY_WALKERS_PER_GPU,
         "MC_NTHERMSTEPS": MC_NTHERMSTEPS, "MC_NINITIALIZATIONSTEPS": MC_NINIT,
         "proposals_per_walker_per_step": STEPS_PER_WALKER,
         "spline_table": "reference SplineFactory::GetWeights3 output (tests/golden/bosonsbulk_n343_equil.npz)"}
    c.update(extra or {})
    return c


# ------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference (oracle/_ref/ref_harness) on the host cores
# ------------------------------------------------------------------------------------------------
def physical_cores():
    allowed = sorted(os.sched_getaffinity(0))
    seen, cpus = set(), []
    for c in allowed:
        try:
            sib = open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list").read().strip()
        except OSError:
            sib = str(c)
        if sib not in seen:
            seen.add(sib)
            cpus.append(c)
    return cpus


def write_reference_case(path, spec, uR, uI, R, seed):
    with open(path, "w") as f:
        f.write("system BosonsBulk\n")
        f.write(f"configdir {ROOT}/oracle/_ref/config/\n")
        for k, v in dict(N=N, DIM=3, LBOX=LBOX, N_PARAM=N_PARAM, MC_STEP=MC_STEP, MC_NSTEPS=MC_NSTEPS,
                         MC_NTHERMSTEPS=MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS=MC_NINIT, seed=seed).items():
            f.write(f"{k} {v!r}\n")
        for k, v in dict(R=R, uR=uR, uI=uI, SYSTEM_PARAMS=SYSTEM_PARAMS).items():
            f.write(k + " " + " ".join(repr(float(x)) for x in np.asarray(v).ravel()) + "\n")


def reference_step(cases, cpus):
    """Every host core runs ONE walker's pass through the reference's own code: 1000 initialization proposals (untimed,
    uncounted), then 2 x 5000 proposals and 2 evaluations + the reference's estimator accumulation (timed, counted);
    returns (counted proposals, seconds of the slowest)."""
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    procs = []
    for case, cpu in zip(cases, cpus):
        cmd = ["taskset", "-c", str(cpu), harness, "bench", case]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
    trials, secs = 0, 0.0
    for p in procs:
        out = p.communicate()[0].split()
        if p.returncode != 0 or len(out) < 3:
            raise RuntimeError("ref_harness failed")
        trials += int(out[0])
        secs = max(secs, float(out[1]))
    return trials, secs


_PORT_WORKER = r"""
import ctypes as C, os, sys, time
import numpy as np
root, seed = sys.argv[1], int(sys.argv[2])
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import bench
from oracle_lib import Oracle
spec, uR, uI, R = bench.golden_spec()
o = Oracle(spec)
t0 = time.perf_counter()
r = o.sample_walker(R, uR, uI, 0.0, seed, seed, 0, bench.MC_NINIT, bench.MC_NSTEPS, bench.MC_NTHERMSTEPS, bench.MC_STEP)
print(bench.STEPS_PER_WALKER, time.perf_counter() - t0, bench.MC_NSTEPS)
"""


def port_step(cpus):
    """Fallback when oracle/_ref was not built (no /root/reference on the build host): the plain-C restatement of the
    reference's algorithm (oracle/tdvmc_oracle.c), one walker's pass per physical core."""
    procs = [subprocess.Popen(["taskset", "-c", str(cpu), sys.executable, "-c", _PORT_WORKER, ROOT, str(i + 1)],
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for i, cpu in enumerate(cpus)]
    trials, secs = 0, 0.0
    for p in procs:
        out = p.communicate()[0].split()
        if p.returncode != 0 or len(out) < 3:
            raise RuntimeError("oracle port failed (build it: make -C oracle)")
        trials += int(out[0])
        secs = max(secs, float(out[1]))
    return trials, secs



def reference_time_steps(uR, uI, R, cpus, samples_target):
    """BASELINE's second metric on the host: `duration for full timestep` (src/TDVMC.cpp:3919-3923) of the unmodified reference
    PROGRAM (oracle/_ref/TDVMC_ref: config file, Euler step with the config's Eigen QR solve, the seven per-step file appends),
    one single-rank process per physical core as in BASELINE.md 3.  Two bounded runs of three time steps each - MC_NSTEPS = 2
    (the config's own count) and MC_NSTEPS = 6 - separate the per-sample cost from the fixed cost of a step, which gives the
    duration of a step whose TOTAL sample count matches the device's (samples = processes x MC_NSTEPS)."""
    from concurrent.futures import ThreadPoolExecutor
    from tdvmc_b200 import driver
    ref = os.path.join(ROOT, "oracle", "_ref", "TDVMC_ref")
    if not os.path.exists(ref):
        return None
    dur = {}
    with tempfile.TemporaryDirectory() as td:
        for ns in (2, 6):
            cfg = driver.headline_config(uR, uI, MC_NSTEPS=ns, TIMESTEP=1e-7, TOTALTIME=2.5e-7, MC_VERY_FIRST_NINITIALIZATIONSTEPS=1000)

            def one(i):
                env = dict(os.environ)
                r = driver.run_driver(ref, cfg, os.path.join(td, f"ts_{ns}_{i}"), R0=R, seed=i + 1, timeout=600,
                                      prefix=["taskset", "-c", str(cpus[i])], env=env)
                return float(np.mean(r.step_ms))

            with ThreadPoolExecutor(max_workers=len(cpus)) as ex:
                dur[ns] = float(np.mean(list(ex.map(one, range(len(cpus))))))
    per_sample = (dur[6] - dur[2]) / 4.0
    fixed = dur[2] - 2.0 * per_sample
    per_core = samples_target / float(len(cpus))
    matched_ms = fixed + per_core * per_sample
    return {"cores": len(cpus), "ms_per_time_step_config_counts": dur[2], "samples_per_time_step_config_counts": 2 * len(cpus),
            "time_steps_per_s_config_counts": 1e3 / dur[2],
            "ms_per_sample_per_core": per_sample, "ms_fixed_per_time_step": fixed,
            "samples_per_time_step_matched": samples_target, "ms_per_time_step_sample_matched": matched_ms,
            "time_steps_per_s_sample_matched": 1e3 / matched_ms,
            "how": "TDVMC_ref, one pinned single-rank process per physical core, 3 time steps at MC_NSTEPS = 2 and at 6 "
                   "(MC_NTHERMSTEPS = 5000, MC_NINITIALIZATIONSTEPS = 1000, Euler, LINEAR_EQUATION_SOLVER_TYPE = 1, file appends): "
                   "step duration = fixed + samples per core x per-sample cost, evaluated at the device arm's samples per time step"}

REFERENCE_KIND = "reference"


def run_reference(steps, warmup, tmpdir, spec, uR, uI, R):
    global REFERENCE_KIND
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    cpus = physical_cores()
    if not os.path.exists(harness):
        REFERENCE_KIND = "port"
        for _ in range(warmup):
            port_step(cpus)
        trials, secs = 0, 0.0
        for _ in range(steps):
            t, sec = port_step(cpus)
            trials += t
            secs += sec
        return trials, secs, len(cpus)
    cases = []
    for i, _ in enumerate(cpus):
        path = os.path.join(tmpdir, f"case_{i}.txt")
        write_reference_case(path, spec, uR, uI, R, seed=i + 1)   # rank-seeded like src/TDVMC.cpp:524
        cases.append(path)
    for _ in range(warmup):
        reference_step(cases, cpus)
    trials, secs = 0, 0.0
    for _ in range(steps):
        t, s = reference_step(cases, cpus)
        trials += t
        secs += s
    return trials, secs, len(cpus)


def reference_main(args, rank):
    if rank != 0:
        return
    spec, uR, uI, R = golden_spec()
    with tempfile.TemporaryDirectory() as td:
        trials, secs, cores = run_reference(args.steps, args.warmup, td, spec, uR, uI, R)
    value = trials / secs
    ts = None
    try:
        ts = reference_time_steps(uR, uI, R, physical_cores(), SURVEY_WALKERS_PER_GPU * MC_NSTEPS * max(args.gpus, 1))
    except Exception as ex:  # a report, never the product
        ts = {"failed": str(ex)[-300:]}
    what = "the unmodified reference" if REFERENCE_KIND == "reference" else "the plain-C port of the reference (oracle/tdvmc_oracle.c)"
    sample = (f"{cores} single-rank processes of {what} (one per physical core, taskset-pinned, "
              f"serial MPI shim), each one walker's pass per step: {MC_NINIT} initialization proposals run BEFORE the clock starts, "
              f"then {MC_NSTEPS * MC_NTHERMSTEPS} timed and counted proposals + {MC_NSTEPS} evaluations with the reference's "
              f"estimator accumulation (ref_harness.cpp ModeMC); the rate is proposals counted / seconds timed")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "walker-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config({"walkers": cores}),
            "time_steps_per_s": (ts or {}).get("time_steps_per_s_sample_matched"),
            "time_step": ts,
            "cpu_baseline": {"value": value, "unit": "walker-steps/s", "cores": cores, "kind": REFERENCE_KIND, "sample": sample},
            "e2e": {"value": value, "unit": "walker-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region.  NVML in-process (nvidia_ml_py), one sample every 20 ms plus
    one at start and one at finish, so that even a half-second region on a busy 8-GPU box has samples (a freshly spawned
    `nvidia-smi -lms` needs longer than that to print its first line; it stays as the fallback)."""
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()
        self.sm, self.mx, self.reasons, self.proc, self.nvml = [], 0.0, set(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample(self):
        try:
            n = self.nvml
            self.sm.append(float(n.nvmlDeviceGetClockInfo(self.dev, n.NVML_CLOCK_SM)))
            try:
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.dev))
            except Exception:
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev))
            for name, bit in self.REASON_BITS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        if self.nvml is not None:
            self._sample()
            while not self.stop_flag.wait(0.02):
                self._sample()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass

    def finish(self):
        if self.nvml is not None:
            self._sample()
        self.stop_flag.set()
        try:
            if self.proc is not None:
                self.proc.terminate()
        except Exception:
            pass
        sm, mx, reasons = list(self.sm), self.mx, set(self.reasons)
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active") and "not" not in v.lower():
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--walkers-per-gpu", type=int, default=0, help="0: 4096 (SURVEY.md 8d)")
    ap.add_argument("--total-walkers", type=int, default=0,
                    help="fixed ensemble split over the GPUs (strong scaling); default: fixed walkers per GPU (weak)")
    ap.add_argument("--no-exhibits", action="store_true", help="skip the K3/K4/K5 roofline exhibits and the CPU baseline")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_main(args, rank)
        return

    import torch
    import torch.distributed as dist
    from tdvmc_b200 import capi

    if not torch.cuda.is_available() or capi.load().tdvmc_gpu_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    spec, uR, uI, R_seed = golden_spec()

    # ensemble size: SURVEY.md 8(d)'s W = 4096 walkers per GPU (weak scaling: fixed walkers per GPU).  That is 1.38 waves of
    # the one-warp-per-walker sweep kernel (20 resident walkers x 148 SMs = 2960); the library time-shares the resident warps
    # in that case (sweep_queue_kernel).  The exact-wave ensemble of round 1 is reported beside it as `full_wave`.
    probe = capi.Handle(spec, 1, device=local_rank)
    per_sm, sms = probe.resident_walkers()
    probe.close()
    W = args.walkers_per_gpu or SURVEY_WALKERS_PER_GPU
    scaling = "weak"
    if args.total_walkers > 0:
        W = max(1, args.total_walkers // world)
        scaling = "strong"
    first = rank * W
    h = capi.Handle(spec, W, seed=1, mc_step=MC_STEP, first_walker=first, max_samples=MC_NSTEPS, device=local_rank)
    if world > 1:
        uid = torch.zeros(capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        h.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)

    def join_comm(hx):
        if world > 1:
            uidx = torch.zeros(capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
            if rank == 0:
                uidx.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(uidx, 0)
            hx.comm_init(bytes(uidx.cpu().numpy().tobytes()), rank, world)

    # synthetic ensemble: jittered 7^3 lattices (src/TDVMC.cpp:727-739), equilibrated by 100 sweeps
    rng = np.random.default_rng(1000 + rank)
    R_host = torch.empty((W, N, 3), dtype=torch.float64).pin_memory()
    R_np = R_host.numpy()
    base = systems.jittered_lattice(N, LBOX, rng)
    R_np[:] = base[None] + rng.uniform(-0.02, 0.02, (W, N, 3))
    h.set_params(uR, uI, 0.0, 0.0, 0.0)
    h.set_positions_raw(R_host.data_ptr(), 0, W)
    h.sweep(100 * N)
    h.wrap_positions()
    h.synchronize()

    out = None

    def step(e2e):
        nonlocal out
        if e2e:
            h.set_positions_raw(R_host.data_ptr(), 0, W)                      # H2D from pinned memory
        h.set_params(uR, uI, 0.0, 0.0, 0.0)                                   # BroadcastNewParameters
        h.sample_and_accumulate(MC_NSTEPS, MC_NTHERMSTEPS, MC_NINIT)          # UpdateExpectationValues
        out = h.allreduce_and_fetch(out)                                      # ReduceToAverage x7 (D2H)
        if e2e:
            h.get_positions(0, W, out=R_np)                                   # D2H into pinned memory
        h.flush_l2()                                                          # next step starts with a cold L2

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()
        h.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(k, e2e):
        barrier()
        h.timer_start()
        for _ in range(k):
            step(e2e)
        ms = h.timer_stop()
        barrier()
        return max_over_ranks(ms)

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    h.profile(True, True)
    launches0 = h.launch_count()
    ms_total = timed(args.steps, False)
    launches = h.launch_count() - launches0
    stats = h.kernel_stats()
    h.profile(False, False)
    clocks = sampler.finish()
    step(True)
    ms_e2e = timed(args.steps, True)

    # BASELINE's second metric, TDVMC time-steps/s: a whole time step of the outer loop (src/TDVMC.cpp:3419-3923) as the
    # driver would run it against this library - parameters in, estimator pass, packed all-reduce, the parameter solve
    # (BuildSystemOfEquations + scaling preconditioner + Cholesky) and the Euler update fed back into the next step.
    # (a) solve on the device (tdvmc_gpu_euler_step, solve.cu): 2P + 5 doubles leave the GPU per step;
    # (b) estimators fetched (328 KB) and solved on the host with LAPACK (numpy.linalg, tdvmc_b200/timestep.py).
    # Every rank solves redundantly in both.
    from tdvmc_b200 import timestep
    cur = {"uR": uR.copy(), "uI": uI.copy(), "phiR": 0.0, "phiI": 0.0}

    def time_step_device():
        h.set_params(cur["uR"], cur["uI"], cur["phiR"], cur["phiI"], 0.0)
        h.sample_and_accumulate(MC_NSTEPS, MC_NTHERMSTEPS, MC_NINIT)
        cur["uR"], cur["uI"], cur["phiR"], cur["phiI"], _ = h.euler_step(1e-7, cur["uR"], cur["uI"], cur["phiR"], cur["phiI"],
                                                                         imaginary_time=0, min_scaling=1e-12)
        h.flush_l2()

    def time_step_host():
        nonlocal out
        h.set_params(cur["uR"], cur["uI"], cur["phiR"], cur["phiI"], 0.0)
        h.sample_and_accumulate(MC_NSTEPS, MC_NTHERMSTEPS, MC_NINIT)
        out = h.allreduce_and_fetch(out)
        est = dict(localOperators=out["O"], localOperatorsMatrix=out["S"], localOperatorlocalEnergyR=out["OER"],
                   localOperatorlocalEnergyI=out["OEI"], localEnergyR=float(out["e_r"][0]), localEnergyI=float(out["e_i"][0]))
        cur["uR"], cur["uI"], cur["phiR"], cur["phiI"] = timestep.euler_step(1e-7, cur["uR"], cur["uI"], cur["phiR"], cur["phiI"],
                                                                             est, imaginary_time=0, lapack=True, min_scaling=1e-12)
        h.flush_l2()

    def timed_time_steps(fn):
        fn()
        barrier()
        h.timer_start()
        for _ in range(args.steps):
            fn()
        ms = max_over_ranks(h.timer_stop())
        barrier()
        return ms

    ms_ts_host = timed_time_steps(time_step_host)
    h.profile(True, True)
    ms_ts = timed_time_steps(time_step_device)
    n_solve, ms_solve = h.kernel_stats()["solve"]
    h.profile(True, True)
    for _ in range(3):
        h.solve_parameters_dot(imaginary_time=0, use_preconditioning=False, solver_type=1)   # solve_qr_kernel on the same estimators
    n_qr, ms_qr = h.kernel_stats()["solve"]
    h.profile(False, False)

    # ---- the same pass at other ensemble sizes (one JSON line carries them all) ----
    def side_run(Wx, k):
        """walker-steps/s of k passes with Wx walkers per GPU (own handle, own communicator, same protocol as the main run)."""
        hx = capi.Handle(spec, Wx, seed=1, mc_step=MC_STEP, first_walker=rank * Wx, max_samples=MC_NSTEPS, device=local_rank)
        join_comm(hx)
        rx = np.random.default_rng(2000 + rank)
        Rx = systems.jittered_lattice(N, LBOX, rx)[None] + rx.uniform(-0.02, 0.02, (Wx, N, 3))
        hx.set_params(uR, uI, 0.0, 0.0, 0.0)
        hx.set_positions(Rx)
        hx.sweep(100 * N)
        hx.wrap_positions()
        ox = None

        def stepx():
            nonlocal ox
            hx.set_params(uR, uI, 0.0, 0.0, 0.0)
            hx.sample_and_accumulate(MC_NSTEPS, MC_NTHERMSTEPS, MC_NINIT)
            ox = hx.allreduce_and_fetch(ox)
            hx.flush_l2()

        stepx()
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        hx.synchronize()
        hx.timer_start()
        for _ in range(k):
            stepx()
        msx = max_over_ranks(hx.timer_stop())
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        hx.close()
        return {"walkers_per_gpu": Wx, "walkers": Wx * world, "steps": k, "ms_per_step": msx / k,
                "value": float(Wx) * world * STEPS_PER_WALKER * k / (msx * 1e-3), "unit": "walker-steps/s"}

    side = {}
    if not args.total_walkers and not args.walkers_per_gpu:
        k_side = max(3, min(args.steps, 5))
        # the exact-wave ensemble (resident walkers x SMs = 2960 per GPU): round 1's headline configuration
        side["full_wave"] = side_run(per_sm * sms, k_side)
        # strong scaling: the 8-GPU weak-scaling ensemble (8 x 2960 = 23 680 walkers) split over the GPUs of THIS run
        side["strong"] = dict(side_run(23680 // world, k_side), total_walkers=23680,
                              note="fixed ensemble of 23 680 walkers (= the weak-scaling ensemble at 8 GPUs) split over n_gpus")

    # ---- TDVMC time-steps/s through the reference's own driver bound to the library (tdvmc_b200/host/build/TDVMC_gpu) ----
    driver_steps = None
    if world == 1 and rank == 0:
        from tdvmc_b200 import driver
        if os.path.exists(driver.TDVMC_GPU):
            driver_steps = {}
            K = max(args.steps, 20)
            for name, over in (("host_solve_qr", dict(LINEAR_EQUATION_SOLVER_TYPE=1, USE_PRECONDITIONING=0)),
                               ("device_solve_qr", dict(LINEAR_EQUATION_SOLVER_TYPE=1, USE_PRECONDITIONING=0, GPU_DEVICE_SOLVE=1)),
                               ("device_solve_cholesky", dict(LINEAR_EQUATION_SOLVER_TYPE=0, USE_PRECONDITIONING=1, GPU_DEVICE_SOLVE=1))):
                try:
                    with tempfile.TemporaryDirectory() as td:
                        cfg = driver.headline_config(uR, uI, TIMESTEP=1e-7, TOTALTIME=1e-7 * (K + 0.5), GPU_WALKERS=W,
                                                     MC_VERY_FIRST_NINITIALIZATIONSTEPS=34300, **over)
                        t0 = time.perf_counter()
                        run = driver.run_driver(driver.TDVMC_GPU, cfg, td, R0=R_seed, timeout=600)
                        wall = time.perf_counter() - t0
                        e_last = float(run.local_energy_r[-1])   # (the .dat files live in the temporary directory)
                    ms = run.step_ms[1:]                      # the first step pays the lazy CUDA initialisations
                    driver_steps[name] = {"time_steps_per_s": 1e3 / float(np.mean(ms)), "ms_per_time_step": float(np.mean(ms)),
                                          "time_steps": int(len(ms)), "samples_per_time_step": W * MC_NSTEPS,
                                          "process_wall_s": wall, "local_energy_r_last": e_last}
                except Exception as ex:
                    driver_steps[name] = {"failed": str(ex)[-300:]}
            driver_steps["what"] = ("the reference's own src/TDVMC.cpp (15 call sites re-pointed, tdvmc_b200/host/driver) run as a program: "
                                    "config file in, per time step BroadcastNewParameters, the estimator pass on the device "
                                    "(GPU_WALKERS walkers x MC_NSTEPS samples), fetch, SolveForParametersDot + Euler (host Eigen QR as the "
                                    "config says, or solve_kernel), AcceptNewParams, the seven AppendDataToFile calls; "
                                    "'duration for full timestep' as the driver prints it (1 ms resolution, mean over the steps)")

    proposals = float(W) * world * STEPS_PER_WALKER * args.steps
    value = proposals / (ms_total * 1e-3)
    e2e_value = proposals / (ms_e2e * 1e-3)
    est_bytes = 8 * (N_PARAM * N_PARAM + 3 * N_PARAM + 2 + spec.n_other)
    h2d = W * N * 3 * 8 + 2 * N_PARAM * 8 + 16
    d2h = W * N * 3 * 8 + est_bytes

    # ---- roofline of the dominant kernel (K1 sweep): FP64 pipe, denominators measured live ----
    dfma_peak, dmma_peak = h.measure_fp64_peak()
    n_sweep, ms_sweep = stats["sweep"]
    sweep_flops = FLOP_PER_WALKER_STEP * float(W) * STEPS_PER_WALKER * args.steps
    sweep_tf = sweep_flops / (ms_sweep * 1e-3) / 1e12
    traffic, traffic_src = sweep_traffic_from_ncu()
    if W != SURVEY_WALKERS_PER_GPU:
        traffic, traffic_src = None, "the committed capture is for 4096 walkers per launch"
    roofline = {"kernel": "sweep_queue_kernel (K1, time-shared launch)" if W != per_sm * sms else "sweep_kernel (K1)", "bound": "fp64", "achieved": sweep_tf, "peak": dfma_peak, "unit": "TFLOP/s",
                "frac": sweep_tf / dfma_peak, "traffic": traffic, "traffic_source": traffic_src,
                "note": "FP64 DFMA-pipe bound (SURVEY 8d): 19 836 algorithmic flop per walker-step; peak = DFMA microbenchmark "
                        "measured in this run (MEASURED_PEAKS.json has no FP64 entry); algorithmic HBM traffic is 16.5 KB per walker per launch",
                "launches": n_sweep, "avg_launch_ms": ms_sweep / max(n_sweep, 1), "share_of_step": ms_sweep / ms_total,
                "sweep_only_walker_steps_per_s": float(W) * STEPS_PER_WALKER * args.steps / (ms_sweep * 1e-3)}   # SURVEY 8(d) metric 1 (i)
    n_ev, ms_ev = stats["evaluate"]
    ev_tf = FLOP_PER_EVALUATION * float(W) * MC_NSTEPS * args.steps / (ms_ev * 1e-3) / 1e12
    kernels = [roofline,
               {"kernel": "evaluate_kernel (K2+K3+K4 fused)", "bound": "fp64", "achieved": ev_tf, "peak": dfma_peak,
                "unit": "TFLOP/s", "frac": ev_tf / dfma_peak, "launches": n_ev, "avg_launch_ms": ms_ev / max(n_ev, 1),
                "share_of_step": ms_ev / ms_total, "samples_per_s": float(W) * MC_NSTEPS * args.steps / (ms_ev * 1e-3)},
               {"kernel": "syrk_kernel (K5) inside the step", "launches": stats["accumulate"][0],
                "avg_launch_ms": stats["accumulate"][1] / max(stats["accumulate"][0], 1),
                "share_of_step": stats["accumulate"][1] / ms_total}]

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"

    cpu_baseline = None
    secondary, ref_ts = [], None
    if not args.no_exhibits and rank == 0:
        # K3 / K4 exhibits: reference table semantics on resident walkers (HBM bound)
        nt = min(W, 1024)
        h.tables_resident(nt)
        h.contract_resident(nt, fetch=False)
        h.flush_l2()
        h.profile(True, True)
        for _ in range(3):
            h.tables_resident(nt)
            h.flush_l2()
            h.contract_resident(nt, fetch=False)
            h.flush_l2()
        st = h.kernel_stats()
        h.profile(False, False)
        t_tab = st["tables"][1] / st["tables"][0] * 1e-3
        t_con = st["contract"][1] / st["contract"][0] * 1e-3
        kernels.append({"kernel": "tables_kernel (K3, reference table form)", "bound": "hbm",
                        "achieved": nt * BYTES_PER_TABLE / t_tab / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": nt * BYTES_PER_TABLE / t_tab / 1e9 / hbm_peak, "peak_source": hbm_src,
                        "samples_per_s": nt / t_tab, "configs_per_launch": nt})
        con_bytes = 8 * 203 * N * 4
        kernels.append({"kernel": "contract_kernel (K4 from tables)", "bound": "hbm", "achieved": nt * con_bytes / t_con / 1e9,
                        "peak": hbm_peak, "unit": "GB/s", "frac": nt * con_bytes / t_con / 1e9 / hbm_peak,
                        "peak_source": hbm_src, "samples_per_s": nt / t_con, "configs_per_launch": nt})
        # K5 exhibit: S/F accumulation on M = 2^17 synthetic samples, against cuBLAS DGEMM and the DMMA probe
        M = 1 << 17
        rngk = np.random.default_rng(2)
        O = rngk.normal(50.0 + np.arange(N_PARAM), 5.0, size=(M, N_PARAM))
        er, ei = rngk.normal(size=M), rngk.normal(size=M)
        h.accumulate_fixed(O[:4096], er[:4096], ei[:4096])
        h.profile(True, True)
        h.accumulate_fixed(O, er, ei)
        st = h.kernel_stats()
        h.profile(False, False)
        t_acc = st["accumulate"][1] / st["accumulate"][0] * 1e-3
        a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        torch.matmul(a, a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            torch.matmul(a, a)
        e1.record()
        torch.cuda.synchronize()
        dgemm_tf = 3 * 2 * 4096 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        syrk_flops = float(M) * (N_PARAM + 3) * (N_PARAM + 4)
        kernels.append({"kernel": "syrk_kernel (K5, M=2^17 synthetic samples)", "bound": "tensor",
                        "achieved": syrk_flops / t_acc / 1e12, "peak": dgemm_tf, "unit": "TFLOP/s",
                        "frac": syrk_flops / t_acc / 1e12 / dgemm_tf, "peak_source": "cuBLAS DGEMM 4096^3 measured in this run",
                        "dmma_probe_tflops": dmma_peak, "flop_convention": "symmetric: M (P+3)(P+4)", "ms": t_acc * 1e3})
        # the same kernel at M = 2^16 and 2^20 (the sample counts of config 4's S-matrix)
        for logm in (16, 20):
            Mx = 1 << logm
            reps = Mx // M if Mx > M else 1
            Ox = np.tile(O, (reps, 1))[:Mx] if Mx > M else O[:Mx]
            erx, eix = np.resize(er, Mx), np.resize(ei, Mx)
            h.profile(True, True)
            h.accumulate_fixed(Ox, erx, eix)
            stx = h.kernel_stats()
            h.profile(False, False)
            tx = stx["accumulate"][1] / stx["accumulate"][0] * 1e-3
            fl = float(Mx) * (N_PARAM + 3) * (N_PARAM + 4)
            kernels.append({"kernel": f"syrk_kernel (K5, M=2^{logm})", "bound": "tensor", "achieved": fl / tx / 1e12, "peak": dgemm_tf,
                            "unit": "TFLOP/s", "frac": fl / tx / 1e12 / dgemm_tf, "ms": tx * 1e3})
            del Ox
        # secondary workloads (BASELINE configs 1, 2, 4, 5): device pass with the config's own step : evaluation ratio and the
        # unmodified reference on EVERY physical core at once; config 4's reference pass is bounded to 2 samples
        secondary = []
        if world == 1:
            sys.path.insert(0, os.path.join(ROOT, "profiles"))
            try:
                import bench_configs
                cpus_all = physical_cores()
                for idx in (1, 2, 4, 5):
                    try:
                        secondary.append(bench_configs.run_config(idx, 3, cpus=cpus_all, dfma_tflops=dfma_peak,
                                                                  ref_samples=2 if idx == 4 else bench_configs.SAMPLES))
                    except Exception as ex:
                        secondary.append({"config": bench_configs.CONFIGS[idx][0], "failed": str(ex)[-300:]})
                try:   # config 3 exactly as shipped (N = 8000): BASELINE scales it to N = 343 for the headline
                    secondary.append(bench_configs.run_config3_as_shipped(1, cpus=cpus_all, dfma_tflops=dfma_peak))
                except Exception as ex:
                    secondary.append({"config": "config/BosonsBulk3D.config as shipped (N = 8000)", "failed": str(ex)[-300:]})
            except Exception as ex:
                secondary.append({"failed": str(ex)[-300:]})
        # CPU baseline: the unmodified reference on this box's host cores, one pass per core
        if world == 1:
            try:
                with tempfile.TemporaryDirectory() as td:
                    trials, secs, cores = run_reference(3, 0, td, spec, uR, uI, R_seed)
                what = "the unmodified reference" if REFERENCE_KIND == "reference" else "the plain-C port of the reference"
                cpu_baseline = {"value": trials / secs, "unit": "walker-steps/s", "cores": cores, "kind": REFERENCE_KIND,
                                "sample": f"{cores} pinned single-rank processes of {what}, three passes of one walker each "
                                          f"({MC_NSTEPS * MC_NTHERMSTEPS} timed proposals + {MC_NSTEPS} evaluations per pass, the {MC_NINIT} "
                                          f"initialization proposals before the clock), {secs:.2f} s wall, "
                                          f"{secs * cores:.0f} core-seconds"}
            except Exception as ex:  # the baseline is a report, never the product
                cpu_baseline = {"value": None, "unit": "walker-steps/s", "cores": 0, "kind": REFERENCE_KIND, "sample": f"failed: {ex}"}
            try:
                ref_ts = reference_time_steps(uR, uI, R_seed, physical_cores(), W * MC_NSTEPS)
            except Exception as ex:
                ref_ts = {"failed": str(ex)[-300:]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "walker-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config({"walkers_per_gpu": W, "walkers": W * world, "parallelism": f"walkers x{world}",
                                           "l2": "flushed between steps (256 MiB memset); walker state is 34 MB per GPU"}),
                "time_steps_per_s": args.steps / (ms_ts * 1e-3),
                "full_wave": side.get("full_wave"), "strong": side.get("strong"), "secondary": secondary,
                "time_step": {"ms": ms_ts / args.steps, "samples_per_time_step": W * world * MC_NSTEPS,
                              "driver_binary": driver_steps, "reference_host": ref_ts,
                              "includes": "set_params, estimator pass, all-reduce, Cholesky solve of S u' = F on the device "
                                          "(P = 201, solve_kernel), Euler update, parameter feedback; every rank solves redundantly",
                              "solve_kernel_ms": ms_solve / max(n_solve, 1), "solve_qr_kernel_ms": ms_qr / max(n_qr, 1),
                              "with_host_lapack_solve": {"ms": ms_ts_host / args.steps,
                                                         "time_steps_per_s": args.steps / (ms_ts_host * 1e-3)}},
                "e2e": {"value": e2e_value, "unit": "walker-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels,
                "fp64_peaks": {"dfma_tflops": dfma_peak, "dmma_tflops": dmma_peak},
                "cpu_baseline": cpu_baseline,
                "estimators": {"local_energy_r": float(out["e_r"][0]), "acceptance": out["n_acceptances"] / max(out["n_trials"], 1)}}
        print(json.dumps(line))
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
