#!/usr/bin/env python3
"""Throughput of the BASELINE configs that are parity cases rather than the headline (configs 1, 2, 4, 5),
through the same C ABI as bench.py.  One JSON line per config: walker-steps/s and samples/s of a
ParallelUpdateExpectationValues pass with the config's own MC_NTHERMSTEPS : 1 evaluation ratio, the share of the sweep and
evaluation kernels, and the unmodified reference timed on ONE host core - or, with --full-host, on every physical core at
once - with the same counts (oracle/_ref/ref_harness).  bench.py imports run_config for its `secondary` list.

    python profiles/bench_configs.py [--configs 1,2,4,5] [--reps 3] > profiles/rNN_configs.jsonl

Not a bench.py line: the headline metric is quoted on config 3 only.  The ensemble sizes follow SURVEY.md 8(d)
(W = 2^20 / N for configs 1, 2, 5; W = 512 for config 4), samples per walker per pass are capped at 8."""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tdvmc_b200 import capi, systems  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")

# config file -> (fixture, system, MC_STEP, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS, walkers)
CONFIGS = {
    1: ("config/drop_6.config", "hedrop_n6_equil", "HeDrop", 0.5, 50, 1000, (1 << 20) // 6),
    2: ("config/bulk_64.config", "hebulk_n64_equil", "HeBulk", 0.3, 100, 100, (1 << 20) // 64),
    4: ("config/NUBosonsBulkPB3D.config", "nubosonsbulkpb_n1728_equil", "NUBosonsBulkPB", 0.5, 200, 400, 512),
    5: ("config/He4He4Na.config", "mixture_he4he4na_equil", "BosonMixtureCluster", 4.0, 20, 1000, (1 << 20) // 3),
    # SURVEY 8(f) rank 4 (not a BASELINE config): the radial + box spline system at its shipped size
    6: ("config/NUBosonsBulkPBBoxAndRadial3D.config", "boxradial_n27_equil", "NUBosonsBulkPBBoxAndRadial", 0.5, 125, 250, (1 << 20) // 27),
    7: ("config/He4He4Na_4thOrder.config", "mixture4_he4he4na_equil", "BosonMixtureCluster_4thorder", 4.0, 20, 1000, (1 << 20) // 3),
    8: ("config/InhContactBosons.config", "inhcontact_n3_equil", "InhContactBosons", 0.5, 100, 1000, (1 << 20) // 3),
}
SAMPLES = 8


def reference_host(g, system, mc_step, n_therm, n_init, n_samples, cpus=None):
    """The unmodified reference with the same counts: one walker's pass on one core (cpus None), or one pass on EVERY core of
    `cpus` at the same time (pinned single-rank processes, rank-seeded) - the full-host figure of BASELINE.md 3.5.
    Returns (proposals/s, samples/s, cores) or None."""
    if not os.path.exists(HARNESS):
        return None
    dim = int(g["DIM"]) if "DIM" in g.files else 3
    scal = dict(N=int(g["N"]), DIM=dim, LBOX=float(g["LBOX"]), N_PARAM=int(g["N_PARAM"]), MC_STEP=mc_step, MC_NSTEPS=n_samples,
                MC_NTHERMSTEPS=n_therm, MC_NINITIALIZATIONSTEPS=n_init, seed=1, phiR=float(g["phiR"]), phiI=float(g["phiI"]))
    arrays = dict(R=g["R"][:, :dim], uR=g["uR"], uI=g["uI"])
    if system == "InhContactBosons":
        scal.update(GR_BIN_COUNT=50, RHO_BIN_COUNT=50)
    for key in ("SYSTEM_PARAMS", "NURBS_GRID", "PARTICLE_TYPES"):
        if key in g.files and np.size(g[key]):
            arrays[key] = g[key]
    if "NURBS_GRID" in arrays:
        scal["USE_NURBS"] = 1
        scal["GR_BIN_COUNT"] = 400 if system.startswith("BosonMixtureCluster") else len(g["other_expectation_values"]) - (
            3 if system == "NUBosonsBulkPBBoxAndRadial" else 9)
    with tempfile.TemporaryDirectory() as td:
        procs = []
        for rank, cpu in enumerate(cpus or [None]):
            case = os.path.join(td, f"case_{rank}.txt")
            with open(case, "w") as f:
                f.write(f"system {system}\nconfigdir {ROOT}/oracle/_ref/config/\n")
                for k, v in dict(scal, seed=rank + 1).items():
                    f.write(f"{k} {v!r}\n")
                for k, v in arrays.items():
                    f.write(k + " " + " ".join(repr(float(x)) for x in np.asarray(v).ravel()) + "\n")
            cmd = ([] if cpu is None else ["taskset", "-c", str(cpu)]) + [HARNESS, "bench", case]
            procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, cwd=td))
        trials = samples = 0.0
        secs = 0.0
        for pr in procs:
            o = pr.communicate()[0].split()
            if pr.returncode != 0 or len(o) < 3:
                return None
            trials += float(o[0])
            samples += float(o[2])
            secs = max(secs, float(o[1]))
        return trials / secs, samples / secs, len(procs)


def run_config(idx, reps, cpus=None, dfma_tflops=None, ref_samples=SAMPLES):
    path, fixture, system, mc_step, n_therm, n_init, W = CONFIGS[idx]
    g = np.load(os.path.join(GOLDEN, fixture + ".npz"))
    spec = systems.from_golden(g)
    N = spec.n_particles
    h = capi.Handle(spec, W, seed=1, mc_step=mc_step, max_samples=SAMPLES)
    h.set_params(g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), float(g["time"]))
    rng = np.random.default_rng(idx)
    R = g["R"][None] + rng.uniform(-0.01, 0.01, (W, N, 3))
    h.set_positions(R)
    h.sweep(20 * N)
    h.synchronize()
    for _ in range(2):
        h.sample_and_accumulate(SAMPLES, n_therm, n_init)
        out = h.allreduce_and_fetch()
    h.profile(True, True)
    h.synchronize()
    h.timer_start()
    for _ in range(reps):
        h.sample_and_accumulate(SAMPLES, n_therm, n_init)
        out = h.allreduce_and_fetch()
    ms = h.timer_stop()
    stats = h.kernel_stats()
    h.profile(False, False)
    per_sm, sms = h.resident_walkers()
    steps = float(W) * (n_init + SAMPLES * n_therm) * reps
    line = {"config": path, "system": system, "N": N, "N_PARAM": spec.n_params, "walkers": W, "MC_STEP": mc_step,
            "MC_NTHERMSTEPS": n_therm, "MC_NINITIALIZATIONSTEPS": n_init, "samples_per_walker_per_pass": SAMPLES,
            "walker_steps_per_s": steps / (ms * 1e-3), "samples_per_s": float(W) * SAMPLES * reps / (ms * 1e-3),
            "ms_per_pass": ms / reps, "sweep_ms": stats["sweep"][1] / reps, "evaluate_ms": stats["evaluate"][1] / reps,
            "accumulate_ms": stats["accumulate"][1] / reps, "resident_walkers_per_sm": per_sm,
            "acceptance": out["n_acceptances"] / out["n_trials"], "local_energy_r": float(out["e_r"][0])}
    h.close()
    # algorithmic work per unit as SURVEY.md 8(d) counts it: 2 (N - 1) x 29 flop per walker-step, 85 flop per pair per sample
    line["sweep_tflops"] = 2.0 * (N - 1) * 29 * steps / (stats["sweep"][1] * 1e-3) / 1e12
    line["evaluate_tflops"] = 85.0 * N * (N - 1) / 2 * float(W) * SAMPLES * reps / (stats["evaluate"][1] * 1e-3) / 1e12
    if dfma_tflops:
        line["sweep_frac_of_dfma_peak"] = line["sweep_tflops"] / dfma_tflops
        line["evaluate_frac_of_dfma_peak"] = line["evaluate_tflops"] / dfma_tflops
    ref = reference_host(g, system, mc_step, n_therm, n_init, ref_samples, cpus)
    if ref:
        key = "reference_full_host" if cpus else "reference_one_core"
        line[key] = {"walker_steps_per_s": ref[0], "samples_per_s": ref[1], "cores": ref[2],
                     "note": "unmodified reference (oracle/_ref/ref_harness bench), one walker's pass per core"
                             f" ({n_init} + {ref_samples} x {n_therm} proposals, {ref_samples} evaluations), pinned single-rank processes"}
        line["walker_steps_ratio_vs_full_host" if cpus else "walker_steps_ratio_vs_one_core"] = line["walker_steps_per_s"] / ref[0]
    return line


def run_config3_as_shipped(reps=1, cpus=None, dfma_tflops=None):
    """config/BosonsBulk3D.config exactly as shipped: N = 8000, LBOX = 20, N_PARAM = 201, all parameters zero, MC_NSTEPS = 2 x
    MC_NTHERMSTEPS = 5000 + 1000 (BASELINE scales it to N = 343 for the headline; round 1 could not run it at all).  One walker per
    SM: positions alone are 192 KB, the sweep gives eight warps to a walker, the evaluation keeps the configuration in a slab of
    global memory.  Reference: the same system through ref_harness on every core, one sample per core."""
    N, L, P = 8000, 20.0, 201
    spec = systems.bosons_bulk(N, L, P, [1.0, 1.0])
    uR = np.zeros(P)
    uI = np.zeros(P)
    probe = capi.Handle(spec, 1)
    _, sms = probe.resident_walkers()
    probe.close()
    W, mc_step, n_therm, n_init, n_samples = sms, 0.5, 5000, 1000, 2
    h = capi.Handle(spec, W, seed=1, mc_step=mc_step, max_samples=n_samples)
    h.set_params(uR, uI, 0.0, 0.0, 0.0)
    rng = np.random.default_rng(3)
    R0 = systems.jittered_lattice(N, L, rng)
    h.set_positions(R0[None] + rng.uniform(-0.01, 0.01, (W, N, 3)))
    h.sample_and_accumulate(n_samples, 64, 64)
    h.allreduce_and_fetch()
    h.profile(True, True)
    h.synchronize()
    h.timer_start()
    for _ in range(reps):
        h.sample_and_accumulate(n_samples, n_therm, n_init)
        out = h.allreduce_and_fetch()
    ms = h.timer_stop()
    stats = h.kernel_stats()
    h.profile(False, False)
    h.close()
    steps = float(W) * (n_init + n_samples * n_therm) * reps
    line = {"config": "config/BosonsBulk3D.config as shipped (N = 8000)", "system": "BosonsBulk", "N": N, "N_PARAM": P, "walkers": W,
            "MC_STEP": mc_step, "MC_NTHERMSTEPS": n_therm, "MC_NINITIALIZATIONSTEPS": n_init, "samples_per_walker_per_pass": n_samples,
            "walker_steps_per_s": steps / (ms * 1e-3), "samples_per_s": float(W) * n_samples * reps / (ms * 1e-3), "ms_per_pass": ms / reps,
            "sweep_ms": stats["sweep"][1] / reps, "evaluate_ms": stats["evaluate"][1] / reps,
            "acceptance": out["n_acceptances"] / out["n_trials"], "local_energy_r": float(out["e_r"][0])}
    line["sweep_tflops"] = 2.0 * (N - 1) * 29 * steps / (stats["sweep"][1] * 1e-3) / 1e12
    line["evaluate_tflops"] = 85.0 * N * (N - 1) / 2 * float(W) * n_samples * reps / (stats["evaluate"][1] * 1e-3) / 1e12
    if dfma_tflops:
        line["sweep_frac_of_dfma_peak"] = line["sweep_tflops"] / dfma_tflops
        line["evaluate_frac_of_dfma_peak"] = line["evaluate_tflops"] / dfma_tflops
    g = dict(N=np.array(N), LBOX=np.array(L), N_PARAM=np.array(P), phiR=np.array(0.0), phiI=np.array(0.0), R=R0, uR=uR, uI=uI,
             SYSTEM_PARAMS=np.array([1.0, 1.0]))

    class G(dict):
        files = list(g)

    ref = reference_host(G(g), "BosonsBulk", mc_step, n_therm, n_init, 1, cpus)
    if ref:
        line["reference_full_host" if cpus else "reference_one_core"] = {
            "walker_steps_per_s": ref[0], "samples_per_s": ref[1], "cores": ref[2],
            "note": "unmodified reference (ref_harness bench), one walker per core, the config's 5000 proposals per evaluation, bounded to ONE sample per core (1000 + 5000 proposals, 1 evaluation)"}
        line["walker_steps_ratio_vs_full_host" if cpus else "walker_steps_ratio_vs_one_core"] = line["walker_steps_per_s"] / ref[0]
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,4,5")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--full-host", action="store_true", help="reference on every physical core instead of one")
    a = ap.parse_args()
    cpus = None
    if a.full_host:
        import bench
        cpus = bench.physical_cores()
    for idx in [int(x) for x in a.configs.split(",")]:
        if idx == 3:
            print(json.dumps(run_config3_as_shipped(1, cpus=cpus)), flush=True)
        else:
            print(json.dumps(run_config(idx, a.reps, cpus=cpus)), flush=True)


if __name__ == "__main__":
    main()
