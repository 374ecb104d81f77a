#!/usr/bin/env python3
"""A/B of the evaluation-kernel variants at the headline size (BosonsBulk N = 343, P = 201), all in one process: the knobs
are read at every launch.  For each setting: outputs of 64 fixed configurations against the plain kernel's (UNIBIN must be
bit-identical) and against the reference fixture, then ms per launch of
4096 configurations.
    python profiles/ab_evaluate_variants.py > gpurun_out/ab_evaluate_variants.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tdvmc_b200 import capi, systems  # noqa: E402

if os.environ.get("TDVMC_LIB"):  # measurement builds (tdvmc_b200/alt/, not shipped)
    capi.LIB_PATH = os.environ["TDVMC_LIB"]

U1 = {"TDVMC_EVAL_UNIBIN": "1"}
SETS = {
    "variants": [
        ("plain", {"TDVMC_EVAL_UNIBIN": "0"}),
        ("unibin", U1),
        ("unibin,u1", {**U1, "TDVMC_EVAL_UCOPIES": "1"}),
        ("plain again", {"TDVMC_EVAL_UNIBIN": "0"}),
    ],
    # knock-outs (TDVMC_LIB=tdvmc_b200/alt/libtdvmc_ko.so): results are wrong by construction, the time is the information
    "knockout": [("ko=%d" % k, {**U1, "TDVMC_EVAL_KO": str(k)}) for k in (0, 1, 2, 4, 16, 1 + 2, 1 + 2 + 4, 1 + 2 + 4 + 16)],
}
VARIANTS = SETS[sys.argv[1] if len(sys.argv) > 1 else "variants"]
KNOBS = ("TDVMC_EVAL_UNIBIN", "TDVMC_EVAL_UCOPIES", "TDVMC_EVAL_KO")


def setenv(kv):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(kv)


def relmax(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


g = np.load(os.path.join(ROOT, "tests", "golden", "bosonsbulk_n343_equil.npz"))
spec = systems.from_golden(g)
W, S, n_therm = 4096, 2, 343
h = capi.Handle(spec, W, seed=1, mc_step=0.5, max_samples=S)
h.set_params(g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), float(g["time"]))
rng = np.random.default_rng(1)
h.set_positions(g["R"][None] + rng.uniform(-0.01, 0.01, (W, spec.n_particles, 3)))
h.sweep(10 * spec.n_particles)
Rfix = np.concatenate([g["R"][None], h.get_positions(0, 63)], axis=0)  # the fixture configuration + 63 equilibrated ones

out = {}
base = None
for name, kv in VARIANTS:
    setenv(kv)
    rec = {}
    try:
        o = h.evaluate_fixed(Rfix)
        if base is None:
            base = o
        rec["vs_plain"] = {k: relmax(o[k], base[k]) for k in ("e_r", "e_i", "O", "ss", "drift_r", "drift_i", "exponent", "other")}
        rec["bit_identical_to_plain"] = all(np.array_equal(o[k], base[k]) for k in base)
        rec["vs_reference"] = {"e_r": abs(o["e_r"][0] / float(g["local_energy_r"]) - 1.0),
                               "e_i": abs(o["e_i"][0] / float(g["local_energy_i"]) - 1.0),
                               "O": relmax(o["O"][0], g["local_operators"]),
                               "ss": relmax(o["ss"][0], g["spline_sums"]),
                               "drift_r": relmax(o["drift_r"][0], g["drift_r"])}
        h.sample_and_accumulate(S, n_therm, 0)  # warm-up of this variant
        h.profile(True, True)
        for _ in range(3):
            h.sample_and_accumulate(S, n_therm, 0)
        st = h.kernel_stats()
        h.profile(False, False)
        n, ms = st["evaluate"]
        rec["launches"] = n
        rec["ms_per_launch_4096"] = ms / n
        rec["samples_per_s"] = n * W / (ms * 1e-3)
    except Exception as e:  # a variant that fails must not hide the others
        rec["error"] = repr(e)
    out[name] = rec
    print(name, json.dumps(rec), file=sys.stderr, flush=True)
h.close()
if len(sys.argv) > 2:  # outputs of the first variant, for comparisons between library builds
    np.savez(os.path.join(ROOT, "gpurun_out", "ab_eval_outputs_%s.npz" % sys.argv[2]), **base)
print(json.dumps(out))
