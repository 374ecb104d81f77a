#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference and the
oracle binaries built by `make -C oracle/ref_build`); the fixtures it writes are committed,
so nothing on the GPU box ever reads /root/reference.

Every number in the fixtures is produced by the reference's own code path
(oracle/_ref/ref_harness calls sys->CalculateWavefunction / CalculateExpectationValues /
CalculateWFQuotient / DoMetropolisStep / VectorDisplacementNIC of mathiasgartner/TDVMC);
this script only chooses the inputs and re-packs the text dumps as .npz.

    python oracle/gen_golden.py            # regenerate everything
"""
import json
import re
import os
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tdvmc_b200 import systems as tsys  # noqa: E402  (host-side set-up only: maps, knots)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("TDVMC_REFERENCE", "/root/reference")
HARNESS = os.path.join(HERE, "_ref", "ref_harness")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def smooth_params(P, rmax, a_r=-0.5, w_r=0.8, a_i=0.05, c_i=1.5, w_i=0.5):
    """Fixed smooth parameter profile: a short-range repulsive real part
    and a small imaginary bump so that kinetic and imaginary estimators are non-trivial."""
    h = rmax / (P - 1)
    k = np.arange(P)
    uR = a_r * np.exp(-((k * h / w_r) ** 2))
    uI = a_i * np.exp(-(((k * h - c_i) / w_i) ** 2))
    return uR, uI


def write_case(path, system, scal, arrays, moves=()):
    with open(path, "w") as f:
        f.write(f"system {system}\n")
        f.write(f"configdir {HERE}/_ref/config/\n")
        for k, v in scal.items():
            f.write(f"{k} {v!r}\n")
        for k, v in arrays.items():
            f.write(k + " " + " ".join(repr(float(x)) for x in np.asarray(v).ravel()) + "\n")
        for m in moves:
            f.write("move " + " ".join(repr(float(x)) for x in m) + "\n")


def parse_dump(path):
    out = {}
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0
    while i + 1 < len(lines):
        head = lines[i].split()
        if not head:
            i += 1
            continue
        name, nd = head[0], int(head[1])
        shape = tuple(int(x) for x in head[2:2 + nd])
        vals = np.array(lines[i + 1].split(), dtype=np.float64)
        out[name] = vals.reshape(shape) if nd else vals.reshape(())
        i += 2
    return out


def run(mode, case_path, out_path=None):
    cmd = [HARNESS, mode, case_path] + ([out_path] if out_path else [])
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=tempfile.gettempdir())
    if r.returncode != 0:
        raise RuntimeError(f"{cmd} failed: {r.stderr}\n{r.stdout}")
    return r.stdout


def read_csv_positions(name):
    with open(os.path.join(REF, "config", name)) as f:
        last = [l for l in f.read().split("\n") if l.strip()][-1]
    return np.array([float(x) for x in last.split(",") if x.strip()])


def jittered_lattice(N, L, seed):
    """Cubic lattice + uniform jitter, the shape of the reference's start-up lattice
    (src/TDVMC.cpp:727-739); exact values do not matter, they are inputs."""
    m = int(round(N ** (1.0 / 3.0)))
    assert m ** 3 == N
    l = L / m
    rng = np.random.default_rng(seed)
    g = (np.arange(m) + 0.5) * l - L / 2
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    R = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    return R + rng.uniform(-0.05, 0.05, R.shape) * l


def contract_drift(d, uR, uI, system):
    """Drift F_n = grad_n ln(psi) from the REFERENCE's sD table (generator-side, long double)."""
    sD = d["sD"].astype(np.longdouble)   # [K][N][D]
    K = sD.shape[0]
    P = len(uR)
    ut_r = np.zeros(K, np.longdouble)
    ut_i = np.zeros(K, np.longdouble)
    rows = bc_rows(d, system, P, K)
    for p, row in enumerate(rows):
        for k, fac in row:
            ut_r[k] += np.longdouble(uR[p]) * np.longdouble(fac)
            ut_i[k] += np.longdouble(uI[p]) * np.longdouble(fac)
    FR = np.einsum("k,kna->na", ut_r, sD).astype(np.float64)
    FI = np.einsum("k,kna->na", ut_i, sD).astype(np.float64)
    return FR, FI


def bc_rows(d, system, P, K):
    """Sparse rows of the boundary-condition map O_p = sum_k M[p][k] ss[k] as the reference applies it
    (BosonsBulk.cpp:158-177 with the dumped bcFactors; NUBosonsBulkPB.cpp:219-232)."""
    rows = []
    if system == "BosonsBulk":
        bs, be = d["bc_start"], d["bc_end"]
        np1 = bs.shape[0]
        np2 = P - be.shape[0]
        for i in range(np1):
            rows.append([(j, bs[i][j]) for j in range(3)])
        for i in range(np1, np2):
            rows.append([(3 + (i - np1), 1.0)])
        for i in range(be.shape[0]):
            rows.append([(K - 3 + j, be[i][j]) for j in range(3)])
    elif system == "NUBosonsBulkPB":
        for i in range(P):
            rows.append([(i + 1, 1.0)])
        rows[1].append((0, 1.0))
        rows[P - 1].append((K - 2, 1.0))
        rows[P - 1].append((K - 1, 1.0))
    else:
        raise ValueError(system)
    return rows


def pad3(a):
    """[N][DIM] -> [N][3], unused coordinates zero: positions and drifts travel three-dimensional over the C ABI."""
    a = np.asarray(a, np.float64)
    out = np.zeros((a.shape[0], 3))
    out[:, :a.shape[1]] = a
    return out


def pack_eval(name, system, scal, arrays, moves, keep_tables="full", subset=(0, 1, 2, 3)):
    with tempfile.TemporaryDirectory() as td:
        cp, op = os.path.join(td, "case.txt"), os.path.join(td, "out.txt")
        write_case(cp, system, scal, arrays, moves)
        run("eval", cp, op)
        d = parse_dump(op)
    out = {
        "system": np.array(system),
        "N": np.array(scal["N"]), "DIM": np.array(scal.get("DIM", 3)), "LBOX": np.array(scal["LBOX"]),
        "N_PARAM": np.array(scal["N_PARAM"]),
        "SYSTEM_PARAMS": np.asarray(arrays["SYSTEM_PARAMS"], np.float64),
        "time": np.array(scal.get("time", 0.0)),
        "R": pad3(np.asarray(arrays["R"], np.float64).reshape(-1, scal.get("DIM", 3))),
        "uR": np.asarray(arrays["uR"], np.float64), "uI": np.asarray(arrays["uI"], np.float64),
        "phiR": np.array(scal.get("phiR", 0.0)), "phiI": np.array(scal.get("phiI", 0.0)),
        "moves": np.asarray(moves, np.float64).reshape(-1, 4),
    }
    if "NURBS_GRID" in arrays:
        out["NURBS_GRID"] = np.asarray(arrays["NURBS_GRID"], np.float64)
    for k in ("exponent", "exponent_wf", "wf", "local_energy_r", "local_energy_i", "local_operators",
              "local_operator_energy_r", "local_operator_energy_i", "other_expectation_values",
              "local_operators_matrix_diag", "local_operators_matrix_row3", "knots", "spline_weights",
              "spline_sums", "outer_sum", "max_distance", "other_local_operators", "move_quotient",
              "move_exponent_new", "bc_start", "bc_end"):
        if k in d:
            out[k] = d[k]
    FR, FI = contract_drift(d, out["uR"], out["uI"], system)
    out["drift_r"], out["drift_i"] = pad3(FR), pad3(FI)
    sD, sD2 = d["sD"], d["sD2"]
    # particle-weighted checksums of the full tables (plain sums cancel pairwise by antisymmetry)
    wn = 1.0 + 0.5 * np.sin(np.arange(sD.shape[1]))
    out["table_checksum_weights"] = wn
    out["sD_checksum"] = np.einsum("n,kna->ka", wn, sD)   # [K][D]
    out["sD2_checksum"] = np.einsum("n,kn->k", wn, sD2)   # [K]
    if keep_tables == "full":
        out["sD"], out["sD2"] = sD, sD2
    else:
        idx = np.array(subset)
        out["table_particles"] = idx
        out["sD_subset"] = sD[:, idx, :]
        out["sD2_subset"] = sD2[:, idx]
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: E_R={float(d['local_energy_r']):.12g} E_I={float(d['local_energy_i']):.12g} "
          f"exponent={float(d['exponent']):.12g} q={d['move_quotient']}")
    return d


def run_mc(system, scal, arrays):
    with tempfile.TemporaryDirectory() as td:
        cp, op = os.path.join(td, "case.txt"), os.path.join(td, "out.txt")
        write_case(cp, system, scal, arrays)
        run("mc", cp, op)
        return parse_dump(op)


def default_moves(R, L, rng, n=6, sigma=0.5):
    moves = []
    for _ in range(n):
        p = int(rng.integers(0, R.shape[0]))
        moves.append([p] + list(R[p] + rng.normal(0, sigma, 3)))
    return moves


def gen_bosonsbulk():
    rng = np.random.default_rng(20261017)
    # (1) reference fixture particleconfiguration_64.csv: L=4, rho=1, small P
    N, L, P = 64, 4.0, 33
    R = read_csv_positions("particleconfiguration_64.csv").reshape(N, 3)
    uR, uI = smooth_params(P, L / 2, w_r=0.6, c_i=1.0, w_i=0.4)
    scal = dict(N=N, LBOX=L, N_PARAM=P, time=0.0, phiR=0.25, phiI=-0.1)
    arr = dict(R=R, uR=uR, uI=uI, SYSTEM_PARAMS=[1.0, 1.0])
    pack_eval("bosonsbulk_n64_fixture", "BosonsBulk", scal, arr, default_moves(R, L, rng))

    # (2) same system after the reference's own sampler has equilibrated it (non-lattice distances),
    #     time-switched potential parameters (BosonsBulk.cpp:237-243)
    mc = run_mc("BosonsBulk", dict(scal, MC_STEP=0.4, MC_NSTEPS=1, MC_NTHERMSTEPS=64 * 200, seed=7), arr)
    R2 = mc["R_final"].reshape(N, 3)
    arr2 = dict(arr, R=R2, SYSTEM_PARAMS=[1.0, 1.0, 0.5, 0.8, 2.5])
    pack_eval("bosonsbulk_n64_equil", "BosonsBulk", dict(scal, time=1.0), arr2, default_moves(R2, L, rng))

    # (3) headline shape: N=343, P=201, natural-box lattice fixture particleconfiguration_343_.csv
    N, P = 343, 201
    R = read_csv_positions("particleconfiguration_343_.csv").reshape(N, 3)
    L = 7.0 * 10.0 ** (1.0 / 3.0)
    uR, uI = smooth_params(P, L / 2, w_r=2.0, c_i=3.5, w_i=1.0)
    scal = dict(N=N, LBOX=L, N_PARAM=P, time=0.0, phiR=0.0, phiI=0.0)
    arr = dict(R=R, uR=uR, uI=uI, SYSTEM_PARAMS=[2.5, 1.0])
    pack_eval("bosonsbulk_n343_lattice", "BosonsBulk", scal, arr, default_moves(R, L, rng),
              keep_tables="subset", subset=(0, 1, 171, 342))

    # (4) headline workload: N=343, L=7 (rho=1), P=201, equilibrated by the reference sampler
    L = 7.0
    R0 = jittered_lattice(N, L, seed=1)
    uR, uI = smooth_params(P, L / 2)
    scal = dict(N=N, LBOX=L, N_PARAM=P, time=0.0, phiR=0.0, phiI=0.0)
    arr = dict(R=R0, uR=uR, uI=uI, SYSTEM_PARAMS=[1.0, 1.0])
    mc = run_mc("BosonsBulk", dict(scal, MC_STEP=0.5, MC_NSTEPS=1, MC_NTHERMSTEPS=343 * 30, seed=3), arr)
    R1 = mc["R_final"].reshape(N, 3)
    pack_eval("bosonsbulk_n343_equil", "BosonsBulk", scal, dict(arr, R=R1), default_moves(R1, L, rng),
              keep_tables="subset", subset=(0, 1, 171, 342))


def gen_bosonsbulk_mc():
    """Statistical golden: the reference sampler at fixed parameters, with its per-sample series."""
    N, L, P = 64, 4.0, 33
    R = read_csv_positions("particleconfiguration_64.csv").reshape(N, 3)
    uR, uI = smooth_params(P, L / 2, w_r=0.6, c_i=1.0, w_i=0.4)
    scal = dict(N=N, LBOX=L, N_PARAM=P, time=0.0, phiR=0.0, phiI=0.0, MC_STEP=0.4,
                MC_NINITIALIZATIONSTEPS=64 * 100, MC_NSTEPS=6000, MC_NTHERMSTEPS=64, seed=11)
    arr = dict(R=R, uR=uR, uI=uI, SYSTEM_PARAMS=[1.0, 1.0])
    d = run_mc("BosonsBulk", scal, arr)
    er = d["energy_r_series"]
    out = dict(N=np.array(N), LBOX=np.array(L), N_PARAM=np.array(P), MC_STEP=np.array(0.4),
               SYSTEM_PARAMS=np.array([1.0, 1.0]), uR=uR, uI=uI, R0=R,
               n_samples=np.array(len(er)), n_therm=np.array(64),
               energy_r_series=er, energy_i_series=d["energy_i_series"],
               local_energy_r=d["local_energy_r"], local_energy_i=d["local_energy_i"],
               local_operators=d["local_operators"], local_operator_energy_r=d["local_operator_energy_r"],
               local_operator_energy_i=d["local_operator_energy_i"],
               local_operators_matrix=d["local_operators_matrix"],
               other_expectation_values=d["other_expectation_values"],
               acceptance=np.array(float(d["n_acceptances"]) / float(d["n_trials"])))
    np.savez_compressed(os.path.join(GOLDEN, "bosonsbulk_n64_mc.npz"), **out)
    print(f"bosonsbulk_n64_mc: <E_R>={float(d['local_energy_r']):.8g} acc={out['acceptance']:.4f}")


def gen_bosonsbulk_mc_headline():
    """The same at the headline size: N = 343, L = 7, N_PARAM = 201, the parameters bench.py uses, MC_STEP = 0.5; the
    reference's single chain, 1500 samples a sweep apart.  (O_k, S, F are left out: 40 401 numbers that the N = 64
    fixture already pins; the energies, the acceptance and a few operators are what a sampler can get wrong.)"""
    g = np.load(os.path.join(GOLDEN, "bosonsbulk_n343_equil.npz"))
    N, L, P = 343, 7.0, 201
    uR, uI = tsys.smooth_params(P, L / 2)
    scal = dict(N=N, LBOX=L, N_PARAM=P, time=0.0, phiR=0.0, phiI=0.0, MC_STEP=0.5,
                MC_NINITIALIZATIONSTEPS=343 * 60, MC_NSTEPS=1500, MC_NTHERMSTEPS=343, seed=12)
    arr = dict(R=g["R"], uR=uR, uI=uI, SYSTEM_PARAMS=[1.0, 1.0])
    d = run_mc("BosonsBulk", scal, arr)
    er = d["energy_r_series"]
    out = dict(N=np.array(N), LBOX=np.array(L), N_PARAM=np.array(P), MC_STEP=np.array(0.5), SYSTEM_PARAMS=np.array([1.0, 1.0]),
               uR=uR, uI=uI, n_samples=np.array(len(er)), n_therm=np.array(343), n_init=np.array(343 * 60),
               energy_r_series=er, energy_i_series=d["energy_i_series"], local_energy_r=d["local_energy_r"],
               local_energy_i=d["local_energy_i"], local_operators=d["local_operators"],
               other_expectation_values=d["other_expectation_values"],
               acceptance=np.array(float(d["n_acceptances"]) / float(d["n_trials"])))
    np.savez_compressed(os.path.join(GOLDEN, "bosonsbulk_n343_mc.npz"), **out)
    print(f"bosonsbulk_n343_mc: <E_R>={float(d['local_energy_r']):.8g} <E_I>={float(d['local_energy_i']):.8g} acc={out['acceptance']:.4f}")


def gen_nubosonsbulkpb():
    rng = np.random.default_rng(4)
    # reduced copy of config/NUBosonsBulkPB3D.config: rho=1, L=6, non-uniform knot grid on [0, 3]
    N, L, P = 216, 6.0, 40
    x = np.linspace(0.0, 1.0, P + 1)
    grid = 3.0 * (0.35 * x + 0.65 * x ** 2)          # finer near the origin
    grid[-1] = 3.0
    R0 = jittered_lattice(N, L, seed=5)
    uR, uI = smooth_params(P, L / 2, w_r=0.8, c_i=1.2, w_i=0.5)
    scal = dict(N=N, LBOX=L, N_PARAM=P, USE_NURBS=1, time=0.0, phiR=0.1, phiI=0.0, GR_BIN_COUNT=50)
    arr = dict(R=R0, uR=uR, uI=uI, SYSTEM_PARAMS=[0.0, 0.0, 0.0, 0.9, 50.0], NURBS_GRID=grid)
    mc = run_mc("NUBosonsBulkPB", dict(scal, MC_STEP=0.5, MC_NSTEPS=1, MC_NTHERMSTEPS=216 * 30, seed=9), arr)
    R1 = mc["R_final"].reshape(N, 3)
    pack_eval("nubosonsbulkpb_n216_equil", "NUBosonsBulkPB", scal, dict(arr, R=R1),
              default_moves(R1, L, rng), keep_tables="subset", subset=(0, 7, 100, 215))


def gen_nubosonsbulkpb_full():
    """config/NUBosonsBulkPB3D.config at its own size: N = 1728, L = 12, N_PARAM = 200 on the config's NURBS_GRID and
    SYSTEM_PARAMS; the shipped parameters are all zero (ideal-gas start), so a smooth non-zero set is used."""
    rng = np.random.default_rng(1728)
    cfg = json.loads(re.sub(r"(\d)\.(\s*[,\]\}])", r"\g<1>.0\2", open(os.path.join(REF, "config", "NUBosonsBulkPB3D.config")).read()))
    N, P = int(cfg["N"]), int(cfg["N_PARAM"])
    L = (N / float(cfg["RHO"])) ** (1.0 / 3.0)
    L = float(round(L, 9))
    grid = np.array(cfg["NURBS_GRID"], dtype=np.float64)
    R0 = jittered_lattice(N, L, seed=6)
    uR, uI = smooth_params(P, L / 2, w_r=0.8, c_i=1.5, w_i=0.5)
    scal = dict(N=N, LBOX=L, N_PARAM=P, USE_NURBS=1, time=0.0, phiR=0.0, phiI=0.0, GR_BIN_COUNT=int(cfg["GR_BIN_COUNT"]))
    arr = dict(R=R0, uR=uR, uI=uI, SYSTEM_PARAMS=cfg["SYSTEM_PARAMS"], NURBS_GRID=grid)
    mc = run_mc("NUBosonsBulkPB", dict(scal, MC_STEP=float(cfg["MC_STEP"]), MC_NSTEPS=1, MC_NTHERMSTEPS=N * 12, seed=11), arr)
    R1 = mc["R_final"].reshape(N, 3)
    pack_eval("nubosonsbulkpb_n1728_equil", "NUBosonsBulkPB", scal, dict(arr, R=R1),
              default_moves(R1, L, rng), keep_tables="subset", subset=(0, 9, 863, 1727))


def gen_observables():
    """g(r) and S(k) (CalculateAdditionalSystemProperties) of stored configurations, and the reference's own sampled
    means (src/TDVMC.cpp:1332-1388) for a statistical check.  Configurations and parameters come from the evaluation
    fixtures, so the observable fixtures stay small."""
    def one(name, src, system, extra_scal, mc):
        g = np.load(os.path.join(GOLDEN, src + ".npz"))
        scal = dict(N=int(g["N"]), LBOX=float(g["LBOX"]), N_PARAM=int(g["N_PARAM"]), time=float(g["time"]),
                    phiR=float(g["phiR"]), phiI=float(g["phiI"]), **extra_scal, **mc)
        arr = dict(R=g["R"], uR=g["uR"], uI=g["uI"], SYSTEM_PARAMS=g["SYSTEM_PARAMS"])
        if "NURBS_GRID" in g.files:
            arr["NURBS_GRID"] = g["NURBS_GRID"]
        with tempfile.TemporaryDirectory() as td:
            cp, op = os.path.join(td, "case.txt"), os.path.join(td, "out.txt")
            write_case(cp, system, scal, arr)
            run("obs", cp, op)
            d = parse_dump(op)
        out = {"source": np.array(src), "system": np.array(system), **{k: np.asarray(v) for k, v in d.items()}}
        for k, v in mc.items():
            out[k] = np.array(v)
        # pair weight: BosonsBulk.cpp:481 (1/(N-1)*DIM), NUBosonsBulkPB.cpp:611 (1)
        out["gr_weight"] = np.array(1.0 / float(scal["N"] - 1) * 3 if system == "BosonsBulk" else 1.0)
        out["GR_BIN_COUNT"] = np.array(extra_scal["GR_BIN_COUNT"])
        out["N"], out["LBOX"] = np.array(scal["N"]), np.array(scal["LBOX"])
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        print(f"{name}: gr_count={int(d['gr_count'])} shells={len(d['k_shell_sizes'])} "
              f"kvecs={int(np.sum(d['k_shell_sizes']))} sk[:3]={d['sk_fixed'][:3]}")

    one("bosonsbulk_n64_obs", "bosonsbulk_n64_equil", "BosonsBulk", dict(GR_BIN_COUNT=50),
        dict(MC_STEP=0.4, MC_NADDITIONALSTEPS=4000, MC_NADDITIONALTHERMSTEPS=64, MC_NADDITIONALINITIALIZATIONSTEPS=6400, seed=5))
    one("bosonsbulk_n343_obs", "bosonsbulk_n343_equil", "BosonsBulk", dict(GR_BIN_COUNT=100), {})
    one("nubosonsbulkpb_n216_obs", "nubosonsbulkpb_n216_equil", "NUBosonsBulkPB", dict(GR_BIN_COUNT=50, USE_NURBS=1),
        dict(MC_STEP=0.5, MC_NADDITIONALSTEPS=1500, MC_NADDITIONALTHERMSTEPS=216, MC_NADDITIONALINITIALIZATIONSTEPS=4320, seed=6))


def gen_he_observables():
    """HeBulk / HeDrop CalculateAdditionalSystemProperties (HeBulk.cpp:413-448, HeDrop.cpp:655-702): the other expectation
    values followed by S(k) per shell (and r2 for the droplet), on the equilibrated evaluation configurations."""
    for name, system in (("hebulk_n64_equil", "HeBulk"), ("hedrop_n6_equil", "HeDrop")):
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        scal = dict(N=int(g["N"]), LBOX=float(g["LBOX"]), N_PARAM=int(g["N_PARAM"]), phiR=float(g["phiR"]), phiI=float(g["phiI"]))
        arr = dict(R=g["R"], uR=g["uR"], uI=g["uI"])
        with tempfile.TemporaryDirectory() as td:
            cp, op = os.path.join(td, "case.txt"), os.path.join(td, "out.txt")
            write_case(cp, system, scal, arr)
            run("obs", cp, op)
            d = parse_dump(op)
        n_other = int(d["n_other"])
        n_sh = len(d["k_shell_sizes"])
        out = dict(source=np.array(name), system=np.array(system), LBOX=g["LBOX"], N=g["N"], n_other=np.array(n_other),
                   other_fixed=d["additional_fixed"][:n_other], sk_fixed=d["additional_fixed"][n_other:n_other + n_sh],
                   tail_fixed=d["additional_fixed"][n_other + n_sh:], k_shell_sizes=d["k_shell_sizes"], k_vectors=d["k_vectors"])
        np.savez_compressed(os.path.join(GOLDEN, name.replace("_equil", "_obs") + ".npz"), **out)
        print(f"{name}: n_other={n_other} shells={n_sh} kvecs={int(np.sum(d['k_shell_sizes']))} tail={out['tail_fixed']}")


def gen_mixture_observables():
    """BosonMixtureCluster::CalculateAdditionalSystemProperties (BosonMixtureCluster.cpp:680-741): r2, corner angles,
    density from the centre of mass, pair distances - on the four evaluation configurations, and the reference's own
    sampled means (config/He4He4Na.config's end-of-run pass, shortened) for the equilibrated one."""
    cfg = json.load(open(os.path.join(REF, "config", "He4He4Na.config")))
    out = {}
    for tag in ("fixture", "compact", "stretched", "equil"):
        g = np.load(os.path.join(GOLDEN, f"mixture_he4he4na_{tag}.npz"))
        scal = dict(N=int(g["N"]), LBOX=float(g["LBOX"]), N_PARAM=int(g["N_PARAM"]), phiR=float(g["phiR"]), phiI=0.0, USE_NURBS=1,
                    GR_BIN_COUNT=int(cfg["GR_BIN_COUNT"]))
        mc = {}
        if tag == "equil":
            mc = dict(MC_STEP=float(cfg["MC_STEP"]), MC_NADDITIONALSTEPS=60000, MC_NADDITIONALTHERMSTEPS=60,
                      MC_NADDITIONALINITIALIZATIONSTEPS=3000, seed=3)
        arr = dict(R=g["R"], uR=g["uR"], uI=g["uI"], NURBS_GRID=g["NURBS_GRID"], PARTICLE_TYPES=g["PARTICLE_TYPES"],
                   SYSTEM_PARAMS=g["SYSTEM_PARAMS"])
        with tempfile.TemporaryDirectory() as td:
            cp, op = os.path.join(td, "case.txt"), os.path.join(td, "out.txt")
            write_case(cp, "BosonMixtureCluster", dict(scal, **mc), arr)
            run("obs", cp, op)
            d = parse_dump(op)
        for k in ("r2_fixed", "angle_fixed", "density_fixed", "distance_fixed"):
            out[f"{tag}_{k}"] = d[k]
        if tag == "equil":
            for k in ("angle_grid", "density_grid", "distance_grid", "density_scaling", "r2_mean", "angle_mean", "density_mean",
                      "distance_mean", "acceptance"):
                out[k] = d[k]
            for k, v in mc.items():
                out[k] = np.array(v)
        print(f"mixture obs {tag}: r2={float(d['r2_fixed']):.6g} angle bins {np.argmax(d['angle_fixed'], axis=1)}")
    np.savez_compressed(os.path.join(GOLDEN, "mixture_he4he4na_obs.npz"), **out)


def gen_evolution(only=None):
    """Time evolution by the reference's own time-step functions (ref_harness evolve): BosonsBulk N = 64, Euler steps
    with the Cholesky solve, 24 RNG seeds -> mean and spread of the parameter trajectories and of the energies
    (north_star level 2: time-evolved parameters agree statistically), plus the first step's estimators and
    derivatives of one seed to pin the host-side restatement of SolveForParametersDot.  Two runs: imaginary time
    (30 steps of 2e-4: the real parts relax) and real time (12 steps of 1e-4 - explicit Euler is only stable for a
    short stretch in real time - in which the imaginary parts grow out of zero)."""
    g = np.load(os.path.join(GOLDEN, "bosonsbulk_n64_equil.npz"))
    P = int(g["N_PARAM"])
    seeds = list(range(1, 25))
    for name, imag, dt, nsteps in (("bosonsbulk_n64_evolution", 1, 2e-4, 30), ("bosonsbulk_n64_evolution_realtime", 0, 1e-4, 12),
                                   # IMAGINARY_TIME = -1: the 1.499 pi time rotation (src/TDVMC.cpp:1475-1504, 1666-1673)
                                   ("bosonsbulk_n64_evolution_rotation", -1, 1e-4, 6),
                                   # LINEAR_EQUATION_SOLVER_TYPE = 1: Eigen FullPivHouseholderQR (:1763-1827), with the scaling +
                                   # 0.002 regularisation of USE_PRECONDITIONING = 1 and without
                                   ("bosonsbulk_n64_evolution_qr", 1, 2e-4, 6), ("bosonsbulk_n64_evolution_qr_raw", 0, 1e-4, 4)):
        if only and name not in only:
            continue
        extra = {}
        if name.endswith("_qr"):
            extra = dict(LINEAR_EQUATION_SOLVER_TYPE=1, USE_PRECONDITIONING=1)
        elif name.endswith("_qr_raw"):
            extra = dict(LINEAR_EQUATION_SOLVER_TYPE=1, USE_PRECONDITIONING=0)
        base = dict(N=64, LBOX=4.0, N_PARAM=P, time=0.0, phiR=0.0, phiI=0.0, MC_STEP=0.4, MC_NSTEPS=1024, MC_NTHERMSTEPS=32,
                    MC_NINITIALIZATIONSTEPS=64, IMAGINARY_TIME=imag, TIMESTEP=dt, time_steps=nsteps, equilibration_steps=6400, **extra)
        arr = dict(R=g["R"], uR=g["uR"], uI=np.zeros(P), SYSTEM_PARAMS=[1.0, 1.0])

        def one(sd):
            with tempfile.TemporaryDirectory() as td:
                cp, op = os.path.join(td, "case.txt"), os.path.join(td, "out.txt")
                write_case(cp, "BosonsBulk", dict(base, seed=sd), arr)
                run("evolve", cp, op)
                return parse_dump(op)

        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=max(1, (os.cpu_count() or 2) - 1)) as ex:   # each run is its own process
            runs = list(ex.map(one, seeds))
        for sd, r in zip(seeds, runs):
            assert float(r["cholesky_failed"]) == 0.0
        print(f"{name}: E {np.mean([r['energy_r_t'][0] for r in runs]):.3f} -> {np.mean([r['energy_r_t'][-1] for r in runs]):.3f}")
        uR_t = np.stack([r["uR_t"] for r in runs])            # [seed][step][P]
        uI_t = np.stack([r["uI_t"] for r in runs])
        e_t = np.stack([r["energy_r_t"] for r in runs])
        first = runs[0]
        out = {k: np.array(v) for k, v in base.items()}
        out.update(system=np.array("BosonsBulk"), source=np.array("bosonsbulk_n64_equil"), seeds=np.array(seeds),
                   SYSTEM_PARAMS=np.array([1.0, 1.0]), uR0=g["uR"], uR_t_mean=uR_t.mean(axis=0), uR_t_std=uR_t.std(axis=0, ddof=1),
                   uI_t_mean=uI_t.mean(axis=0), uI_t_std=uI_t.std(axis=0, ddof=1),
                   energy_r_t_mean=e_t.mean(axis=0), energy_r_t_std=e_t.std(axis=0, ddof=1),
                   phiR_t_mean=np.stack([r["phiR_t"] for r in runs]).mean(axis=0),
                   acceptance=np.mean([float(r["acceptance"]) for r in runs]))
        for k in ("first_O", "first_S", "first_OER", "first_OEI", "first_ER", "first_EI", "first_uDotR", "first_uDotI",
                  "first_phiDotR", "first_phiDotI"):
            out[k] = first[k]
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)


def hebulk_drift(d, uR, uI):
    """Drift from the REFERENCE's HeBulk tables with its parameter map (HeBulk.cpp:319-356), long double."""
    sD = d["sD"].astype(np.longdouble)
    mc = d["mcmillan_sum_d"].astype(np.longdouble)
    f11, f12, f21, f22, fSL, fL, fSLP, fLP = [np.longdouble(x) for x in d["bc_factors"]]
    K, P = sD.shape[0], len(uR)
    out = []
    for u in (uR, uI):
        u = u.astype(np.longdouble)
        F = u[0] * (mc + f11 * sD[0] + f21 * sD[1]) + u[1] * (sD[2] + f12 * sD[0] + f22 * sD[1])
        for k in range(2, P - 2):
            F = F + u[k] * sD[k + 1]
        F = F + u[P - 2] * (sD[K - 6] + fSL * sD[K - 5] + fL * sD[K - 4])
        F = F + u[P - 1] * (1 + fSLP * sD[K - 5] + fLP * sD[K - 4])
        out.append(F.astype(np.float64))
    return out


def pack_eval_hebulk(name, scal, arrays, moves):
    with tempfile.TemporaryDirectory() as td:
        cp, op = os.path.join(td, "case.txt"), os.path.join(td, "out.txt")
        write_case(cp, "HeBulk", scal, arrays, moves)
        run("eval", cp, op)
        d = parse_dump(op)
    out = {"system": np.array("HeBulk"), "N": np.array(scal["N"]), "DIM": np.array(3), "LBOX": np.array(scal["LBOX"]),
           "N_PARAM": np.array(scal["N_PARAM"]), "SYSTEM_PARAMS": np.zeros(0), "time": np.array(0.0),
           "R": np.asarray(arrays["R"], np.float64).reshape(-1, 3), "uR": np.asarray(arrays["uR"], np.float64),
           "uI": np.asarray(arrays["uI"], np.float64), "phiR": np.array(scal.get("phiR", 0.0)),
           "phiI": np.array(scal.get("phiI", 0.0)), "moves": np.asarray(moves, np.float64).reshape(-1, 4)}
    for k in ("exponent", "exponent_wf", "wf", "local_energy_r", "local_energy_i", "local_operators",
              "local_operator_energy_r", "local_operator_energy_i", "other_expectation_values",
              "local_operators_matrix_diag", "local_operators_matrix_row3", "spline_sums", "mcmillan_sum", "sD", "sD2",
              "mcmillan_sum_d", "mcmillan_sum_d2", "rij_split", "node_point_spacing", "max_distance", "bc_factors",
              "move_quotient", "move_exponent_new"):
        out[k] = d[k]
    out["drift_r"], out["drift_i"] = hebulk_drift(d, out["uR"], out["uI"])
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: E_R={float(d['local_energy_r']):.12g} E_I={float(d['local_energy_i']):.12g} "
          f"exponent={float(d['exponent']):.12g} q={d['move_quotient']}")
    return d


def gen_hebulk():
    """config/bulk_64.config: 64 He-4 atoms at rho = 0.0219 (L = 14.30), N_PARAM = 49, the config's own PARAMS_REAL."""
    rng = np.random.default_rng(64)
    cfg = json.load(open(os.path.join(REF, "config", "bulk_64.config")))
    N, P = int(cfg["N"]), int(cfg["N_PARAM"])
    L = (N / float(cfg["RHO"])) ** (1.0 / 3.0)
    uR = np.array(cfg["PARAMS_REAL"], dtype=np.float64)
    uI = 0.02 * np.sin(0.3 * np.arange(P))           # the config is imaginary-time (uI = 0); exercise the imaginary part
    # (1) the reference's 64-particle fixture (L=4 lattice-like), scaled to the He box
    R = read_csv_positions("particleconfiguration_64.csv").reshape(N, 3) * (L / 4.0)
    scal = dict(N=N, LBOX=L, N_PARAM=P, phiR=float(cfg["PARAM_PHIR"]), phiI=0.0)
    arr = dict(R=R, uR=uR, uI=uI)
    pack_eval_hebulk("hebulk_n64_fixture", scal, arr, default_moves(R, L, rng, sigma=0.3))
    # (2) equilibrated by the reference's own sampler
    mc = run_mc("HeBulk", dict(scal, MC_STEP=float(cfg["MC_STEP"]), MC_NSTEPS=1, MC_NTHERMSTEPS=64 * 300, seed=5), arr)
    R2 = mc["R_final"].reshape(N, 3)
    pack_eval_hebulk("hebulk_n64_equil", scal, dict(arr, R=R2), default_moves(R2, L, rng, sigma=0.3))
    # (3) sampler statistics at the config's parameters
    d = run_mc("HeBulk", dict(scal, MC_STEP=float(cfg["MC_STEP"]), MC_NINITIALIZATIONSTEPS=64 * 300, MC_NSTEPS=4000,
                              MC_NTHERMSTEPS=64, seed=13), dict(arr, uI=np.zeros(P)))
    np.savez_compressed(os.path.join(GOLDEN, "hebulk_n64_mc.npz"), N=np.array(N), LBOX=np.array(L), N_PARAM=np.array(P),
                        MC_STEP=np.array(float(cfg["MC_STEP"])), uR=uR, uI=np.zeros(P), R0=R, n_therm=np.array(64),
                        energy_r_series=d["energy_r_series"], local_energy_r=d["local_energy_r"],
                        local_operators=d["local_operators"], other_expectation_values=d["other_expectation_values"],
                        acceptance=np.array(float(d["n_acceptances"]) / float(d["n_trials"])))
    print(f"hebulk_n64_mc: <E_R>={float(d['local_energy_r']):.8g} acc={float(d['n_acceptances']) / float(d['n_trials']):.4f}")


def he_drift_from_tables(spec, d, uR, uI):
    """Drift from the REFERENCE's derivative tables through the parameter map, long double."""
    sD = d["sD"].astype(np.longdouble)
    N = sD.shape[1]
    zero = np.zeros((1, N, 3), np.longdouble)
    lin = d["linear_sum_d"].astype(np.longdouble)[None] if "linear_sum_d" in d else zero
    tab = np.concatenate([sD, d["mcmillan_sum_d"].astype(np.longdouble)[None], zero, lin])
    out = []
    for u in (uR, uI):
        F = np.zeros((N, 3), np.longdouble)
        for p, row in enumerate(spec.map_rows()):
            t = np.full((N, 3), np.longdouble(spec.grad_const[p]))
            for k, f in row:
                t = t + np.longdouble(f) * tab[k]
            F = F + np.longdouble(u[p]) * t
        out.append(F.astype(np.float64))
    return out


def pack_eval_hedrop(name, scal, arrays, moves):
    with tempfile.TemporaryDirectory() as td:
        cp, op = os.path.join(td, "case.txt"), os.path.join(td, "out.txt")
        write_case(cp, "HeDrop", scal, arrays, moves)
        run("eval", cp, op)
        d = parse_dump(op)
    out = {"system": np.array("HeDrop"), "N": np.array(scal["N"]), "DIM": np.array(3), "LBOX": np.array(scal["LBOX"]),
           "N_PARAM": np.array(scal["N_PARAM"]), "SYSTEM_PARAMS": np.zeros(0), "time": np.array(0.0),
           "R": np.asarray(arrays["R"], np.float64).reshape(-1, 3), "uR": np.asarray(arrays["uR"], np.float64),
           "uI": np.asarray(arrays["uI"], np.float64), "phiR": np.array(scal.get("phiR", 0.0)),
           "phiI": np.array(scal.get("phiI", 0.0)), "moves": np.asarray(moves, np.float64).reshape(-1, 4)}
    for k in ("exponent", "exponent_wf", "wf", "local_energy_r", "local_energy_i", "local_operators",
              "local_operator_energy_r", "local_operator_energy_i", "other_expectation_values",
              "local_operators_matrix_diag", "local_operators_matrix_row3", "spline_sums", "mcmillan_sum", "const_sum",
              "linear_sum", "sD", "sD2", "mcmillan_sum_d", "mcmillan_sum_d2", "linear_sum_d", "linear_sum_d2", "rij_split",
              "rij_spline_split", "rij_tail", "bc_factors", "move_quotient", "move_exponent_new"):
        out[k] = d[k]
    spec = tsys.he_drop(int(scal["N"]), int(scal["N_PARAM"]))
    out["drift_r"], out["drift_i"] = he_drift_from_tables(spec, d, out["uR"], out["uI"])
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: E_R={float(d['local_energy_r']):.12g} E_I={float(d['local_energy_i']):.12g} "
          f"exponent={float(d['exponent']):.12g} q={d['move_quotient']}")
    return d


def gen_hedrop():
    """config/drop_6.config: 6 He-4 atoms, open boundary, N_PARAM = 93, the config's own PARAMS_REAL."""
    rng = np.random.default_rng(6)
    cfg = json.load(open(os.path.join(REF, "config", "drop_6.config")))
    N, P = int(cfg["N"]), int(cfg["N_PARAM"])
    uR = np.array(cfg["PARAMS_REAL"], dtype=np.float64)
    uI = 0.01 * np.cos(0.2 * np.arange(P))
    R = read_csv_positions("particleconfiguration_6.csv").reshape(N, 3)      # the reference's 6-particle fixture
    scal = dict(N=N, LBOX=40.0, N_PARAM=P, phiR=float(cfg["PARAM_PHIR"]), phiI=0.0, GR_BIN_COUNT=200, RHO_BIN_COUNT=200)
    arr = dict(R=R, uR=uR, uI=uI)
    pack_eval_hedrop("hedrop_n6_fixture", scal, arr, default_moves(R, 0.0, rng, sigma=0.5))
    # a spread-out configuration that reaches the 0.5-spacing grid and the const/linear tails (r >= r_tail = 21.2)
    R2 = R * 3.4
    pack_eval_hedrop("hedrop_n6_spread", scal, dict(arr, R=R2), default_moves(R2, 0.0, rng, sigma=1.5))
    mc = run_mc("HeDrop", dict(scal, MC_STEP=float(cfg["MC_STEP"]), MC_NSTEPS=1, MC_NTHERMSTEPS=6 * 2000, seed=4), arr)
    R3 = mc["R_final"].reshape(N, 3)
    pack_eval_hedrop("hedrop_n6_equil", scal, dict(arr, R=R3), default_moves(R3, 0.0, rng, sigma=0.5))
    # no sampler-statistics fixture for HeDrop: with HBAR2_2M = 1 (src/Constants.h:12) the 6-atom droplet is unbound and
    # the chain is not stationary; sampling is pinned by replaying the oracle chain move for move instead


def pack_eval_mixture(name, scal, arrays, moves, order=3):
    system = "BosonMixtureCluster" if order == 3 else "BosonMixtureCluster_4thorder"
    with tempfile.TemporaryDirectory() as td:
        cp, op = os.path.join(td, "case.txt"), os.path.join(td, "out.txt")
        write_case(cp, system, scal, arrays, moves)
        run("eval", cp, op)
        d = parse_dump(op)
    out = {"system": np.array(system), "N": np.array(scal["N"]), "DIM": np.array(3), "LBOX": np.array(scal["LBOX"]),
           "N_PARAM": np.array(scal["N_PARAM"]), "SYSTEM_PARAMS": np.zeros(0), "time": np.array(0.0),
           "PARTICLE_TYPES": np.asarray(arrays["PARTICLE_TYPES"], np.float64), "NURBS_GRID": np.asarray(arrays["NURBS_GRID"]),
           "R": np.asarray(arrays["R"], np.float64).reshape(-1, 3), "uR": np.asarray(arrays["uR"], np.float64),
           "uI": np.asarray(arrays["uI"], np.float64), "phiR": np.array(scal.get("phiR", 0.0)),
           "phiI": np.array(scal.get("phiI", 0.0)), "moves": np.asarray(moves, np.float64).reshape(-1, 4)}
    for k, v in d.items():
        out[k] = v
    T = int(d["n_pair_types"])
    spec = tsys.boson_mixture_cluster(arrays["PARTICLE_TYPES"], [d[f"knots_{t}"] for t in range(T)],
                                      [d[f"spline_weights_{t}"] for t in range(T)], [d[f"bc_factors_{t}"] for t in range(T)],
                                      order=order)
    N = int(scal["N"])
    tabs = []
    for t in range(T):
        z = np.zeros((1, N, 3), np.longdouble)
        tabs += [d[f"sD_{t}"].astype(np.longdouble), d[f"mcmillan_sum_d_{t}"].astype(np.longdouble)[None], z,
                 d[f"linear_sum_d_{t}"].astype(np.longdouble)[None], d[f"log_sum_d_{t}"].astype(np.longdouble)[None]]
    tab = np.concatenate(tabs)
    for key, u in (("drift_r", out["uR"]), ("drift_i", out["uI"])):
        F = np.zeros((N, 3), np.longdouble)
        for p, row in enumerate(spec.map_rows()):
            tt = np.zeros((N, 3), np.longdouble)
            for k, f in row:
                tt = tt + np.longdouble(f) * tab[k]
            F = F + np.longdouble(u[p]) * tt
        out[key] = F.astype(np.float64)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: E_R={float(d['local_energy_r']):.12g} E_I={float(d['local_energy_i']):.12g} "
          f"exponent={float(d['exponent']):.12g} q={d['move_quotient']}")
    return d


def gen_mixture():
    """config/He4He4Na.config: two He-4 atoms and one Na, N_PARAM = 52 = 2 x 26, the config's own parameters."""
    rng = np.random.default_rng(3)
    cfg = json.load(open(os.path.join(REF, "config", "He4He4Na.config")))
    N, P = int(cfg["N"]), int(cfg["N_PARAM"])
    uR = np.array(cfg["PARAMS_REAL"], dtype=np.float64)
    uI = 0.01 * np.sin(0.4 * np.arange(P))
    R = read_csv_positions("particleconfiguration_3.csv").reshape(N, 3)      # the reference's 3-particle fixture
    scal = dict(N=N, LBOX=float(cfg["LBOX"]), N_PARAM=P, phiR=float(cfg["PARAM_PHIR"]), phiI=0.0, USE_NURBS=1,
                GR_BIN_COUNT=int(cfg["GR_BIN_COUNT"]))
    arr = dict(R=R, uR=uR, uI=uI, NURBS_GRID=cfg["NURBS_GRID"], PARTICLE_TYPES=cfg["PARTICLE_TYPES"],
               SYSTEM_PARAMS=cfg["SYSTEM_PARAMS"])
    pack_eval_mixture("mixture_he4he4na_fixture", scal, arr, default_moves(R, 0.0, rng, sigma=1.0))
    # compact and stretched triangles: McMillan cores (r < 2.0 / 4.0) and the tails (r >= 15.8 / 17.8)
    R2 = np.array([[0.0, 0.0, 0.0], [1.7, 0.3, -0.2], [3.1, 2.0, 0.5]])
    pack_eval_mixture("mixture_he4he4na_compact", scal, dict(arr, R=R2), default_moves(R2, 0.0, rng, sigma=1.0))
    R3 = np.array([[0.0, 0.0, 0.0], [16.5, 1.0, -2.0], [-9.0, 17.0, 4.0]])
    pack_eval_mixture("mixture_he4he4na_stretched", scal, dict(arr, R=R3), default_moves(R3, 0.0, rng, sigma=4.0))
    mc = run_mc("BosonMixtureCluster", dict(scal, MC_STEP=float(cfg["MC_STEP"]), MC_NSTEPS=1, MC_NTHERMSTEPS=3 * 3000, seed=2), arr)
    R4 = mc["R_final"].reshape(N, 3)
    pack_eval_mixture("mixture_he4he4na_equil", scal, dict(arr, R=R4), default_moves(R4, 0.0, rng, sigma=2.0))


def gen_mixture_4th():
    """config/He4He4Na_4thOrder.config (BosonMixtureCluster_4thorder): the same He4-He4-Na trimer on quartic splines
    (SplineFactory::GetWeights4, 28 splines and three boundary factors per pair type), the config's own grid.  Its shipped
    PARAMS_REAL is a placeholder, so the cubic config's optimised parameters are used (same 2 x 26 layout)."""
    rng = np.random.default_rng(34)
    cfg = json.load(open(os.path.join(REF, "config", "He4He4Na_4thOrder.config")))
    cfg3 = json.load(open(os.path.join(REF, "config", "He4He4Na.config")))
    N, P = int(cfg["N"]), int(cfg["N_PARAM"])
    uR = np.array(cfg["PARAMS_REAL"], dtype=np.float64)
    if len(uR) != P or not np.any(uR):
        uR = np.array(cfg3["PARAMS_REAL"], dtype=np.float64)
    uI = 0.01 * np.sin(0.4 * np.arange(P))
    R = read_csv_positions("particleconfiguration_3.csv").reshape(N, 3)
    scal = dict(N=N, LBOX=float(cfg["LBOX"]), N_PARAM=P, phiR=float(cfg.get("PARAM_PHIR", 0.0)), phiI=0.0, USE_NURBS=1,
                GR_BIN_COUNT=int(cfg["GR_BIN_COUNT"]))
    arr = dict(R=R, uR=uR, uI=uI, NURBS_GRID=cfg["NURBS_GRID"], PARTICLE_TYPES=cfg["PARTICLE_TYPES"],
               SYSTEM_PARAMS=cfg["SYSTEM_PARAMS"])
    pack_eval_mixture("mixture4_he4he4na_fixture", scal, arr, default_moves(R, 0.0, rng, sigma=1.0), order=4)
    R2 = np.array([[0.0, 0.0, 0.0], [1.7, 0.3, -0.2], [3.1, 2.0, 0.5]])          # inside the McMillan cores
    pack_eval_mixture("mixture4_he4he4na_compact", scal, dict(arr, R=R2), default_moves(R2, 0.0, rng, sigma=1.0), order=4)
    R3 = np.array([[0.0, 0.0, 0.0], [22.5, 1.0, -2.0], [-9.0, 24.0, 4.0]])       # out on the tails
    pack_eval_mixture("mixture4_he4he4na_stretched", scal, dict(arr, R=R3), default_moves(R3, 0.0, rng, sigma=4.0), order=4)
    mc = run_mc("BosonMixtureCluster_4thorder", dict(scal, MC_STEP=float(cfg["MC_STEP"]), MC_NSTEPS=1, MC_NTHERMSTEPS=3 * 3000, seed=2), arr)
    R4 = mc["R_final"].reshape(N, 3)
    pack_eval_mixture("mixture4_he4he4na_equil", scal, dict(arr, R=R4), default_moves(R4, 0.0, rng, sigma=2.0), order=4)


def gen_more_configs():
    """Further shipped configs of systems that are already covered, for the paths the BASELINE configs do not reach:
    config/He3He4Cs.config (BosonMixtureCluster with THREE pair types, N_PARAM = 78, He-3 mass, KTTY He-Cs potential) and
    config/drop_20.config (HeDrop with 20 atoms: a whole warp per walker in the sweep instead of 8 lanes)."""
    rng = np.random.default_rng(78)
    cfg = json.load(open(os.path.join(REF, "config", "He3He4Cs.config")))
    N, P = int(cfg["N"]), int(cfg["N_PARAM"])
    uR = np.array(cfg["PARAMS_REAL"], dtype=np.float64)
    uI = 0.01 * np.sin(0.3 * np.arange(P))
    scal = dict(N=N, LBOX=float(cfg.get("LBOX", 10.0)), N_PARAM=P, phiR=float(cfg.get("PARAM_PHIR", 0.0)), phiI=0.0, USE_NURBS=1,
                GR_BIN_COUNT=int(cfg["GR_BIN_COUNT"]))
    R = np.array([[0.0, 0.0, 0.0], [6.5, 1.0, -1.5], [-2.0, 7.5, 3.0]])
    arr = dict(R=R, uR=uR, uI=uI, NURBS_GRID=cfg["NURBS_GRID"], PARTICLE_TYPES=cfg["PARTICLE_TYPES"], SYSTEM_PARAMS=cfg["SYSTEM_PARAMS"])
    mc = run_mc("BosonMixtureCluster", dict(scal, MC_STEP=float(cfg["MC_STEP"]), MC_NSTEPS=1, MC_NTHERMSTEPS=3 * 3000, seed=5), arr)
    R1 = mc["R_final"].reshape(N, 3)
    pack_eval_mixture("mixture_he3he4cs_equil", scal, dict(arr, R=R1), default_moves(R1, 0.0, rng, sigma=2.0))
    cfg = json.load(open(os.path.join(REF, "config", "drop_20.config")))
    N, P = int(cfg["N"]), int(cfg["N_PARAM"])
    uR = np.array(cfg["PARAMS_REAL"], dtype=np.float64)
    uI = 0.01 * np.cos(0.2 * np.arange(P))
    g = (np.arange(3) - 1.0) * 4.2
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    R = (np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1) + rng.uniform(-0.4, 0.4, (27, 3)))[:N]
    scal = dict(N=N, LBOX=40.0, N_PARAM=P, phiR=float(cfg.get("PARAM_PHIR", 0.0)), phiI=0.0, GR_BIN_COUNT=200, RHO_BIN_COUNT=200)
    arr = dict(R=R, uR=uR, uI=uI)
    mc = run_mc("HeDrop", dict(scal, MC_STEP=float(cfg["MC_STEP"]), MC_NSTEPS=1, MC_NTHERMSTEPS=N * 300, seed=6), arr)
    R1 = mc["R_final"].reshape(N, 3)
    pack_eval_hedrop("hedrop_n20_equil", scal, dict(arr, R=R1), default_moves(R1, 0.0, rng, sigma=0.5))


def pack_eval_inhcontact(name, scal, arrays, moves):
    system = "InhContactBosons"
    x = np.asarray(arrays["R"], np.float64).reshape(-1)
    with tempfile.TemporaryDirectory() as td:
        cp, op = os.path.join(td, "case.txt"), os.path.join(td, "out.txt")
        write_case(cp, system, scal, dict(arrays, R=x), moves)
        run("eval", cp, op)
        d = parse_dump(op)
    R3 = np.zeros((len(x), 3))
    R3[:, 0] = x
    out = {"system": np.array(system), "N": np.array(scal["N"]), "DIM": np.array(1), "LBOX": np.array(scal["LBOX"]),
           "N_PARAM": np.array(scal["N_PARAM"]), "SYSTEM_PARAMS": np.asarray(arrays["SYSTEM_PARAMS"], np.float64),
           "time": np.array(scal.get("time", 0.0)), "R": R3, "uR": np.asarray(arrays["uR"], np.float64),
           "uI": np.asarray(arrays["uI"], np.float64), "phiR": np.array(scal.get("phiR", 0.0)), "phiI": np.array(scal.get("phiI", 0.0)),
           "moves": np.asarray(moves, np.float64).reshape(-1, 4)}
    for k, v in d.items():
        out[k] = v
    spec = tsys.from_golden(out)
    tab = np.concatenate([d["sD_spf"][:, :, 0], d["sD_pc"][:, :, 0]]).astype(np.longdouble)      # [K1 + K2][N]
    cg = np.longdouble(-2.0 * float(d["gamma"]) * float(d["node_spacing_pc"]))
    K1 = spec.extra["n_splines_spf"]
    for key, u, contact in (("drift_r", out["uR"], cg), ("drift_i", out["uI"], np.longdouble(0))):
        F = contact * tab[K1]                                   # InhContactBosons.cpp:595-602: real part only
        for p, row in enumerate(spec.map_rows()):
            t = np.zeros(tab.shape[1], np.longdouble)
            for k, f in row:
                t = t + np.longdouble(f) * tab[k]
            F = F + np.longdouble(u[p]) * t
        D = np.zeros((len(x), 3))
        D[:, 0] = F.astype(np.float64)
        out[key] = D
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: E_R={float(d['local_energy_r']):.12g} E_I={float(d['local_energy_i']):.12g} "
          f"exponent={float(d['exponent']):.12g} q={d['move_quotient']}")
    return d


def gen_inhcontact():
    """InhContactBosons (the one-body WFParts system of config/InhContactBosons.config, the only
    shipped config of the current format): N = 3, L = 3, N_PARAM = 62 as shipped, with the contact interaction and the
    lattice potential switched on (the shipped SYSTEM_PARAMS leave both at zero), a square-well variant, and N = 20."""
    rng = np.random.default_rng(62)
    cfg = json.loads(re.sub(r"(\d)\.(\s*[,\]\}])", r"\g<1>.0\2",
                            open(os.path.join(REF, "config", "InhContactBosons.config")).read()))   # "0." literals
    N, P, L = int(cfg["N"]), int(cfg["N_PARAM"]), float(cfg["LBOX"])
    k = np.arange(P)
    uR = np.where(k < P // 2, 0.3 * np.cos(2 * np.pi * k / (P // 2)), -0.4 * np.exp(-((k - P // 2) / 6.0) ** 2))
    uI = 0.03 * np.sin(0.37 * k)

    def moves1d(x, n=6, sigma=0.5):
        out = []
        for _ in range(n):
            p = int(rng.integers(0, len(x)))
            out.append([p, x[p] + rng.normal(0, sigma), 0.0, 0.0])
        return out

    scal = dict(N=N, DIM=1, LBOX=L, N_PARAM=P, phiR=0.02, phiI=0.0, GR_BIN_COUNT=int(cfg["GR_BIN_COUNT"]),
                RHO_BIN_COUNT=int(cfg["RHO_BIN_COUNT"]), time=0.0)
    x0 = np.array([-1.1, 0.2, 0.9])
    arr = dict(R=x0, uR=uR, uI=uI, SYSTEM_PARAMS=[0.0, 10.0, 1.0, 2.0])
    pack_eval_inhcontact("inhcontact_n3_fixture", scal, arr, moves1d(x0))
    pack_eval_inhcontact("inhcontact_n3_well", scal, dict(arr, SYSTEM_PARAMS=[0.8, 5.0, 2.0, 1.5]), moves1d(x0))
    mcs = dict(scal, MC_STEP=float(cfg["MC_STEP"]), MC_NSTEPS=1, MC_NTHERMSTEPS=N * 500, seed=31)
    mc = run_mc("InhContactBosons", mcs, arr)
    x1 = mc["R_final"].reshape(N)
    pack_eval_inhcontact("inhcontact_n3_equil", scal, dict(arr, R=x1), moves1d(x1))
    mc2 = run_mc("InhContactBosons", dict(mcs, MC_NSTEPS=6000, MC_NTHERMSTEPS=3 * N, MC_NINITIALIZATIONSTEPS=N * 200, seed=32), dict(arr, R=x1))
    np.savez_compressed(os.path.join(GOLDEN, "inhcontact_n3_mc.npz"), source=np.array("inhcontact_n3_equil"),
                        MC_STEP=np.array(mcs["MC_STEP"]), MC_NSTEPS=np.array(6000), MC_NTHERMSTEPS=np.array(3 * N),
                        local_energy_r=mc2["local_energy_r"], local_energy_i=mc2["local_energy_i"],
                        local_operators=mc2["local_operators"], energy_r_series=mc2["energy_r_series"],
                        acceptance=np.array(float(mc2["n_acceptances"]) / float(mc2["n_trials"])))
    print("inhcontact_n3_mc: E_R=%.6f acceptance=%.3f" % (float(mc2["local_energy_r"]), float(mc2["n_acceptances"]) / float(mc2["n_trials"])))
    N2, L2 = 20, 20.0
    x2 = (np.arange(N2) + 0.5) * (L2 / N2) - L2 / 2 + rng.uniform(-0.2, 0.2, N2)
    scal2 = dict(scal, N=N2, LBOX=L2)
    arr2 = dict(arr, R=x2, SYSTEM_PARAMS=[0.0, 4.0, 1.0, 1.0])
    mc3 = run_mc("InhContactBosons", dict(scal2, MC_STEP=0.5, MC_NSTEPS=1, MC_NTHERMSTEPS=N2 * 200, seed=33), arr2)
    x3 = mc3["R_final"].reshape(N2)
    pack_eval_inhcontact("inhcontact_n20_equil", scal2, dict(arr2, R=x3), moves1d(x3))


def gen_lowdim():
    """The two spline-table systems in one and two dimensions, from the reference's own low-dimensional configs:
    config/BosonsBulk2D.config (N = 16, L = 4, N_PARAM = 40), config/NUBosonsBulkPB2D.config (N = 25, N_PARAM = 100, grid up
    to the half diagonal), config/Rydberg2D.config (NUBosonsBulkPB, N = 50, grid to L/2: the reflection rule is active) and
    config/BosonsBulk1D.config (N = 20, N_PARAM = 100)."""
    rng = np.random.default_rng(2)

    def load(name):
        t = open(os.path.join(REF, "config", name)).read()
        return json.loads(re.sub(r"(\d)\.(\s*[,\]\}])", r"\g<1>.0\2", t))

    def lattice(N, L, D):
        m = int(np.ceil(N ** (1.0 / D) - 1e-9))
        g = (np.arange(m) + 0.5) * (L / m) - L / 2
        X = np.stack(np.meshgrid(*([g] * D), indexing="ij"), axis=-1).reshape(-1, D)[:N]      # first N sites if N != m^D
        return X + rng.uniform(-0.05, 0.05, X.shape) * (L / m)

    def moves_d(R, D, n=6, sigma=0.4):
        out = []
        for _ in range(n):
            p = int(rng.integers(0, R.shape[0]))
            new = np.zeros(3)
            new[:D] = R[p] + rng.normal(0, sigma, D)
            out.append([p] + list(new))
        return out

    for cfgname, system, tag, L in (("BosonsBulk2D.config", "BosonsBulk", "bosonsbulk2d_n16", 4.0),
                                    ("NUBosonsBulkPB2D.config", "NUBosonsBulkPB", "nubosonsbulkpb2d_n25", 5.0),
                                    ("Rydberg2D.config", "NUBosonsBulkPB", "rydberg2d_n50", None),
                                    ("BosonsBulk1D.config", "BosonsBulk", "bosonsbulk1d_n20", 20.0)):
        cfg = load(cfgname)
        N, D, P = int(cfg["N"]), int(cfg["DIM"]), int(cfg["N_PARAM"])
        L = float(cfg["LBOX"]) if L is None else L
        nurbs = system == "NUBosonsBulkPB"
        grid = np.array(cfg["NURBS_GRID"], dtype=np.float64) if nurbs else None
        rmax = float(grid[-1]) if nurbs else L / 2
        uR, uI = smooth_params(P, rmax, w_r=0.3 * rmax, c_i=0.5 * rmax, w_i=0.2 * rmax)
        scal = dict(N=N, DIM=D, LBOX=L, N_PARAM=P, USE_NURBS=1 if nurbs else 0, time=0.0, phiR=0.03, phiI=0.0,
                    GR_BIN_COUNT=int(cfg.get("GR_BIN_COUNT") or 50))
        arr = dict(uR=uR, uI=uI, SYSTEM_PARAMS=cfg["SYSTEM_PARAMS"])
        if nurbs:
            arr["NURBS_GRID"] = grid
        R0 = lattice(N, L, D)
        mc = run_mc(system, dict(scal, MC_STEP=float(cfg["MC_STEP"]), MC_NSTEPS=1, MC_NTHERMSTEPS=N * 100, seed=3), dict(arr, R=R0))
        R1 = mc["R_final"].reshape(N, D)
        mv = moves_d(R1, D)
        # (the harness reads DIM coordinates of a move and ignores the zero padding)
        pack_eval(tag + "_equil", system, scal, dict(arr, R=R1), mv)


def gen_boxradial_2d():
    """config/NUBosonsBulkPBBoxAndRadial2D.config at its own size: N = 25, DIM = 2, rho = 1 -> L = 5, N_PARAM = 100."""
    rng = np.random.default_rng(25)
    cfg = json.loads(re.sub(r"(\d)\.(\s*[,\]\}])", r"\g<1>.0\2",
                            open(os.path.join(REF, "config", "NUBosonsBulkPBBoxAndRadial2D.config")).read()))
    N, P, D = int(cfg["N"]), int(cfg["N_PARAM"]), int(cfg["DIM"])
    L = float(round((N / float(cfg["RHO"])) ** (1.0 / D), 9))
    grid = np.array(cfg["NURBS_GRID"], dtype=np.float64)
    uR, uI = boxradial_params(P, L / 2)
    scal = dict(N=N, DIM=D, LBOX=L, N_PARAM=P, USE_NURBS=1, time=0.0, phiR=0.0, phiI=0.0, GR_BIN_COUNT=int(cfg["GR_BIN_COUNT"]))
    arr = dict(uR=uR, uI=uI, SYSTEM_PARAMS=cfg["SYSTEM_PARAMS"], NURBS_GRID=grid)
    g1 = (np.arange(5) + 0.5) * (L / 5) - L / 2
    X, Y = np.meshgrid(g1, g1, indexing="ij")
    R0 = np.stack([X.ravel(), Y.ravel()], axis=1) + rng.uniform(-0.05, 0.05, (N, 2))
    mc = run_mc("NUBosonsBulkPBBoxAndRadial", dict(scal, MC_STEP=float(cfg["MC_STEP"]), MC_NSTEPS=1, MC_NTHERMSTEPS=N * 200, seed=15),
                dict(arr, R=R0))
    R1 = mc["R_final"].reshape(N, D)
    moves = []
    for _ in range(6):
        p = int(rng.integers(0, N))
        moves.append([p] + list(R1[p] + rng.normal(0, 0.4, D)) + [0.0])
    pack_eval_boxradial("boxradial2d_n25_equil", scal, dict(arr, R=R1), moves)


def gen_min_image():
    """Reference minimum-image displacement on edge cases + random inputs (Utils.cpp:266-281, 352-382)."""
    rng = np.random.default_rng(99)
    rows = []
    for L in (4.0, 5.0, 7.0, 15.081):
        for _ in range(64):
            a = rng.uniform(-2.5 * L, 2.5 * L, 3)
            b = rng.uniform(-2.5 * L, 2.5 * L, 3)
            rows.append((L, a, b))
        # exact half-box ties and multiples of the box
        for s in (0.5, -0.5, 1.0, -1.0, 1.5, -1.5, 2.5, 0.0):
            rows.append((L, np.array([s * L, 0.0, -s * L]), np.zeros(3)))
            rows.append((L, np.array([0.1, s * L / 2, 0.3]), np.array([0.1, -s * L / 2, 0.3])))
    with tempfile.TemporaryDirectory() as td:
        ip, op = os.path.join(td, "in.txt"), os.path.join(td, "out.txt")
        with open(ip, "w") as f:
            for L, a, b in rows:
                f.write(f"{float(L)!r} 3 " + " ".join(repr(float(x)) for x in [*a, *b]) + "\n")
        run("nic", ip, op)
        res = np.loadtxt(op)
    np.savez_compressed(os.path.join(GOLDEN, "min_image_reference.npz"),
                        L=np.array([r[0] for r in rows]), a=np.array([r[1] for r in rows]),
                        b=np.array([r[2] for r in rows]), norm=res[:, 0], disp=res[:, 1:4])
    print(f"min_image_reference: {len(rows)} cases")


def boxradial_drift(d, uR, uI):
    """Drift from the REFERENCE's four tables, contracted as NUBosonsBulkPBBoxAndRadial.cpp:461-523 does (long double),
    including its use of the box table sD[K-1] for the last radial parameter's gradient (:493-497)."""
    sD, sDr = d["sD"].astype(np.longdouble), d["sD_rad"].astype(np.longdouble)   # [K][N][3]
    K = sD.shape[0]
    P = len(uR)
    PR = P // 2
    out = []
    for u in (uR, uI):
        u = np.asarray(u, np.longdouble)
        F = np.einsum("k,kna->na", u[:PR], sDr[1:PR + 1]) + u[1] * sDr[0]
        F = F + u[PR - 1] * sDr[K - 2] / np.longdouble(-2.0) + u[PR - 1] * sD[K - 1]
        F = F + np.einsum("k,kna->na", u[PR:], sD[1:PR + 1]) + u[PR + 1] * sD[0] + u[P - 1] * sD[K - 2] + u[P - 1] * sD[K - 1]
        out.append(F.astype(np.float64))
    return out


def boxradial_params(P, half):
    """Smooth non-zero parameters for both bases (the shipped config starts from zeros)."""
    PR = P // 2
    x = np.arange(PR) * half / (PR - 1)
    uR = np.concatenate([-0.5 * np.exp(-((x / 0.5) ** 2)), 0.04 * np.exp(-(((x - 0.6 * half) / (0.25 * half)) ** 2))])
    uI = np.concatenate([0.05 * np.exp(-(((x - 0.5 * half) / (0.2 * half)) ** 2)), -0.02 * np.exp(-((x / (0.4 * half)) ** 2))])
    return uR, uI


def pack_eval_boxradial(name, scal, arrays, moves):
    system = "NUBosonsBulkPBBoxAndRadial"
    with tempfile.TemporaryDirectory() as td:
        cp, op = os.path.join(td, "case.txt"), os.path.join(td, "out.txt")
        write_case(cp, system, scal, arrays, moves)
        run("eval", cp, op)
        d = parse_dump(op)
    D = int(scal.get("DIM", 3))
    out = {"system": np.array(system), "N": np.array(scal["N"]), "DIM": np.array(D), "LBOX": np.array(scal["LBOX"]),
           "N_PARAM": np.array(scal["N_PARAM"]), "SYSTEM_PARAMS": np.asarray(arrays["SYSTEM_PARAMS"], np.float64),
           "NURBS_GRID": np.asarray(arrays["NURBS_GRID"], np.float64), "time": np.array(scal.get("time", 0.0)),
           "R": pad3(np.asarray(arrays["R"], np.float64).reshape(-1, D)), "uR": np.asarray(arrays["uR"], np.float64),
           "uI": np.asarray(arrays["uI"], np.float64), "phiR": np.array(scal.get("phiR", 0.0)),
           "phiI": np.array(scal.get("phiI", 0.0)), "moves": np.asarray(moves, np.float64).reshape(-1, 4)}
    for k in ("exponent", "exponent_wf", "wf", "local_energy_r", "local_energy_i", "local_operators", "local_operator_energy_r",
              "local_operator_energy_i", "other_expectation_values", "local_operators_matrix_diag", "local_operators_matrix_row3",
              "knots", "knots_rad", "spline_weights", "spline_weights_rad", "spline_sums", "spline_sums_rad", "max_distance_rad",
              "half_length", "other_local_operators", "gr_bins", "gr_bin_volumes", "gr_node_point_spacing", "move_quotient",
              "move_exponent_new", "sD", "sD2", "sD_rad", "sD2_rad"):
        out[k] = d[k]
    out["drift_r"], out["drift_i"] = (pad3(x) for x in boxradial_drift(d, out["uR"], out["uI"]))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: E_R={float(d['local_energy_r']):.12g} E_I={float(d['local_energy_i']):.12g} "
          f"exponent={float(d['exponent']):.12g} q={d['move_quotient']}")
    return d


def gen_boxradial():
    """NUBosonsBulkPBBoxAndRadial (config/NUBosonsBulkPBBoxAndRadial3D.config at its own size: N = 27,
    rho = 1 -> L = 3, N_PARAM = 100 on the config's 51-point grid, SYSTEM_PARAMS = [0.1, 50], GR_BIN_COUNT = 400), plus a
    64-particle case on a non-uniform grid.  Fixtures: the jittered start-up lattice, an equilibrated snapshot, a sampling run."""
    rng = np.random.default_rng(27)
    cfg = json.loads(re.sub(r"(\d)\.(\s*[,\]\}])", r"\g<1>.0\2",
                            open(os.path.join(REF, "config", "NUBosonsBulkPBBoxAndRadial3D.config")).read()))
    N, P = int(cfg["N"]), int(cfg["N_PARAM"])
    L = float(round((N / float(cfg["RHO"])) ** (1.0 / 3.0), 9))
    grid = np.array(cfg["NURBS_GRID"], dtype=np.float64)
    uR, uI = boxradial_params(P, L / 2)
    scal = dict(N=N, LBOX=L, N_PARAM=P, USE_NURBS=1, time=0.0, phiR=0.05, phiI=0.0, GR_BIN_COUNT=int(cfg["GR_BIN_COUNT"]))
    arr = dict(uR=uR, uI=uI, SYSTEM_PARAMS=cfg["SYSTEM_PARAMS"], NURBS_GRID=grid)
    # (A perfect lattice cannot be a fixture: a displacement component of exactly zero sends the reference's
    #  lower_bound to bin 2 and its loop to splineWeights[-1] (:258-266) - it crashes.  Jittered start instead.)
    Rj = jittered_lattice(N, L, seed=27)
    pack_eval_boxradial("boxradial_n27_jittered", scal, dict(arr, R=Rj), default_moves(Rj, L, rng))
    R0 = jittered_lattice(N, L, seed=28)
    mcs = dict(scal, MC_STEP=float(cfg["MC_STEP"]), MC_NSTEPS=1, MC_NTHERMSTEPS=N * 200, seed=12)
    mc = run_mc("NUBosonsBulkPBBoxAndRadial", mcs, dict(arr, R=R0))
    R1 = mc["R_final"].reshape(N, 3)
    pack_eval_boxradial("boxradial_n27_equil", scal, dict(arr, R=R1), default_moves(R1, L, rng))
    # sampling run of the reference's own Metropolis loop at these parameters: estimators + series for error bars
    mc2 = run_mc("NUBosonsBulkPBBoxAndRadial", dict(mcs, MC_NSTEPS=4000, MC_NTHERMSTEPS=N, MC_NINITIALIZATIONSTEPS=N * 50, seed=13),
                 dict(arr, R=R1))
    np.savez_compressed(os.path.join(GOLDEN, "boxradial_n27_mc.npz"), source=np.array("boxradial_n27_equil"),
                        MC_STEP=np.array(mcs["MC_STEP"]), MC_NSTEPS=np.array(4000), MC_NTHERMSTEPS=np.array(N),
                        local_energy_r=mc2["local_energy_r"], local_energy_i=mc2["local_energy_i"],
                        local_operators=mc2["local_operators"], energy_r_series=mc2["energy_r_series"],
                        energy_i_series=mc2["energy_i_series"], other_expectation_values=mc2["other_expectation_values"][:3],
                        acceptance=np.array(float(mc2["n_acceptances"]) / float(mc2["n_trials"])))
    print("boxradial_n27_mc: E_R=%.6f acceptance=%.3f" % (float(mc2["local_energy_r"]), float(mc2["n_acceptances"]) / float(mc2["n_trials"])))
    # 64 particles (more than one warp's worth), non-uniform grid, time-switched potential
    N2, L2, P2 = 64, 4.0, 60
    x = np.linspace(0.0, 1.0, P2 // 2 + 1)
    grid2 = 2.0 * (0.4 * x + 0.6 * x ** 2)
    grid2[-1] = 2.0
    uR2, uI2 = boxradial_params(P2, L2 / 2)
    scal2 = dict(N=N2, LBOX=L2, N_PARAM=P2, USE_NURBS=1, time=1.0, phiR=0.0, phiI=0.0, GR_BIN_COUNT=50)
    arr2 = dict(uR=uR2, uI=uI2, SYSTEM_PARAMS=[0.1, 50.0, 0.5, 0.3, 20.0], NURBS_GRID=grid2)
    R2 = jittered_lattice(N2, L2, seed=29)
    mc3 = run_mc("NUBosonsBulkPBBoxAndRadial", dict(scal2, MC_STEP=0.4, MC_NSTEPS=1, MC_NTHERMSTEPS=N2 * 100, seed=14), dict(arr2, R=R2))
    R3 = mc3["R_final"].reshape(N2, 3)
    pack_eval_boxradial("boxradial_n64_equil", scal2, dict(arr2, R=R3), default_moves(R3, L2, rng))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    if not os.path.exists(HARNESS):
        sys.exit("build the oracle first: make -C oracle/ref_build")
    which = sys.argv[1:] or ["min_image", "bosonsbulk", "bosonsbulk_mc", "bosonsbulk_mc_headline", "nubosonsbulkpb", "nubosonsbulkpb_full", "hebulk", "hedrop",
                             "mixture", "observables", "he_observables", "mixture_observables", "evolution", "boxradial", "mixture_4th", "more_configs", "inhcontact", "lowdim", "boxradial_2d"]
    for w in which:
        globals()["gen_" + w]()


if __name__ == "__main__":
    main()
