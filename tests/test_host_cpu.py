"""CPU-only tests: host logic, C-ABI surface, multi-rank reduction semantics (gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle_lib import Oracle
from tdvmc_b200 import estimators, splines, systems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib_path():
    path = os.path.join(ROOT, "tdvmc_b200", "libtdvmc_b200.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tdvmc_b200", "csrc"), "-j8"])
    return path


def test_library_exports_every_declared_symbol():
    """Every function include/tdvmc_gpu.h declares is exported by the built library and bound in capi.py."""
    from tdvmc_b200 import capi

    header = open(os.path.join(ROOT, "include", "tdvmc_gpu.h")).read()
    declared = set(re.findall(r"\b(tdvmc_gpu_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = C.CDLL(_lib_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == {n for n, _, _ in capi.SYMBOLS}
    assert capi.load().tdvmc_gpu_abi_version() == 2


def test_struct_layouts_match_header_sizes():
    from tdvmc_b200 import capi

    # 8 int32 + 2 double + 6 pointers + 2 int32 ; uint32 + 5 int32 + uint64 + double ; 7 pointers + 3 int64
    assert C.sizeof(capi.SystemDesc) == 8 * 4 + 2 * 8 + 6 * 8 + 4 * 4 + 3 * 8
    assert C.sizeof(capi.MixtureDesc) == 2 * 4 + 7 * 8
    assert C.sizeof(capi.ObservableDesc) == 2 * 4 + 3 * 8 + 3 * 8
    assert C.sizeof(capi.ClusterObservableDesc) == 4 * 4 + 5 * 8 + 8
    assert C.sizeof(capi.EnsembleDesc) == 6 * 4 + 8 + 8
    assert C.sizeof(capi.Estimators) == 7 * 8 + 3 * 8
    assert C.sizeof(capi.SolverDesc) == 4 * 4 + 2 * 8 + 2 * 4
    assert C.sizeof(capi.ParametersDot) == 2 * 8 + 4 * 8 + 8


def test_no_cpu_fallback_without_a_device(golden):
    from tdvmc_b200 import capi

    lib = capi.load()
    if lib.tdvmc_gpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    spec = systems.from_golden(golden("bosonsbulk_n64_fixture"))
    with pytest.raises(capi.TdvmcError, match="no CUDA device"):
        capi.Handle(spec, 4)


def test_product_package_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tdvmc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "oracle_lib" not in text and "tdvmc_oracle.h" not in text, f


@pytest.mark.parametrize("name", ["bosonsbulk_n343_equil", "nubosonsbulkpb_n216_equil", "bosonsbulk_n64_fixture"])
def test_spline_builder_against_reference_table(golden, name):
    g = golden(name)
    w_ref = g["spline_weights"]
    w = splines.bspline_monomial_weights(g["knots"])
    assert w.shape == w_ref.shape
    # coefficient-level agreement; the reference's closed forms lose ~3e-12 relative to cancellation
    scale = np.abs(w_ref).max(axis=2, keepdims=True)
    assert np.max(np.abs(w - w_ref) / scale) < 1e-10
    # partition of unity of our table on (0, r_max)
    knots = g["knots"]
    K = len(knots) - 4
    for r in np.linspace(knots[3] + 1e-6, knots[K] - 1e-6, 97):
        b = int(np.searchsorted(knots, r, side="left")) - 1
        tot = sum(np.polyval(w[b - p, p, ::-1], r) for p in range(4))
        assert abs(tot - 1.0) < 1e-7


def test_system_builders_reproduce_reference_setup(golden):
    for name in ("bosonsbulk_n343_equil", "nubosonsbulkpb_n216_equil"):
        g = golden(name)
        spec = systems.from_golden(g)                    # asserts bit-equal knots
        assert spec.r_max == float(g["max_distance"])
        o = Oracle(spec)
        assert np.allclose(o.local_operators(g["spline_sums"]), g["local_operators"], rtol=1e-15, atol=0)
    with pytest.raises(ValueError):
        systems.bosons_bulk(8, 4.0, 10, nurbs_grid=np.linspace(0, 2, 5))


def test_shard_walkers_is_a_partition():
    for total in (1, 7, 8, 4096, 4736):
        for world in (1, 2, 3, 8):
            got = [estimators.shard_walkers(total, r, world) for r in range(world)]
            ids = [i for f, n in got for i in range(f, f + n)]
            assert ids == list(range(total))
            assert max(n for _, n in got) - min(n for _, n in got) <= 1


def test_estimator_layout_roundtrip():
    lay = estimators.EstimatorLayout(5, 9)
    rng = np.random.default_rng(0)
    S = rng.normal(size=(5, 5))
    buf = lay.pack(S, rng.normal(size=5), rng.normal(size=5), np.arange(5.0), 3.0, -1.0, np.arange(9.0), 10, 40, 4)
    assert buf.size == lay.size == 25 + 15 + 2 + 9 + 3
    av = lay.averages(buf)
    assert np.allclose(av["localOperatorsMatrix"], S / 4)
    assert av["localEnergyR"] == 0.75 and av["nTrials"] == 40 and av["nSamples"] == 4
    assert np.allclose(av["localOperators"], np.arange(5.0) / 4)


_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from oracle_lib import Oracle
from tdvmc_b200 import estimators, systems
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "bosonsbulk_n64_equil.npz"))
spec = systems.from_golden(g)
o = Oracle(spec, time=float(g["time"]))
W, n_samples, n_therm = 6, 2, 32
first, n_local = estimators.shard_walkers(W, rank, world)
lay = estimators.EstimatorLayout(spec.n_params, 9)
est = np.zeros(o.est_size()); acc = 0
for w in range(first, first + n_local):
    r = o.sample_walker(g["R"] + 0.002 * w, g["uR"], g["uI"], float(g["phiR"]), 9, w, 0, 16, n_samples, n_therm, 0.4, est)
    acc += r["accepted"]
u = o.unpack_est(est, 1.0)
buf = lay.pack(u["S"], u["OER"], u["OEI"], u["O"], u["e_r"], u["e_i"], u["other"], acc, n_local * (16 + n_samples * n_therm),
               n_local * n_samples)
buf = estimators.allreduce_sum(buf)
if rank == 0:
    np.save(sys.argv[2], buf)
dist.destroy_process_group()
"""


def _run_world(tmp_path, world, tag):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    out = tmp_path / f"est_{tag}.npy"
    port = 29500 + (os.getpid() % 2000) + world
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, str(out)], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    return np.load(out)


def test_two_rank_reduction_equals_single_rank(tmp_path):
    """Walkers sharded over 2 gloo ranks + one packed all-reduce == the same ensemble on one rank
    (Philox streams are keyed by GLOBAL walker id, averages divide by the GLOBAL sample count)."""
    one = _run_world(tmp_path, 1, "w1")
    two = _run_world(tmp_path, 2, "w2")
    lay = estimators.EstimatorLayout(33, 9)
    a, b = lay.averages(one), lay.averages(two)
    assert a["nSamples"] == b["nSamples"] == 12 and a["nTrials"] == b["nTrials"] and a["nAcceptances"] == b["nAcceptances"]
    for k in ("localOperators", "localOperatorsMatrix", "localOperatorlocalEnergyR", "localOperatorlocalEnergyI",
              "otherExpectationValues"):
        assert np.allclose(a[k], b[k], rtol=1e-13, atol=1e-13), k
    assert abs(a["localEnergyR"] - b["localEnergyR"]) < 1e-12 * abs(a["localEnergyR"])


def _write_tables(path, g):
    with open(path, "w") as f:
        f.write(f"{int(g['N'])} {float(g['LBOX'])!r} {int(g['N_PARAM'])} {len(g['knots'])}\n")
        f.write(" ".join(repr(float(x)) for x in g["knots"]) + "\n")
        f.write(" ".join(repr(float(x)) for x in g["spline_weights"].ravel()) + "\n")


def test_cpp_host_adapter_builds_and_refuses_without_gpu(tmp_path, golden):
    """The C++ adapter (tdvmc_b200/host) compiles with plain g++ against the C ABI; without a CUDA device
    the driver stops with the library's error instead of computing anything on the CPU."""
    from tdvmc_b200 import capi

    _lib_path()
    host = os.path.join(ROOT, "tdvmc_b200", "host")
    subprocess.check_call(["make", "-C", host, "example_driver"])
    tables = tmp_path / "tables.txt"
    _write_tables(tables, golden("bosonsbulk_n64_fixture"))
    r = subprocess.run([os.path.join(host, "example_driver"), str(tables)], capture_output=True, text=True)
    if capi.load().tdvmc_gpu_device_count() > 0:
        assert r.returncode == 0 and "E_R=" in r.stdout
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


def test_cpp_he_table_builders_match_python_specs():
    """MakeHeBulkTables / MakeHeDropTables / MakeNUBosonsBulkPBBoxAndRadialTables (C++ adapter) build the same CSR parameter
    map, bit for bit, as tdvmc_b200.systems.he_bulk / he_drop / nu_bosons_bulk_pb_box_and_radial (which the golden fixtures
    pin against the reference)."""
    from tdvmc_b200 import systems as tsys
    host = os.path.join(ROOT, "tdvmc_b200", "host")
    subprocess.check_call(["make", "-C", host, "example_driver"])
    exe = os.path.join(host, "example_driver")
    br = tsys.nu_bosons_bulk_pb_box_and_radial(27, 3.0, 100, np.linspace(0.0, 1.5, 51), weights=np.zeros((53, 4, 4)))
    for args, spec in ((["hebulk", "64", "14.5", "40"], tsys.he_bulk(64, 14.5, 40)),
                       (["hedrop", "6", "150"], tsys.he_drop(6, 150)),
                       (["boxradial", "27", "3.0", "100"], br)):
        out = subprocess.run([exe, "--map"] + args, capture_output=True, text=True, check=True).stdout.splitlines()
        kind, P, n_ext, n_other, K = (int(x) for x in out[0].split())
        assert (kind, P, n_ext, n_other) == (spec.kind, spec.n_params, spec.n_ext, spec.n_other)
        assert K == spec.extra["n_splines"]
        np.testing.assert_array_equal(np.array(out[1].split(), dtype=np.int64), spec.map_ptr)
        np.testing.assert_array_equal(np.array(out[2].split(), dtype=np.int64), spec.map_col)
        np.testing.assert_array_equal(np.array(out[3].split(), dtype=np.float64), spec.map_val)
        for line, want in ((out[4], spec.map_const), (out[5], spec.grad_const)):   # empty in the adapter = zeros
            got = np.array(line.split(), dtype=np.float64)
            np.testing.assert_array_equal(got if got.size else np.zeros(P), want if want is not None else np.zeros(P))


@pytest.mark.parametrize("name", ["bosonsbulk_n64_evolution", "bosonsbulk_n64_evolution_realtime", "bosonsbulk_n64_evolution_rotation",
                                  "bosonsbulk_n64_evolution_qr", "bosonsbulk_n64_evolution_qr_raw"])
def test_timestep_solver_matches_reference(golden, name):
    """tdvmc_b200.timestep (host mirror of SolveForParametersDot, Cholesky branch) reproduces the derivatives the
    reference computed from its own first-step estimators (ref_harness evolve)."""
    from tdvmc_b200 import timestep
    g = golden(name)
    est = dict(localOperators=g["first_O"], localOperatorsMatrix=g["first_S"], localOperatorlocalEnergyR=g["first_OER"],
               localOperatorlocalEnergyI=g["first_OEI"], localEnergyR=float(g["first_ER"]), localEnergyI=float(g["first_EI"]))
    kw = {}
    if "LINEAR_EQUATION_SOLVER_TYPE" in g.files:       # the Eigen FullPivHouseholderQR branch (src/TDVMC.cpp:1763-1827)
        kw = dict(solver_type=int(g["LINEAR_EQUATION_SOLVER_TYPE"]), use_preconditioning=bool(int(g["USE_PRECONDITIONING"])))
    u_r, u_i, p_r, p_i = timestep.solve_for_parameters_dot(est, imaginary_time=int(g["IMAGINARY_TIME"]), **kw)
    # in real time the first step starts from uI = 0, so E^I = 0 and uDotR vanishes identically: one common scale
    scale = max(np.max(np.abs(g["first_uDotR"])), np.max(np.abs(g["first_uDotI"])))
    pscale = max(abs(float(g["first_phiDotR"])), abs(float(g["first_phiDotI"])))
    assert scale > 0 and pscale > 0
    assert np.max(np.abs(u_r - g["first_uDotR"])) <= 1e-9 * scale
    assert np.max(np.abs(u_i - g["first_uDotI"])) <= 1e-9 * scale
    assert abs(p_r - float(g["first_phiDotR"])) <= 1e-9 * pscale
    assert abs(p_i - float(g["first_phiDotI"])) <= 1e-9 * pscale
    if kw:
        return                                         # (what follows checks the Cholesky branch against LAPACK)
    v_r, v_i, q_r, q_i = timestep.solve_for_parameters_dot(est, imaginary_time=int(g["IMAGINARY_TIME"]), lapack=True)
    assert np.max(np.abs(v_r - u_r)) <= 1e-9 * scale and np.max(np.abs(v_i - u_i)) <= 1e-9 * scale
    assert abs(q_r - p_r) <= 1e-9 * pscale and abs(q_i - p_i) <= 1e-9 * pscale
    # numpy's own solver agrees with the hand-written Cholesky
    A, b_r, _ = timestep.build_system_of_equations(est, int(g["IMAGINARY_TIME"]))
    s = np.sqrt(np.diag(A))
    x = np.linalg.solve(A / np.outer(s, s) + 0.001 * np.eye(len(s)), b_r / s) / s
    assert np.max(np.abs(x - u_r)) <= 1e-8 * scale


def test_recorded_bench_line_has_the_contract_keys():
    """The last committed bench line (profiles/) carries every key the measurement contract asks for, with consistent
    values: guards the format bench.py prints (it cannot run without a GPU)."""
    import glob
    import json
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench.json")))
    assert paths, "no committed bench line under profiles/"
    d = json.loads(open(paths[-1]).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "walker-steps/s" and d["unit"] == "walker-steps/s" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] in ("weak", "strong")
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] <= d["value"] * 1.02
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["gpu_launches"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # r02 additions: ensemble-size records, secondary configs against the full host, and the
    # time-steps/s of the reference's own driver program next to a sample-matched reference figure
    assert d["config"]["walkers_per_gpu"] == d["config"]["walkers_per_gpu_survey"] == 4096
    assert d["roofline"]["traffic"] and "profiles/" in d["roofline"]["traffic_source"]
    for rec in ("full_wave", "strong"):
        assert d[rec]["value"] > 0 and d[rec]["walkers"] > 0
    assert d["strong"]["total_walkers"] == 23680
    assert [c["config"] for c in d["secondary"]][:4] == ["config/drop_6.config", "config/bulk_64.config", "config/NUBosonsBulkPB3D.config",
                                                          "config/He4He4Na.config"]
    assert all(c["reference_full_host"]["cores"] >= 1 and c["walker_steps_ratio_vs_full_host"] > 1 for c in d["secondary"])
    ts = d["time_step"]
    assert ts["driver_binary"]["host_solve_qr"]["time_steps_per_s"] > 0 and ts["driver_binary"]["device_solve_cholesky"]["time_steps_per_s"] > 0
    assert ts["reference_host"]["samples_per_time_step_matched"] == ts["driver_binary"]["host_solve_qr"]["samples_per_time_step"]
    # value = proposals of all walkers / time
    steps = d["config"]["proposals_per_walker_per_step"] * d["config"]["walkers"] * d["steps"]
    assert abs(steps / (d["ms_per_step"] * d["steps"] * 1e-3) - d["value"]) < 1e-6 * d["value"]


@pytest.mark.parametrize("name", ["mixture_he4he4na_equil", "mixture_he3he4cs_equil", "mixture4_he4he4na_equil", "inhcontact_n20_equil"])
def test_cpp_table_builders_with_reference_boundary_factors(golden, tmp_path, name):
    """MakeBosonMixtureClusterTables (cubic and quartic) and MakeInhContactBosonsTables (C++ adapter), fed the reference's
    own knots and boundary factors from the fixtures, build the same CSR parameter map bit for bit as the Python specs
    (which the golden fixtures pin against the reference)."""
    from tdvmc_b200 import systems as tsys
    g = golden(name)
    spec = tsys.from_golden(g)
    host = os.path.join(ROOT, "tdvmc_b200", "host")
    subprocess.check_call(["make", "-C", host, "example_driver"])
    f = tmp_path / "in.txt"
    fl = lambda a: " ".join(repr(float(x)) for x in np.asarray(a).ravel())
    with open(f, "w") as out:
        if spec.kind == tsys.KIND_MIXTURE:
            T, order = int(g["n_pair_types"]), int(spec.extra["order"])
            out.write(f"mixture {order} {spec.n_particles} {T}\n")
            out.write(" ".join(str(int(x)) for x in g["correlation_types"]) + "\n")
            pots = spec.extra["type_potential"]
            for t in range(T):
                out.write(f"{len(g[f'knots_{t}'])} {fl(g[f'knots_{t}'])}\n{fl(g[f'bc_factors_{t}'])}\n{float(g[f'extras_{t}'][6])!r} {int(pots[t])}\n")
        else:
            out.write(f"inhcontact {spec.n_particles} {spec.lbox!r} {spec.n_params}\n{fl(g['SYSTEM_PARAMS'])}\n")
            for q in ("spf", "pc"):
                bs, be, npq = g["bc_start_" + q], g["bc_end_" + q], g["np_" + q]
                out.write(f"{len(g['knots_' + q])} {fl(g['knots_' + q])}\n{len(bs)} {fl(bs)}\n{len(be)} {fl(be)}\n"
                          f"{int(npq[0])} {int(npq[1])} {int(npq[2])}\n")
    r = subprocess.run([os.path.join(host, "example_driver"), "--map", "file", str(f)], capture_output=True, text=True, check=True)
    o = r.stdout.splitlines()
    kind, P, n_ext, n_other, K = (int(x) for x in o[0].split())
    assert (kind, P, n_ext) == (spec.kind, spec.n_params, spec.n_ext) and K == spec.n_splines
    np.testing.assert_array_equal(np.array(o[1].split(), dtype=np.int64), spec.map_ptr)
    np.testing.assert_array_equal(np.array(o[2].split(), dtype=np.int64), spec.map_col)
    np.testing.assert_array_equal(np.array(o[3].split(), dtype=np.float64), spec.map_val)


# ---------------------------------------------------------------------------------------------------
# the reference's own driver bound to the library (tdvmc_b200/host/driver): host-side checks
# ---------------------------------------------------------------------------------------------------
TDVMC_REF = os.path.join(ROOT, "oracle", "_ref", "TDVMC_ref")


def test_driver_patch_applies_to_the_reference_source(tmp_path):
    """patch_driver.py finds every anchor exactly once in the reference's src/TDVMC.cpp (only where /root/reference exists:
    the build container) and changes nothing but the documented call sites."""
    src = "/root/reference/src/TDVMC.cpp"
    if not os.path.exists(src):
        pytest.skip("the reference tree is not on this machine (GPU box): the binary was built in the build container")
    out = tmp_path / "TDVMC_gpu.cpp"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tdvmc_b200", "host", "driver", "patch_driver.py"), src, str(out)])
    import difflib
    a = open(src).read().splitlines()
    b = open(out).read().splitlines()[2:]
    added = [l[2:] for l in difflib.ndiff(a, b) if l.startswith("+ ")]
    removed = [l[2:] for l in difflib.ndiff(a, b) if l.startswith("- ")]
    assert len(removed) == 2 and len(added) <= 40                  # two lines re-pointed in place, the rest are insertions
    assert all("gpu" in l.lower() or l.strip() in ("{", "}", "return;") for l in added), added


def test_driver_with_gpu_walkers_zero_is_the_reference_byte_for_byte(golden, tmp_path):
    """GPU_WALKERS = 0 (or absent): every hook falls through and TDVMC_gpu IS the reference - its .dat files equal those of
    the unmodified TDVMC_ref byte for byte (same host RNG stream, same CPU path).  No GPU needed."""
    from tdvmc_b200 import driver
    if not (os.path.exists(TDVMC_REF) and os.path.exists(driver.TDVMC_GPU)):
        pytest.skip("driver binaries not built (needs /root/reference at build time)")
    g = golden("bosonsbulk_n64_equil")
    cfg = driver.base_config(N=64, LBOX=4.0, N_PARAM=33, MC_STEP=0.4, MC_NSTEPS=64, MC_NTHERMSTEPS=32, MC_NINITIALIZATIONSTEPS=64,
                             MC_VERY_FIRST_NINITIALIZATIONSTEPS=640, TIMESTEP=1e-5, TOTALTIME=1e-5 * 2.5, IMAGINARY_TIME=1,
                             USE_PRECONDITIONING=1, PARAMS_REAL=[float(x) for x in g["uR"]], SYSTEM_PARAMS=[1.0, 1.0],
                             MC_NADDITIONALSTEPS=4, MC_NADDITIONALTHERMSTEPS=16, MC_NADDITIONALINITIALIZATIONSTEPS=16)
    a = driver.run_driver(TDVMC_REF, cfg, str(tmp_path / "ref"), R0=g["R"], seed=5)
    b = driver.run_driver(driver.TDVMC_GPU, cfg, str(tmp_path / "gpu_off"), R0=g["R"], seed=5)
    assert len(a.local_energy_r) == 3 and np.all(np.isfinite(a.parameters_r))
    for name in ("LocalEnergyR", "LocalEnergyI", "LocalOperators", "OtherExpectationValues", "ParametersR", "ParametersI", "timesSystem",
                 "AdditionalObservables_pairDistribution", "AdditionalObservables_structureFactor"):
        pa, pb = os.path.join(a.out_dir, name + ".dat"), os.path.join(b.out_dir, name + ".dat")
        if name.startswith("Additional") and not os.path.exists(pa):
            continue
        assert open(pa, "rb").read() == open(pb, "rb").read(), name


def test_driver_refuses_gpu_walkers_without_a_device(golden, tmp_path):
    """GPU_WALKERS > 0 on a machine without a CUDA device: the driver stops with the library's message, it does not fall
    back to the CPU path silently."""
    from tdvmc_b200 import capi, driver
    if not os.path.exists(driver.TDVMC_GPU):
        pytest.skip("driver binary not built")
    if capi.load().tdvmc_gpu_device_count() > 0:
        pytest.skip("this machine has a CUDA device")
    g = golden("bosonsbulk_n64_equil")
    cfg = driver.base_config(N=64, LBOX=4.0, N_PARAM=33, PARAMS_REAL=[float(x) for x in g["uR"]], GPU_WALKERS=8)
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        driver.run_driver(driver.TDVMC_GPU, cfg, str(tmp_path / "nogpu"), R0=g["R"])
