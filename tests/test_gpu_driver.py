"""The reference's own driver bound to the device (tdvmc_b200/host/build/TDVMC_gpu = src/TDVMC.cpp with the call sites of
INTEGRATION.md section 2 re-pointed, built by tdvmc_b200/host/driver/Makefile) against the unmodified reference program
(oracle/_ref/TDVMC_ref): whole time evolutions - config file in, LocalEnergyR.dat / ParametersR.dat out - compared within
error bars (north_star level 2).  The reference runs ONE Markov chain per process; its error bars come from the spread over
RNG seeds (RNG state files), the device's from the spread over GPU_SEEDs.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from tdvmc_b200 import driver, systems

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TDVMC_REF = os.path.join(ROOT, "oracle", "_ref", "TDVMC_ref")


@pytest.fixture(scope="module")
def binaries():
    from tdvmc_b200 import capi

    if capi.load().tdvmc_gpu_device_count() <= 0:
        pytest.skip("no CUDA device on this machine")
    assert os.path.exists(driver.TDVMC_GPU), "build it: make -C tdvmc_b200/host/driver (done by __graft_entry__.build())"
    assert os.path.exists(TDVMC_REF), "build it: make -C oracle/ref_build"
    return driver.TDVMC_GPU, TDVMC_REF


def run_seeds(binary, cfg, tag, R0, seeds, tmp_path, gpu_seed=False, workers=None):
    """One process per seed; for the device arm the seed is GPU_SEED and the runs go one after the other."""
    def one(sd):
        c = dict(cfg, GPU_SEED=sd) if gpu_seed else cfg
        return driver.run_driver(binary, c, str(tmp_path / f"{tag}_{sd}"), R0=R0, seed=sd, timeout=1800)

    with ThreadPoolExecutor(max_workers=1 if gpu_seed else (workers or min(len(seeds), os.cpu_count() or 1))) as ex:
        return list(ex.map(one, seeds))


def compare_trajectories(ref, dev, n_par, parity_log, case, z_max=6.5, gauge=None):
    """Mean trajectories of the two arms against the standard error of their difference, each arm with its own seed-to-seed
    spread (Welch): z = (mean_dev - mean_ref) / sqrt(s_ref^2 / n_ref + s_dev^2 / n_dev).  With 6-8 seeds per arm z follows a
    t distribution with ~10 degrees of freedom, hence the bound of 6.5 on the maximum over some thousand (strongly
    correlated) values; the normalised deviations must ALSO be of unit size on average - a test that could not fail proves
    nothing, and spreads that differed between the arms would show here.  `gauge`: a direction in parameter space that leaves
    the wave function unchanged; the parameters are compared with their component along it removed."""
    out = {}
    if gauge is not None:
        gauge = np.asarray(gauge, float) / np.linalg.norm(gauge)
    for name, get in (("E_R", lambda r: r.local_energy_r[:, None]), ("uR", lambda r: r.parameters_r[:, :n_par]),
                      ("uI", lambda r: r.parameters_i[:, :n_par])):
        a = np.stack([get(r) for r in ref])            # [seed][step][cols]
        b = np.stack([get(r) for r in dev])
        assert a.shape[1:] == b.shape[1:], (name, a.shape, b.shape)
        assert np.all(np.isfinite(a)) and np.all(np.isfinite(b)), name
        if name != "E_R":
            a, b = a[:, 1:], b[:, 1:]                    # step 0 holds the start parameters: identical, no spread
            if gauge is not None:
                a = a - (a @ gauge)[..., None] * gauge
                b = b - (b @ gauge)[..., None] * gauge
        sa, sb = a.std(axis=0, ddof=1), b.std(axis=0, ddof=1)
        live = (sa > 0) & (sb > 0)
        if not np.any(live):
            continue
        z = (b.mean(axis=0) - a.mean(axis=0))[live] / np.sqrt(sa[live] ** 2 / len(ref) + sb[live] ** 2 / len(dev))
        out[name] = z
        rms = np.sqrt(np.mean(z ** 2))
        parity_log.check("driver_evolution", case, f"max|z| {name}", np.max(np.abs(z)), 1.0, z_max,
                         f"{z.size} values, rms z = {rms:.2f}, spread dev/ref = {np.median(sb[live] / sa[live]):.2f}")
        if z.size >= 100:
            assert 0.4 < rms < 1.8, (name, rms)
        else:                                            # a handful of strongly correlated energies: effectively ONE t-distributed
            assert rms < 3.5, (name, rms)                # value (24 seeds per arm on this case: |z| <= 1.6, profiles/r02_driver_stability.txt)
    return out


def test_driver_evolution_n64_matches_reference_program(binaries, golden, parity_log, tmp_path):
    """BosonsBulk N = 64: ten imaginary-time Euler steps (Cholesky solve with preconditioning) of the reference program,
    eight seeds, against the same config through TDVMC_gpu (GPU_WALKERS = 512, two samples each = the reference's 1024
    samples per step), eight device seeds."""
    gpu_bin, ref_bin = binaries
    g = golden("bosonsbulk_n64_equil")
    cfg = driver.base_config(N=64, LBOX=4.0, N_PARAM=33, MC_STEP=0.4, MC_NSTEPS=1024, MC_NTHERMSTEPS=64, MC_NINITIALIZATIONSTEPS=64,
                             MC_VERY_FIRST_NINITIALIZATIONSTEPS=6400, TIMESTEP=2e-4, TOTALTIME=2e-4 * 9.5, IMAGINARY_TIME=1,
                             USE_PRECONDITIONING=1, PARAMS_REAL=[float(x) for x in g["uR"]], SYSTEM_PARAMS=[1.0, 1.0])
    ref = run_seeds(ref_bin, cfg, "ref", g["R"], list(range(1, 9)), tmp_path)
    dev = run_seeds(gpu_bin, dict(cfg, GPU_WALKERS=512, MC_NSTEPS=2), "gpu", g["R"], list(range(1, 9)), tmp_path, gpu_seed=True)
    assert len(dev[0].local_energy_r) == 10 and dev[0].parameters_r.shape == (10, 34)
    z = compare_trajectories(ref, dev, 33, parity_log, "bosonsbulk_n64_euler_imaginary")
    assert "uR" in z and "E_R" in z
    # the evolution is not noise against noise: the parameters moved by many standard errors over the run
    a = np.stack([r.parameters_r[:, :33] for r in ref])
    moved = np.abs(a[:, -1] - a[:, 0]).mean(axis=0) / (a[:, -1].std(axis=0, ddof=1) + 1e-300)
    assert np.max(moved) > 10.0
    # acceptance rate and the per-step log line of the driver survive the binding
    assert abs(np.mean(dev[0].acceptance) - np.mean(ref[0].acceptance)) < 1.0
    assert len(dev[0].step_ms) == 10


def test_driver_real_time_qr_branch_n64(binaries, golden, parity_log, tmp_path):
    """Real time with LINEAR_EQUATION_SOLVER_TYPE = 1: the driver's own Eigen FullPivHouseholderQR branch
    (src/TDVMC.cpp:1763-1827) solves on the estimators the device fetched - the branch the device solver does not offer."""
    gpu_bin, ref_bin = binaries
    g = golden("bosonsbulk_n64_equil")
    cfg = driver.base_config(N=64, LBOX=4.0, N_PARAM=33, MC_STEP=0.4, MC_NSTEPS=1024, MC_NTHERMSTEPS=64, MC_NINITIALIZATIONSTEPS=64,
                             MC_VERY_FIRST_NINITIALIZATIONSTEPS=6400, TIMESTEP=1e-4, TOTALTIME=1e-4 * 5.5, IMAGINARY_TIME=0,
                             LINEAR_EQUATION_SOLVER_TYPE=1, USE_PRECONDITIONING=0, PARAMS_REAL=[float(x) for x in g["uR"]],
                             SYSTEM_PARAMS=[1.0, 1.0])
    ref = run_seeds(ref_bin, cfg, "ref", g["R"], list(range(1, 9)), tmp_path)
    dev = run_seeds(gpu_bin, dict(cfg, GPU_WALKERS=512, MC_NSTEPS=2), "gpu", g["R"], list(range(1, 9)), tmp_path, gpu_seed=True)
    z = compare_trajectories(ref, dev, 33, parity_log, "bosonsbulk_n64_euler_realtime_qr")
    assert "uI" in z                                     # the imaginary parts grow out of zero in real time


def test_driver_sample_reuse_nubosons_n216(binaries, golden, parity_log, tmp_path):
    """config 4's mode at reduced size: NUBosonsBulkPB (non-uniform knots, reflection rule), UPDATE_SAMPLES_EVERY_NTH_STEP = 1
    - step 0 samples and stores, every later step advances the stored samples (UpdateSamplesConsecutive) and re-evaluates them
    (ParallelUpdateExpectationValuesForGivenSamples), here with stored R recomputed on the device instead of 11 MB tables."""
    gpu_bin, ref_bin = binaries
    g = golden("nubosonsbulkpb_n216_equil")
    P = int(g["N_PARAM"])
    cfg = driver.nubosons_config(g["NURBS_GRID"], g["uR"], np.zeros(P), N=216, LBOX=float(g["LBOX"]), N_PARAM=P,
                                 SYSTEM_PARAMS=[float(x) for x in g["SYSTEM_PARAMS"]], MC_STEP=0.35, MC_NSTEPS=512, MC_NTHERMSTEPS=108,
                                 MC_NINITIALIZATIONSTEPS=216, MC_VERY_FIRST_NINITIALIZATIONSTEPS=21600, TIMESTEP=2e-5,
                                 TOTALTIME=2e-5 * 5.5, IMAGINARY_TIME=1, UPDATE_SAMPLES_PERCENT=100.0)
    ref = run_seeds(ref_bin, cfg, "ref", g["R"], list(range(1, 9)), tmp_path)
    dev = run_seeds(gpu_bin, dict(cfg, GPU_WALKERS=256, MC_NSTEPS=2), "gpu", g["R"], list(range(1, 9)), tmp_path, gpu_seed=True)
    assert len(dev[0].local_energy_r) == 6
    compare_trajectories(ref, dev, P, parity_log, "nubosonsbulkpb_n216_sample_reuse")


def test_driver_config3_own_size_against_reference_trajectories(binaries, golden, parity_log, tmp_path):
    """BASELINE configs[2] at its own size (N = 343, N_PARAM = 201): the device-bound driver against trajectories of the
    reference program recorded by oracle/gen_driver_fixtures.py (six seeds, ten Euler steps, 4096 samples per step - 160 s
    per step and host core, which is why they are a fixture and not run here), then the run continued to 50 steps."""
    gpu_bin, _ = binaries
    f = golden("driver_cfg3_reference")
    g = golden("bosonsbulk_n343_equil")
    uR, uI = systems.smooth_params(201, 3.5)
    n_steps = f["e_r"].shape[1]
    cfg = driver.headline_config(uR, uI, MC_NSTEPS=2, MC_NTHERMSTEPS=int(f["MC_NTHERMSTEPS"]), MC_NINITIALIZATIONSTEPS=1000,
                                 MC_VERY_FIRST_NINITIALIZATIONSTEPS=34300, TIMESTEP=float(f["TIMESTEP"]),
                                 TOTALTIME=float(f["TIMESTEP"]) * (n_steps - 0.5), LINEAR_EQUATION_SOLVER_TYPE=0, USE_PRECONDITIONING=1,
                                 GPU_WALKERS=int(f["MC_NSTEPS"]) // 2)
    dev = run_seeds(gpu_bin, cfg, "gpu", g["R"], list(range(1, 9)), tmp_path, gpu_seed=True)

    class Ref:                                            # the fixture's series under the DriverRun attribute names
        def __init__(self, i):
            self.local_energy_r, self.parameters_r, self.parameters_i = f["e_r"][i], f["p_r"][i], f["p_i"][i]

    compare_trajectories([Ref(i) for i in range(f["e_r"].shape[0])], dev, 201, parity_log, "bosonsbulk_n343_p201_euler")
    # 50 Euler steps with the ensemble the headline benchmark uses: stable, energy relaxing in imaginary time
    long = driver.run_driver(gpu_bin, dict(cfg, GPU_WALKERS=4096, TOTALTIME=float(f["TIMESTEP"]) * 49.5), str(tmp_path / "long"),
                             R0=g["R"])
    e = long.local_energy_r
    assert len(e) == 50 and np.all(np.isfinite(e)) and np.all(np.isfinite(long.parameters_r))
    assert e[-10:].mean() < e[:10].mean()
    assert np.median(long.step_ms) < 200.0


def test_driver_config4_own_size_runs(binaries, golden, tmp_path):
    """BASELINE configs[3] at its own size and counts (NUBosonsBulkPB N = 1728, N_PARAM = 200, 50 samples x 200 steps, sample
    reuse on) through the device-bound driver: three time steps, finite, and the step-0 energy agrees with the fixed-parameter
    sampler statistics of the same ensemble."""
    gpu_bin, _ = binaries
    g = golden("nubosonsbulkpb_n1728_equil")
    cfg = driver.nubosons_config(g["NURBS_GRID"], g["uR"], g["uI"], MC_VERY_FIRST_NINITIALIZATIONSTEPS=17280, GPU_WALKERS=64,
                                 TOTALTIME=1e-5 * 2.5, SYSTEM_PARAMS=[float(x) for x in g["SYSTEM_PARAMS"]])
    r = driver.run_driver(gpu_bin, cfg, str(tmp_path / "cfg4"), R0=g["R"])
    assert len(r.local_energy_r) == 3 and np.all(np.isfinite(r.local_energy_r)) and np.all(np.isfinite(r.parameters_r))
    assert abs(r.local_energy_r[1] - r.local_energy_r[0]) < 0.02 * abs(r.local_energy_r[0])


@pytest.mark.parametrize("solver_type,tol", [(0, 1e-9), (1, 1e-7)])
def test_device_solve_option_matches_host_solve(binaries, golden, tmp_path, solver_type, tol):
    """GPU_DEVICE_SOLVE = 1: SolveForParametersDot of the Euler step served by the device - solve_kernel (the hand-written
    Cholesky, bit-identical) or solve_qr_kernel (Eigen's FullPivHouseholderQR step by step) - against the driver's own host
    solve on the same ensemble: the trajectories coincide to rounding (times the condition number for QR)."""
    gpu_bin, _ = binaries
    g = golden("bosonsbulk_n64_equil")
    cfg = driver.base_config(N=64, LBOX=4.0, N_PARAM=33, MC_STEP=0.4, MC_NSTEPS=2, MC_NTHERMSTEPS=64, MC_NINITIALIZATIONSTEPS=64,
                             MC_VERY_FIRST_NINITIALIZATIONSTEPS=6400, TIMESTEP=2e-4, TOTALTIME=2e-4 * 5.5, IMAGINARY_TIME=1,
                             LINEAR_EQUATION_SOLVER_TYPE=solver_type, USE_PRECONDITIONING=1, PARAMS_REAL=[float(x) for x in g["uR"]],
                             SYSTEM_PARAMS=[1.0, 1.0], GPU_WALKERS=512)
    a = driver.run_driver(gpu_bin, cfg, str(tmp_path / "host"), R0=g["R"])
    b = driver.run_driver(gpu_bin, dict(cfg, GPU_DEVICE_SOLVE=1), str(tmp_path / "dev"), R0=g["R"])
    assert len(a.local_energy_r) == 6 and np.max(np.abs(a.parameters_r[-1, :33] - a.parameters_r[0, :33])) > 1e-4
    assert np.max(np.abs(a.parameters_r[:, :33] - b.parameters_r[:, :33])) < tol
    assert np.max(np.abs(a.local_energy_r - b.local_energy_r)) < 10 * tol * np.max(np.abs(a.local_energy_r))


def test_driver_hebulk_config2(binaries, golden, parity_log, tmp_path):
    """BASELINE configs[1], config/bulk_64.config with its own values (HeBulk N = 64, N_PARAM = 49, MC_NSTEPS = 2000 x
    MC_NTHERMSTEPS = 100, imaginary-time Euler steps of 1e-5, Cholesky without preconditioning, the shipped PARAMS_REAL) through
    both programs: five time steps, eight seeds each."""
    gpu_bin, ref_bin = binaries
    g = golden("hebulk_n64_equil")
    cfg = driver.base_config(SYSTEM_TYPE="HeBulk", N=64, LBOX=float(g["LBOX"]), N_PARAM=49, RHO=0.0219, RC=8.8, MC_STEP=0.3,
                             MC_NSTEPS=2000, MC_NTHERMSTEPS=100, MC_NINITIALIZATIONSTEPS=100, MC_VERY_FIRST_NINITIALIZATIONSTEPS=6400,
                             TIMESTEP=1e-5, TOTALTIME=1e-5 * 4.5, IMAGINARY_TIME=1, ODE_SOLVER_TYPE=0, LINEAR_EQUATION_SOLVER_TYPE=0,
                             USE_PRECONDITIONING=0, USE_NORMALIZE_WF=0, SYSTEM_PARAMS=[0.0], PARAMS_REAL=[float(x) for x in g["uR"]],
                             PARAMS_IMAGINARY=[0.0] * 49, PARAM_PHIR=float(g["phiR"]))
    ref = run_seeds(ref_bin, cfg, "ref", g["R"], list(range(1, 9)), tmp_path)
    dev = run_seeds(gpu_bin, dict(cfg, GPU_WALKERS=1000, MC_NSTEPS=2), "gpu", g["R"], list(range(1, 9)), tmp_path, gpu_seed=True)
    assert len(dev[0].local_energy_r) == 5
    compare_trajectories(ref, dev, 49, parity_log, "hebulk_n64_config2")


def test_driver_hedrop_config1_first_pass(binaries, golden, parity_log, tmp_path):
    """BASELINE configs[0], config/drop_6.config (HeDrop, six atoms, open boundary, N_PARAM = 93, MC_NSTEPS = 10000 x
    MC_NTHERMSTEPS = 50): the estimator pass of the first time step through both programs - <E^R> and <O_k> as the driver
    writes them.  (With the shipped parameters the droplet is unbound and the reference's own evolution drifts, DESIGN.md 2,
    so only the fixed-parameter pass is compared.)"""
    gpu_bin, ref_bin = binaries
    g = golden("hedrop_n6_equil")
    cfg = driver.base_config(SYSTEM_TYPE="HeDrop", N=6, LBOX=float(g["LBOX"]), N_PARAM=93, RHO=0.015, RC=8.8, MC_STEP=0.5,
                             MC_NSTEPS=10000, MC_NTHERMSTEPS=50, MC_NINITIALIZATIONSTEPS=1000, MC_VERY_FIRST_NINITIALIZATIONSTEPS=10000,
                             TIMESTEP=2e-5, TOTALTIME=0.0, IMAGINARY_TIME=1, USE_PRECONDITIONING=0, USE_NORMALIZE_WF=0, SYSTEM_PARAMS=[0.0],
                             RHO_BIN_COUNT=200, GR_BIN_COUNT=200, PARAMS_REAL=[float(x) for x in g["uR"]], PARAMS_IMAGINARY=[0.0] * 93,
                             PARAM_PHIR=float(g["phiR"]))
    ref = run_seeds(ref_bin, cfg, "ref", g["R"], list(range(1, 9)), tmp_path)
    dev = run_seeds(gpu_bin, dict(cfg, GPU_WALKERS=5000, MC_NSTEPS=2), "gpu", g["R"], list(range(1, 9)), tmp_path, gpu_seed=True)
    for name, get in (("E_R", lambda r: np.atleast_1d(r.local_energy_r)[:1]), ("O_k", lambda r: np.atleast_2d(r.local_operators)[0])):
        a, b = np.stack([get(r) for r in ref]), np.stack([get(r) for r in dev])
        sa, sb = a.std(axis=0, ddof=1), b.std(axis=0, ddof=1)
        live = (sa > 0) & (sb > 0)
        z = (b.mean(axis=0) - a.mean(axis=0))[live] / np.sqrt(sa[live] ** 2 / len(ref) + sb[live] ** 2 / len(dev))
        parity_log.check("driver_first_pass", "hedrop_n6_config1", f"max|z| {name}", np.max(np.abs(z)), 1.0, 6.5, f"{z.size} values")


def _mixture_base(g, system_type):
    P = int(g["N_PARAM"])
    return driver.base_config(SYSTEM_TYPE=system_type, N=3, LBOX=float(g["LBOX"]), N_PARAM=P, RHO=0.015, RC=8.8, MC_STEP=4.0,
                              MC_NTHERMSTEPS=20, MC_NINITIALIZATIONSTEPS=1000, MC_VERY_FIRST_NINITIALIZATIONSTEPS=100000,
                              TIMESTEP=1e-4, IMAGINARY_TIME=1, ODE_SOLVER_TYPE=0, LINEAR_EQUATION_SOLVER_TYPE=0, USE_PRECONDITIONING=1,
                              USE_NORMALIZE_WF=1, GR_BIN_COUNT=400, USE_NURBS=1, NURBS_GRID=[float(x) for x in g["NURBS_GRID"]],
                              PARTICLE_TYPES=[int(x) for x in g["PARTICLE_TYPES"]], SYSTEM_PARAMS=[float(x) for x in g["SYSTEM_PARAMS"]],
                              PARAMS_REAL=[float(x) for x in g["uR"]], PARAMS_IMAGINARY=[0.0] * P, PARAM_PHIR=float(g["phiR"]))


def _mixture_observable_pass(binaries, base, g, case, parity_log, tmp_path):
    """The config as shipped (TOTALTIME < 0): no time loop, only the end-of-run observable pass (counts reduced 25-fold on the
    reference arm), AdditionalObservables_*.dat of the two programs within error bars."""
    gpu_bin, ref_bin = binaries
    obs = dict(base, TOTALTIME=-1e-4, MC_NSTEPS=100, MC_NADDITIONALSTEPS=20000, MC_NADDITIONALTHERMSTEPS=100,
               MC_NADDITIONALINITIALIZATIONSTEPS=10000)
    ref = run_seeds(ref_bin, obs, "ref_obs", g["R"], list(range(1, 9)), tmp_path)
    dev = run_seeds(gpu_bin, dict(obs, GPU_WALKERS=2000, MC_NADDITIONALSTEPS=10), "gpu_obs", g["R"], list(range(1, 9)), tmp_path, gpu_seed=True)
    assert "GPU ensemble: 2000 walkers" in dev[0].log

    def observable(run, name):
        path = os.path.join(run.out_dir, f"AdditionalObservables_{name}.dat")
        if name == "r2":
            return np.array([float(open(path).read().split()[0])])
        return driver.read_dat(path)[:, 1:].ravel()          # first column: the grid

    for name in ("r2", "angularDistribution", "densityFromCOM", "particleDistances"):
        a, b = np.stack([observable(r, name) for r in ref]), np.stack([observable(r, name) for r in dev])
        assert a.shape == b.shape and np.all(np.isfinite(b))
        sa, sb = a.std(axis=0, ddof=1), b.std(axis=0, ddof=1)
        live = (sa > 0) & (sb > 0) & (a.mean(axis=0) > 0.02 * a.mean(axis=0).max())   # bins that are actually populated
        z = (b.mean(axis=0) - a.mean(axis=0))[live] / np.sqrt(sa[live] ** 2 / len(ref) + sb[live] ** 2 / len(dev))
        parity_log.check("driver_observables", case, f"max|z| {name}", np.max(np.abs(z)), 1.0, 6.5,
                         f"{z.size} values, rms z = {np.sqrt(np.mean(z ** 2)):.2f}")
        assert np.sqrt(np.mean(z ** 2)) < 1.8


def _mixture_evolution(binaries, base, g, case, parity_log, tmp_path):
    """Four imaginary-time Euler steps at the config's ratio of 20 steps per sample, without preconditioning."""
    gpu_bin, ref_bin = binaries
    evo = dict(base, TOTALTIME=1e-4 * 3.5, MC_NSTEPS=20000, USE_PRECONDITIONING=0)
    ref = run_seeds(ref_bin, evo, "ref_evo", g["R"], list(range(1, 9)), tmp_path)
    dev = run_seeds(gpu_bin, dict(evo, GPU_WALKERS=2000, MC_NSTEPS=10), "gpu_evo", g["R"], list(range(1, 9)), tmp_path, gpu_seed=True)
    assert len(dev[0].local_energy_r) == 4
    compare_trajectories(ref, dev, int(g["N_PARAM"]), parity_log, case)


def test_driver_mixture_config5(binaries, golden, parity_log, tmp_path):
    """BASELINE configs[4], config/He4He4Na.config (BosonMixtureCluster: two He-4 and one Na, two pair types, HFD-B and KTTY
    potentials).  The config ships TOTALTIME < 0: its whole workload is the end-of-run observable pass (r^2, corner angles,
    density from the centre of mass, pair distances; BosonMixtureCluster.cpp:680-741), written to AdditionalObservables_*.dat -
    compared between the two programs, and so is a short imaginary-time evolution with the config's sampling ratio."""
    gpu_bin, ref_bin = binaries
    g = golden("mixture_he4he4na_equil")
    base = _mixture_base(g, "BosonMixtureCluster")
    _mixture_observable_pass(binaries, base, g, "mixture_he4he4na_config5", parity_log, tmp_path)
    # With the config's USE_PRECONDITIONING = 1 the reference divides by the zero variance of operators whose knot interval is
    # never visited and stops with "Energy not finite" after the first step - and so does the device-bound program, fed the same
    # estimators; without it both evolve
    both = [driver.run_driver(b, dict(base, TOTALTIME=1e-4 * 3.5, MC_NSTEPS=ns, GPU_WALKERS=w), str(tmp_path / f"nan_{i}"), R0=g["R"], seed=2)
            for i, (b, ns, w) in enumerate(((ref_bin, 20000, 0), (gpu_bin, 10, 2000)))]
    for r in both:
        assert len(r.local_energy_r) == 2 and np.isfinite(r.local_energy_r[0]) and np.isnan(r.local_energy_r[1])
        assert "Energy not finite" in r.log
    _mixture_evolution(binaries, base, g, "mixture_he4he4na_config5", parity_log, tmp_path)


def test_driver_mixture_4thorder(binaries, golden, parity_log, tmp_path):
    """config/He4He4Na_4thOrder.config (BosonMixtureCluster_4thorder: quartic splines, 28 per pair type, 5 x 3 boundary-condition
    factors): the same two comparisons as for config 5."""
    g = golden("mixture4_he4he4na_equil")
    base = _mixture_base(g, "BosonMixtureCluster_4thorder")
    _mixture_observable_pass(binaries, base, g, "mixture4_he4he4na", parity_log, tmp_path)
    _mixture_evolution(binaries, base, g, "mixture4_he4he4na", parity_log, tmp_path)


def test_driver_inhcontact_bosons(binaries, golden, parity_log, tmp_path):
    """config/InhContactBosons.config with its own values (three bosons on a ring of length 3 in one dimension, N_PARAM = 62 =
    31 single-particle + 31 pair parameters plus the phase pair, MC_NSTEPS = 1000 x MC_NTHERMSTEPS = 100, imaginary-time Euler
    steps of 2e-4 from all-zero parameters, FullPivHouseholderQR, acceptance check 5, USE_NORMALIZE_WF) - the shipped
    SYSTEM_PARAMS {0, 10, 0, 0} switch both the contact interaction and the lattice off (gamma = 10 k pi with k = 0,
    InhContactBosons.cpp:26-29: every estimator is identically zero in both programs), so k = 1, V0 = 2 of the golden fixture are
    used: ten time steps through both programs."""
    gpu_bin, ref_bin = binaries
    g = golden("inhcontact_n3_equil")
    cfg = driver.base_config(SYSTEM_TYPE="InhContactBosons", N=3, LBOX=3.0, DIM=1, N_PARAM=62, MC_STEP=0.5, MC_NSTEPS=1000,
                             MC_NTHERMSTEPS=100, MC_NINITIALIZATIONSTEPS=1000, MC_VERY_FIRST_NINITIALIZATIONSTEPS=100000, TIMESTEP=2e-4,
                             TOTALTIME=2e-4 * 9.5, IMAGINARY_TIME=1, ODE_SOLVER_TYPE=0, LINEAR_EQUATION_SOLVER_TYPE=1, USE_PRECONDITIONING=0,
                             USE_PARAMETER_ACCEPTANCE_CHECK=1, PARAMETER_ACCEPTANCE_CHECK_TYPE=5, USE_NORMALIZE_WF=1, GR_BIN_COUNT=50,
                             RHO_BIN_COUNT=50, SYSTEM_PARAMS=[0.0, 10.0, 1.0, 2.0])
    R0 = np.asarray(g["R"])[:, :1]
    ref = run_seeds(ref_bin, cfg, "ref", R0, list(range(1, 9)), tmp_path)
    dev = run_seeds(gpu_bin, dict(cfg, GPU_WALKERS=500, MC_NSTEPS=2), "gpu", R0, list(range(1, 9)), tmp_path, gpu_seed=True)
    assert "GPU ensemble: 500 walkers" in dev[0].log
    assert len(dev[0].local_energy_r) == 10
    # Sum_k B_k = 1 on both knot vectors: a constant added to the 31 single-particle parameters and subtracted from the 31
    # pair parameters ((N - 1) / 2 = 1 pair per particle) changes nothing but the norm.  S is exactly singular along that
    # direction, the QR solution's component along it is rounding noise over rounding noise - the reference's own seeds
    # end +-12 apart there after ten steps - so it is projected out before the comparison
    compare_trajectories(ref, dev, 62, parity_log, "inhcontact_n3", gauge=[1.0] * 31 + [-1.0] * 31)
    # the shipped values themselves: both programs report zero energy and leave the parameters at zero
    flat = dict(cfg, SYSTEM_PARAMS=[0.0, 10.0, 0.0, 0.0], TOTALTIME=2e-4 * 2.5)
    for r in (driver.run_driver(ref_bin, flat, str(tmp_path / "flat_ref"), R0=R0, seed=1),
              driver.run_driver(gpu_bin, dict(flat, GPU_WALKERS=500, MC_NSTEPS=2), str(tmp_path / "flat_gpu"), R0=R0, seed=1)):
        assert np.all(r.local_energy_r == 0.0) and np.all(r.parameters_r[:, :62] == 0.0)


def test_driver_boxandradial_n27(binaries, golden, parity_log, tmp_path):
    """config/NUBosonsBulkPBBoxAndRadial3D.config (N = 27, N_PARAM = 100 = 50 box + 50 radial parameters, MC_NSTEPS = 10 x
    MC_NTHERMSTEPS = 125, imaginary-time Euler, preconditioned Cholesky, acceptance check 5); the shipped file leaves LBOX, the
    NURBS grid, SYSTEM_PARAMS and the parameters empty, the golden fixture's are used (LBOX = 3, 51 knots, {0.1, 50}).  With that
    parameter set the reference's own evolution blows up at the config's step of 1e-4 ("NOT POSITIVE SEMI DEFINITE" after two
    steps), and with 1000 samples one seed in eight never visits the innermost knot interval (zero variance, the preconditioner
    divides by it), so six steps of 1e-6 with 4000 samples each go through both programs."""
    gpu_bin, ref_bin = binaries
    g = golden("boxradial_n27_equil")
    cfg = driver.base_config(SYSTEM_TYPE="NUBosonsBulkPBBoxAndRadial", N=27, LBOX=float(g["LBOX"]), DIM=3, N_PARAM=100, MC_STEP=0.5,
                             MC_NSTEPS=4000, MC_NTHERMSTEPS=125, MC_NINITIALIZATIONSTEPS=250, MC_VERY_FIRST_NINITIALIZATIONSTEPS=10000,
                             TIMESTEP=1e-6, TOTALTIME=1e-6 * 5.5, IMAGINARY_TIME=1, ODE_SOLVER_TYPE=0, LINEAR_EQUATION_SOLVER_TYPE=0,
                             USE_PRECONDITIONING=1, USE_PARAMETER_ACCEPTANCE_CHECK=1, PARAMETER_ACCEPTANCE_CHECK_TYPE=5, USE_NORMALIZE_WF=1,
                             GR_BIN_COUNT=400, USE_NURBS=1, NURBS_GRID=[float(x) for x in g["NURBS_GRID"]],
                             SYSTEM_PARAMS=[float(x) for x in g["SYSTEM_PARAMS"]], PARAMS_REAL=[float(x) for x in g["uR"]],
                             PARAMS_IMAGINARY=[float(x) for x in g["uI"]], PARAM_PHIR=float(g["phiR"]))
    ref = run_seeds(ref_bin, cfg, "ref", g["R"], list(range(1, 9)), tmp_path)
    dev = run_seeds(gpu_bin, dict(cfg, GPU_WALKERS=2000, MC_NSTEPS=2), "gpu", g["R"], list(range(1, 9)), tmp_path, gpu_seed=True)
    assert "GPU ensemble: 2000 walkers" in dev[0].log
    assert len(dev[0].local_energy_r) == 6
    compare_trajectories(ref, dev, 100, parity_log, "boxradial_n27")


def test_driver_config3_as_shipped_n8000(binaries, tmp_path):
    """config/BosonsBulk3D.config with the values it ships with - N = 8000, LBOX = 20, N_PARAM = 201, all parameters zero,
    MC_NSTEPS = 2 x MC_NTHERMSTEPS = 5000 + 1000, TIMESTEP = 2e-4, Eigen QR solve - through the device-bound driver (round 1
    could not hold one such configuration on an SM).  With u = 0 the walk is free (every move accepted) and E_L is the
    square-well energy b x #(pairs closer than a), about N rho (4 pi / 3) a^3 / 2 = 16 755 for uniformly distributed particles."""
    gpu_bin, _ = binaries
    z = [0.0] * 201
    cfg = driver.headline_config(z, z, N=8000, LBOX=20.0, TIMESTEP=2e-4, TOTALTIME=2e-4 * 1.5, MC_VERY_FIRST_NINITIALIZATIONSTEPS=100000,
                                 LINEAR_EQUATION_SOLVER_TYPE=1, USE_PRECONDITIONING=0, GPU_WALKERS=148)
    r = driver.run_driver(gpu_bin, cfg, str(tmp_path / "n8000"), timeout=900)
    e = r.local_energy_r
    assert len(e) == 2 and np.all(np.isfinite(e)) and np.all(np.isfinite(r.parameters_r))
    assert abs(e[0] - 16755.0) < 0.05 * 16755.0
    assert np.all(r.acceptance[:1] > 99.9)
