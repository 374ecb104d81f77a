"""GPU parity tests: the CUDA path (through the C ABI) against reference fixtures and the pinned oracle.

Tolerances: fixed-configuration E_L, drift and O_k within 1e-10 relative (north_star); the
single-move quotient within 1e-7 -- the sweep evaluates the SAME polynomial the reference's table
defines, but in a well-conditioned local form, while the reference's own monomial evaluation
carries ~1e-10 absolute error per pair term (coefficients up to 4e6), which accumulates over the
2(N-1) terms of a move.
"""
import json
import os

import numpy as np
import pytest

from oracle_lib import Oracle
from tdvmc_b200 import systems

pytestmark = pytest.mark.gpu

EVAL_CASES = ["bosonsbulk_n64_fixture", "bosonsbulk_n64_equil", "bosonsbulk_n343_lattice", "bosonsbulk_n343_equil",
              "nubosonsbulkpb_n216_equil", "nubosonsbulkpb_n1728_equil"]
RTOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def drift_term_scale(spec, R, u):
    """max_n sum_i |u'(r_ni)|: the magnitude of the terms whose (partly cancelling) sum is the drift F_n.
    On the perfect-lattice fixture F_n itself cancels to ~1e-6, so 'relative to max |F|' would measure
    summation-order noise; the error bound that means something is relative to the summed terms."""
    ut = spec.spline_space(np.asarray(u))
    knots, w, L = spec.knots, spec.weights, spec.lbox
    d = R[:, None, :] - R[None, :, :]
    d -= L * np.round(d / L)
    r = np.sqrt((d ** 2).sum(-1))
    if spec.pair_rule == systems.PAIR_RULE_REFLECT:
        r = np.where(r < spec.r_max, r, 2 * spec.r_max - r)
    np.fill_diagonal(r, 2 * spec.r_max + 1.0)
    inside = r <= spec.r_max
    rb = np.where(inside, r, 0.5 * spec.r_max)
    b = np.searchsorted(knots, rb, side="left") - 1
    up = np.zeros_like(r)
    for p in range(4):
        wp = w[b - p, p]
        up += ut[b - p] * (wp[..., 1] + 2 * wp[..., 2] * rb + 3 * wp[..., 3] * rb ** 2)
    return float(np.max(np.where(inside, np.abs(up), 0.0).sum(axis=1)))


def record_core(parity_log, test, name, r, g, exponent_floor=0.0):
    """north_star level 1 for one fixture: O_k, exponent, E^R, E^I and the drift within RTOL of the reference's own
    evaluation, each relative to its own magnitude (the drift: to max |F| over particles and components)."""
    parity_log.check(test, name, "O_k", np.max(np.abs(r["O"][0] - g["local_operators"])), np.max(np.abs(g["local_operators"])), RTOL)
    parity_log.check(test, name, "exponent", abs(r["exponent"][0] - float(g["exponent"])), max(abs(float(g["exponent"])), exponent_floor), RTOL)
    for key, gk in (("e_r", "local_energy_r"), ("e_i", "local_energy_i")):
        if float(g[gk]) == 0.0:
            assert r[key][0] == 0.0
            continue
        parity_log.check(test, name, key, abs(r[key][0] - float(g[gk])), abs(float(g[gk])), RTOL)
    for key in ("drift_r", "drift_i"):
        fmax = np.max(np.abs(g[key]))
        if fmax == 0.0:
            assert np.all(r[key][0] == 0.0)
            continue
        parity_log.check(test, name, key, np.max(np.abs(r[key][0] - g[key])), fmax, RTOL, "max|F|")


@pytest.fixture(scope="module")
def capi():
    from tdvmc_b200 import capi as c

    lib = c.load()                       # a missing or unloadable libtdvmc_b200.so is an ERROR, never a skip
    if lib.tdvmc_gpu_device_count() <= 0:
        if os.environ.get("TDVMC_REQUIRE_GPU"):
            raise AssertionError("no CUDA device: the GPU tests have no fallback")
        pytest.skip("no CUDA device on this machine (the library has no CPU fallback; run with -m gpu on the B200 box)")
    return c


def make_handle(capi, g, n_walkers=4, **kw):
    spec = systems.from_golden(g)
    h = capi.Handle(spec, n_walkers, **kw)
    h.set_params(g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), float(g["time"]))
    return spec, h


def test_min_image_known_answers_and_reference_bits(capi, golden):
    g0 = golden("bosonsbulk_n64_fixture")
    _, h = make_handle(capi, g0)
    here = os.path.dirname(os.path.abspath(__file__))
    cases = [c for c in json.load(open(os.path.join(here, "golden", "min_image_known_answers.json")))["cases"]]
    for L in (4.0, 5.0):
        sub = [c for c in cases if c["L"] == L]
        a = np.array([c["a"] + [0.0] * (3 - len(c["a"])) for c in sub])
        b = np.array([c["b"] + [0.0] * (3 - len(c["b"])) for c in sub])
        norm, disp = h.min_image(L, a, b)
        for c, n, d in zip(sub, norm, disp):
            assert abs(n - c["norm"]) < 1e-9, c           # src/test/Tests.h:11
            for k, want in enumerate(c["disp"]):
                assert abs(d[k] - want) < 1e-9, c
    g = golden("min_image_reference")
    for L in np.unique(g["L"]):
        m = g["L"] == L
        norm, disp = h.min_image(float(L), g["a"][m], g["b"][m])
        assert np.array_equal(norm, g["norm"][m])         # same IEEE operations in the same order
        assert np.array_equal(disp, g["disp"][m])
    h.close()


# the perfect 7^3 lattice wrapped into L = 7: every drift component cancels to ~1e-6 of its terms, so "relative to
# max |F|" would measure summation-order noise; only there the bound is relative to the summed terms (drift_term_scale)
LATTICE_CASES = {"bosonsbulk_n343_lattice"}


@pytest.mark.parametrize("name", EVAL_CASES)
def test_fixed_configuration_energy_drift_operators(capi, golden, parity_log, name):
    """north_star level 1: local energy, drift and O_k within 1e-10 relative of the reference's own CPU evaluation;
    the drift against max |F| of the fixture, other[4,5,8] (sum |F_R|^2, sum |F_I|^2, 2 sum F_R.F_I) against
    their own magnitude."""
    g = golden(name)
    spec, h = make_handle(capi, g)
    r = h.evaluate_fixed(g["R"])
    t = "fixed_configuration"
    parity_log.check(t, name, "spline_sums", np.max(np.abs(r["ss"][0] - g["spline_sums"])), np.max(np.abs(g["spline_sums"])), 1e-13)
    assert r["outer"][0] == float(g["outer_sum"])
    parity_log.check(t, name, "O_k", np.max(np.abs(r["O"][0] - g["local_operators"])), np.max(np.abs(g["local_operators"])), RTOL)
    for key, gk in (("exponent", "exponent"), ("e_r", "local_energy_r"), ("e_i", "local_energy_i")):
        parity_log.check(t, name, key, abs(r[key][0] - float(g[gk])), abs(float(g[gk])), RTOL)
    lattice = name in LATTICE_CASES
    scale = {}
    for key, u in (("drift_r", g["uR"]), ("drift_i", g["uI"])):
        fmax = np.max(np.abs(g[key]))
        scale[key] = max(fmax, drift_term_scale(spec, g["R"], u)) if lattice else fmax
        parity_log.check(t, name, key, np.max(np.abs(r[key][0] - g[key])), scale[key], RTOL,
                         "sum_i |u'(r_ni)| (perfect lattice, max|F| = %.1e)" % fmax if lattice else "max|F|")
    want = g["other_expectation_values"]
    got = r["other"][0]
    assert got.shape == want.shape
    # on the lattice R1 = sum |F_R|^2, I1, R1I1 are sums of squares of cancelled quantities: d(sum F^2) <= 2 N max|F| dF
    N = spec.n_particles
    fr, fi = np.max(np.abs(g["drift_r"])), np.max(np.abs(g["drift_i"]))
    quad = {4: 2 * N * fr * scale["drift_r"], 5: 2 * N * fi * scale["drift_i"],
            8: 2 * N * (fr * scale["drift_i"] + fi * scale["drift_r"])} if lattice else {}
    for k in range(9):
        if want[k] == 0.0 and not lattice:
            assert got[k] == 0.0, k
            continue
        parity_log.check(t, name, "other[%d]" % k, abs(got[k] - want[k]), max(abs(want[k]), quad.get(k, 0.0)), RTOL)
    assert np.all(got[9:] == 0.0) and np.all(want[9:] == 0.0)
    h.close()


@pytest.mark.parametrize("name", EVAL_CASES)
def test_tables_match_reference(capi, golden, name):
    g = golden(name)
    spec, h = make_handle(capi, g)
    sD, sD2 = h.tables_fixed(g["R"])
    wn = g["table_checksum_weights"]
    assert rel(np.einsum("n,kna->ka", wn, sD), g["sD_checksum"]) < 1e-12
    assert rel(np.einsum("n,kn->k", wn, sD2), g["sD2_checksum"]) < 1e-12
    if "sD" in g:
        assert rel(sD, g["sD"]) < 1e-13 and rel(sD2, g["sD2"]) < 1e-13
    else:
        idx = g["table_particles"]
        assert rel(sD[:, idx, :], g["sD_subset"]) < 1e-13
        assert rel(sD2[:, idx], g["sD2_subset"]) < 1e-13
    h.close()


@pytest.mark.parametrize("name", EVAL_CASES)
def test_scripted_move_quotient(capi, golden, name):
    g = golden(name)
    spec, h = make_handle(capi, g)
    q, d = h.quotient_fixed(g["R"], g["moves"])
    d_ref = g["move_exponent_new"] - float(g["exponent"])
    assert np.max(np.abs(d - d_ref)) < 5e-8
    assert np.max(np.abs(q / g["move_quotient"] - 1.0)) < 1e-7
    h.close()


def test_contract_from_tables_matches_fused(capi, golden):
    g = golden("bosonsbulk_n343_equil")
    spec, h = make_handle(capi, g, n_walkers=3)
    R = np.stack([g["R"], g["R"] + 0.01, np.roll(g["R"], 1, axis=0)])
    h.set_positions(R)
    r = h.evaluate_fixed(R)
    h.tables_resident(3)
    e_r, e_i = h.contract_resident(3)
    assert rel(e_r, r["e_r"]) < 1e-11 and rel(e_i, r["e_i"]) < 1e-11
    h.close()


def test_proposal_stream_matches_oracle(capi, golden):
    g = golden("bosonsbulk_n64_fixture")
    spec, h = make_handle(capi, g, seed=1234567890123, mc_step=0.4)
    o = Oracle(spec)
    p, d, lu = h.proposals(global_walker=5, first_step=(1 << 32) - 3, n=64)
    for i in range(64):
        p2, d2, lu2 = o.proposal(1234567890123, 5, (1 << 32) - 3 + i, 0.4)
        assert p[i] == p2
        assert np.allclose(d[i], d2, rtol=1e-13, atol=1e-15)
        assert abs(lu[i] - lu2) <= 1e-13 * max(1.0, abs(lu2))
    h.close()


@pytest.mark.parametrize("name,n_steps", [("bosonsbulk_n64_equil", 640), ("nubosonsbulkpb_n216_equil", 432),
                                          ("nubosonsbulkpb_n1728_equil", 300)])
def test_sweep_replays_oracle_chain(capi, golden, name, n_steps):
    """Same proposal stream, same accept rule: the device chain follows the oracle chain move for move."""
    g = golden(name)
    W, seed, mc_step, first = 3, 77, 0.35, 10
    spec, h = make_handle(capi, g, n_walkers=W, seed=seed, mc_step=mc_step, first_walker=first)
    o = Oracle(spec, time=float(g["time"]))
    R0 = np.stack([g["R"] + 0.003 * w for w in range(W)])
    h.set_positions(R0)
    h.sweep(n_steps // 2)
    h.sweep(n_steps - n_steps // 2)          # the stream is a function of the step counter, not of the launch
    R_gpu = h.get_positions()
    acc_gpu = None
    for w in range(W):
        R_ref, acc = o.sweep(R0[w], g["uR"], seed, first + w, 0, n_steps, mc_step)
        d = R_gpu[w] - R_ref                       # the device keeps positions wrapped into the first cell
        d -= spec.lbox * np.round(d / spec.lbox)
        assert np.max(np.abs(d)) < 1e-9, w
        assert np.all(np.abs(R_gpu[w]) <= spec.lbox / 2 + 1e-12)
    h.close()


def test_sample_and_accumulate_matches_oracle(capi, golden):
    g = golden("bosonsbulk_n64_equil")
    W, seed, mc_step = 5, 9, 0.4
    n_samples, n_therm, n_init = 3, 64, 32
    spec, h = make_handle(capi, g, n_walkers=W, seed=seed, mc_step=mc_step, max_samples=n_samples)
    o = Oracle(spec, time=float(g["time"]))
    R0 = np.stack([g["R"] + 0.002 * w for w in range(W)])
    h.set_positions(R0)
    h.sample_and_accumulate(n_samples, n_therm, n_init)
    got = h.allreduce_and_fetch()
    est = np.zeros(o.est_size())
    acc = 0
    for w in range(W):
        r = o.sample_walker(R0[w], g["uR"], g["uI"], float(g["phiR"]), seed, w, 0, n_init, n_samples, n_therm, mc_step, est)
        acc += r["accepted"]
    want = o.unpack_est(est, W * n_samples)
    assert got["n_samples"] == W * n_samples
    assert got["n_trials"] == W * (n_init + n_samples * n_therm)
    assert got["n_acceptances"] == acc
    assert rel(got["O"], want["O"]) < 1e-9
    assert abs(got["e_r"][0] - want["e_r"]) < 1e-9 * abs(want["e_r"])
    assert abs(got["e_i"][0] - want["e_i"]) < 1e-9 * abs(want["e_i"])
    assert rel(got["S"], want["S"]) < 1e-9
    assert rel(got["OER"], want["OER"]) < 1e-9
    assert rel(got["OEI"], want["OEI"]) < 1e-9
    assert rel(got["other"][[0, 1, 3, 4, 5, 6, 7, 8]], want["other"][[0, 1, 3, 4, 5, 6, 7, 8]]) < 1e-9
    assert abs(h.last_exponent() - o.evaluate(h.get_positions()[0], g["uR"], g["uI"])["exponent"]) < 1e-9
    h.close()


@pytest.mark.parametrize("M,P_case", [(1, "bosonsbulk_n64_fixture"), (33, "bosonsbulk_n64_fixture"),
                                       (4096 + 17, "bosonsbulk_n343_equil"), (20000, "nubosonsbulkpb_n216_equil")])
def test_accumulate_fixed_matches_numpy(capi, golden, M, P_case):
    g = golden(P_case)
    spec, h = make_handle(capi, g)
    P = spec.n_params
    rng = np.random.default_rng(2)
    O = rng.normal(50.0 + np.arange(P), 5.0, size=(M, P))
    e_r = rng.normal(-3.0, 1.0, M)
    e_i = rng.normal(0.5, 0.2, M)
    S, fr, fi, o = h.accumulate_fixed(O, e_r, e_i)
    Ol = O.astype(np.longdouble)
    assert rel(S, (Ol.T @ Ol).astype(float)) < 1e-13
    assert rel(fr, (Ol.T @ e_r.astype(np.longdouble)).astype(float)) < 1e-13
    assert rel(fi, (Ol.T @ e_i.astype(np.longdouble)).astype(float)) < 1e-13
    assert rel(o, Ol.sum(axis=0).astype(float)) < 1e-13
    assert np.array_equal(S, S.T)
    h.close()


def test_ensemble_statistics_match_reference_sampler(capi, golden):
    """Energies agree with the reference's own sampler within stated error bars (north_star level 2)."""
    g = golden("bosonsbulk_n64_mc")
    spec = systems.bosons_bulk(int(g["N"]), float(g["LBOX"]), int(g["N_PARAM"]), g["SYSTEM_PARAMS"])
    W = 512
    h = capi.Handle(spec, W, seed=2024, mc_step=float(g["MC_STEP"]), max_samples=8)
    h.set_params(g["uR"], g["uI"], 0.0, 0.0, 0.0)
    h.set_positions(np.broadcast_to(g["R0"], (W, spec.n_particles, 3)).copy())
    h.sample_and_accumulate(8, int(g["n_therm"]), 64 * 100)
    got = h.allreduce_and_fetch()
    er = g["energy_r_series"]
    nb = 20
    b = er[:len(er) // nb * nb].reshape(nb, -1).mean(axis=1)
    m_ref, s_ref = b.mean(), b.std(ddof=1) / np.sqrt(nb)
    var = np.var(er)                                              # per-sample variance from the reference series
    s_gpu = np.sqrt(var / (W * 8)) * 2.0                          # x2: 8 consecutive samples are correlated
    assert abs(got["e_r"][0] - m_ref) < 4.0 * np.hypot(s_ref, s_gpu), (got["e_r"][0], m_ref, s_ref, s_gpu)
    acc = got["n_acceptances"] / got["n_trials"]
    assert abs(acc - float(g["acceptance"])) < 0.01
    scale = np.abs(g["local_operators"]).max()
    assert np.max(np.abs(got["O"] - g["local_operators"])) / scale < 0.01
    h.close()


def test_headline_size_statistics_match_reference_sampler(capi, golden):
    """The headline workload's size and parameters (N = 343, L = 7, N_PARAM = 201, MC_STEP = 0.5): ensemble means of
    E^R, E^I, the acceptance rate and the operators O_k against the reference's own single-chain sampler, within stated
    error bars (blocking of the reference series into 15 blocks; the device ensemble has 4096 samples from 1024
    independent walkers, its error is taken from the reference's per-sample variance, x2 for correlation)."""
    g = golden("bosonsbulk_n343_mc")
    src = golden("bosonsbulk_n343_equil")
    spec = systems.bosons_bulk(343, 7.0, 201, g["SYSTEM_PARAMS"], weights=src["spline_weights"])
    W, n_samples = 1024, 4
    h = capi.Handle(spec, W, seed=31337, mc_step=float(g["MC_STEP"]), max_samples=n_samples)
    h.set_params(g["uR"], g["uI"], 0.0, 0.0, 0.0)
    h.set_positions(np.broadcast_to(src["R"], (W, 343, 3)).copy())
    h.sample_and_accumulate(n_samples, int(g["n_therm"]), int(g["n_init"]))
    got = h.allreduce_and_fetch()
    h.close()
    nb = 15
    for key, series in (("e_r", g["energy_r_series"]), ("e_i", g["energy_i_series"])):
        b = series[:len(series) // nb * nb].reshape(nb, -1).mean(axis=1)
        m_ref, s_ref = b.mean(), b.std(ddof=1) / np.sqrt(nb)
        s_gpu = np.sqrt(np.var(series) / (W * n_samples)) * 2.0
        assert abs(got[key][0] - m_ref) < 4.5 * np.hypot(s_ref, s_gpu), (key, got[key][0], m_ref, s_ref, s_gpu)
    assert abs(got["n_acceptances"] / got["n_trials"] - float(g["acceptance"])) < 0.005
    # operators: relative agreement where the reference's mean is well away from zero (pair counts per interval)
    ref_o = g["local_operators"]
    big = np.abs(ref_o) > 0.05 * np.max(np.abs(ref_o))
    assert np.max(np.abs(got["O"][big] - ref_o[big]) / np.abs(ref_o[big])) < 0.03


def test_full_size_properties(capi, golden):
    """BASELINE size (N=343, P=201): size-independent properties of the resident path."""
    g = golden("bosonsbulk_n343_equil")
    W = 64
    spec, h = make_handle(capi, g, n_walkers=W, seed=3, mc_step=0.5, max_samples=2, keep_sample_positions=True)
    R0 = np.broadcast_to(g["R"], (W, 343, 3)).copy()
    h.set_positions(R0)
    assert np.array_equal(h.get_positions(), R0)                  # layout round trip
    h.sample_and_accumulate(2, 343, 100)
    a = h.allreduce_and_fetch()
    assert a["n_samples"] == 2 * W and a["n_trials"] == W * (100 + 2 * 343)
    assert 0.2 < a["n_acceptances"] / a["n_trials"] < 0.95
    assert np.array_equal(a["S"], a["S"].T)
    # Cauchy-Schwarz / positivity of the second-moment matrix
    ev = np.linalg.eigvalsh(a["S"] - np.outer(a["O"], a["O"]))
    assert ev.min() > -1e-8 * ev.max()
    # sum_k O_k = number of pairs inside the cut (partition of unity of the B-spline basis) -> bounded by N(N-1)/2
    assert a["O"].sum() <= 343 * 342 / 2 * (1 + 1e-4)
    # re-evaluating the stored samples at unchanged parameters reproduces the estimators
    h.reevaluate_stored()
    b = h.allreduce_and_fetch()
    for k in ("O", "S", "OER", "OEI", "e_r", "e_i"):
        assert rel(b[k], a[k]) < 1e-13, k
    # fixed evaluation of the final positions equals the last stored sample rows on average
    Rf = h.get_positions()
    ev2 = h.evaluate_fixed(Rf[:4])
    o = Oracle(spec)
    ref = o.evaluate(Rf[0], g["uR"], g["uI"], float(g["phiR"]))
    assert abs(ev2["e_r"][0] - ref["e_r"]) < RTOL * abs(ref["e_r"])
    assert rel(ev2["O"][0], ref["O"]) < RTOL
    # periodic images: shifting every particle by a lattice vector leaves E_L unchanged
    shift = np.array([7.0, -14.0, 7.0])
    ev3 = h.evaluate_fixed(Rf[:1] + shift)
    assert abs(ev3["e_r"][0] - ev2["e_r"][0]) < 1e-9 * abs(ev2["e_r"][0])
    # wrap keeps distances: energies after MoveCoordinatesToFirstCell are unchanged
    h.wrap_positions()
    Rw = h.get_positions()
    assert np.all(np.abs(Rw) <= 3.5 + 1e-9)
    ev4 = h.evaluate_fixed(Rw[:1])
    assert abs(ev4["e_r"][0] - ev2["e_r"][0]) < 1e-9 * abs(ev2["e_r"][0])
    h.close()


def test_full_size_properties_config4(capi, golden):
    """config/NUBosonsBulkPB3D.config at its own size (N=1728, L=12, P=200, reflection rule) with its own sample counts
    (MC_NSTEPS=50, MC_NTHERMSTEPS=200) and sample reuse (UPDATE_SAMPLES_EVERY_NTH_STEP=1): the resident path."""
    g = golden("nubosonsbulkpb_n1728_equil")
    W, n_samples = 32, 50
    spec, h = make_handle(capi, g, n_walkers=W, seed=5, mc_step=0.5, max_samples=n_samples, keep_sample_positions=True)
    R0 = np.broadcast_to(g["R"], (W, 1728, 3)).copy()
    h.set_positions(R0)
    assert np.array_equal(h.get_positions(), R0)
    h.sample_and_accumulate(n_samples, 200, 400)
    a = h.allreduce_and_fetch()
    assert a["n_samples"] == n_samples * W and a["n_trials"] == W * (400 + n_samples * 200)
    assert 0.2 < a["n_acceptances"] / a["n_trials"] < 0.95
    assert np.array_equal(a["S"], a["S"].T)
    ev = np.linalg.eigvalsh(a["S"] - np.outer(a["O"], a["O"]))
    assert ev.min() > -1e-8 * ev.max()
    h.reevaluate_stored()                                        # src/TDVMC.cpp:1222-1303 at unchanged parameters
    b = h.allreduce_and_fetch()
    for k in ("O", "S", "OER", "OEI", "e_r", "e_i"):
        assert rel(b[k], a[k]) < 1e-13, k
    # re-evaluation at CHANGED parameters = fresh fixed evaluation of the same stored configurations (the last sample
    # of every walker is the walker's current position)
    u2 = g["uR"] * 1.01
    h.set_params(u2, g["uI"], float(g["phiR"]), 0.0, float(g["time"]))
    h.reevaluate_stored()
    c = h.allreduce_and_fetch()
    assert abs(c["e_r"][0] - a["e_r"][0]) > 1e-9 * abs(a["e_r"][0])
    assert rel(c["O"], a["O"]) < 1e-13                             # O_k do not depend on the parameters
    Rf = h.get_positions()
    o = Oracle(spec, time=float(g["time"]))
    ref = o.evaluate(Rf[0], u2, g["uI"], float(g["phiR"]))
    ev2 = h.evaluate_fixed(Rf[:2])
    assert abs(ev2["e_r"][0] - ref["e_r"]) < RTOL * abs(ref["e_r"])
    assert abs(ev2["e_i"][0] - ref["e_i"]) < RTOL * abs(ref["e_i"])
    assert rel(ev2["O"][0], ref["O"]) < RTOL
    ev3 = h.evaluate_fixed(Rf[:1] + np.array([12.0, -24.0, 12.0]))
    assert abs(ev3["e_r"][0] - ev2["e_r"][0]) < 1e-9 * abs(ev2["e_r"][0])
    h.close()


def test_update_stored_samples_replays_oracle(capi, golden):
    """UpdateSamplesConsecutive + UpdateExpectationValuesForGivenSamples (src/TDVMC.cpp:975-983, 1222-1303): stored
    samples advance at NEW parameters in ring order, then all are re-evaluated; the oracle replays every chain."""
    g = golden("bosonsbulk_n64_equil")
    W, seed, mc_step = 3, 13, 0.4
    n_samples, n_therm, n_init = 3, 48, 30
    spec, h = make_handle(capi, g, n_walkers=W, seed=seed, mc_step=mc_step, max_samples=n_samples, keep_sample_positions=True)
    o = Oracle(spec, time=float(g["time"]))
    R0 = np.stack([g["R"] + 0.002 * w for w in range(W)])
    h.set_positions(R0)
    h.sample_and_accumulate(n_samples, n_therm, n_init)
    u2, ui2 = g["uR"] * 1.03, g["uI"] * 0.9
    h.set_params(u2, ui2, float(g["phiR"]), float(g["phiI"]), float(g["time"]))
    h.update_stored(4, n_therm)                      # ring: slots 0, 1, 2, 0
    h.reevaluate_stored()
    got = h.allreduce_and_fetch()
    P = spec.n_params
    O = np.zeros(P)
    S = np.zeros((P, P))
    er = ei = 0.0
    acc = 0
    for w in range(W):
        R, a = o.sweep(R0[w], g["uR"], seed, w, 0, n_init, mc_step)
        acc += a
        step, stored = n_init, []
        for m in range(n_samples):
            R, a = o.sweep(R, g["uR"], seed, w, step, n_therm, mc_step)
            acc += a
            step += n_therm
            stored.append(R.copy())
        for slot in (0, 1, 2, 0):
            stored[slot], a = o.sweep(stored[slot], u2, seed, w, step, n_therm, mc_step)
            acc += a
            step += n_therm
        for Rm in stored:
            ev = o.evaluate(Rm, u2, ui2, float(g["phiR"]))
            O += ev["O"]
            S += np.outer(ev["O"], ev["O"])
            er += ev["e_r"]
            ei += ev["e_i"]
    M = W * n_samples
    assert got["n_samples"] == M
    assert got["n_trials"] == W * (n_init + n_samples * n_therm + 4 * n_therm)
    assert got["n_acceptances"] == acc
    assert rel(got["O"], O / M) < 1e-9
    assert rel(got["S"], S / M) < 1e-9
    assert abs(got["e_r"][0] - er / M) < 1e-9 * abs(er / M)
    assert abs(got["e_i"][0] - ei / M) < 1e-9 * abs(ei / M)
    h.close()


@pytest.mark.parametrize("N", [2, 3, 33, 65, 96, 128, 400, 416, 700])
def test_evaluate_tile_schedule_edge_sizes(capi, golden, N):
    """The evaluation kernel's 32x32 tile schedule at particle counts that hit its corner cases - a single partial tile,
    odd and even tile counts (the half shift), partial last tiles, more tiles than warps (N = 400, 416: several tiles
    per warp and shift; 700: the 24-warp variant) - against the pinned oracle on random configurations in the N = 64
    fixture's box and spline table (density is irrelevant to the arithmetic)."""
    g = golden("bosonsbulk_n64_equil")
    L, P = float(g["LBOX"]), int(g["N_PARAM"])
    spec = systems.bosons_bulk(N, L, P, g["SYSTEM_PARAMS"], weights=g["spline_weights"])
    h = capi.Handle(spec, 2)
    h.set_params(g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), float(g["time"]))
    rng = np.random.default_rng(N)
    R = rng.uniform(-L / 2, L / 2, (2, N, 3))
    if N <= 3:
        R *= 0.3                                                  # keep the few particles inside each other's cut
    R[1] += rng.integers(-2, 3, (N, 3)) * L                      # second configuration: unwrapped images
    ev = h.evaluate_fixed(R)
    o = Oracle(spec, time=float(g["time"]))
    for c in range(2):
        ref = o.evaluate(R[c], g["uR"], g["uI"], float(g["phiR"]))
        assert rel(ev["O"][c], ref["O"]) < RTOL
        assert abs(ref["e_r"]) > 0 and abs(ev["e_r"][c] - ref["e_r"]) < RTOL * abs(ref["e_r"])
        assert abs(ev["e_i"][c] - ref["e_i"]) < RTOL * max(abs(ref["e_i"]), abs(ref["e_r"]))
        assert abs(ev["exponent"][c] - ref["exponent"]) < RTOL * abs(ref["exponent"])
        assert np.max(np.abs(ev["drift_r"][c] - ref["drift_r"])) < RTOL * np.max(np.abs(ref["drift_r"]))
        assert np.max(np.abs(ev["drift_i"][c] - ref["drift_i"])) < RTOL * np.max(np.abs(ref["drift_i"]))
    # the sweep at the same sizes: chain replay for a few steps
    h.set_positions(R[:, :, :])
    h.sweep(40)
    Rg = h.get_positions()
    for c in range(2):
        Rr, _ = o.sweep(R[c], g["uR"], 1, c, 0, 40, 0.5)
        d = Rg[c] - Rr
        d -= L * np.round(d / L)
        assert np.max(np.abs(d)) < 1e-9
    h.close()


def test_errors_are_loud(capi, golden):
    g = golden("bosonsbulk_n64_fixture")
    spec = systems.from_golden(g)
    h = capi.Handle(spec, 2)
    with pytest.raises(capi.TdvmcError):
        h.sweep(10)                                   # parameters not set
    h.set_params(g["uR"], g["uI"])
    with pytest.raises(capi.TdvmcError):
        h.sample_and_accumulate(5, 1, 0)              # exceeds max_samples_per_walker
    with pytest.raises(capi.TdvmcError):
        h.allreduce_and_fetch()                       # nothing accumulated
    with pytest.raises(capi.TdvmcError):
        h.reevaluate_stored()                         # no stored samples
    with pytest.raises(capi.TdvmcError, match="nothing accumulated"):
        h.solve_parameters_dot()                      # the device solve needs estimators
    with pytest.raises(capi.TdvmcError, match="nothing accumulated"):
        h.euler_step(1e-3, g["uR"], g["uI"], 0.0, 0.0)
    h.close()
    # system descriptions the library does not offer are refused at creation, not later
    bad = systems.from_golden(g)
    bad.dim = 4
    with pytest.raises(capi.TdvmcError, match="invalid system"):
        capi.Handle(bad, 2)
    inh = systems.from_golden(golden("inhcontact_n3_equil"))
    inh.system_params = np.concatenate([inh.system_params, [0.0] * 5])     # the pulse extensions of GetExternalPotential
    with pytest.raises(capi.TdvmcError, match="invalid system"):
        capi.Handle(inh, 2)


def test_cpp_host_adapter_matches_python_path(capi, golden, tmp_path):
    """tdvmc_b200/host/example_driver (C++ over the C ABI) reproduces the ctypes path on the same inputs."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "tdvmc_b200", "host")
    subprocess.check_call(["make", "-C", host, "example_driver"])
    g = golden("bosonsbulk_n64_fixture")
    tables = tmp_path / "tables.txt"
    with open(tables, "w") as f:
        f.write(f"{int(g['N'])} {float(g['LBOX'])!r} {int(g['N_PARAM'])} {len(g['knots'])}\n")
        f.write(" ".join(repr(float(x)) for x in g["knots"]) + "\n")
        f.write(" ".join(repr(float(x)) for x in g["spline_weights"].ravel()) + "\n")
    r = subprocess.run([os.path.join(host, "example_driver"), str(tables)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    e_cpp = float(r.stdout.split("E_R=")[1].split()[0])
    # same ensemble through the Python binding
    N, L, P = int(g["N"]), float(g["LBOX"]), int(g["N_PARAM"])
    spec = systems.bosons_bulk(N, L, P, [1.0, 1.0], weights=g["spline_weights"])
    W, m = 64, round(N ** (1 / 3))
    R = np.zeros((W, N, 3))
    n = np.arange(N)
    for w in range(W):
        R[w, :, 0] = ((n % m) + 0.5 + 0.01 * w / W) * L / m - L / 2
        R[w, :, 1] = (((n // m) % m) + 0.5) * L / m - L / 2
        R[w, :, 2] = ((n // (m * m)) + 0.5) * L / m - L / 2
    k = np.arange(P)
    uR = -0.5 * np.exp(-((k * (L / 2) / (P - 1) / 0.8) ** 2))
    h = capi.Handle(spec, W, seed=1, mc_step=0.5, max_samples=2)
    h.set_positions(R)
    h.wrap_positions()
    h.set_params(uR, np.zeros(P))
    h.sample_and_accumulate(2, N, 10 * N)
    e_py = h.allreduce_and_fetch()["e_r"][0]
    assert abs(e_cpp - e_py) < 1e-9 * abs(e_py)
    # the Euler step the driver took with the device solve: same chain, same estimators -> same new parameters
    u2, _, phi2, _, dot = h.euler_step(1e-4, uR, np.zeros(P), 0.0, 0.0, time=1e-4, imaginary_time=1)
    euler = dict(kv.split("=") for kv in r.stdout.split("EULER ")[1].splitlines()[0].split())
    # (the driver builds uR with std::exp/std::pow, numpy may differ in the last bit: same tolerance as E_R above)
    for key, want in (("uR0", u2[0]), ("uRlast", u2[-1]), ("phiR", phi2), ("E_R", dot["e_r"])):
        assert abs(float(euler[key]) - want) < 1e-9 * abs(want), key
    assert abs(u2[0] - uR[0]) > 1e-9 * abs(uR[0]) and int(euler["notPD"]) == 0    # the step did move the parameters
    # ... and the Runge-Kutta step that followed (four device stages in the C++ adapter and in the Python mirror)
    from tdvmc_b200.ensemble import GpuEnsembleSystem
    ens = GpuEnsembleSystem.__new__(GpuEnsembleSystem)
    ens.handle = h
    ens.SampleExpectationValues(u2, np.zeros(P), phi2, 0.0, 2, N, 0, time=1e-4)
    u4, _, phi4, _, _ = ens.CalculateNextParametersRK4(1e-4, u2, np.zeros(P), phi2, 0.0, 2, N, 0, IMAGINARY_TIME=1, time=2e-4)
    rk4 = dict(kv.split("=") for kv in r.stdout.split("RK4 ")[1].splitlines()[0].split())
    for key, want in (("uR0", u4[0]), ("uRlast", u4[-1]), ("phiR", phi4)):
        assert abs(float(rk4[key]) - want) < 1e-9 * abs(want), key
    assert abs(u4[0] - u2[0]) > 1e-9 * abs(u2[0]) and int(rk4["notPD"]) == 0
    h.close()


_MG_WORKER = r"""
import os, sys, time
import numpy as np
root, tmp, rank, world = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
sys.path.insert(0, root)
from tdvmc_b200 import capi, systems
from tdvmc_b200.ensemble import GpuEnsembleSystem
g = np.load(os.path.join(root, "tests", "golden", "bosonsbulk_n64_equil.npz"))
spec = systems.from_golden(g)
idf = os.path.join(tmp, "nccl_id.bin")
if rank == 0:
    uid = capi.comm_unique_id()
    open(idf + ".tmp", "wb").write(uid)
    os.rename(idf + ".tmp", idf)
else:
    while not os.path.exists(idf):
        time.sleep(0.05)
    uid = open(idf, "rb").read()
W = 10
ens = GpuEnsembleSystem(spec, W, mc_step=0.4, seed=21, rank=rank, world=world, device=rank, mc_nsteps=2, unique_id=uid)
R = np.stack([g["R"] + 0.002 * w for w in range(W)])
ens.SetPositions(R[ens.first_walker:ens.first_walker + ens.n_local])
e = ens.ParallelUpdateExpectationValues(g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), 2, 64, 32, float(g["time"]))
from tdvmc_b200 import observables
go = np.load(os.path.join(root, "tests", "golden", "bosonsbulk_n64_obs.npz"))
o = ens.ParallelCalculateAdditionalSystemProperties(g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]),
                                                    observables.from_golden(go), 3, 40, 20, float(g["time"]))
e["pairDistribution"], e["structureFactor"] = o["pairDistribution"], o["structureFactor"]
# a second pass whose estimators are NOT fetched: the Euler step all-reduces them itself and solves on every rank;
# the fetch that follows must see the same (once-reduced) sums
ens.SampleExpectationValues(g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), 2, 64, 0, float(g["time"]))
u2, _, p2, _, info = ens.CalculateNextParametersEuler(1e-4, g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), IMAGINARY_TIME=1)
e2 = ens._fetch()
e["euler_uR"], e["euler_phiR"], e["euler_e_r"], e["second_e_r"], e["second_n"] = u2, p2, info["e_r"], e2["localEnergyR"], e2["nSamples"]
np.savez(os.path.join(tmp, f"rank{rank}.npz"), **{k: np.asarray(v) for k, v in e.items()})
ens.close()
"""


def test_two_gpu_allreduce_matches_single_gpu(capi, golden, tmp_path):
    """Walkers sharded over 2 GPUs + the packed NCCL all-reduce == the same ensemble on one GPU, on every rank."""
    import subprocess
    import sys

    if capi.load().tdvmc_gpu_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(_MG_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), root, str(tmp_path), str(r), "2"]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    g = golden("bosonsbulk_n64_equil")
    spec, h = make_handle(capi, g, n_walkers=10, seed=21, mc_step=0.4, max_samples=2)
    h.set_positions(np.stack([g["R"] + 0.002 * w for w in range(10)]))
    h.sample_and_accumulate(2, 64, 32)
    one = h.allreduce_and_fetch()
    from tdvmc_b200 import observables
    gr1, sk1 = h.sample_observables(observables.from_golden(golden("bosonsbulk_n64_obs")), 3, 40, 20)
    h.sample_and_accumulate(2, 64, 0)
    u1, _, p1, _, info1 = h.euler_step(1e-4, g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), imaginary_time=1)
    h.close()
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    assert np.array_equal(ranks[0]["euler_uR"], ranks[1]["euler_uR"])          # every rank solved the same system
    for r in range(2):
        two = ranks[r]
        # device solve after the in-library all-reduce: same derivatives as one GPU holding all walkers (the sums
        # differ by the all-reduce's rounding), and the later fetch saw the once-reduced estimators
        du2, du1 = (two["euler_uR"] - g["uR"]) / 1e-4, (u1 - g["uR"]) / 1e-4
        assert np.max(np.abs(du2 - du1)) < 1e-8 * np.max(np.abs(du1))
        assert abs(float(two["euler_phiR"]) - p1) < 1e-9 * abs(p1)
        assert float(two["euler_e_r"]) == float(two["second_e_r"]) and int(two["second_n"]) == 20
        assert abs(float(two["second_e_r"]) - info1["e_r"]) < 1e-13 * abs(info1["e_r"])
        assert rel(two["pairDistribution"], gr1) < 1e-13 and rel(two["structureFactor"], sk1) < 1e-12
        assert int(two["nSamples"]) == 20 and int(two["nTrials"]) == one["n_trials"]
        assert int(two["nAcceptances"]) == one["n_acceptances"]
        assert rel(two["localOperators"], one["O"]) < 1e-13
        assert rel(two["localOperatorsMatrix"], one["S"]) < 1e-13
        assert rel(two["localOperatorlocalEnergyR"], one["OER"]) < 1e-13
        assert abs(float(two["localEnergyR"]) - one["e_r"][0]) < 1e-13 * abs(one["e_r"][0])


# ---------------------------------------------------------------------------------------------------
# HeBulk (BASELINE configs[1], config/bulk_64.config): McMillan core + uniform splines + Aziz potential
# ---------------------------------------------------------------------------------------------------
HE_CASES = ["hebulk_n64_fixture", "hebulk_n64_equil"]


@pytest.mark.parametrize("name", HE_CASES)
def test_hebulk_fixed_configuration(capi, golden, parity_log, name):
    g = golden(name)
    spec, h = make_handle(capi, g)
    r = h.evaluate_fixed(g["R"])
    K = spec.n_splines
    assert rel(r["ss"][0][:K], g["spline_sums"]) < 1e-13
    assert abs(r["ss"][0][K] - float(g["mcmillan_sum"])) <= 1e-13 * abs(float(g["mcmillan_sum"]))
    record_core(parity_log, "hebulk", name, r, g)
    want, got = g["other_expectation_values"], r["other"][0]
    assert got.shape == want.shape == (103,)
    assert abs(got[0] - want[0]) < RTOL * abs(want[0]) and abs(got[1] - want[1]) < RTOL * abs(want[1])
    assert abs(got[2] - want[2]) <= 1e-9 * abs(want[2])                      # wf = exp(exponent)
    assert rel(got[3:], want[3:]) < 1e-12                                     # g(r)
    q, d = h.quotient_fixed(g["R"], g["moves"])
    d_ref = g["move_exponent_new"] - float(g["exponent"])
    assert np.max(np.abs(d - d_ref) / np.maximum(1.0, np.abs(d_ref))) < 1e-9
    h.close()


def test_hebulk_sweep_and_estimators_match_oracle(capi, golden):
    from oracle_lib import OracleHeBulk

    g = golden("hebulk_n64_equil")
    W, seed, mc_step = 4, 31, 0.3
    n_samples, n_therm, n_init = 2, 64, 64
    spec, h = make_handle(capi, g, n_walkers=W, seed=seed, mc_step=mc_step, max_samples=n_samples)
    o = OracleHeBulk(spec)
    R0 = np.stack([g["R"] + 0.002 * w for w in range(W)])
    h.set_positions(R0)
    h.sample_and_accumulate(n_samples, n_therm, n_init)
    got = h.allreduce_and_fetch()
    est = np.zeros(o.est_size())
    acc = 0
    for w in range(W):
        r = o.sample_walker(R0[w], g["uR"], g["uI"], 0.0, seed, w, 0, n_init, n_samples, n_therm, mc_step, est)
        acc += r["accepted"]
    want = o.unpack_est(est, W * n_samples)
    assert got["n_acceptances"] == acc and got["n_trials"] == W * (n_init + n_samples * n_therm)
    assert rel(got["O"], want["O"]) < 1e-9
    assert abs(got["e_r"][0] - want["e_r"]) < 1e-9 * abs(want["e_r"])
    assert rel(got["S"], want["S"]) < 1e-9 and rel(got["OER"], want["OER"]) < 1e-9
    assert rel(got["other"][[0, 1]], want["other"][[0, 1]]) < 1e-9
    assert rel(got["other"][3:], want["other"][3:]) < 1e-9
    Rg = h.get_positions()
    for w in range(W):
        d = Rg[w] - r["R"] if w == W - 1 else None
    d = Rg[W - 1] - r["R"]
    d -= spec.lbox * np.round(d / spec.lbox)
    assert np.max(np.abs(d)) < 1e-9
    h.close()


def test_hebulk_statistics_match_reference_sampler(capi, golden):
    g = golden("hebulk_n64_mc")
    spec = systems.he_bulk(int(g["N"]), float(g["LBOX"]), int(g["N_PARAM"]))
    W = 512
    h = capi.Handle(spec, W, seed=77, mc_step=float(g["MC_STEP"]), max_samples=8)
    h.set_params(g["uR"], g["uI"])
    h.set_positions(np.broadcast_to(g["R0"], (W, spec.n_particles, 3)).copy())
    h.sample_and_accumulate(8, int(g["n_therm"]), 64 * 300)
    got = h.allreduce_and_fetch()
    er = g["energy_r_series"]
    nb = 20
    b = er[:len(er) // nb * nb].reshape(nb, -1).mean(axis=1)
    m_ref, s_ref = b.mean(), b.std(ddof=1) / np.sqrt(nb)
    s_gpu = np.sqrt(np.var(er) / (W * 8)) * 2.0
    assert abs(got["e_r"][0] - m_ref) < 4.0 * np.hypot(s_ref, s_gpu), (got["e_r"][0], m_ref, s_ref, s_gpu)
    assert abs(got["n_acceptances"] / got["n_trials"] - float(g["acceptance"])) < 0.015
    h.close()


@pytest.mark.parametrize("name", ["hebulk_n64_equil", "hedrop_n6_equil"])
def test_stored_sample_reevaluation_he_systems(capi, golden, name):
    """UpdateExpectationValuesForGivenSamples (src/TDVMC.cpp:1222-1303) for the He systems: stored configurations,
    re-evaluated at unchanged parameters (idempotent) and at new parameters (equals a fresh evaluation of them)."""
    g = golden(name)
    W, n_samples = 6, 3
    spec, h = make_handle(capi, g, n_walkers=W, seed=21, mc_step=float(g["MC_STEP"]) if "MC_STEP" in g else 0.3,
                          max_samples=n_samples, keep_sample_positions=True)
    h.set_positions(np.stack([g["R"] + 0.001 * w for w in range(W)]))
    h.sample_and_accumulate(n_samples, 40, 20)
    a = h.allreduce_and_fetch()
    h.reevaluate_stored()
    b = h.allreduce_and_fetch()
    for k in ("O", "S", "OER", "OEI", "e_r", "e_i", "other"):
        assert rel(b[k], a[k]) < 1e-13, k
    u2, ui2 = g["uR"] * 0.97, g["uI"] + 0.01
    h.set_params(u2, ui2, float(g["phiR"]), float(g["phiI"]), float(g["time"]))
    h.reevaluate_stored()
    c = h.allreduce_and_fetch()
    # the last stored sample of every walker is its current position: compare its share through a second handle
    Rf = h.get_positions()
    ev = h.evaluate_fixed(Rf)
    h2 = capi.Handle(spec, W, seed=21, max_samples=1)
    h2.set_params(u2, ui2, float(g["phiR"]), float(g["phiI"]), float(g["time"]))
    h2.set_positions(Rf)
    h2.sample_and_accumulate(1, 0, 0)                                 # zero moves: evaluates the given positions
    d = h2.allreduce_and_fetch()
    assert abs(d["e_r"][0] - ev["e_r"].mean()) < 1e-12 * abs(d["e_r"][0])
    assert abs(c["e_r"][0] - a["e_r"][0]) > 1e-9 * abs(a["e_r"][0])  # parameters did change the energy
    P = spec.n_params
    free = np.arange(P - 1) if name.startswith("hebulk") else np.arange(P)   # HeBulk's last operator is the constant 1
    assert rel(c["O"][free], a["O"][free]) < 1e-13                            # O_k do not depend on the parameters
    h.close()
    h2.close()


# ---------------------------------------------------------------------------------------------------
# HeDrop (BASELINE configs[0], config/drop_6.config): open boundary, two spline grids, const/linear tails, LJ
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["hedrop_n6_fixture", "hedrop_n6_spread", "hedrop_n6_equil", "hedrop_n20_equil"])
def test_hedrop_fixed_configuration(capi, golden, parity_log, name):
    g = golden(name)
    spec, h = make_handle(capi, g)
    r = h.evaluate_fixed(g["R"])
    K = spec.n_splines
    assert rel(r["ss"][0][:K], g["spline_sums"]) < 1e-13
    assert abs(r["ss"][0][K] - float(g["mcmillan_sum"])) <= 1e-13 * max(abs(float(g["mcmillan_sum"])), 1e-300)
    assert r["ss"][0][K + 1] == float(g["const_sum"])
    assert abs(r["ss"][0][K + 2] - float(g["linear_sum"])) <= 1e-13 * max(abs(float(g["linear_sum"])), 1e-300)
    record_core(parity_log, "hedrop", name, r, g)
    want, got = g["other_expectation_values"], r["other"][0]
    assert got.shape == want.shape == (403,)
    assert abs(got[0] - want[0]) < RTOL * abs(want[0]) and abs(got[1] - want[1]) < RTOL * abs(want[1])
    assert abs(got[2] - want[2]) <= 1e-9 * abs(want[2])                      # wf = exp(exponent + phiR)
    assert rel(got[3:], want[3:]) < 1e-12                                     # g(r) and the density profile
    q, d = h.quotient_fixed(g["R"], g["moves"])
    d_ref = g["move_exponent_new"] - float(g["exponent"])
    assert np.max(np.abs(d - d_ref) / np.maximum(1.0, np.abs(d_ref))) < 1e-9
    h.close()


@pytest.mark.parametrize("name", ["hedrop_n6_fixture", "hedrop_n20_equil"])   # 8 lanes per walker / a whole warp per walker
def test_hedrop_chain_estimators_and_com(capi, golden, name):
    from oracle_lib import OracleHe

    g = golden(name)
    W, seed, mc_step = 8, 5, 0.5
    n_samples, n_therm, n_init = 3, 30, 60
    spec, h = make_handle(capi, g, n_walkers=W, seed=seed, mc_step=mc_step, max_samples=n_samples)
    o = OracleHe(spec)
    R0 = np.stack([g["R"] * (1.0 + 0.05 * w) for w in range(W)])
    h.set_positions(R0)
    h.sample_and_accumulate(n_samples, n_therm, n_init)
    got = h.allreduce_and_fetch()
    est = np.zeros(o.est_size())
    acc, Rf = 0, []
    for w in range(W):
        r = o.sample_walker(R0[w], g["uR"], g["uI"], float(g["phiR"]), seed, w, 0, n_init, n_samples, n_therm, mc_step, est)
        acc += r["accepted"]
        Rf.append(r["R"])
    want = o.unpack_est(est, W * n_samples)
    assert got["n_acceptances"] == acc and got["n_trials"] == W * (n_init + n_samples * n_therm)
    assert np.max(np.abs(h.get_positions() - np.stack(Rf))) < 1e-9          # open boundary: no wrapping, same chain
    assert rel(got["O"], want["O"]) < 1e-9
    assert abs(got["e_r"][0] - want["e_r"]) < 1e-9 * abs(want["e_r"])
    assert rel(got["S"], want["S"]) < 1e-9 and rel(got["OER"], want["OER"]) < 1e-9
    assert rel(got["other"], want["other"]) < 1e-9
    # AlignCoordinates for USE_MOVE_COM_TO_ZERO systems (src/TDVMC.cpp:2569-2582): centre of mass to zero
    Rb = h.get_positions()
    h.wrap_positions()
    Ra = h.get_positions()
    assert np.max(np.abs(Ra.mean(axis=1))) < 1e-12
    assert np.max(np.abs((Ra - Rb) + Rb.mean(axis=1, keepdims=True))) < 1e-12
    h.close()


# ---------------------------------------------------------------------------------------------------
# BosonMixtureCluster (BASELINE configs[4], config/He4He4Na.config): species, pair types, log term, HFDB / KTTY
# ---------------------------------------------------------------------------------------------------
MIX_CASES = ["mixture_he4he4na_fixture", "mixture_he4he4na_compact", "mixture_he4he4na_stretched", "mixture_he4he4na_equil",
             "mixture_he3he4cs_equil"]   # config/He3He4Cs.config: three pair types (N_PARAM = 78), He-3, KTTY He-Cs
# BosonMixtureCluster_4thorder (config/He4He4Na_4thOrder.config): quartic splines, same kernels
MIX4_CASES = ["mixture4_he4he4na_fixture", "mixture4_he4he4na_compact", "mixture4_he4he4na_stretched", "mixture4_he4he4na_equil"]


@pytest.mark.parametrize("name", MIX_CASES + MIX4_CASES)
def test_mixture_fixed_configuration(capi, golden, parity_log, name):
    from oracle_lib import OracleMix

    g = golden(name)
    spec, h = make_handle(capi, g)
    r = h.evaluate_fixed(g["R"])
    o = OracleMix(spec).evaluate(g["R"], g["uR"], g["uI"], float(g["phiR"]))
    assert rel(r["ss"][0], o["ext"]) < 1e-13
    record_core(parity_log, "mixture", name, r, g)
    want, got = g["other_expectation_values"], r["other"][0]
    assert got.shape == want.shape
    for k in (0, 1, 2, 3, 5):
        assert abs(got[k] - want[k]) <= RTOL * abs(want[k]), k
    assert abs(got[4] - want[4]) <= 1e-9 * abs(want[4])
    assert np.all(got[6:] == 0.0)
    q, d = h.quotient_fixed(g["R"], g["moves"])
    d_ref = g["move_exponent_new"] - float(g["exponent"])
    assert np.max(np.abs(d - d_ref) / np.maximum(1.0, np.abs(d_ref))) < 1e-9
    h.close()


@pytest.mark.parametrize("name", ["mixture_he4he4na_equil", "mixture4_he4he4na_equil", "mixture_he3he4cs_equil"])
def test_mixture_chain_estimators_and_com(capi, golden, name):
    from oracle_lib import OracleMix

    g = golden(name)
    W, seed, mc_step = 70, 12, 2.0
    n_samples, n_therm, n_init = 3, 10, 30
    spec, h = make_handle(capi, g, n_walkers=W, seed=seed, mc_step=mc_step, max_samples=n_samples)
    o = OracleMix(spec)
    R0 = np.stack([g["R"] * (1.0 + 0.01 * w) for w in range(W)])
    h.set_positions(R0)
    h.sample_and_accumulate(n_samples, n_therm, n_init)
    got = h.allreduce_and_fetch()
    est = np.zeros(o.est_size())
    acc, Rf = 0, []
    for w in range(W):
        r = o.sample_walker(R0[w], g["uR"], g["uI"], float(g["phiR"]), seed, w, 0, n_init, n_samples, n_therm, mc_step, est)
        acc += r["accepted"]
        Rf.append(r["R"])
    want = o.unpack_est(est, W * n_samples)
    assert got["n_acceptances"] == acc and got["n_trials"] == W * (n_init + n_samples * n_therm)
    assert np.max(np.abs(h.get_positions() - np.stack(Rf))) < 1e-9
    assert rel(got["O"], want["O"]) < 1e-9
    assert abs(got["e_r"][0] - want["e_r"]) < 1e-9 * abs(want["e_r"])
    assert rel(got["S"], want["S"]) < 1e-9 and rel(got["OER"], want["OER"]) < 1e-9
    assert rel(got["other"][:6], want["other"][:6]) < 1e-9
    Rb = h.get_positions()
    h.wrap_positions()                                   # mass-weighted centre of mass to zero
    Ra = h.get_positions()
    m = spec.extra["mass"]
    assert np.max(np.abs((Ra * m[None, :, None]).sum(axis=1))) < 1e-10
    com = (Rb * m[None, :, None]).sum(axis=1) / m.sum()
    assert np.max(np.abs(Ra - (Rb - com[:, None, :]))) < 1e-12
    assert np.allclose(com[0], o.center_of_mass(Rb[0]), rtol=1e-14, atol=1e-14)
    h.close()


# ---------------------------------------------------------------------------------------------------
# additional observables g(r), S(k): CalculateAdditionalSystemProperties (BosonsBulk.cpp:474-520,
# NUBosonsBulkPB.cpp:597-639) and the driver loop around it (src/TDVMC.cpp:1332-1388, 1438-1444)
# ---------------------------------------------------------------------------------------------------
OBS_CASES = ["bosonsbulk_n64_obs", "bosonsbulk_n343_obs", "nubosonsbulkpb_n216_obs"]


@pytest.mark.parametrize("name", OBS_CASES)
def test_observables_fixed_match_reference(capi, golden, name):
    from tdvmc_b200 import observables
    g = golden(name)
    src = golden(str(g["source"]))
    obs = observables.from_golden(g)
    spec, h = make_handle(capi, src)
    R = np.stack([src["R"], src["R"][::-1].copy()])          # particle order does not matter
    gr, sk = h.observables_fixed(obs, R)
    for c in range(2):
        assert np.max(np.abs(gr[c] - g["gr_fixed"])) <= 1e-12 * np.max(np.abs(g["gr_fixed"]))
        assert np.max(np.abs(sk[c] - g["sk_fixed"])) <= 1e-10 * np.max(np.abs(g["sk_fixed"]))
    h.close()


def test_sample_observables_replay_oracle(capi, golden):
    """The observable pass over the resident walkers equals the oracle chain + oracle observables, sample by sample."""
    from oracle_lib import oracle_observables
    from tdvmc_b200 import observables
    g = golden("bosonsbulk_n64_obs")
    src = golden(str(g["source"]))
    obs = observables.from_golden(g)
    W, seed, mc_step = 3, 31, 0.4
    n_samples, n_therm, n_init = 4, 40, 25
    spec, h = make_handle(capi, src, n_walkers=W, seed=seed, mc_step=mc_step)
    o = Oracle(spec, time=float(src["time"]))
    R0 = np.stack([src["R"] + 0.002 * w for w in range(W)])
    h.set_positions(R0)
    gr, sk = h.sample_observables(obs, n_samples, n_therm, n_init)
    gr_ref, sk_ref = np.zeros(obs.gr_count), np.zeros(obs.n_shells)
    for w in range(W):
        R, step = R0[w].copy(), 0
        R, _ = o.sweep(R, src["uR"], seed, w, step, n_init, mc_step)
        step += n_init
        for m in range(n_samples):
            R, _ = o.sweep(R, src["uR"], seed, w, step, n_therm, mc_step)
            step += n_therm
            a, b = oracle_observables(spec.lbox, R, obs)
            gr_ref += a
            sk_ref += b
    gr_ref /= W * n_samples
    sk_ref /= W * n_samples
    assert np.max(np.abs(gr - gr_ref)) <= 1e-12 * np.max(gr_ref)
    assert np.max(np.abs(sk - sk_ref)) <= 1e-9 * np.max(sk_ref)
    h.close()


@pytest.mark.parametrize("name", ["bosonsbulk_n64_obs", "nubosonsbulkpb_n216_obs"])
def test_sample_observables_statistics_match_reference(capi, golden, name):
    """Mean g(r) and S(k) agree with the reference's own end-of-run pass (its sampler, its RNG) within error bars:
    pair counts per bin are close to Poisson, |rho_k|^2 close to exponential; x4 for autocorrelation."""
    from tdvmc_b200 import observables
    g = golden(name)
    src = golden(str(g["source"]))
    obs = observables.from_golden(g)
    W, n_samples = 256, 8
    spec, h = make_handle(capi, src, n_walkers=W, seed=77, mc_step=float(g["MC_STEP"]))
    h.set_positions(np.broadcast_to(src["R"], (W, spec.n_particles, 3)).copy())
    gr, sk = h.sample_observables(obs, n_samples, int(g["MC_NADDITIONALTHERMSTEPS"]), int(g["MC_NADDITIONALINITIALIZATIONSTEPS"]))
    n_ref, n_gpu = float(g["MC_NADDITIONALSTEPS"]) / 4.0, W * n_samples / 4.0
    q = obs.gr_weight / obs.gr_scaling
    sig_gr = np.sqrt(np.maximum(g["gr_mean"], q) * q * (1.0 / n_ref + 1.0 / n_gpu))
    assert np.all(np.abs(gr - g["gr_mean"]) < 6.0 * sig_gr), np.max(np.abs(gr - g["gr_mean"]) / sig_gr)
    n_k = np.diff(obs.shell_ptr)
    sig_sk = np.maximum(g["sk_mean"], 0.05) * np.sqrt(1.0 / n_ref + 1.0 / n_gpu)
    assert np.all(np.abs(sk - g["sk_mean"]) < 6.0 * sig_sk), np.max(np.abs(sk - g["sk_mean"]) / sig_sk)
    h.close()


# ---------------------------------------------------------------------------------------------------
# north_star level 2: time-evolved variational parameters agree statistically with the reference
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["bosonsbulk_n64_evolution", "bosonsbulk_n64_evolution_realtime"])
def test_time_evolution_matches_reference_statistically(capi, golden, name):
    """Euler time steps (the reference's ParallelUpdateExpectationValues -> SolveForParametersDot ->
    CalculateNextParametersEuler loop) with the estimators coming from the GPU ensemble, in imaginary time (30 steps, the
    real parts relax) and in real time (12 steps, the imaginary parts grow out of zero); the reference ran the same
    loops with its own sampler for 24 seeds.  With the same number of samples per step a device run is one more draw
    from the same distribution of trajectories.  Stated error bars: sigma = the reference's seed-to-seed spread per
    parameter, widened for the uncertainty of its mean.  (a) every parameter (real and imaginary part) of every device
    run within 6 sigma (Student t, 23 degrees of freedom: p ~ 4e-6 per comparison); (b) over five device seeds the
    normalised deviations have unit spread (rms within [0.7, 1.4]) - the device trajectories scatter like the
    reference's, no more, no less; (c) the energy of the last steps agrees within its error bar; (d) the parameters
    moved by > 20 sigma, so this is not noise compared with noise."""
    from tdvmc_b200 import timestep
    from tdvmc_b200.ensemble import GpuEnsembleSystem
    g = golden(name)
    src = golden(str(g["source"]))
    spec = systems.bosons_bulk(int(g["N"]), float(g["LBOX"]), int(g["N_PARAM"]), g["SYSTEM_PARAMS"], weights=src["spline_weights"])
    W = 64
    n_samples = int(g["MC_NSTEPS"]) // W                            # same samples per time step as the reference
    dt, steps, imag = float(g["TIMESTEP"]), int(g["time_steps"]), int(g["IMAGINARY_TIME"])

    def evolve(seed, on_device):
        """on_device: the estimators never leave the GPU - SolveForParametersDot + CalculateNextParametersEuler run in
        solve_kernel (tdvmc_gpu_euler_step); otherwise they are fetched and the host mirror of the reference's solve
        steps the parameters."""
        ens = GpuEnsembleSystem(spec, W, mc_step=float(g["MC_STEP"]), mc_nsteps=n_samples, seed=seed)
        ens.SetPositions(np.broadcast_to(src["R"], (W, spec.n_particles, 3)).copy())
        uR, uI, phiR, phiI = g["uR0"].copy(), np.zeros(spec.n_params), 0.0, 0.0
        ens.BroadcastNewParameters(uR, uI, phiR, phiI)
        ens.DoMetropolisSteps(int(g["equilibration_steps"]))
        tr, ti, energies = [], [], []
        for _ in range(steps):
            if on_device:
                ens.SampleExpectationValues(uR, uI, phiR, phiI, n_samples, int(g["MC_NTHERMSTEPS"]), int(g["MC_NINITIALIZATIONSTEPS"]))
                uR, uI, phiR, phiI, info = ens.CalculateNextParametersEuler(dt, uR, uI, phiR, phiI, IMAGINARY_TIME=imag)
                assert not info["not_positive_definite"]
                energies.append(info["e_r"])
            else:
                est = ens.ParallelUpdateExpectationValues(uR, uI, phiR, phiI, n_samples, int(g["MC_NTHERMSTEPS"]),
                                                          int(g["MC_NINITIALIZATIONSTEPS"]))
                energies.append(est["localEnergyR"])
                uR, uI, phiR, phiI = timestep.euler_step(dt, uR, uI, phiR, phiI, est, imaginary_time=imag)
            tr.append(uR.copy())
            ti.append(uI.copy())
        ens.handle.close()
        return np.array(tr), np.array(ti), np.array(energies)

    runs = [evolve(seed, on_device=k % 2 == 0) for k, seed in enumerate((4242, 100, 101, 102, 103))]
    # same seed, both routes: the chains are identical and the device solve equals the reference-order solve bit for
    # bit, so the trajectories may differ only by the numpy mirror's summation order (np.dot) - far below the noise
    host_twin = evolve(4242, on_device=False)
    assert np.max(np.abs(host_twin[0] - runs[0][0])) < 1e-9 * np.max(np.abs(runs[0][0]))
    K = len(g["seeds"])
    widen = np.sqrt(1.0 + 1.0 / K)
    checked = (steps // 3, 2 * steps // 3, steps - 1)
    parts = [("uR", 0)] + ([("uI", 1)] if imag == 0 else [])
    for key, idx in parts:
        for t in checked:
            sig = np.maximum(g[key + "_t_std"][t], 1e-7) * widen
            z = np.stack([(run[idx][t] - g[key + "_t_mean"][t]) / sig for run in runs])          # [seed][param]
            assert np.abs(z).max() < 6.0, (key, t, float(np.abs(z).max()))
            rms = float(np.sqrt(np.mean(z ** 2)))
            assert 0.7 < rms < 1.4, (key, t, rms)
    moving = "uI" if imag == 0 else "uR"
    start = np.zeros(spec.n_params) if imag == 0 else g["uR0"]
    moved = np.abs(g[moving + "_t_mean"][-1] - start) / np.maximum(g[moving + "_t_std"][-1], 1e-7)
    assert moved.max() > 20.0
    n_tail = min(10, steps // 2)
    tail = slice(steps - n_tail, steps)
    e_sig = np.sqrt(np.mean(g["energy_r_t_std"][tail] ** 2) / n_tail) * widen
    e_gpu = np.mean([run[2][tail].mean() for run in runs])
    assert abs(e_gpu - g["energy_r_t_mean"][tail].mean()) < 6.0 * e_sig * np.sqrt(1.0 / len(runs) + 1.0 / K)



# ---------------------------------------------------------------------------------------------------
# cluster observables of the He4-He4-Na mixture (BosonMixtureCluster.cpp:680-741): config 5's end-of-run pass
# ---------------------------------------------------------------------------------------------------
def test_cluster_observables_fixed_match_reference(capi, golden):
    g = golden("mixture_he4he4na_obs")
    tags = ["fixture", "compact", "stretched", "equil"]
    src = [golden(f"mixture_he4he4na_{t}") for t in tags]
    spec, h = make_handle(capi, src[0])
    r2, angle, density, distance = h.cluster_observables_fixed(g, np.stack([s["R"] for s in src]))
    for c, t in enumerate(tags):
        assert abs(r2[c] - float(g[f"{t}_r2_fixed"])) < 1e-13 * r2[c]
        assert np.array_equal(angle[c], g[f"{t}_angle_fixed"])
        assert np.array_equal(distance[c], g[f"{t}_distance_fixed"])
        assert np.max(np.abs(density[c] - g[f"{t}_density_fixed"])) <= 1e-15 * np.max(g[f"{t}_density_fixed"])
    h.close()


def test_cluster_observables_sampled_replay_oracle_and_match_reference(capi, golden):
    """(a) the pass over resident walkers equals oracle chain + oracle observables; (b) with many walkers the means agree
    with the reference's own end-of-run pass (its sampler) within error bars: r2 within 5 standard errors estimated from
    the histogram of distances from the centre of mass, histogram bins within 6 sigma of binomial counting noise (x4 for
    autocorrelation)."""
    from oracle_lib import OracleMix, oracle_mix_observables
    g = golden("mixture_he4he4na_obs")
    src = golden("mixture_he4he4na_equil")
    mc_step = float(g["MC_STEP"])
    W, seed, n_samples, n_therm, n_init = 3, 17, 4, 30, 20
    spec, h = make_handle(capi, src, n_walkers=W, seed=seed, mc_step=mc_step)
    o = OracleMix(spec)
    R0 = np.stack([src["R"] + 0.01 * w for w in range(W)])
    h.set_positions(R0)
    r2, angle, density, distance = h.sample_cluster_observables(g, n_samples, n_therm, n_init)
    acc = [0.0, 0, 0, 0]
    for w in range(W):
        R, _ = o.sweep(R0[w], src["uR"], seed, w, 0, n_init, mc_step)
        step = n_init
        for m in range(n_samples):
            R, _ = o.sweep(R, src["uR"], seed, w, step, n_therm, mc_step)
            step += n_therm
            a = oracle_mix_observables(o, R, g)
            acc = [x + y for x, y in zip(acc, a)]
    M = W * n_samples
    assert abs(r2 - acc[0] / M) < 1e-10 * r2
    assert np.max(np.abs(angle - acc[1] / M)) < 1e-12
    assert np.max(np.abs(distance - acc[3] / M)) < 1e-12
    assert np.max(np.abs(density - acc[2] / M)) <= 1e-12 * np.max(acc[2] / M)
    h.close()

    W, n_samples = 4096, 8
    spec, h = make_handle(capi, src, n_walkers=W, seed=99, mc_step=mc_step)
    h.set_positions(np.broadcast_to(src["R"], (W, 3, 3)).copy())
    r2, angle, density, distance = h.sample_cluster_observables(g, n_samples, int(g["MC_NADDITIONALTHERMSTEPS"]),
                                                                int(g["MC_NADDITIONALINITIALIZATIONSTEPS"]))
    n_ref, n_gpu = float(g["MC_NADDITIONALSTEPS"]) / 4.0, W * n_samples / 4.0
    for got, want in ((angle, g["angle_mean"]), (distance, g["distance_mean"])):
        sig = np.sqrt(np.maximum(want, 1.0 / n_ref) * (1.0 / n_ref + 1.0 / n_gpu))      # bin frequencies: binomial
        assert np.all(np.abs(got - want) < 6.0 * sig), float(np.max(np.abs(got - want) / sig))
    assert abs(angle.sum(axis=1) - 1.0).max() < 1e-12 and abs(distance.sum(axis=1) - 1.0).max() < 1e-3
    assert abs(r2 - float(g["r2_mean"])) < 0.05 * float(g["r2_mean"])
    h.close()


@pytest.mark.parametrize("name", ["hebulk_n64_obs", "hedrop_n6_obs"])
def test_he_structure_factor_matches_reference(capi, golden, name):
    """S(k) of the He systems' CalculateAdditionalSystemProperties (HeBulk.cpp:432-447: 300 shells, 3976 wave vectors;
    HeDrop.cpp:674-689) through the same observables kernel (no g(r) grid: their g(r) lives in the other values)."""
    from tdvmc_b200 import observables
    g = golden(name)
    src = golden(str(g["source"]))
    ptr = np.concatenate([[0], np.cumsum(g["k_shell_sizes"].astype(np.int64))]).astype(np.int32)
    obs = observables.ObservableSpec(0, 1.0, 1.0, 1.0, np.ones(1), ptr, g["k_vectors"].reshape(-1, 3))
    spec, h = make_handle(capi, src)
    _, sk = h.observables_fixed(obs, src["R"][None])
    assert np.max(np.abs(sk[0] - g["sk_fixed"])) <= 1e-10 * np.max(np.abs(g["sk_fixed"]))
    h.close()


# ---------------------------------------------------------------------------------------------------
# SolveForParametersDot / CalculateNextParametersEuler on the device
# ---------------------------------------------------------------------------------------------------
def _solve_in_reference_order(O, S, OER, OEI, ER, EI, imaginary_time, eps=0.001):
    """SolveForParametersDot, Cholesky branch, in scalar Python floats with the reference's loop order
    (src/TDVMC.cpp:1506-1537, 1560-1622, 1658-1711): IEEE double, no fused operations - what g++ -O3 computes."""
    import math
    P = len(O)
    O = [float(x) for x in O]
    if imaginary_time == 0:
        bR = [float(OEI[i]) - EI * O[i] for i in range(P)]
        bI = [-float(OER[i]) + ER * O[i] for i in range(P)]
    else:
        bR = [-float(OER[i]) + ER * O[i] for i in range(P)]
        bI = [-float(OEI[i]) for i in range(P)]
    A = [[0.0] * P for _ in range(P)]
    for i in range(P):
        for j in range(i + 1):
            A[i][j] = float(S[i][j]) - O[i] * O[j]
            A[j][i] = A[i][j]
    sc = [math.sqrt(A[i][i]) for i in range(P)]
    for i in range(P):
        for j in range(P):
            A[i][j] /= (sc[i] * sc[j])
        bR[i] /= sc[i]
        bI[i] /= sc[i]
    for i in range(P):
        A[i][i] += eps
    for i in range(P):
        Ai = A[i]
        for j in range(i + 1):
            s = Ai[j]
            Aj = A[j]
            for k in range(j):
                s -= Ai[k] * Aj[k]
            if i > j:
                Ai[j] = s / Aj[j]
            elif s > 0:
                Ai[i] = math.sqrt(s)
    sols = []
    for rhs in (bR, bI):
        tmp = [0.0] * P
        for i in range(P):
            s = 0.0
            for j in range(i):
                s += A[i][j] * tmp[j]
            tmp[i] = 1.0 / A[i][i] * (rhs[i] - s)
        x = [0.0] * P
        for i in range(P - 1, -1, -1):
            s = 0.0
            for j in range(P - 1, i, -1):
                s += A[j][i] * x[j]
            x[i] = 1.0 / A[i][i] * (tmp[i] - s)
        sols.append(x)
    pr = pi = 0.0
    for i in range(P):
        pr -= O[i] * sols[0][i]
        pi -= O[i] * sols[1][i]
    if imaginary_time == 0:
        pi -= ER
    else:
        pr -= ER
    return (np.array([sols[0][i] / sc[i] for i in range(P)]), np.array([sols[1][i] / sc[i] for i in range(P)]), pr, pi)


@pytest.mark.parametrize("name", ["bosonsbulk_n64_evolution", "bosonsbulk_n64_evolution_realtime", "bosonsbulk_n64_evolution_rotation"])
def test_device_solver_matches_reference(capi, golden, name):
    """tdvmc_gpu_solve_fixed on the reference's own first-step estimators reproduces the derivatives the reference's
    SolveForParametersDot computed from them (ref_harness evolve, printed at 17 digits): the kernel applies the
    reference's operations in the reference's order, so the agreement is to the last bit, in shared memory and in the
    global-memory variant alike."""
    g = golden(name)
    src = golden(str(g["source"]))
    spec = systems.bosons_bulk(int(g["N"]), float(g["LBOX"]), int(g["N_PARAM"]), g["SYSTEM_PARAMS"], weights=src["spline_weights"])
    h = capi.Handle(spec, 4)
    est = dict(O=g["first_O"], S=g["first_S"], OER=g["first_OER"], OEI=g["first_OEI"], e_r=g["first_ER"], e_i=g["first_EI"])
    imag = int(g["IMAGINARY_TIME"])
    for force_global in (False, True):
        d = h.solve_fixed(est, imaginary_time=imag, force_global=force_global)
        assert not d["not_positive_definite"]
        if imag == -1:
            # the 1.499 pi time rotation (src/TDVMC.cpp:1475-1504): cos / sin of the device's libm may differ from glibc's in
            # the last bit, so the right-hand sides do too
            assert rel(d["u_dot_r"], g["first_uDotR"]) < 1e-12 and rel(d["u_dot_i"], g["first_uDotI"]) < 1e-12
            assert abs(d["phi_dot_r"] - float(g["first_phiDotR"])) < 1e-12 * abs(float(g["first_phiDotR"]))
            assert abs(d["phi_dot_i"] - float(g["first_phiDotI"])) < 1e-12 * abs(float(g["first_phiDotI"]))
        else:
            assert np.array_equal(d["u_dot_r"], g["first_uDotR"]) and np.array_equal(d["u_dot_i"], g["first_uDotI"]), \
                (rel(d["u_dot_r"], g["first_uDotR"]), rel(d["u_dot_i"], g["first_uDotI"]))
            assert d["phi_dot_r"] == float(g["first_phiDotR"]) and d["phi_dot_i"] == float(g["first_phiDotI"])
        assert d["e_r"] == float(g["first_ER"]) and d["e_i"] == float(g["first_EI"])
    # a matrix that is not positive definite is reported, as the reference's doNotAcceptStep
    bad = dict(est, S=np.outer(g["first_O"], g["first_O"]) - np.eye(len(g["first_O"])))
    d = h.solve_fixed(bad, imaginary_time=imag, use_preconditioning=False)
    assert d["not_positive_definite"]
    with pytest.raises(capi.TdvmcError, match="IMAGINARY_TIME"):
        h.solve_fixed(est, imaginary_time=2)
    with pytest.raises(capi.TdvmcError, match="LINEAR_EQUATION_SOLVER_TYPE"):
        h.solve_fixed(est, imaginary_time=imag, solver_type=2)
    h.close()


@pytest.mark.parametrize("name", ["bosonsbulk_n64_evolution_qr", "bosonsbulk_n64_evolution_qr_raw"])
def test_device_qr_solver_matches_eigen(capi, golden, name):
    """LINEAR_EQUATION_SOLVER_TYPE = 1 on the device (solve_qr_kernel: Eigen's FullPivHouseholderQR step by step, then the mean
    subtraction of src/TDVMC.cpp:1800-1809) against the derivatives the reference's Eigen branch computed from its own
    first-step estimators - with the scaling + 0.002 regularisation of USE_PRECONDITIONING = 1 (condition ~500) and without
    (condition 1.4e5): equal to rounding times the condition number."""
    g = golden(name)
    src = golden(str(g["source"]))
    spec = systems.bosons_bulk(int(g["N"]), float(g["LBOX"]), int(g["N_PARAM"]), g["SYSTEM_PARAMS"], weights=src["spline_weights"])
    h = capi.Handle(spec, 4)
    est = dict(O=g["first_O"], S=g["first_S"], OER=g["first_OER"], OEI=g["first_OEI"], e_r=g["first_ER"], e_i=g["first_EI"])
    pre = bool(int(g["USE_PRECONDITIONING"]))
    d = h.solve_fixed(est, imaginary_time=int(g["IMAGINARY_TIME"]), use_preconditioning=pre, solver_type=1)
    tol = 1e-11 if pre else 1e-9
    scale = max(np.max(np.abs(g["first_uDotR"])), np.max(np.abs(g["first_uDotI"])))
    assert scale > 0
    assert np.max(np.abs(d["u_dot_r"] - g["first_uDotR"])) < tol * scale and np.max(np.abs(d["u_dot_i"] - g["first_uDotI"])) < tol * scale
    pscale = max(abs(float(g["first_phiDotR"])), abs(float(g["first_phiDotI"])))
    assert abs(d["phi_dot_r"] - float(g["first_phiDotR"])) < tol * pscale and abs(d["phi_dot_i"] - float(g["first_phiDotI"])) < tol * pscale
    assert abs(np.sum(d["u_dot_r"] * (1.0 if not pre else 0.0))) < 1e-9 * scale * len(d["u_dot_r"])   # mean subtracted (unscaled case)
    # the two-CTA cluster kernel (matrix in distributed shared memory, the default where it fits) and the one-CTA kernel with
    # the matrix in global memory form every sum in the same order: bit-identical
    d1 = h.solve_fixed(est, imaginary_time=int(g["IMAGINARY_TIME"]), use_preconditioning=pre, solver_type=1, force_global=True)
    for key in ("u_dot_r", "u_dot_i"):
        assert np.array_equal(d[key], d1[key]), (key, np.max(np.abs(d[key] - d1[key])))
    assert d["phi_dot_r"] == d1["phi_dot_r"] and d["phi_dot_i"] == d1["phi_dot_i"]
    h.close()


def test_device_euler_step_at_headline_size(capi, golden):
    """P = 201 (BASELINE config 3): SolveForParametersDot on the device-resident estimators of a sampling pass equals
    the reference-order scalar solve of the fetched estimators bit for bit and the numpy mirror to rounding; the Euler
    step leaves the device at the new parameters (evaluation equals a handle set to them by hand)."""
    from tdvmc_b200 import timestep
    g = golden("bosonsbulk_n343_equil")
    W = 296
    spec, h = make_handle(capi, g, n_walkers=W, seed=11, mc_step=0.5, max_samples=2)
    h.set_positions(np.broadcast_to(g["R"], (W, 343, 3)).copy())
    h.sample_and_accumulate(2, 343, 343)
    d = h.solve_parameters_dot(imaginary_time=1, min_scaling=1e-12)
    a = h.allreduce_and_fetch()                                   # fetch after the solve: the sums are still there
    assert d["e_r"] == a["e_r"][0] and d["e_i"] == a["e_i"][0]
    diag = np.diag(a["S"]) - a["O"] ** 2
    if diag.min() > 0:                                            # (an operator that never varied has no reference answer)
        ur, ui, pr, pi = _solve_in_reference_order(a["O"], a["S"], a["OER"], a["OEI"], float(a["e_r"][0]), float(a["e_i"][0]), 1)
        assert np.array_equal(d["u_dot_r"], ur) and np.array_equal(d["u_dot_i"], ui), (rel(d["u_dot_r"], ur), rel(d["u_dot_i"], ui))
        assert d["phi_dot_r"] == pr and d["phi_dot_i"] == pi
    est = dict(localOperators=a["O"], localOperatorsMatrix=a["S"], localOperatorlocalEnergyR=a["OER"],
               localOperatorlocalEnergyI=a["OEI"], localEnergyR=float(a["e_r"][0]), localEnergyI=float(a["e_i"][0]))
    vr, vi, qr, qi = timestep.solve_for_parameters_dot(est, imaginary_time=1, min_scaling=1e-12, lapack=True)
    scale = np.max(np.abs(vr))
    assert np.max(np.abs(d["u_dot_r"] - vr)) < 1e-8 * scale and np.max(np.abs(d["u_dot_i"] - vi)) < 1e-8 * max(np.max(np.abs(vi)), scale)
    assert abs(d["phi_dot_r"] - qr) < 1e-8 * abs(qr)
    dt = 1e-4
    uR, uI, phiR, phiI, d2 = h.euler_step(dt, g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), imaginary_time=1, min_scaling=1e-12)
    assert np.array_equal(d2["u_dot_r"], d["u_dot_r"])
    assert np.array_equal(uR, g["uR"] + d["u_dot_r"] * dt) and np.array_equal(uI, g["uI"] + d["u_dot_i"] * dt)
    assert phiR == float(g["phiR"]) + d["phi_dot_r"] * dt
    e_new = h.evaluate_fixed(g["R"][None])
    h2 = capi.Handle(spec, 4)
    h2.set_params(uR, uI, phiR, phiI, 0.0)
    e_ref = h2.evaluate_fixed(g["R"][None])
    assert e_new["e_r"][0] == e_ref["e_r"][0] and e_new["exponent"][0] == e_ref["exponent"][0]
    assert e_new["e_r"][0] != float(g["local_energy_r"])          # the parameters did move
    h.close()
    h2.close()


# ---------------------------------------------------------------------------------------------------
# NUBosonsBulkPBBoxAndRadial (radial + box spline bases) through the same C ABI
# ---------------------------------------------------------------------------------------------------
BR_CASES = ["boxradial_n27_jittered", "boxradial_n27_equil", "boxradial_n64_equil",
            "boxradial2d_n25_equil"]   # config/NUBosonsBulkPBBoxAndRadial2D.config (DIM = 2)


@pytest.mark.parametrize("name", BR_CASES)
def test_boxradial_fixed_configuration(capi, golden, parity_log, name):
    """Local energy, drift, O_k, basis sums, g(r) bins and the scripted-move quotient against the reference's own
    evaluation (config/NUBosonsBulkPBBoxAndRadial3D.config at its own size, and a 64-particle non-uniform-grid case)."""
    from oracle_lib import OracleBR

    g = golden(name)
    spec, h = make_handle(capi, g)
    K = spec.extra["n_splines"]
    r = h.evaluate_fixed(g["R"][None])
    record_core(parity_log, "boxradial", name, r, g)
    assert rel(r["ss"][0][:K], g["spline_sums_rad"]) < 1e-12 and rel(r["ss"][0][K:], g["spline_sums"]) < 1e-12
    assert rel(r["other"][0][:3], g["other_expectation_values"][:3]) < RTOL
    assert rel(r["other"][0][3:], g["gr_bins"]) < 1e-13 and g["gr_bins"].sum() > 0
    scale = np.abs(g["drift_r"]).max()
    # the drift really carries the reference's box-for-radial substitution (NUBosonsBulkPBBoxAndRadial.cpp:493-497):
    # contracting the reference's own tables the "corrected" way moves it by far more than the tolerance
    PR = spec.n_params // 2
    fixed = g["drift_r"].copy()
    fixed[:, :spec.dim] += g["uR"][PR - 1] * (g["sD_rad"][K - 1] - g["sD"][K - 1])
    if abs(g["uR"][PR - 1]) > 1e-6:          # (the 2-D fixture's last radial parameter is ~1e-11: nothing to see there)
        assert np.abs(fixed - g["drift_r"]).max() > 1e3 * RTOL * scale
    q, d = h.quotient_fixed(g["R"], g["moves"])
    assert rel(q, g["move_quotient"]) < 1e-9
    o = OracleBR(spec, time=float(g["time"]))
    for m, dd in zip(g["moves"], d):
        _, en, ex = o.quotient(g["R"], int(m[0]), m[1:4], g["uR"])
        assert abs(dd - (en - ex)) < 1e-10 * max(abs(en), 1.0)
    h.close()


@pytest.mark.parametrize("name,n_steps", [("boxradial_n27_equil", 27 * 40), ("boxradial_n64_equil", 64 * 10),
                                          ("boxradial2d_n25_equil", 25 * 40)])
def test_boxradial_sweep_replays_oracle_chain(capi, golden, name, n_steps):
    from oracle_lib import OracleBR

    g = golden(name)
    W, seed, mc_step, first = 3, 31, 0.35, 7
    spec, h = make_handle(capi, g, n_walkers=W, seed=seed, mc_step=mc_step, first_walker=first)
    o = OracleBR(spec, time=float(g["time"]))
    R0 = np.stack([g["R"] + 0.003 * w for w in range(W)])
    h.set_positions(R0)
    h.sweep(n_steps // 3)
    h.sweep(n_steps - n_steps // 3)
    R_gpu = h.get_positions()
    for w in range(W):
        R_ref, acc = o.sweep(R0[w], g["uR"], seed, first + w, 0, n_steps, mc_step)
        d = R_gpu[w] - R_ref
        d -= spec.lbox * np.round(d / spec.lbox)
        assert np.max(np.abs(d)) < 1e-9, w
        assert 0.2 * n_steps < acc < n_steps
    h.close()


def test_boxradial_estimators_match_oracle(capi, golden):
    """A whole UpdateExpectationValues pass (sweeps, evaluations, S / F accumulation, g(r) bins, counters) against the
    oracle driven by the same proposal stream."""
    from oracle_lib import OracleBR

    g = golden("boxradial_n27_equil")
    W, seed, mc_step = 6, 17, 0.5
    n_samples, n_therm, n_init = 3, 27, 54
    spec, h = make_handle(capi, g, n_walkers=W, seed=seed, mc_step=mc_step, max_samples=n_samples)
    o = OracleBR(spec, time=float(g["time"]))
    R0 = np.stack([g["R"] + 0.002 * w for w in range(W)])
    h.set_positions(R0)
    h.sample_and_accumulate(n_samples, n_therm, n_init)
    got = h.allreduce_and_fetch()
    est = np.zeros(o.est_size())
    acc = 0
    for w in range(W):
        acc += o.sample_walker(R0[w], g["uR"], g["uI"], float(g["phiR"]), seed, w, 0, n_init, n_samples, n_therm, mc_step, est)["accepted"]
    want = o.unpack_est(est, W * n_samples)
    assert got["n_samples"] == W * n_samples and got["n_trials"] == W * (n_init + n_samples * n_therm)
    assert got["n_acceptances"] == acc
    for k in ("O", "S", "OER", "OEI"):
        assert rel(got[k], want[k]) < 1e-9, k
    assert abs(got["e_r"][0] - want["e_r"]) < 1e-9 * abs(want["e_r"])
    assert abs(got["e_i"][0] - want["e_i"]) < 1e-9 * abs(want["e_i"])
    assert rel(got["other"][:2], want["other"][:2]) < 1e-9 and rel(got["other"][3:], want["other"][3:]) < 1e-12
    # the solve of section 8(f) rank 3 runs on this system's estimators as well (P = 100)
    d = h.solve_parameters_dot(imaginary_time=1, min_scaling=1e-12)
    assert np.all(np.isfinite(d["u_dot_r"])) and d["e_r"] == got["e_r"][0]
    h.close()


def test_boxradial_statistics_match_reference_sampler(capi, golden):
    """Ensemble energy, acceptance and <O_k> against the reference's own Metropolis run of this system (mt19937_64
    stream, 4000 samples): within 4 combined standard errors / 0.01 / 2 % of the profile's scale."""
    g = golden("boxradial_n27_mc")
    src = golden(str(g["source"]))
    W = 1024
    spec, h = make_handle(capi, src, n_walkers=W, seed=99, mc_step=float(g["MC_STEP"]), max_samples=4)
    h.set_positions(np.broadcast_to(src["R"], (W, 27, 3)).copy())
    h.sample_and_accumulate(4, 27 * 4, 27 * 100)
    got = h.allreduce_and_fetch()
    er = g["energy_r_series"]
    nb = 20
    b = er[:len(er) // nb * nb].reshape(nb, -1).mean(axis=1)
    m_ref, s_ref = b.mean(), b.std(ddof=1) / np.sqrt(nb)
    s_gpu = np.std(er) / np.sqrt(W)            # W independent walkers: at least W independent samples
    assert abs(got["e_r"][0] - m_ref) < 4.0 * np.hypot(s_ref, s_gpu), (got["e_r"][0], m_ref, s_ref, s_gpu)
    assert abs(got["n_acceptances"] / got["n_trials"] - float(g["acceptance"])) < 0.01
    assert np.max(np.abs(got["O"] - g["local_operators"])) / np.abs(g["local_operators"]).max() < 0.02
    h.close()


# ---------------------------------------------------------------------------------------------------
# InhContactBosons - the one-dimensional one-body + pair-spline system of the shipped
# config/InhContactBosons*.config - through the same C ABI (system kind 5, dim = 1)
# ---------------------------------------------------------------------------------------------------
INH_CASES = ["inhcontact_n3_fixture", "inhcontact_n3_well", "inhcontact_n3_equil", "inhcontact_n20_equil"]


@pytest.mark.parametrize("name", INH_CASES)
def test_inhcontact_fixed_configuration(capi, golden, parity_log, name):
    from oracle_lib import OracleInh

    g = golden(name)
    spec, h = make_handle(capi, g)
    K1 = spec.extra["n_splines_spf"]
    r = h.evaluate_fixed(g["R"][None])
    record_core(parity_log, "inhcontact", name, r, g, exponent_floor=1.0)
    assert rel(r["ss"][0][:K1], g["spline_sums_spf"]) < 1e-13 and rel(r["ss"][0][K1:], g["spline_sums_pc"]) < 1e-13
    assert rel(r["other"][0], g["other_expectation_values"]) < RTOL
    assert np.all(r["drift_r"][0][:, 1:] == 0.0)
    q, d = h.quotient_fixed(g["R"], g["moves"])
    d_ref = g["move_exponent_new"] - float(g["exponent"])
    assert np.max(np.abs(d - d_ref) / np.maximum(1.0, np.abs(d_ref))) < 1e-10
    assert rel(q, g["move_quotient"]) < 1e-9
    h.close()


@pytest.mark.parametrize("name", ["inhcontact_n3_equil", "inhcontact_n20_equil"])
def test_inhcontact_chain_and_estimators_match_oracle(capi, golden, name):
    """Chain replay (same proposal stream: the first Gaussian component moves the one coordinate) and a whole
    UpdateExpectationValues pass against the oracle."""
    from oracle_lib import OracleInh

    g = golden(name)
    N = int(g["N"])
    W, seed, mc_step = 40, 23, 0.5
    n_samples, n_therm, n_init = 3, 4 * N, 10 * N
    spec, h = make_handle(capi, g, n_walkers=W, seed=seed, mc_step=mc_step, max_samples=n_samples)
    o = OracleInh(spec)
    R0 = np.stack([g["R"] + np.array([0.004 * w, 0.0, 0.0]) for w in range(W)])
    h.set_positions(R0)
    h.sample_and_accumulate(n_samples, n_therm, n_init)
    got = h.allreduce_and_fetch()
    est = np.zeros(o.est_size())
    acc, Rf = 0, []
    for w in range(W):
        r = o.sample_walker(R0[w], g["uR"], g["uI"], float(g["phiR"]), seed, w, 0, n_init, n_samples, n_therm, mc_step, est)
        acc += r["accepted"]
        Rf.append(r["R"])
    want = o.unpack_est(est, W * n_samples)
    assert got["n_acceptances"] == acc and got["n_trials"] == W * (n_init + n_samples * n_therm)
    assert 0.3 * got["n_trials"] < acc < got["n_trials"]
    assert np.max(np.abs(h.get_positions() - np.stack(Rf))) < 1e-9
    for k in ("O", "S", "OER", "OEI"):
        assert rel(got[k], want[k]) < 1e-9, k
    assert abs(got["e_r"][0] - want["e_r"]) < 1e-9 * abs(want["e_r"])
    assert abs(got["e_i"][0] - want["e_i"]) < 1e-9 * max(abs(want["e_i"]), 1e-3 * abs(want["e_r"]))
    assert rel(got["other"], want["other"]) < 1e-9
    d = h.solve_parameters_dot(imaginary_time=1, min_scaling=1e-12)      # the device solve runs on this system too (P = 62)
    assert d["e_r"] == got["e_r"][0] and np.all(np.isfinite(d["u_dot_r"]))
    h.close()


def test_inhcontact_statistics_match_reference_sampler(capi, golden):
    """Ensemble energy and acceptance against the reference's own Metropolis run (mt19937_64 stream, 6000 samples)."""
    g = golden("inhcontact_n3_mc")
    src = golden(str(g["source"]))
    W = 4096
    spec, h = make_handle(capi, src, n_walkers=W, seed=5, mc_step=float(g["MC_STEP"]), max_samples=4)
    h.set_positions(np.broadcast_to(src["R"], (W, 3, 3)).copy())
    h.sample_and_accumulate(4, 3 * 10, 3 * 300)
    got = h.allreduce_and_fetch()
    er = g["energy_r_series"]
    nb = 20
    b = er[:len(er) // nb * nb].reshape(nb, -1).mean(axis=1)
    m_ref, s_ref = b.mean(), b.std(ddof=1) / np.sqrt(nb)
    s_gpu = np.std(er) / np.sqrt(W)
    assert abs(got["e_r"][0] - m_ref) < 4.0 * np.hypot(s_ref, s_gpu), (got["e_r"][0], m_ref, s_ref, s_gpu)
    assert abs(got["n_acceptances"] / got["n_trials"] - float(g["acceptance"])) < 0.015
    h.close()


@pytest.mark.parametrize("method", ["PC", "PCReuseSamples", "RK4", "RK4ReuseSamples"])
def test_multistage_integrators_on_device_match_host_mirror(capi, golden, method):
    """CalculateNextParametersPC / PCReuseSamples / RK4 / RK4ReuseSamples (src/TDVMC.cpp:1855-2103) composed from device
    stages (sampling or stored-sample re-evaluation + solve_kernel, nothing fetched) against the same composition with
    every stage's estimators fetched and solved by the host mirror of the reference's SolveForParametersDot: two
    ensembles with the same seed walk the same chains, so the results agree to the solver's rounding."""
    from tdvmc_b200 import timestep
    from tdvmc_b200.ensemble import GpuEnsembleSystem
    g = golden("bosonsbulk_n64_evolution")
    src = golden(str(g["source"]))
    spec = systems.bosons_bulk(int(g["N"]), float(g["LBOX"]), int(g["N_PARAM"]), g["SYSTEM_PARAMS"], weights=src["spline_weights"])
    W, n_samples, n_therm, n_init, dt = 64, 4, 64, 32, 2e-3
    reuse = method.endswith("ReuseSamples")
    mc = (n_samples, n_therm, n_init)
    uR0, uI0 = g["uR0"].copy(), 0.01 * np.sin(np.arange(spec.n_params))

    def make():
        ens = GpuEnsembleSystem(spec, W, mc_step=float(g["MC_STEP"]), mc_nsteps=n_samples, seed=8,
                                update_samples_every_nth_step=1 if reuse else 0)
        ens.SetPositions(np.broadcast_to(src["R"], (W, spec.n_particles, 3)).copy())
        ens.BroadcastNewParameters(uR0, uI0, 0.0, 0.0)
        ens.DoMetropolisSteps(640)
        return ens

    dev = make()
    dev.SampleExpectationValues(uR0, uI0, 0.0, 0.0, *mc)
    args = () if reuse else mc
    uR1, uI1, pR1, pI1, _ = getattr(dev, "CalculateNextParameters" + method)(dt, uR0, uI0, 0.0, 0.0, *args, IMAGINARY_TIME=0)
    e_dev = dev.handle.evaluate_fixed(src["R"][None])                    # the new parameters are current on the device
    dev.close()

    host = make()
    est = host.ParallelUpdateExpectationValues(uR0, uI0, 0.0, 0.0, *mc)
    solve = lambda e: timestep.solve_for_parameters_dot(e, imaginary_time=0)

    def stage(u_r, u_i, p_r, p_i):
        if reuse:
            return solve(host.ParallelUpdateExpectationValuesForGivenSamples(u_r, u_i, p_r, p_i))
        return solve(host.ParallelUpdateExpectationValues(u_r, u_i, p_r, p_i, *mc))

    d0 = solve(est)
    if method.startswith("PC"):
        t = (uR0 + d0[0] * dt, uI0 + d0[1] * dt, 0.0 + d0[2] * dt, 0.0 + d0[3] * dt)
        for _ in range(6 if reuse else 1):
            d1 = stage(*t)
            t = (uR0 + (d0[0] + d1[0]) * (dt / 2), uI0 + (d0[1] + d1[1]) * (dt / 2), (d0[2] + d1[2]) * (dt / 2), (d0[3] + d1[3]) * (dt / 2))
        want = t
    else:
        d = [d0]
        for step in (dt / 2, dt / 2, dt):
            k = d[-1]
            d.append(stage(uR0 + k[0] * step, uI0 + k[1] * step, k[2] * step, k[3] * step))
        comb = lambda j: (d[0][j] + d[1][j] * 2.0 + d[2][j] * 2.0 + d[3][j]) / 6.0
        want = (uR0 + comb(0) * dt, uI0 + comb(1) * dt, comb(2) * dt, comb(3) * dt)
    host.close()
    moved = np.max(np.abs(want[0] - uR0)) + np.max(np.abs(want[1] - uI0))
    assert moved > 1e-6
    assert np.max(np.abs(uR1 - want[0])) < 1e-8 * moved and np.max(np.abs(uI1 - want[1])) < 1e-8 * moved
    assert abs(pR1 - want[2]) < 1e-8 * max(abs(want[2]), 1e-12) and abs(pI1 - want[3]) < 1e-8 * max(abs(want[3]), 1e-12)
    h2 = capi.Handle(spec, 4)
    h2.set_params(uR1, uI1, pR1, pI1, 0.0)
    assert h2.evaluate_fixed(src["R"][None])["e_r"][0] == e_dev["e_r"][0]
    h2.close()


# ---------------------------------------------------------------------------------------------------
# BosonsBulk / NUBosonsBulkPB with DIM = 2 and DIM = 1 (config/BosonsBulk2D, NUBosonsBulkPB2D, Rydberg2D, BosonsBulk1D):
# same kernels, unused coordinates zero, secondDerivativeFactor = DIM - 1, DIM Gaussian components per move
# ---------------------------------------------------------------------------------------------------
LOWDIM_CASES = ["bosonsbulk2d_n16_equil", "nubosonsbulkpb2d_n25_equil", "rydberg2d_n50_equil", "bosonsbulk1d_n20_equil"]


@pytest.mark.parametrize("name", LOWDIM_CASES)
def test_low_dimensional_fixed_configuration(capi, golden, parity_log, name):
    g = golden(name)
    D = int(g["DIM"])
    spec, h = make_handle(capi, g)
    r = h.evaluate_fixed(g["R"])
    assert rel(r["ss"][0], g["spline_sums"]) < 1e-13 and r["outer"][0] == float(g["outer_sum"])
    assert rel(r["O"][0], g["local_operators"]) < RTOL
    assert abs(r["exponent"][0] - float(g["exponent"])) < RTOL * abs(float(g["exponent"]))
    assert abs(r["e_r"][0] - float(g["local_energy_r"])) < RTOL * abs(float(g["local_energy_r"]))
    assert abs(r["e_i"][0] - float(g["local_energy_i"])) < RTOL * abs(float(g["local_energy_i"]))
    for key in ("drift_r", "drift_i"):
        if np.max(np.abs(g[key])) == 0.0:                          # imaginary parameters all zero
            assert np.all(r[key][0] == 0.0)
            continue
        parity_log.check("low_dimensional", name, key, np.max(np.abs(r[key][0] - g[key])), np.max(np.abs(g[key])), RTOL, "max|F|")
        assert np.all(r[key][0][:, D:] == 0.0)
    want, got = g["other_expectation_values"], r["other"][0]
    for k in range(9):
        if want[k] == 0.0:
            assert got[k] == 0.0, k
            continue
        parity_log.check("low_dimensional", name, "other[%d]" % k, abs(got[k] - want[k]), abs(want[k]), RTOL)
    sD, sD2 = h.tables_fixed(g["R"])
    assert rel(sD[:, :, :D], g["sD"]) < 1e-13 and np.all(sD[:, :, D:] == 0.0) and rel(sD2, g["sD2"]) < 1e-13
    q, d = h.quotient_fixed(g["R"], g["moves"])
    d_ref = g["move_exponent_new"] - float(g["exponent"])
    assert np.max(np.abs(d - d_ref)) < 5e-8 and np.max(np.abs(q / g["move_quotient"] - 1.0)) < 1e-7
    with pytest.raises(capi.TdvmcError, match="DIM = 3"):
        from tdvmc_b200 import observables
        h.observables_fixed(observables.from_golden(golden("bosonsbulk_n64_obs")), g["R"][None])
    h.close()


@pytest.mark.parametrize("name", ["bosonsbulk2d_n16_equil", "rydberg2d_n50_equil", "bosonsbulk1d_n20_equil"])
def test_low_dimensional_chain_and_estimators_match_oracle(capi, golden, name):
    g = golden(name)
    D, N = int(g["DIM"]), int(g["N"])
    W, seed, mc_step = 5, 19, 0.4
    n_samples, n_therm, n_init = 3, 2 * N, 4 * N
    spec, h = make_handle(capi, g, n_walkers=W, seed=seed, mc_step=mc_step, max_samples=n_samples)
    o = Oracle(spec, time=float(g["time"]))
    shift = np.zeros(3)
    shift[0] = 0.002
    R0 = np.stack([g["R"] + shift * w for w in range(W)])
    h.set_positions(R0)
    h.sample_and_accumulate(n_samples, n_therm, n_init)
    got = h.allreduce_and_fetch()
    est = np.zeros(o.est_size())
    acc, Rf = 0, []
    for w in range(W):
        r = o.sample_walker(np.ascontiguousarray(R0[w][:, :D]), g["uR"], g["uI"], float(g["phiR"]), seed, w, 0, n_init, n_samples,
                            n_therm, mc_step, est)
        acc += r["accepted"]
        Rf.append(r["R"])
    want = o.unpack_est(est, W * n_samples)
    assert got["n_acceptances"] == acc and got["n_trials"] == W * (n_init + n_samples * n_therm) and acc > 0.2 * got["n_trials"]
    Rg = h.get_positions()
    assert np.all(Rg[:, :, D:] == 0.0)                             # the unused coordinates never move
    dlt = Rg[:, :, :D] - np.stack(Rf)
    dlt -= spec.lbox * np.round(dlt / spec.lbox)
    assert np.max(np.abs(dlt)) < 1e-9
    for k in ("O", "S", "OER", "OEI"):
        assert rel(got[k], want[k]) < 1e-9, k
    assert abs(got["e_r"][0] - want["e_r"]) < 1e-9 * abs(want["e_r"])
    assert abs(got["e_i"][0] - want["e_i"]) < 1e-9 * max(abs(want["e_i"]), 1e-3 * abs(want["e_r"]))
    h.close()


def _sweep_with_env(capi, spec, g, W, n_steps, env):
    """Positions and acceptance counts after n_steps (two launches) of W walkers under the given tuning knobs."""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        h = capi.Handle(spec, W, seed=21, mc_step=0.5, max_samples=1)
        h.set_params(g["uR"], g["uI"], float(g["phiR"]), float(g["phiI"]), float(g["time"]))
        rng = np.random.default_rng(3)
        h.set_positions(g["R"][None] + rng.uniform(-0.01, 0.01, (W, spec.n_particles, 3)))
        h.sweep(n_steps // 2)
        h.sweep(n_steps - n_steps // 2)
        R = h.get_positions()
        h.sample_and_accumulate(1, 0, 0)
        acc = h.allreduce_and_fetch()["n_acceptances"]
        h.close()
        return R, acc
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_time_shared_sweep_equals_plain_launch(capi, golden):
    """sweep_queue_kernel (ensembles that are not a whole number of waves: walkers time-share the resident warps, their
    state travelling through HBM between chunks) against the plain one-warp-per-walker launch: same proposal stream, same
    arithmetic per walker - bit-identical configurations and acceptance counts."""
    g = golden("bosonsbulk_n343_equil")
    spec = systems.from_golden(g)
    per_sm, sms = capi.Handle(spec, 1).resident_walkers()
    W = per_sm * sms + sms // 2 + 7                               # 1.03 waves: the plain launch idles most of the second
    Rq, aq = _sweep_with_env(capi, spec, g, W, 700, {"TDVMC_SWEEP_QUEUE": "1"})
    Rp, ap = _sweep_with_env(capi, spec, g, W, 700, {"TDVMC_SWEEP_QUEUE": "0"})
    assert aq == ap and 0.5 < aq / (W * 700.0) < 0.95
    assert np.array_equal(Rq, Rp)


def test_time_shared_sweep_small_ensemble_forced(capi, golden):
    """The same comparison on an ensemble small enough for compute-sanitizer's racecheck (TDVMC_SWEEP_QUEUE = 2 takes the
    time-shared launch whatever the wave count): 333 walkers of N = 64, most SMs holding two or three, chunks handed from warp
    to warp through the per-walker flags."""
    g = golden("bosonsbulk_n64_equil")
    spec = systems.from_golden(g)
    Rq, aq = _sweep_with_env(capi, spec, g, 333, 640, {"TDVMC_SWEEP_QUEUE": "2", "TDVMC_SWEEP_SPLIT": "1"})
    Rp, ap = _sweep_with_env(capi, spec, g, 333, 640, {"TDVMC_SWEEP_QUEUE": "0", "TDVMC_SWEEP_SPLIT": "1"})
    assert aq == ap and 0.3 < aq / (333 * 640.0) < 0.98
    assert np.array_equal(Rq, Rp)


@pytest.mark.parametrize("name,W", [("bosonsbulk_n343_equil", 300), ("nubosonsbulkpb_n1728_equil", 40)])
def test_split_sweep_equals_one_warp_per_walker(capi, golden, name, W):
    """sweep_split_kernel (several warps per walker for ensembles that do not fill the machine) against the one-warp launch:
    same stream, same accept rule; the exponent change is summed in a different order, so configurations agree to rounding
    and the acceptance counts exactly (a decision flips only if 2 delta sits within 1e-15 of log U)."""
    g = golden(name)
    spec = systems.from_golden(g)
    Rs, a_s = _sweep_with_env(capi, spec, g, W, 400, {"TDVMC_SWEEP_SPLIT": "4"})
    R1, a_1 = _sweep_with_env(capi, spec, g, W, 400, {"TDVMC_SWEEP_SPLIT": "1"})
    assert a_s == a_1
    assert np.max(np.abs(Rs - R1)) < 1e-12


@pytest.mark.parametrize("N,L", [(2744, 14.0), (8000, 20.0)])
def test_large_systems_beyond_one_sm_of_shared_memory(capi, N, L):
    """N = 2744: the configuration (positions + forces, 198 KB) no longer fits an SM next to the tables, the evaluation kernel
    keeps it in a per-block slab of global memory.  N = 8000, LBOX = 20, N_PARAM = 201 - config/BosonsBulk3D.config exactly as
    shipped: also one walker's positions (192 KB) nearly fill an SM, the sweep runs eight warps per walker with four table
    replicas.  Both against the pinned oracle: E_L, drift, O_k at 1e-10, chain replay move for move."""
    P = 201
    spec = systems.bosons_bulk(N, L, P, [1.0, 1.0])
    uR, uI = systems.smooth_params(P, L / 2)
    rng = np.random.default_rng(N)
    R = np.stack([systems.jittered_lattice(N, L, rng) + rng.uniform(-0.05, 0.05, (N, 3)) for _ in range(2)])
    h = capi.Handle(spec, 2, seed=3, mc_step=0.5)
    h.set_params(uR, uI, 0.0, 0.0, 0.0)
    ev = h.evaluate_fixed(R)
    o = Oracle(spec)
    for c in range(2):
        ref = o.evaluate(R[c], uR, uI, 0.0)
        assert rel(ev["O"][c], ref["O"]) < RTOL
        assert abs(ev["e_r"][c] - ref["e_r"]) < RTOL * abs(ref["e_r"]) and abs(ev["e_i"][c] - ref["e_i"]) < RTOL * abs(ref["e_i"])
        assert abs(ev["exponent"][c] - ref["exponent"]) < RTOL * abs(ref["exponent"])
        assert np.max(np.abs(ev["drift_r"][c] - ref["drift_r"])) < RTOL * np.max(np.abs(ref["drift_r"]))
        assert np.max(np.abs(ev["drift_i"][c] - ref["drift_i"])) < RTOL * np.max(np.abs(ref["drift_i"]))
    h.set_positions(R)
    h.sweep(60)
    Rg = h.get_positions()
    for c in range(2):
        Rr, _ = o.sweep(R[c], uR, 3, c, 0, 60, 0.5)
        d = Rg[c] - Rr
        d -= L * np.round(d / L)
        assert np.max(np.abs(d)) < 1e-9
    # a whole estimator pass runs (two samples per walker)
    h2 = capi.Handle(spec, 2, seed=3, mc_step=0.5, max_samples=2)
    h2.set_params(uR, uI, 0.0, 0.0, 0.0)
    h2.set_positions(R)
    h2.sample_and_accumulate(2, 32, 0)
    got = h2.allreduce_and_fetch()
    assert got["n_samples"] == 4 and np.isfinite(got["e_r"][0]) and got["n_trials"] == 2 * 64
    h.close()
    h2.close()


def _evaluate_with_env(h, R, env):
    """evaluate_fixed under tuning knobs (the evaluation kernel reads them at every launch)."""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return h.evaluate_fixed(R)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("name", ["bosonsbulk_n64_equil", "bosonsbulk_n343_equil", "nubosonsbulkpb_n1728_equil"])
def test_uniform_knot_interval_index_is_the_exact_search(capi, golden, name):
    """Uniform knots (BosonsBulk.cpp:61-67): the evaluation kernel takes the knot interval from r / h alone and consults the
    knots only for a distance within ~1e-10 h of one (find_bin_uniform).  Bit-identical outputs to the exact search
    (std::lower_bound, BosonsBulk.cpp:197-198, the kernel pinned to the reference above) on equilibrated configurations and
    on configurations with a pair placed ON knots, a few ulp and 1e-12 ... 1e-8 h beside them, and at the cut itself
    (third case: the reflection rule on the uniform NURBS_GRID of config/NUBosonsBulkPB3D.config, N = 1728)."""
    g = golden(name)
    spec, h = make_handle(capi, g, 2)
    N = spec.n_particles
    knots = np.asarray(g["knots"])
    hsp = knots[4] - knots[3]
    rng = np.random.default_rng(11)
    cfgs = [g["R"], g["R"] + rng.uniform(-0.02, 0.02, (N, 3))]
    K = len(knots) - 4
    for k in sorted({1, 2, min(17, K - 3), max(1, K - 40), K - 4, K - 3}):   # knots[3 + j] = j h; K - 3 is r_max
        for delta in (0.0, 2.3e-16, -2.3e-16, 1e-12, -1e-12, 3e-11, -3e-11, 1e-8, -1e-8):
            R = g["R"].copy()
            d = knots[3 + k] * (1.0 + delta) if delta != 0.0 and abs(delta) < 1e-15 else knots[3 + k] + delta * hsp
            R[1] = R[0] + np.array([d, 0.0, 0.0])        # (the minimum image folds it back where d > L/2 - never here)
            cfgs.append(R)
    R = np.stack(cfgs)
    exact = _evaluate_with_env(h, R, {"TDVMC_EVAL_UNIBIN": "0"})
    fast = _evaluate_with_env(h, R, {"TDVMC_EVAL_UNIBIN": "1"})
    for key in exact:
        assert np.array_equal(exact[key], fast[key]), key
    assert rel(fast["O"][0], g["local_operators"]) < RTOL
    h.close()


@pytest.mark.parametrize("name", ["bosonsbulk_n64_equil", "bosonsbulk_n343_equil"])
def test_sampler_exponent_change_against_exact_arithmetic(capi, golden, parity_log, name):
    """What the Metropolis sweep evaluates per proposal - minimum image of wrapped points, the one-Newton-step square root
    (sweep_math.cuh), the per-interval cubic in the local coordinate - against the polynomial the caller's spline table
    DEFINES, u(r) = sum_p u~[bin - p] (w0 + w1 r + w2 r^2 + w3 r^3)_{bin - p, p}, evaluated in 60-digit decimal arithmetic
    with an exact square root.  The reference's own double evaluation of the same expression carries ~1e-10 per pair term
    (the 1e-7 of test_scripted_move_quotient); the device stays within 1e-11 of the exact value."""
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    g = golden(name)
    spec, h = make_handle(capi, g)
    N, L = spec.n_particles, float(g["LBOX"])
    knots = np.asarray(g["knots"], np.float64)
    w = np.asarray(g["spline_weights"], np.float64)
    K = w.shape[0]
    ut = spec.spline_space(np.asarray(g["uR"], np.float64))
    u_tail = float(g["uR"][spec.tail_param]) if spec.tail_param >= 0 else 0.0
    rmax = knots[K]
    Ld = Decimal(L)

    def u_exact(a, b):
        s = Decimal(0)
        for c in range(3):
            d = Decimal(float(a[c])) - Decimal(float(b[c]))
            d -= Ld * (d / Ld).to_integral_value()          # nearest image
            s += d * d
        r = s.sqrt()
        rf = float(r)
        if rf > rmax:
            return Decimal(u_tail)
        b_ = int(np.searchsorted(knots, rf, side="left")) - 1   # knots[b] < r <= knots[b + 1]
        tot = Decimal(0)
        for p in range(4):
            cf = w[b_ - p, p]
            val = Decimal(float(cf[0])) + r * (Decimal(float(cf[1])) + r * (Decimal(float(cf[2])) + r * Decimal(float(cf[3]))))
            tot += Decimal(float(ut[b_ - p])) * val
        return tot

    R = np.asarray(g["R"], np.float64)
    moves = np.asarray(g["moves"], np.float64)
    rng = np.random.default_rng(5)
    extra = np.array([[p, *(R[p] + rng.normal(0.0, 0.5, 3))] for p in rng.integers(0, N, 6)])
    moves = np.concatenate([moves, extra])
    _, d = h.quotient_fixed(R, moves)
    worst = 0.0
    for m, dm in zip(moves, d):
        p = int(m[0])
        exact = sum((u_exact(R[i], m[1:]) - u_exact(R[i], R[p]) for i in range(N) if i != p), Decimal(0))
        worst = max(worst, abs(float(Decimal(float(dm)) - exact)))
    parity_log.check("test_sampler_exponent_change_against_exact_arithmetic", name, "delta_exponent", worst, 1.0, 1e-11,
                     "absolute, against 60-digit arithmetic on the table's polynomial")
    h.close()
