"""Pins oracle/tdvmc_oracle.c against fixtures produced by the UNMODIFIED reference
(oracle/_ref/ref_harness via oracle/gen_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle_lib import Oracle
from tdvmc_b200 import systems

EVAL_CASES = ["bosonsbulk_n64_fixture", "bosonsbulk_n64_equil", "bosonsbulk_n343_lattice", "bosonsbulk_n343_equil",
              "nubosonsbulkpb_n216_equil", "nubosonsbulkpb_n1728_equil"]
RTOL = 1e-10  # north_star: fixed-configuration E_L, drift, O_k within 1e-10 relative


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def test_min_image_known_answers():
    """The reference's 3-D known-answer cases (src/test/Tests.h:76-129), tolerance 1e-9 as there (:11)."""
    here = os.path.dirname(os.path.abspath(__file__))
    cases = json.load(open(os.path.join(here, "golden", "min_image_known_answers.json")))["cases"]
    spec = systems.bosons_bulk(8, 4.0, 9)
    o = Oracle(spec)
    assert len(cases) == 48
    for c in cases:
        a = np.array(c["a"] + [0.0] * (3 - len(c["a"])))
        b = np.array(c["b"] + [0.0] * (3 - len(c["b"])))
        n, d = o.min_image(c["L"], a, b)
        assert abs(n - c["norm"]) < 1e-9, c
        for k, want in enumerate(c["disp"]):
            assert abs(d[k] - want) < 1e-9, c


def test_min_image_bit_exact_vs_reference(golden):
    g = golden("min_image_reference")
    o = Oracle(systems.bosons_bulk(8, 4.0, 9))
    for L, a, b, n, d in zip(g["L"], g["a"], g["b"], g["norm"], g["disp"]):
        n2, d2 = o.min_image(float(L), a, b)
        assert n2 == n and np.array_equal(d2, d)


@pytest.mark.parametrize("name", EVAL_CASES)
def test_fixed_configuration_matches_reference(golden, name):
    g = golden(name)
    spec = systems.from_golden(g)
    o = Oracle(spec, time=float(g["time"]))
    r = o.evaluate(g["R"], g["uR"], g["uI"], float(g["phiR"]))
    # same arithmetic in the same order: the restatement reproduces the reference to the last bits
    assert rel(r["ss"], g["spline_sums"]) < 1e-14
    assert r["outer"] == float(g["outer_sum"])
    assert rel(r["O"], g["local_operators"]) < 1e-14
    assert abs(r["exponent"] - float(g["exponent"])) < 1e-12 * abs(float(g["exponent"]))
    assert rel(r["e_r"], g["local_energy_r"]) < 1e-12
    assert rel(r["e_i"], g["local_energy_i"]) < 1e-12
    assert rel(r["other"], g["other_expectation_values"][:9]) < 1e-12
    assert rel(r["drift_r"], g["drift_r"]) < RTOL
    assert rel(r["drift_i"], g["drift_i"]) < RTOL
    wn = g["table_checksum_weights"]
    assert rel(np.einsum("n,kna->ka", wn, r["sD"]), g["sD_checksum"]) < 1e-12
    assert rel(np.einsum("n,kn->k", wn, r["sD2"]), g["sD2_checksum"]) < 1e-12
    if "sD" in g:
        assert np.array_equal(r["sD"], g["sD"]) and np.array_equal(r["sD2"], g["sD2"])
    else:
        idx = g["table_particles"]
        assert np.array_equal(r["sD"][:, idx, :], g["sD_subset"])
        assert np.array_equal(r["sD2"][:, idx], g["sD2_subset"])


@pytest.mark.parametrize("name", EVAL_CASES)
def test_scripted_move_quotient_matches_reference(golden, name):
    g = golden(name)
    spec = systems.from_golden(g)
    o = Oracle(spec, time=float(g["time"]))
    for m, q_ref, en_ref in zip(g["moves"], g["move_quotient"], g["move_exponent_new"]):
        q, en, _ = o.quotient(g["R"], int(m[0]), m[1:4], g["uR"])
        assert abs(en - en_ref) < 1e-12 * abs(en_ref)
        assert abs(q - q_ref) < 1e-10 * q_ref


def test_sampler_statistics_match_reference(golden):
    """The oracle sampler (Philox proposals) against the reference sampler (mt19937_64): same
    distribution, different streams -> agreement within combined error bars (4 sigma)."""
    g = golden("bosonsbulk_n64_mc")
    spec = systems.bosons_bulk(int(g["N"]), float(g["LBOX"]), int(g["N_PARAM"]), g["SYSTEM_PARAMS"])
    o = Oracle(spec)
    n_therm, mc_step = int(g["n_therm"]), float(g["MC_STEP"])
    n_samples = 1500
    r = o.sample_walker(g["R0"], g["uR"], g["uI"], 0.0, seed=5, walker=0, step0=0, n_init=64 * 100,
                        n_samples=n_samples, n_therm=n_therm, mc_step=mc_step)
    er = r["rows"][:, spec.n_params]

    def blocked(x, nb=20):
        b = x[:len(x) // nb * nb].reshape(nb, -1).mean(axis=1)
        return b.mean(), b.std(ddof=1) / np.sqrt(nb)

    m1, s1 = blocked(er)
    m2, s2 = blocked(g["energy_r_series"])
    assert abs(m1 - m2) < 4.0 * np.hypot(s1, s2), (m1, s1, m2, s2)
    acc = r["accepted"] / r["steps"]
    assert abs(acc - float(g["acceptance"])) < 0.01
    # <O_k> profile
    est = o.unpack_est(r["est"], n_samples)
    O_ref = g["local_operators"]
    scale = np.abs(O_ref).max()
    assert np.max(np.abs(est["O"] - O_ref)) / scale < 0.02


HE_CASES = ["hebulk_n64_fixture", "hebulk_n64_equil", "hedrop_n6_fixture", "hedrop_n6_spread", "hedrop_n6_equil",
            "hedrop_n20_equil"]   # config/drop_20.config


@pytest.mark.parametrize("name", HE_CASES)
def test_he_family_fixed_configuration_matches_reference(golden, name):
    from oracle_lib import OracleHe

    g = golden(name)
    spec = systems.from_golden(g)           # also checks every boundary factor against the reference object, bit for bit
    o = OracleHe(spec)
    K = spec.n_splines
    r = o.evaluate(g["R"], g["uR"], g["uI"], float(g["phiR"]))
    assert rel(r["ss"], g["spline_sums"]) < 1e-14
    assert abs(r["mcm"] - float(g["mcmillan_sum"])) <= 1e-14 * abs(float(g["mcmillan_sum"]))
    if "const_sum" in g:
        assert r["ext"][K + 1] == float(g["const_sum"])
        assert abs(r["ext"][K + 2] - float(g["linear_sum"])) <= 1e-14 * abs(float(g["linear_sum"]))
    assert rel(r["O"], g["local_operators"]) < 1e-13
    assert abs(r["exponent"] - float(g["exponent"])) < 1e-13 * abs(float(g["exponent"]))
    assert abs(r["e_r"] - float(g["local_energy_r"])) < 1e-11 * abs(float(g["local_energy_r"]))
    assert abs(r["e_i"] - float(g["local_energy_i"])) < 1e-11 * abs(float(g["local_energy_i"]))
    assert rel(r["other"], g["other_expectation_values"]) < 1e-12
    assert rel(r["drift_r"], g["drift_r"]) < RTOL and rel(r["drift_i"], g["drift_i"]) < RTOL
    assert rel(r["tabD"][:K], g["sD"]) < 1e-14 and rel(r["tabD2"][:K], g["sD2"]) < 1e-14
    assert rel(r["tabD"][K], g["mcmillan_sum_d"]) < 1e-14 and rel(r["tabD2"][K], g["mcmillan_sum_d2"]) < 1e-14
    if "linear_sum_d" in g:
        assert rel(r["tabD"][K + 2], g["linear_sum_d"]) < 1e-14 and rel(r["tabD2"][K + 2], g["linear_sum_d2"]) < 1e-14
    for m, q_ref, en_ref in zip(g["moves"], g["move_quotient"], g["move_exponent_new"]):
        q, en, _ = o.quotient(g["R"], int(m[0]), m[1:4], g["uR"])
        assert abs(en - en_ref) < 1e-12 * abs(en_ref)
        assert abs(q - q_ref) < 1e-9 * q_ref


def test_hebulk_sampler_statistics_match_reference(golden):
    from oracle_lib import OracleHe

    g = golden("hebulk_n64_mc")
    spec = systems.he_bulk(int(g["N"]), float(g["LBOX"]), int(g["N_PARAM"]))
    o = OracleHe(spec)
    n_samples = 1200
    r = o.sample_walker(g["R0"], g["uR"], g["uI"], 0.0, seed=3, walker=0, step0=0, n_init=64 * 300, n_samples=n_samples,
                        n_therm=int(g["n_therm"]), mc_step=float(g["MC_STEP"]))
    er = r["rows"][:, spec.n_params]

    def blocked(x, nb=20):
        b = x[:len(x) // nb * nb].reshape(nb, -1).mean(axis=1)
        return b.mean(), b.std(ddof=1) / np.sqrt(nb)

    m1, s1 = blocked(er)
    m2, s2 = blocked(g["energy_r_series"])
    assert abs(m1 - m2) < 4.0 * np.hypot(s1, s2), (m1, s1, m2, s2)
    assert abs(r["accepted"] / r["steps"] - float(g["acceptance"])) < 0.015


MIX_CASES = ["mixture_he4he4na_fixture", "mixture_he4he4na_compact", "mixture_he4he4na_stretched", "mixture_he4he4na_equil",
             "mixture_he3he4cs_equil"]   # config/He3He4Cs.config: three pair types, KTTY He-Cs


MIX4_CASES = ["mixture4_he4he4na_fixture", "mixture4_he4he4na_compact", "mixture4_he4he4na_stretched", "mixture4_he4he4na_equil"]


@pytest.mark.parametrize("name", MIX_CASES + MIX4_CASES)
def test_mixture_fixed_configuration_matches_reference(golden, name):
    """BosonMixtureCluster (cubic splines) and BosonMixtureCluster_4thorder (quartic)."""
    from oracle_lib import OracleMix

    g = golden(name)
    spec = systems.from_golden(g)     # checks the pair-type numbering and hbar^2/2m against the reference object
    o = OracleMix(spec)
    K, T = spec.extra["n_splines"], int(g["n_pair_types"])
    assert K == (26 if name in MIX_CASES else 28) and g[f"spline_sums_0"].shape == (K,)
    r = o.evaluate(g["R"], g["uR"], g["uI"], float(g["phiR"]))
    for t in range(T):
        e = r["ext"][t * (K + 4):(t + 1) * (K + 4)]
        x = g[f"extras_{t}"]
        assert rel(e[:K], g[f"spline_sums_{t}"]) < 1e-14 or np.all(g[f"spline_sums_{t}"] == 0)
        for got, want in ((e[K], x[0]), (e[K + 1], x[1]), (e[K + 2], x[2]), (e[K + 3], x[3])):
            assert abs(got - want) <= 1e-14 * max(abs(want), 1e-300)
        b = t * (K + 4)
        assert rel(r["tabD"][b:b + K], g[f"sD_{t}"]) < 1e-14 or np.all(g[f"sD_{t}"] == 0)
        assert rel(r["tabD"][b + K + 3], g[f"log_sum_d_{t}"]) < 1e-14
        assert rel(r["tabD2"][b + K + 3], g[f"log_sum_d2_{t}"]) < 1e-14
    assert rel(r["O"], g["local_operators"]) < 1e-13
    assert abs(r["exponent"] - float(g["exponent"])) < 1e-13 * abs(float(g["exponent"]))
    assert abs(r["e_r"] - float(g["local_energy_r"])) < 1e-11 * abs(float(g["local_energy_r"]))
    assert abs(r["e_i"] - float(g["local_energy_i"])) < 1e-11 * abs(float(g["local_energy_i"]))
    assert rel(r["other"], g["other_expectation_values"]) < 1e-12
    assert rel(r["drift_r"], g["drift_r"]) < RTOL and rel(r["drift_i"], g["drift_i"]) < RTOL
    for m, q_ref, en_ref in zip(g["moves"], g["move_quotient"], g["move_exponent_new"]):
        q, en, _ = o.quotient(g["R"], int(m[0]), m[1:4], g["uR"])
        assert abs(en - en_ref) < 1e-12 * abs(en_ref)
        assert abs(q - q_ref) <= 1e-9 * abs(q_ref)


# ---------------------------------------------------------------------------------------------------
# additional observables g(r), S(k) (BosonsBulk.cpp:474-520, NUBosonsBulkPB.cpp:597-639)
# ---------------------------------------------------------------------------------------------------
OBS_CASES = ["bosonsbulk_n64_obs", "bosonsbulk_n343_obs", "nubosonsbulkpb_n216_obs"]


@pytest.mark.parametrize("name", OBS_CASES)
def test_observables_oracle_matches_reference(golden, name):
    from oracle_lib import oracle_observables
    from tdvmc_b200 import observables
    g = golden(name)
    src = golden(str(g["source"]))
    obs = observables.from_golden(g)
    gr, sk = oracle_observables(float(g["LBOX"]), src["R"], obs)
    assert np.max(np.abs(gr - g["gr_fixed"])) <= 1e-12 * np.max(np.abs(g["gr_fixed"]))
    assert np.max(np.abs(sk - g["sk_fixed"])) <= 1e-11 * np.max(np.abs(g["sk_fixed"]))
    # every pair inside the grid is counted once: sum_b gr[b] scaling[b] / weight = number of pairs with r < max
    # (minus those the reference drops past the end of its vector)
    n_pairs = np.sum(gr * obs.gr_scaling) / obs.gr_weight
    assert abs(n_pairs - round(n_pairs)) < 1e-6 and 0 < n_pairs <= int(g["N"]) * (int(g["N"]) - 1) / 2


@pytest.mark.parametrize("name", OBS_CASES)
def test_observable_grid_builder_matches_reference(golden, name):
    """pair_distribution_grid / wave_vectors rebuild the reference's grid, shell volumes and scaled wave vectors."""
    from tdvmc_b200 import observables
    g = golden(name)
    count, spacing, scaling = observables.pair_distribution_grid(float(g["gr_max"]), int(g["GR_BIN_COUNT"]) if str(g["system"]) == "BosonsBulk"
                                                                 else int(g["gr_count"]))
    assert count == int(g["gr_count"]) and spacing == float(g["gr_spacing"])
    assert np.max(np.abs(scaling - g["gr_scaling"])) <= 1e-15 * np.max(g["gr_scaling"])
    obs = observables.from_golden(g)
    L = float(g["LBOX"])
    ints = np.rint(obs.kvec * L / (2 * np.pi)).astype(int)
    shells = [ints[obs.shell_ptr[k]:obs.shell_ptr[k + 1]].tolist() for k in range(obs.n_shells)]
    ptr, kv = observables.wave_vectors(shells, L)
    assert np.array_equal(ptr, obs.shell_ptr) and np.array_equal(kv, obs.kvec)


@pytest.mark.parametrize("tag", ["fixture", "compact", "stretched", "equil"])
def test_mixture_observables_oracle_matches_reference(golden, tag):
    from oracle_lib import OracleMix, oracle_mix_observables
    g = golden("mixture_he4he4na_obs")
    src = golden(f"mixture_he4he4na_{tag}")
    o = OracleMix(systems.from_golden(src))
    r2, angle, density, distance = oracle_mix_observables(o, src["R"], g)
    assert abs(r2 - float(g[f"{tag}_r2_fixed"])) < 1e-13 * r2
    assert np.array_equal(angle, g[f"{tag}_angle_fixed"])
    assert np.array_equal(distance, g[f"{tag}_distance_fixed"])
    assert np.max(np.abs(density - g[f"{tag}_density_fixed"])) <= 1e-15 * np.max(g[f"{tag}_density_fixed"])
    assert angle.sum() == 3 and distance.sum() == 3


@pytest.mark.parametrize("name", ["hebulk_n64_obs", "hedrop_n6_obs"])
def test_he_structure_factor_oracle_matches_reference(golden, name):
    """S(k) part of HeBulk / HeDrop CalculateAdditionalSystemProperties (HeBulk.cpp:432-447, HeDrop.cpp:674-689)."""
    from oracle_lib import oracle_observables
    from tdvmc_b200 import observables
    g = golden(name)
    src = golden(str(g["source"]))
    ptr = np.concatenate([[0], np.cumsum(g["k_shell_sizes"].astype(np.int64))]).astype(np.int32)
    obs = observables.ObservableSpec(0, 1.0, 1.0, 1.0, np.ones(1), ptr, g["k_vectors"].reshape(-1, 3))
    _, sk = oracle_observables(float(g["LBOX"]), src["R"], obs)
    assert np.max(np.abs(sk - g["sk_fixed"])) <= 1e-11 * np.max(np.abs(g["sk_fixed"]))
    assert np.array_equal(g["other_fixed"], src["other_expectation_values"])      # the first block is the ordinary "other" values


# ---------------------------------------------------------------------------------------------------
# NUBosonsBulkPBBoxAndRadial: radial + box spline bases
# ---------------------------------------------------------------------------------------------------
BR_CASES = ["boxradial_n27_jittered", "boxradial_n27_equil", "boxradial_n64_equil",
            "boxradial2d_n25_equil"]   # config/NUBosonsBulkPBBoxAndRadial2D.config (DIM = 2)


@pytest.mark.parametrize("name", BR_CASES)
def test_boxradial_fixed_configuration_matches_reference(golden, name):
    from oracle_lib import OracleBR, br_shell_volumes

    g = golden(name)
    spec = systems.from_golden(g)     # checks knots (and that the reference's two bases share knots and table)
    o = OracleBR(spec, time=float(g["time"]))
    K = spec.extra["n_splines"]
    assert spec.r_max == float(g["max_distance_rad"]) and spec.extra["half"] == float(g["half_length"])
    vol, spacing = br_shell_volumes(spec.extra["half"], spec.extra["gr_bins"], spec.dim)
    assert np.array_equal(vol, g["gr_bin_volumes"]) and spacing == float(g["gr_node_point_spacing"])
    r = o.evaluate(g["R"], g["uR"], g["uI"], float(g["phiR"]))
    assert rel(r["ext"][:K], g["spline_sums_rad"]) < 1e-13 and rel(r["ext"][K:], g["spline_sums"]) < 1e-13
    # the four tables: same operations in the same order as the reference -> bit for bit
    D = spec.dim
    assert np.array_equal(r["tabD"][:K, :, :D], g["sD_rad"]) and np.array_equal(r["tabD"][K:, :, :D], g["sD"])
    assert np.all(r["tabD"][:, :, D:] == 0.0)
    assert np.array_equal(r["tabD2"][:K], g["sD2_rad"]) and np.array_equal(r["tabD2"][K:], g["sD2"])
    assert rel(r["O"], g["local_operators"]) < 1e-13
    assert abs(r["exponent"] - float(g["exponent"])) < 1e-13 * abs(float(g["exponent"]))
    assert abs(r["e_r"] - float(g["local_energy_r"])) < 1e-12 * abs(float(g["local_energy_r"]))
    assert abs(r["e_i"] - float(g["local_energy_i"])) < 1e-12 * abs(float(g["local_energy_i"]))
    assert rel(r["other"], g["other_expectation_values"]) < 1e-12
    assert np.array_equal(r["other"][3:], g["gr_bins"])
    assert rel(r["drift_r"], g["drift_r"]) < RTOL and rel(r["drift_i"], g["drift_i"]) < RTOL
    for m, q_ref, en_ref in zip(g["moves"], g["move_quotient"], g["move_exponent_new"]):
        q, en, _ = o.quotient(g["R"], int(m[0]), m[1:4], g["uR"])
        assert abs(en - en_ref) < 1e-12 * abs(en_ref)
        assert abs(q - q_ref) <= 1e-9 * abs(q_ref)


def test_boxradial_sampler_statistics_match_reference(golden):
    """Oracle sampler (Philox proposals) against the reference's own Metropolis run (mt19937_64) at the config's size:
    energy within 4 combined standard errors (blocked), acceptance within 0.01, <O_k> profile within 2 % of its scale."""
    from oracle_lib import OracleBR

    g = golden("boxradial_n27_mc")
    src = golden(str(g["source"]))
    spec = systems.from_golden(src)
    o = OracleBR(spec, time=float(src["time"]))
    n_samples = 3000
    r = o.sample_walker(src["R"], src["uR"], src["uI"], float(src["phiR"]), seed=21, walker=0, step0=0, n_init=27 * 50,
                        n_samples=n_samples, n_therm=int(g["MC_NTHERMSTEPS"]), mc_step=float(g["MC_STEP"]))

    def blocked(x, nb=20):
        b = x[:len(x) // nb * nb].reshape(nb, -1).mean(axis=1)
        return b.mean(), b.std(ddof=1) / np.sqrt(nb)

    m1, s1 = blocked(r["rows"][:, spec.n_params])
    m2, s2 = blocked(g["energy_r_series"])
    assert abs(m1 - m2) < 4.0 * np.hypot(s1, s2), (m1, s1, m2, s2)
    assert abs(r["accepted"] / r["steps"] - float(g["acceptance"])) < 0.01
    est = o.unpack_est(r["est"], n_samples)
    assert np.max(np.abs(est["O"] - g["local_operators"])) / np.abs(g["local_operators"]).max() < 0.02


# ---------------------------------------------------------------------------------------------------
# InhContactBosons: one-dimensional, single-particle function + pair correlation
# ---------------------------------------------------------------------------------------------------
INH_CASES = ["inhcontact_n3_fixture", "inhcontact_n3_well", "inhcontact_n3_equil", "inhcontact_n20_equil"]


@pytest.mark.parametrize("name", INH_CASES)
def test_inhcontact_fixed_configuration_matches_reference(golden, name):
    from oracle_lib import OracleInh

    g = golden(name)
    spec = systems.from_golden(g)     # checks gamma and maxDistance against the reference object
    o = OracleInh(spec)
    K1 = spec.extra["n_splines_spf"]
    r = o.evaluate(g["R"], g["uR"], g["uI"], float(g["phiR"]))
    assert rel(r["ext"][:K1], g["spline_sums_spf"]) < 1e-13 and rel(r["ext"][K1:], g["spline_sums_pc"]) < 1e-13
    assert np.array_equal(r["tabD"][:K1], g["sD_spf"][:, :, 0]) and np.array_equal(r["tabD"][K1:], g["sD_pc"][:, :, 0])
    assert np.array_equal(r["tabD2"][:K1], g["sD2_spf"]) and np.array_equal(r["tabD2"][K1:], g["sD2_pc"])
    assert rel(r["O"], g["local_operators"]) < 1e-13
    assert abs(r["exponent"] - float(g["exponent"])) < 1e-12 * max(abs(float(g["exponent"])), 1.0)
    assert abs(r["e_r"] - float(g["local_energy_r"])) < 1e-12 * abs(float(g["local_energy_r"]))
    assert abs(r["e_i"] - float(g["local_energy_i"])) < 1e-12 * abs(float(g["local_energy_i"]))
    assert rel(r["other"], g["other_expectation_values"]) < 1e-12
    assert rel(r["drift_r"], g["drift_r"]) < RTOL and rel(r["drift_i"], g["drift_i"]) < RTOL
    for m, q_ref, en_ref in zip(g["moves"], g["move_quotient"], g["move_exponent_new"]):
        q, en, _ = o.quotient(g["R"], int(m[0]), m[1:4], g["uR"])
        assert abs(en - en_ref) < 1e-12 * max(abs(en_ref), 1.0)
        assert abs(q - q_ref) <= 1e-9 * abs(q_ref)


def test_inhcontact_sampler_statistics_match_reference(golden):
    from oracle_lib import OracleInh

    g = golden("inhcontact_n3_mc")
    src = golden(str(g["source"]))
    spec = systems.from_golden(src)
    o = OracleInh(spec)
    n_samples = 6000
    r = o.sample_walker(src["R"], src["uR"], src["uI"], float(src["phiR"]), seed=77, walker=0, step0=0, n_init=600,
                        n_samples=n_samples, n_therm=int(g["MC_NTHERMSTEPS"]), mc_step=float(g["MC_STEP"]))

    def blocked(x, nb=20):
        b = x[:len(x) // nb * nb].reshape(nb, -1).mean(axis=1)
        return b.mean(), b.std(ddof=1) / np.sqrt(nb)

    m1, s1 = blocked(r["rows"][:, spec.n_params])
    m2, s2 = blocked(g["energy_r_series"])
    assert abs(m1 - m2) < 4.0 * np.hypot(s1, s2), (m1, s1, m2, s2)
    assert abs(r["accepted"] / r["steps"] - float(g["acceptance"])) < 0.015
    # <O_k>: three particles give noisy operators - compare each with its own blocked standard error (both runs have
    # the same number of samples, hence sqrt(2)), 5 sigma
    rows = r["rows"][:, :spec.n_params]
    nb = 20
    blocks = rows[:len(rows) // nb * nb].reshape(nb, -1, spec.n_params).mean(axis=1)
    sem = blocks.std(axis=0, ddof=1) / np.sqrt(nb)
    dev = np.abs(rows.mean(axis=0) - g["local_operators"]) / (np.sqrt(2.0) * np.maximum(sem, 1e-6))
    assert dev.max() < 5.0, dev.max()


# ---------------------------------------------------------------------------------------------------
# BosonsBulk / NUBosonsBulkPB in one and two dimensions (the reference's own low-dimensional configs)
# ---------------------------------------------------------------------------------------------------
LOWDIM_CASES = ["bosonsbulk2d_n16_equil", "nubosonsbulkpb2d_n25_equil", "rydberg2d_n50_equil", "bosonsbulk1d_n20_equil"]


@pytest.mark.parametrize("name", LOWDIM_CASES)
def test_low_dimensional_fixed_configuration_matches_reference(golden, name):
    g = golden(name)
    spec = systems.from_golden(g)
    D = int(g["DIM"])
    assert spec.dim == D and D < 3 and np.all(g["R"][:, D:] == 0.0)
    o = Oracle(spec, time=float(g["time"]))
    R = np.ascontiguousarray(g["R"][:, :D])
    r = o.evaluate(R, g["uR"], g["uI"], float(g["phiR"]))
    assert rel(r["ss"], g["spline_sums"]) < 1e-14 and r["outer"] == float(g["outer_sum"])
    assert rel(r["O"], g["local_operators"]) < 1e-14
    assert abs(r["exponent"] - float(g["exponent"])) < 1e-12 * abs(float(g["exponent"]))
    assert rel(r["e_r"], g["local_energy_r"]) < 1e-12 and rel(r["e_i"], g["local_energy_i"]) < 1e-12
    assert rel(r["other"], g["other_expectation_values"][:9]) < 1e-12
    assert rel(r["drift_r"], g["drift_r"][:, :D]) < RTOL and rel(r["drift_i"], g["drift_i"][:, :D]) < RTOL
    assert r["sD"].shape == g["sD"].shape == (spec.n_splines, spec.n_particles, D)
    assert np.array_equal(r["sD"], g["sD"]) and np.array_equal(r["sD2"], g["sD2"])
    for m, q_ref, en_ref in zip(g["moves"], g["move_quotient"], g["move_exponent_new"]):
        q, en, _ = o.quotient(R, int(m[0]), m[1:1 + D], g["uR"])
        assert abs(en - en_ref) < 1e-12 * abs(en_ref) and abs(q - q_ref) < 1e-10 * q_ref
