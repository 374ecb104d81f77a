/*
 * tdvmc_oracle — plain-C restatement of the reference's walker hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library, and only as
 * the checker.  The product (tdvmc_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against
 * fixtures produced by the unmodified reference (oracle/_ref/ref_harness, oracle/gen_golden.py).
 *
 * Each function cites the reference code it follows (paths relative to mathiasgartner/TDVMC).
 * Arithmetic is IEEE double in the reference's association order; compile without FMA
 * contraction (see oracle/Makefile: -ffp-contract=off).
 */
#ifndef TDVMC_ORACLE_H
#define TDVMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum
{
    ORACLE_PAIR_RULE_CUT = 0,    /* BosonsBulk: r <= r_max spline, else tail count (BosonsBulk.cpp:195-210) */
    ORACLE_PAIR_RULE_REFLECT = 1 /* NUBosonsBulkPB: r -> 2 r_max - r beyond r_max (NUBosonsBulkPB.cpp:249-269) */
};

typedef struct oracle_system
{
    int32_t n_particles;   /* N */
    int32_t dim;           /* D (3) */
    int32_t n_params;      /* P = N_PARAM */
    int32_t n_splines;     /* K = #knots - 4 */
    int32_t pair_rule;     /* ORACLE_PAIR_RULE_* */
    int32_t tail_param;    /* parameter multiplying the tail count in the exponent (P-1) */
    double lbox;           /* LBOX */
    double r_max;          /* maxDistance = knots[#knots-4] */
    double hbar2_2m;       /* HBAR2_2M (Constants.h:12) */
    double pot_a, pot_b;   /* square well width/height, already time-switched (BosonsBulk.cpp:237-243) */
    const double* knots;   /* [K+4] */
    const double* weights; /* [K][4][4] monomial coefficients, SplineFactory::GetWeights3 layout */
    const int32_t* map_ptr;/* [P+1] CSR rows of the boundary-condition map O_p = sum_j val_j ss[col_j] */
    const int32_t* map_col;
    const double* map_val;
} oracle_system;

/* Utils.cpp:266-281, 329-338, 368-374: minimum image displacement a - b and its norm */
double oracle_min_image(double lbox, int dim, const double* a, const double* b, double* disp);

/* BosonsBulk.cpp:179-218 / NUBosonsBulkPB.cpp:234-277: basis sums ss[K] and tail count */
void oracle_basis_sums(const oracle_system* s, const double* R, double* ss, double* outer);

/* BosonsBulk.cpp:158-177 / NUBosonsBulkPB.cpp:219-232: O_p from ss through the map */
void oracle_local_operators(const oracle_system* s, const double* ss, double* O);

/* BosonsBulk.cpp:522-545: exponent = sum_p uR_p O_p + uR[tail] * outer */
double oracle_exponent(const oracle_system* s, const double* O, double outer, const double* uR);

/* BosonsBulk.cpp:220-336 / NUBosonsBulkPB.cpp:279-414: sD[K][N][D], sD2[K][N], V_int */
void oracle_tables(const oracle_system* s, const double* R, double* sD, double* sD2, double* v_int);

/* BosonsBulk.cpp:349-458 / NUBosonsBulkPB.cpp:427-560: contraction to drift, E_L and other[9].
 * drift_r/drift_i ([N][D], may be NULL) receive vecKineticSumR1/I1 per particle. */
void oracle_expectation(const oracle_system* s, const double* O, const double* sD, const double* sD2,
                        double v_int, double exponent, double phiR, const double* uR, const double* uI,
                        double* e_r, double* e_i, double* other9, double* drift_r, double* drift_i);

/* BosonsBulk.cpp:553-657 / NUBosonsBulkPB.cpp:672-773: single-particle move, returns the quotient
 * exp(2 (exponent_new - exponent)); ss_new[K], outer_new, exponent_new are the proposal state. */
double oracle_wf_quotient(const oracle_system* s, const double* R, int particle, const double* old_pos,
                          const double* ss, double outer, double exponent, const double* uR,
                          double* ss_new, double* outer_new, double* exponent_new);

/* Counter-based proposal stream shared with the CUDA path (Philox4x32-10, key = seed,
 * counter = (step, walker, call)).  NOT the reference's mt19937_64 stream (Utils.cpp:5-46):
 * sampled trajectories are only statistically comparable with the reference. */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void oracle_proposal(uint64_t seed, uint32_t walker, uint64_t step, int n_particles, double mc_step,
                     int* particle, double disp[3], double* log_u);

/* src/TDVMC.cpp:858-916 with oracle_proposal() as the random source.  State per walker:
 * R[N][D], ss[K], outer, exponent.  Returns number of accepted moves. */
int64_t oracle_sweep(const oracle_system* s, double* R, double* ss, double* outer, double* exponent,
                     const double* uR, uint64_t seed, uint32_t walker, uint64_t first_step, int64_t n_steps,
                     double mc_step);

/* src/TDVMC.cpp:1038-1150 for one walker: n_init steps, then n_samples x (n_therm steps + evaluation).
 * Adds (unnormalised) sums into est: [O(P) | E_R | E_I | S(P*P) | OE_R(P) | OE_I(P) | other(9)].
 * sample_rows (may be NULL) receives per sample [O(P), E_R, E_I]. Returns accepted moves. */
int64_t oracle_sample_walker(const oracle_system* s, double* R, const double* uR, const double* uI, double phiR,
                             uint64_t seed, uint32_t walker, uint64_t* step_counter, int n_init, int n_samples,
                             int n_therm, double mc_step, double* est, double* sample_rows);

#ifdef __cplusplus
}
#endif
#endif

/* ------------------------------------------------------------------------------------------------
 * HeBulk (src/PhysicalSystems/HeBulk.cpp): periodic He-4, McMillan r^-5 core + uniform cubic
 * B-splines in the local coordinate, Aziz HFD-B(He) potential inline, g(r) in other[3..102].
 * ------------------------------------------------------------------------------------------------ */
#ifndef TDVMC_ORACLE_HE_H
#define TDVMC_ORACLE_HE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_hebulk
{
    int32_t n_particles, n_params, n_splines, gr_bins;
    double lbox, rij_split, h, max_distance, hbar2_2m;
    double f[8]; /* factorFirstSpline1, FirstSpline2, SecondSpline1, SecondSpline2, SecondLastSpline, LastSpline,
                    SecondLastSplinePhi, LastSplinePhi (HeBulk.cpp:57-67) */
} oracle_hebulk;

/* HeBulk::InitSystem (HeBulk.cpp:40-70) */
void oracle_hebulk_init(oracle_hebulk* s, int n_particles, double lbox, int n_params);
/* value sums of CalculateWavefunction (HeBulk.cpp:462-489) */
void oracle_hebulk_values(const oracle_hebulk* s, const double* R, double* ss, double* mcm);
/* localOperators (HeBulk.cpp:376-383) and the exponent (:491-498) */
void oracle_hebulk_operators(const oracle_hebulk* s, const double* ss, double mcm, double* O);
double oracle_hebulk_exponent(const oracle_hebulk* s, const double* ss, double mcm, const double* uR);
/* CalculateExpectationValues (HeBulk.cpp:166-405). other: [3 + gr_bins]; drift_*: [N][3] or NULL;
 * sD [K][N][3], sD2 [K][N], mcD [N][3], mcD2 [N]: caller-provided scratch (also outputs). */
void oracle_hebulk_expectation(const oracle_hebulk* s, const double* R, double wf, const double* uR, const double* uI,
                               double* e_r, double* e_i, double* other, double* drift_r, double* drift_i, double* sD,
                               double* sD2, double* mcD, double* mcD2);
/* CalculateWFChange / Quotient (HeBulk.cpp:512-604) */
double oracle_hebulk_quotient(const oracle_hebulk* s, const double* R, int particle, const double* old_pos, const double* ss,
                              double mcm, double exponent, const double* uR, double* ss_new, double* mcm_new,
                              double* exponent_new);
int64_t oracle_hebulk_sweep(const oracle_hebulk* s, double* R, double* ss, double* mcm, double* exponent, const double* uR,
                            uint64_t seed, uint32_t walker, uint64_t first_step, int64_t n_steps, double mc_step);
/* est: [O(P) | E_R | E_I | S(P*P) | OE_R(P) | OE_I(P) | other(3 + gr_bins)] sums; rows: per sample [O(P), E_R, E_I] */
int64_t oracle_hebulk_sample_walker(const oracle_hebulk* s, double* R, const double* uR, const double* uI, uint64_t seed,
                                    uint32_t walker, uint64_t* step_counter, int n_init, int n_samples, int n_therm,
                                    double mc_step, double* est, double* sample_rows);

#ifdef __cplusplus
}
#endif
#endif
