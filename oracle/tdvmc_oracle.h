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

/* ------------------------------------------------------------------------------------------------
 * NUBosonsBulkPBBoxAndRadial (tdvmc_oracle_br.c): radial + box spline bases on one knot vector
 * ---------------------------------------------------------------------------------------------- */
typedef struct oracle_br
{
    int32_t n_particles, n_params, n_splines, gr_bins; /* K = N_PARAM/2 + 3 splines per basis */
    int32_t dim, pad;  /* DIM (3, or 2 / 1: R stays [N][3] with the unused coordinates zero) */
    double lbox;
    double r_max;       /* maxDistanceRad = nodesRad[size - 4] (NUBosonsBulkPBBoxAndRadial.cpp:96) */
    double pot_a, pot_b; /* Gauss potential b exp(-(r/a)^2/2), time switch applied (:290-297) */
    double gr_max, gr_spacing; /* halfLength, halfLength / grBinCount (:146-147) */
    const double* gr_volumes;  /* [gr_bins] shell volumes (:149-169) */
    const double* knots;       /* [K + 4] */
    const double* weights;     /* [K][4][4] */
    const int32_t* map_ptr;    /* CSR of RefreshLocalOperators (:193-211) over ext = [ssRad | ss] */
    const int32_t* map_col;
    const double* map_val;
} oracle_br;
void oracle_br_values(const oracle_br* s, const double* R, double* ext);
void oracle_br_operators(const oracle_br* s, const double* ext, double* O);
double oracle_br_exponent(const oracle_br* s, const double* ext, const double* uR);
void oracle_br_expectation(const oracle_br* s, const double* R, double wf, const double* uR, const double* uI, double* e_r,
                           double* e_i, double* other, double* drift_r, double* drift_i, double* tabD, double* tabD2);
double oracle_br_quotient(const oracle_br* s, const double* R, int particle, const double* old_pos, const double* ext,
                          double exponent, const double* uR, double* ext_new, double* exponent_new);
int64_t oracle_br_sweep(const oracle_br* s, double* R, double* ext, double* exponent, const double* uR, uint64_t seed,
                        uint32_t walker, uint64_t first_step, int64_t n_steps, double mc_step);
int64_t oracle_br_sample_walker(const oracle_br* s, double* R, const double* uR, const double* uI, double phiR, uint64_t seed,
                                uint32_t walker, uint64_t* step_counter, int n_init, int n_samples, int n_therm, double mc_step,
                                double* est, double* sample_rows);

/* ------------------------------------------------------------------------------------------------
 * InhContactBosons (tdvmc_oracle_inh.c): one-dimensional, single-particle + pair-correlation splines
 * ---------------------------------------------------------------------------------------------- */
typedef struct oracle_inh
{
    int32_t n_particles, n_params, n_splines_spf, n_splines_pc; /* K1, K2 */
    double lbox;
    double r_max;      /* maxDistance = pc.nodes[size - 4] (InhContactBosons.cpp:100) */
    double h_pc;       /* pc.nodeSpacing */
    double gamma;      /* contact strength (:25-29) */
    double pot_range, pot_strength; /* square well, SYSTEM_PARAMS[0], [1] (:327-334) */
    double ext_k, ext_v0;           /* lattice potential k^2 V0 sin^2(k x), SYSTEM_PARAMS[2], [3] (:448-467) */
    double hbar2_2m;
    const double* knots_spf;   /* [K1 + 4] */
    const double* weights_spf; /* [K1][4][4] */
    const double* knots_pc;    /* [K2 + 4] */
    const double* weights_pc;  /* [K2][4][4] */
    const int32_t* map_ptr;    /* CSR of RefreshLocalOperators (:208-247) over ext = [ss_spf | ss_pc] */
    const int32_t* map_col;
    const double* map_val;
} oracle_inh;
void oracle_inh_values(const oracle_inh* s, const double* R, double* ext);
void oracle_inh_operators(const oracle_inh* s, const double* ext, double* O);
double oracle_inh_exponent(const oracle_inh* s, const double* ext, const double* uR);
void oracle_inh_expectation(const oracle_inh* s, const double* R, double wf, double exponent, const double* uR, const double* uI,
                            double* e_r, double* e_i, double* other, double* drift_r, double* drift_i, double* tabD, double* tabD2);
double oracle_inh_quotient(const oracle_inh* s, const double* R, int particle, const double* old_pos, const double* ext,
                           double exponent, const double* uR, double* ext_new, double* exponent_new);
int64_t oracle_inh_sweep(const oracle_inh* s, double* R, double* ext, double* exponent, const double* uR, uint64_t seed,
                         uint32_t walker, uint64_t first_step, int64_t n_steps, double mc_step);
int64_t oracle_inh_sample_walker(const oracle_inh* s, double* R, const double* uR, const double* uI, double phiR, uint64_t seed,
                                 uint32_t walker, uint64_t* step_counter, int n_init, int n_samples, int n_therm, double mc_step,
                                 double* est, double* sample_rows);

#ifdef __cplusplus
}
#endif
#endif

/* ------------------------------------------------------------------------------------------------
 * He family: HeBulk (src/PhysicalSystems/HeBulk.cpp) and HeDrop (src/PhysicalSystems/HeDrop.cpp).
 * McMillan core r^m below rs, uniform cubic B-splines written in the local coordinate on one (HeBulk)
 * or two (HeDrop: spacing 0.1 then 0.5) grids, const + linear tails beyond r_tail (HeDrop), Aziz HFD-B
 * (HeBulk.cpp:187-195) or Lennard-Jones (HeDrop.cpp:389-394) potential, g(r) (and, for HeDrop, the
 * density profile around the centre of mass) carried in otherExpectationValues.
 * Extended basis sums: ext = [ss_0 .. ss_{K-1} | mcMillanSum | constSum | linearSum]; the parameter map
 * (HeBulk.cpp:376-383, HeDrop.cpp:609-626) is passed as CSR rows over ext plus per-parameter constants.
 * ------------------------------------------------------------------------------------------------ */
#ifndef TDVMC_ORACLE_HE_H
#define TDVMC_ORACLE_HE_H
#ifdef __cplusplus
extern "C" {
#endif

/* BosonsBulk.cpp:474-520 / NUBosonsBulkPB.cpp:597-639 (CalculateAdditionalSystemProperties) for one configuration:
 *   gr[bin] += weight / scaling[bin] for every pair i > j with r < gr_max, bin = floor(r / spacing)
 *              (Grid.cpp:52-61, ObservableVsOnGridWithScaling.cpp:47-52); pairs whose bin falls beyond gr_count
 *              (possible because count = (max-min)/spacing truncates, Grid.cpp:21) are dropped - the reference
 *              writes them past the end of its vector;
 *   sk[k]   = ((sum_{i,kn} cos k_kn.R_i)^2 + (sum_{i,kn} sin k_kn.R_i)^2) / (N n_k).
 * kvec holds the wave vectors already multiplied by 2 pi / L (BosonsBulk.cpp:127-137), shell by shell. */
void oracle_observables(double lbox, int n_particles, const double* R, int gr_count, double gr_spacing, double gr_max,
                        double gr_weight, const double* gr_scaling, int n_shells, const int32_t* shell_ptr,
                        const double* kvec, double* gr, double* sk);

typedef struct oracle_he
{
    int32_t n_particles, n_params, n_splines, n_short; /* n_short: numberOfShortSplines (HeBulk: = n_splines) */
    int32_t periodic;      /* 1: minimum image in a box of lbox (HeBulk), 0: open (HeDrop) */
    int32_t potential;     /* 0: Aziz HFD-B(He) inline (HeBulk), 1: LJ sigma=4 eps=3.56 (HeDrop) */
    int32_t gr_bins, rho_bins, use_phi;
    int32_t pad;
    double lbox, rs, r_split2, r_tail, h_short, h_large, max_distance, mcm, gr_max, hbar2_2m;
    const int32_t* map_ptr; /* [P+1] rows over ext columns (K + 3) */
    const int32_t* map_col;
    const double* map_val;
    const double* map_const;  /* [P] */
    const double* grad_const; /* [P] the literal gradient constant of HeBulk.cpp:351 */
} oracle_he;

/* value sums of CalculateWavefunction (HeBulk.cpp:462-489, HeDrop.cpp:718-763): ext[K+3] */
void oracle_he_values(const oracle_he* s, const double* R, double* ext);
void oracle_he_operators(const oracle_he* s, const double* ext, double* O);
double oracle_he_exponent(const oracle_he* s, const double* ext, const double* uR);
/* CalculateExpectationValues (HeBulk.cpp:166-405, HeDrop.cpp:277-648).  other: [3 + gr_bins + rho_bins];
 * tabD [K+3][N][3], tabD2 [K+3][N]: derivative tables of every ext column (outputs / scratch). */
void oracle_he_expectation(const oracle_he* s, const double* R, double wf, const double* uR, const double* uI, double* e_r,
                           double* e_i, double* other, double* drift_r, double* drift_i, double* tabD, double* tabD2);
/* CalculateWFChange / Quotient (HeBulk.cpp:512-604, HeDrop.cpp:796-938) incl. the max(0, .) clamps */
double oracle_he_quotient(const oracle_he* s, const double* R, int particle, const double* old_pos, const double* ext,
                          double exponent, const double* uR, double* ext_new, double* exponent_new);
int64_t oracle_he_sweep(const oracle_he* s, double* R, double* ext, double* exponent, const double* uR, uint64_t seed,
                        uint32_t walker, uint64_t first_step, int64_t n_steps, double mc_step);
/* est: [O(P) | E_R | E_I | S(P*P) | OE_R(P) | OE_I(P) | other(n_other)] sums; rows: per sample [O(P), E_R, E_I] */
int64_t oracle_he_sample_walker(const oracle_he* s, double* R, const double* uR, const double* uI, double phiR, uint64_t seed,
                                uint32_t walker, uint64_t* step_counter, int n_init, int n_samples, int n_therm,
                                double mc_step, double* est, double* sample_rows);

#ifdef __cplusplus
}
#endif
#endif

/* ------------------------------------------------------------------------------------------------
 * BosonMixtureCluster (src/PhysicalSystems/BosonMixtureCluster.cpp): open-boundary cluster of several
 * species.  One basis set per pair type: McMillan core r^m below nodes[3], monomial-table cubic B-splines
 * (SplineFactory::GetWeights3) on a non-uniform knot vector, constant + linear tails beyond
 * nodes[#nodes-4], and a log term for every pair; per-species hbar^2/2m; pair potentials through
 * IPairPotential (HFDB_He_He, KTTY_He_Na, KTTY_He_Cs).
 * Extended sums per pair type t: ext[t*EXT + j], EXT = K + 4: [ss_0..ss_{K-1} | mcMillan | const | linear | log].
 * ------------------------------------------------------------------------------------------------ */
#ifndef TDVMC_ORACLE_MIX_H
#define TDVMC_ORACLE_MIX_H
#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_POT_HFDB_HE_HE = 0, ORACLE_POT_KTTY_HE_NA = 1, ORACLE_POT_KTTY_HE_CS = 2 };

typedef struct oracle_mix
{
    int32_t n_particles, n_params, n_types, n_splines; /* n_splines per pair type (26) */
    int32_t n_other, order;     /* spline order: 3 (0 also means 3) or 4 (BosonMixtureCluster_4thorder) */
    const int32_t* pair_type;   /* [N][N] correlationTypes */
    const double* hbar;         /* [N] hbarOver2m of each particle's species (BosonMixtureCluster.cpp:113-134) */
    const double* mass;         /* [N] */
    const double* knots;        /* [T][K+order+1] */
    const double* weights;      /* [T][K][order+1][order+1] */
    const double* mcm;          /* [T] mcMillanFactor */
    const int32_t* potential;   /* [T] ORACLE_POT_* */
    const int32_t* map_ptr;     /* [P+1] rows over the T*(K+4) extended sums (BosonMixtureCluster.cpp:636-645) */
    const int32_t* map_col;
    const double* map_val;
} oracle_mix;

double oracle_pair_potential(int id, double r); /* HFDB.cpp:23-47, KTTY.cpp:42-87 */
void oracle_mix_values(const oracle_mix* s, const double* R, double* ext);
void oracle_mix_operators(const oracle_mix* s, const double* ext, double* O);
double oracle_mix_exponent(const oracle_mix* s, const double* ext, const double* uR);
/* CalculateExpectationValues (BosonMixtureCluster.cpp:375-673).  other: [n_other] (first six set);
 * tabD [T*EXT][N][3], tabD2 [T*EXT][N] outputs/scratch. */
void oracle_mix_expectation(const oracle_mix* s, const double* R, double wf, double exponent, const double* uR, const double* uI,
                            double* e_r, double* e_i, double* other, double* drift_r, double* drift_i, double* tabD, double* tabD2);
double oracle_mix_quotient(const oracle_mix* s, const double* R, int particle, const double* old_pos, const double* ext,
                           double exponent, const double* uR, double* ext_new, double* exponent_new);
int64_t oracle_mix_sweep(const oracle_mix* s, double* R, double* ext, double* exponent, const double* uR, uint64_t seed,
                         uint32_t walker, uint64_t first_step, int64_t n_steps, double mc_step);
int64_t oracle_mix_sample_walker(const oracle_mix* s, double* R, const double* uR, const double* uI, double phiR, uint64_t seed,
                                 uint32_t walker, uint64_t* step_counter, int n_init, int n_samples, int n_therm,
                                 double mc_step, double* est, double* sample_rows);
/* mass-weighted centre of mass (BosonMixtureCluster.cpp:348-368) */
void oracle_mix_center_of_mass(const oracle_mix* s, const double* R, double* com);

/* BosonMixtureCluster::CalculateAdditionalSystemProperties (BosonMixtureCluster.cpp:680-741) for one three-particle
 * configuration.  grid = {count, spacing, max} of each histogram (Grid.cpp:16-31).
 *   r2                     mean squared distance from the (mass-weighted) centre of mass
 *   angle[3][n_angle]      GetCornerAngle (Utils.cpp:384-396) of 1-2-3, 1-3-2, 2-1-3 in degrees, one count each
 *   density[3][n_density]  |R_i - com| < max: 1 / scaling[bin] (ObservableVsOnGridWithScaling.cpp:47-52)
 *   distance[3][n_dist]    pairs (1,0), (2,0), (2,1) with r < max: one count each
 * Bins beyond a grid's count are dropped (the reference writes them past the end of its vector). */
void oracle_mix_observables(const oracle_mix* s, const double* R, const double* angle_grid, const double* density_grid,
                            const double* density_scaling, const double* distance_grid, double* r2, double* angle,
                            double* density, double* distance);

#ifdef __cplusplus
}
#endif
#endif
