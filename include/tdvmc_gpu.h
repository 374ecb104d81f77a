/*
 * tdvmc_gpu.h — C ABI of the B200 walker-ensemble library (libtdvmc_b200.so).
 *
 * Drop-in boundary for the walker-parallel sampling + evaluation path of mathiasgartner/TDVMC.
 * The reference calls its IPhysicalSystem plugin once per Metropolis proposal
 * (src/TDVMC.cpp:876), which is far too fine-grained for a device; the seam therefore sits one
 * level up, at the three per-rank loops whose contract is "fill the seven estimator arrays and
 * the acceptance counters for parameters (uR, uI, phiR, phiI)":
 *
 *   UpdateExpectationValues                    src/TDVMC.cpp:1038-1150
 *   UpdateExpectationValuesForGivenSamples     src/TDVMC.cpp:1222-1303
 *   MPIMethods::ReduceToAverage x7             src/TDVMC.cpp:1182-1188, src/MPIMethods.h:132-206,329-362
 *
 * Conventions: plain pointers and sizes, caller-owned host buffers, row-major doubles.  Every
 * function returns 0 on success and a non-zero status otherwise; tdvmc_gpu_last_error() gives the
 * message.  One handle per process/GPU; not thread-safe.  All device memory lives behind the
 * opaque handle.  There is NO CPU fallback: without a CUDA device every entry point fails.
 *
 * Positions on the host are array-of-structs R[walker][particle][dim], exactly the
 * vector<vector<double>> of the reference flattened; on the device they are SoA per walker.
 */
#ifndef TDVMC_GPU_H
#define TDVMC_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDVMC_GPU_ABI_VERSION 2

typedef struct tdvmc_gpu_handle tdvmc_gpu_handle;

enum tdvmc_system_kind
{
    /* BosonsBulk, NUBosonsBulkPB: the caller's monomial spline table + boundary-condition map */
    TDVMC_SYSTEM_SPLINE_TABLE = 0,
    /* HeBulk (HeBulk.cpp): McMillan r^-5 core below rijSplit = 1.95, uniform cubic B-splines in the local coordinate
     * above it, Aziz HFD-B(He) inline, g(r) in other[3..102]; knots / spline_weights are not used (may be NULL) */
    TDVMC_SYSTEM_HE_BULK = 1,
    /* HeDrop (HeDrop.cpp): open boundary, McMillan r^-4.7 core below 3.0, 70 splines of spacing 0.1 then spacing 0.5 up
     * to rijTail, constant + linear tails beyond, Lennard-Jones inline, g(r) and the density profile in other[3..402];
     * lbox unused; wrap_positions moves the centre of mass to zero (src/TDVMC.cpp:798-809) */
    TDVMC_SYSTEM_HE_DROP = 2,
    /* BosonMixtureCluster (BosonMixtureCluster.cpp): open boundary, several species; one basis per pair type (see
     * tdvmc_mixture_desc), 26 parameters per pair type, other[0..5] = {kinR part 1, part 2, kinR, V, wf, exponent};
     * wrap_positions moves the mass-weighted centre of mass to zero */
    TDVMC_SYSTEM_MIXTURE = 3,
    /* NUBosonsBulkPBBoxAndRadial (NUBosonsBulkPBBoxAndRadial.cpp): periodic; a radial spline basis in r_ij inside
     * maxDistanceRad = knots[K] AND a "box" spline basis in |x_ij|, |y_ij|, |z_ij|, both on the caller's knots / spline
     * table (SetNodes mirrors the same grid into both, :36-62); n_splines = K = N_PARAM/2 + 3, n_ext = 2K with the sums
     * ordered [ssRad | ss], map = RefreshLocalOperators (:193-211); Gauss pair potential b exp(-(r/a)^2/2) from
     * system_params {a, b [, t, a2, b2]}; other[0..2] = {kinetic, potential, wf}, other[3..] = g(r) bins on
     * (0, lbox/2) weighted by 1/shell volume (n_other = 3 + GR_BIN_COUNT) */
    TDVMC_SYSTEM_BOX_RADIAL = 4,
    /* InhContactBosons (InhContactBosons.cpp): ONE-dimensional (dim = 1; positions still travel as [N][3], coordinate in
     * component 0), periodic; a single-particle spline function of the coordinate shifted into [0, lbox] and a
     * pair-correlation spline function of the minimum-image distance.  knots = [spf.nodes | pc.nodes], spline_weights =
     * [spf.splineWeights | pc.splineWeights], n_splines = K1 + K2 with K1 in n_splines_first, n_ext = K1 + K2, map =
     * RefreshLocalOperators (:208-247); system_params = the config's four {range, strength, k, V0} (square well or, with
     * range 0, contact strength gamma = strength k pi, :25-29; lattice potential, :448-467); n_other = 9; n_particles <= 32 */
    TDVMC_SYSTEM_INH_CONTACT = 5
};

/* Per-pair-type data of BosonMixtureCluster::InitSystem (BosonMixtureCluster.cpp:104-346), as data. */
typedef struct tdvmc_mixture_desc
{
    int32_t n_pair_types;          /* T = corrFuncData.size() */
    int32_t spline_order;          /* 3 (or 0): BosonMixtureCluster, cubic; 4: BosonMixtureCluster_4thorder (quartic splines,
                                    * SplineFactory::GetWeights4, rijSplit = nodes[4], rijTail = nodes[size - 5];
                                    * BosonMixtureCluster_4thorder.cpp:138-153).  For order 4 pass K = numberOfSplines = 28:
                                    * the 27 splines of the table plus a zero spline, and one padding knot after the 32 nodes */
    const int32_t* pair_type;      /* [N][N] correlationTypes (:58-102) */
    const double* hbar_over_2m;    /* [N] pp.hbarOver2m of each particle's species (:108-135) */
    const double* mass;            /* [N] pp.mass */
    const double* knots;           /* [T][K+order+1] cfd.nodes */
    const double* spline_weights;  /* [T][K][order+1][order+1] cfd.splineWeights (SplineFactory::GetWeights3 / GetWeights4) */
    const double* mcmillan_factor; /* [T] cfd.mcMillanFactor */
    const int32_t* potential;      /* [T] 0 HFDB_He_He, 1 KTTY_He_Na, 2 KTTY_He_Cs (:233-282, src/Potentials) */
} tdvmc_mixture_desc;

enum tdvmc_pair_rule
{
    TDVMC_PAIR_RULE_CUT = 0,    /* BosonsBulk.cpp:195-210: r <= r_max spline, else tail count */
    TDVMC_PAIR_RULE_REFLECT = 1 /* NUBosonsBulkPB.cpp:249-269: r -> 2 r_max - r beyond r_max */
};

/* What IPhysicalSystem::InitSystem() sets up, as data (BosonsBulk.cpp:49-156, NUBosonsBulkPB.cpp:53-216). */
typedef struct tdvmc_system_desc
{
    uint32_t struct_size;      /* sizeof(tdvmc_system_desc), for ABI evolution */
    int32_t n_particles;       /* N */
    int32_t dim;               /* DIM: 3; also 1 or 2 for TDVMC_SYSTEM_SPLINE_TABLE and TDVMC_SYSTEM_BOX_RADIAL (config/BosonsBulk2D,
                                * NUBosonsBulkPB2D, Rydberg2D, BosonsBulk1D, NUBosonsBulkPBBoxAndRadial2D ...) and 1 for
                                * TDVMC_SYSTEM_INH_CONTACT.  Positions always travel as [N][3]; the
                                * unused coordinates must be zero and stay zero */
    int32_t n_params;          /* N_PARAM */
    int32_t n_splines;         /* K = #knots - 4 (BosonsBulk.cpp:71) */
    int32_t pair_rule;         /* enum tdvmc_pair_rule */
    int32_t tail_param;        /* parameter that multiplies the tail count in the exponent (BosonsBulk.cpp:532-534), -1: none */
    int32_t n_other;           /* length of otherExpectationValues (>= 9) */
    double lbox;               /* LBOX */
    double hbar2_2m;           /* HBAR2_2M (src/Constants.h:12) */
    const double* knots;       /* [K+4], nodes (BosonsBulk.cpp:61-67 or SetNodes :35-44) */
    const double* spline_weights; /* [K][4][4] as returned by SplineFactory::GetWeights3 (SplineFactory.cpp:54-101) */
    const int32_t* map_ptr;    /* [N_PARAM+1] CSR rows of the boundary-condition map (BosonsBulk.cpp:158-177) */
    const int32_t* map_col;    /*   O_p = sum_j map_val[j] * splineSums[map_col[j]] */
    const double* map_val;
    const double* system_params; /* SYSTEM_PARAMS: a, b [, t_switch, a2, b2] (BosonsBulk.cpp:237-243) */
    int32_t n_system_params;
    int32_t system_kind;       /* enum tdvmc_system_kind */
    /* Columns of the map: the K spline sums plus analytic basis sums.  He family: n_ext = K + 3, columns K, K+1, K+2 are
     * the McMillan, constant and linear sums (HeBulk.cpp:376-383, HeDrop.cpp:609-626).  Spline-table systems: n_ext = K. */
    int32_t n_ext;
    int32_t n_splines_first;   /* TDVMC_SYSTEM_INH_CONTACT: K1, the splines of the single-particle function; else 0 */
    const double* map_const;   /* [N_PARAM] constant part of O_p (HeBulk.cpp:383: 1.0 + ...), NULL = zeros */
    const double* grad_const;  /* [N_PARAM] constant added to every gradient component of parameter p
                                  (the literal 1 of HeBulk.cpp:351), NULL = zeros */
    const tdvmc_mixture_desc* mixture; /* TDVMC_SYSTEM_MIXTURE only, else NULL; n_ext = T * (K + 4) */
} tdvmc_system_desc;

/* The walker ensemble owned by this rank.  The reference runs one walker per MPI rank, seeded
 * rank+1 (src/TDVMC.cpp:524); here a rank owns n_walkers chains with global ids
 * first_walker .. first_walker + n_walkers - 1, each with its own Philox4x32-10 stream
 * keyed by (seed, global id), so results do not depend on how walkers are split over GPUs. */
typedef struct tdvmc_ensemble_desc
{
    uint32_t struct_size;
    int32_t device;            /* CUDA device ordinal */
    int32_t n_walkers;         /* walkers resident on this GPU */
    int32_t first_walker;      /* global id of the first local walker */
    int32_t max_samples_per_walker; /* capacity of the sample store (MC_NSTEPS) */
    int32_t keep_sample_positions;  /* != 0: keep R of every sample for tdvmc_gpu_reevaluate_stored */
    uint64_t seed;
    double mc_step;            /* MC_STEP */
} tdvmc_ensemble_desc;

/* Averages as the reference's root rank holds them after ReduceToAverage (src/TDVMC.cpp:147-153). */
typedef struct tdvmc_estimators
{
    double* local_operators;            /* [P]    <O_k>      */
    double* local_energy_r;             /* [1]    <E^R>      */
    double* local_energy_i;             /* [1]    <E^I>      */
    double* local_operators_matrix;     /* [P*P]  <O_k O_j>  row-major */
    double* local_operator_energy_r;    /* [P]    <O_k E^R>  */
    double* local_operator_energy_i;    /* [P]    <O_k E^I>  */
    double* other_expectation_values;   /* [n_other]         */
    int64_t n_acceptances;              /* summed over all ranks (src/TDVMC.cpp:3727) */
    int64_t n_trials;
    int64_t n_samples;                  /* samples in the averages, all ranks */
} tdvmc_estimators;

int tdvmc_gpu_abi_version(void);
int tdvmc_gpu_device_count(void);

/* Supported envelope (tdvmc_gpu_create refuses anything outside it with a message, nothing is truncated):
 *   all systems          N_PARAM + 3 <= 208 (register-resident S matrix of the accumulation kernel), DIM as stated below
 *   SPLINE_TABLE         any N up to ~8300 (one walker's positions, 24 N bytes, must fit an SM's shared memory next to four
 *                        replicas of the sweep table: config/BosonsBulk3D.config as shipped, N = 8000, runs); up to N ~ 2000 a
 *                        configuration is evaluated out of shared memory, beyond that out of a per-block slab in global memory;
 *                        DIM 1, 2 or 3 (the large-system sweep: DIM = 3)
 *   HE_BULK / HE_DROP    DIM = 3
 *   MIXTURE              N <= 8 particles, n_ext <= 96, spline order 3 or 4, DIM = 3
 *   BOX_RADIAL           DIM 2 or 3, one walker's tables must fit shared memory
 *   INH_CONTACT          DIM = 1, N <= 32
 *   device solver        N_PARAM <= 1024; LINEAR_EQUATION_SOLVER_TYPE 0 and 1; IMAGINARY_TIME -1, 0, 1; all parameters in use
 *                        (USE_PARAM_START = USE_PARAM_END = 0)
 *   observables          g(r) / S(k) for DIM = 3 */

int tdvmc_gpu_create(const tdvmc_system_desc* system, const tdvmc_ensemble_desc* ensemble, tdvmc_gpu_handle** out);
void tdvmc_gpu_destroy(tdvmc_gpu_handle* h);
const char* tdvmc_gpu_last_error(const tdvmc_gpu_handle* h); /* h may be NULL: error of the last failed create */

/* ---- state ---- */
/* R[n_walkers][N][3] host <-> device SoA (the reference's R, src/TDVMC.cpp:3095). */
int tdvmc_gpu_set_positions(tdvmc_gpu_handle* h, const double* R, int32_t first_local_walker, int32_t n_walkers);
int tdvmc_gpu_get_positions(tdvmc_gpu_handle* h, double* R, int32_t first_local_walker, int32_t n_walkers);
/* BroadcastNewParameters + sys->SetTime (src/TDVMC.cpp:506-512, 3437). */
int tdvmc_gpu_set_params(tdvmc_gpu_handle* h, const double* uR, const double* uI, double phiR, double phiI, double time);
/* MoveCoordinatesToFirstCell (src/TDVMC.cpp:787-796). */
int tdvmc_gpu_wrap_positions(tdvmc_gpu_handle* h);

/* nAcceptances = 0; nTrials = 0 at the start of every time step (src/TDVMC.cpp:3428-3429, 3976-3977): the counters
 * tdvmc_gpu_allreduce_and_fetch returns count from the last reset (or from creation). */
int tdvmc_gpu_reset_counters(tdvmc_gpu_handle* h);
/* MC_STEP is a runtime-changeable config item of the reference (./param file, src/TDVMC.cpp:2718-2744). */
int tdvmc_gpu_set_mc_step(tdvmc_gpu_handle* h, double mc_step);

/* ---- sampling ---- */
/* DoMetropolisSteps for every local walker (src/TDVMC.cpp:858-924). */
int tdvmc_gpu_sweep(tdvmc_gpu_handle* h, int64_t n_steps);
/* UpdateExpectationValues (src/TDVMC.cpp:1038-1150): n_init steps, then n_samples x (n_therm steps +
 * evaluation of O_k, E_L, other) per walker, then S / F accumulation over all local samples.
 * Results stay on the device until tdvmc_gpu_allreduce_and_fetch. */
int tdvmc_gpu_sample_and_accumulate(tdvmc_gpu_handle* h, int32_t n_samples, int32_t n_therm, int32_t n_init);
/* UpdateExpectationValuesForGivenSamples (src/TDVMC.cpp:1222-1303): re-evaluate the stored samples at the
 * current parameters (recomputed from the stored R) and accumulate.  Needs keep_sample_positions. */
int tdvmc_gpu_reevaluate_stored(tdvmc_gpu_handle* h);
/* UpdateSamplesConsecutive (src/TDVMC.cpp:975-983) with UpdateSample (:948-961): the next n_update stored samples
 * of every walker (ring order, cursor = currentSampleIndexForUpdate) each advance by n_therm Metropolis steps at the
 * current parameters.  Follow with tdvmc_gpu_reevaluate_stored (the driver does, :3675-3677).  Every slot draws its
 * own stretch of the walker's proposal stream. */
int tdvmc_gpu_update_stored(tdvmc_gpu_handle* h, int32_t n_update, int32_t n_therm);
/* ReduceToAverage x7 + nAcceptances (src/TDVMC.cpp:1182-1188, 3727): one packed all-reduce over the
 * communicator (if any), division by the global sample count; every rank receives the averages. */
int tdvmc_gpu_allreduce_and_fetch(tdvmc_gpu_handle* h, tdvmc_estimators* out);
/* sys->GetExponent() of the first local walker's last sample, for NormalizeWavefunction (src/TDVMC.cpp:3763). */
int tdvmc_gpu_last_exponent(tdvmc_gpu_handle* h, double* exponent);

/* ---- parameter derivatives and the Euler step on the device ---- */
/* Options of SolveForParametersDot (src/TDVMC.cpp:1713-1828), both branches, all three values of IMAGINARY_TIME:
 * solver_type 0 = the hand-written Cholesky (:1733-1763; bit-identical to the host code), solver_type 1 = Eigen's
 * FullPivHouseholderQR followed by the mean subtraction of :1800-1809 (:1763-1827; Eigen's algorithm step by step, results
 * equal to rounding x condition number).  With use_preconditioning the reference regularises the scaled matrix by 0.001
 * (type 0) or 0.002 (type 1, :1770) - pass that value in `regularization`; type 1 without preconditioning is not regularised. */
typedef struct tdvmc_solver_desc
{
    uint32_t struct_size;
    int32_t imaginary_time;      /* IMAGINARY_TIME: 0 real time, 1 imaginary time, -1 the 1.499 pi time rotation (:1475-1504, 1666-1673) */
    int32_t use_preconditioning; /* USE_PRECONDITIONING: scale by sqrt(diag), src/TDVMC.cpp:1684-1701 */
    int32_t force_global_scratch;/* tests: factorise in global memory even when P fits shared memory */
    double regularization;       /* RegularizeEquationSystem; the reference hard-codes 0.001 (src/TDVMC.cpp:1737) */
    double min_scaling;          /* 0 = reference behaviour; > 0 floors the scalings (a parameter whose operator never varied) */
    int32_t solver_type;         /* LINEAR_EQUATION_SOLVER_TYPE: 0 Cholesky, 1 full-pivoting Householder QR */
    int32_t reserved;            /* 0 */
} tdvmc_solver_desc;
/* What the root rank of the reference holds after SolveForParametersDot (+ the energies the driver logs each step). */
typedef struct tdvmc_parameters_dot
{
    double* u_dot_r;             /* [P] */
    double* u_dot_i;             /* [P] */
    double phi_dot_r, phi_dot_i;
    double local_energy_r, local_energy_i; /* <E^R>, <E^I> of the estimators the system was built from */
    int32_t not_positive_definite; /* the reference logs "NOT POSITIVE SEMI DEFINITE" and sets doNotAcceptStep (:1577-1589) */
} tdvmc_parameters_dot;
/* SolveForParametersDot (Cholesky branch: BuildSystemOfEquationsForParametersIncludePhi src/TDVMC.cpp:1506-1537,
 * PreconditionEquationSystemByScaling :1684-1701, RegularizeEquationSystem :1703-1711, PerformCholeskyDecomposition
 * :1560-1592, SolveCholeskyDecomposedEquationSystem :1594-1622, CalculatePhiDot :1658-1682) on the estimators of the
 * last accumulation, all-reduced in place first if a communicator is bound; nothing but 2P + 5 doubles leaves the
 * device.  Every rank solves redundantly and receives the same result (the reference solves on the root and
 * broadcasts, :506-512). */
int tdvmc_gpu_solve_parameters_dot(tdvmc_gpu_handle* h, const tdvmc_solver_desc* sd, tdvmc_parameters_dot* out);
/* CalculateNextParametersEuler (src/TDVMC.cpp:1834-1853) + BroadcastNewParameters (:506-512): solve as above,
 * uR += uDotR dt, uI += uDotI dt, phi += phiDot dt on the caller's arrays (in place), and make the new parameters
 * current on the device as tdvmc_gpu_set_params(uR, uI, phiR, phiI, time) would.  If the matrix was not positive
 * definite the parameters are still updated, as in the reference; the flag is in dot->not_positive_definite.
 * dot may be NULL. */
int tdvmc_gpu_euler_step(tdvmc_gpu_handle* h, const tdvmc_solver_desc* sd, double dt, double time, double* uR, double* uI,
                         double* phiR, double* phiI, tdvmc_parameters_dot* dot);
/* The same solve on estimators given by the caller (averages, as tdvmc_estimators holds them): parity entry point
 * against the reference's own SolveForParametersDot. */
int tdvmc_gpu_solve_fixed(tdvmc_gpu_handle* h, const tdvmc_solver_desc* sd, const tdvmc_estimators* est,
                          tdvmc_parameters_dot* out);

/* ---- communicator (replaces MPI_COMM_WORLD for the reduce; src/MPIMethods.h) ---- */
#define TDVMC_GPU_UNIQUE_ID_BYTES 128
int tdvmc_gpu_comm_unique_id(uint8_t id[TDVMC_GPU_UNIQUE_ID_BYTES]);          /* rank 0, then broadcast by the host */
int tdvmc_gpu_comm_init(tdvmc_gpu_handle* h, const uint8_t id[TDVMC_GPU_UNIQUE_ID_BYTES], int32_t rank, int32_t n_ranks);

/* ---- fixed-configuration entry points (parity tests; no random numbers) ---- */
/* CalculateWavefunction + CalculateExpectationValues on n_cfg given configurations R[n_cfg][N][3]
 * (BosonsBulk.cpp:460-466, 541-545).  Any output pointer may be NULL.
 * e_r/e_i/exponent: [n_cfg]; O: [n_cfg][P]; other: [n_cfg][n_other]; drift_r/drift_i: [n_cfg][N][3]
 * (vecKineticSumR1/I1 per particle, BosonsBulk.cpp:354-420); spline_sums: [n_cfg][n_ext]; outer: [n_cfg]. */
int tdvmc_gpu_evaluate_fixed(tdvmc_gpu_handle* h, const double* R, int32_t n_cfg, double* e_r, double* e_i,
                             double* O, double* other, double* exponent, double* drift_r, double* drift_i,
                             double* spline_sums, double* outer);
/* CalculateWFQuotient (BosonsBulk.cpp:649-657) for n_moves scripted single-particle moves of ONE configuration
 * R[N][3]: moves[n_moves][4] = {particle, x, y, z}.  quotient/delta: [n_moves]; delta = exponentNew - exponent. */
int tdvmc_gpu_quotient_fixed(tdvmc_gpu_handle* h, const double* R, const double* moves, int32_t n_moves,
                             double* quotient, double* delta);
/* CalculateOtherLocalOperators tables (BosonsBulk.cpp:220-336) in the reference's layout:
 * sD[K][N][3], sD2[K][N] for ONE configuration. */
int tdvmc_gpu_tables_fixed(tdvmc_gpu_handle* h, const double* R, double* sD, double* sD2);
/* Minimum-image displacement a - b and norm for n vector pairs (Utils.cpp:266-281, 368-374). */
int tdvmc_gpu_min_image(tdvmc_gpu_handle* h, double lbox, const double* a, const double* b, int32_t n, double* norm,
                        double* disp);
/* S/F accumulation alone (src/TDVMC.cpp:1103-1109) on caller-provided samples: O[M][P], e_r[M], e_i[M] ->
 * sums (not averages) S[P*P], f_r[P], f_i[P], o[P]. */
int tdvmc_gpu_accumulate_fixed(tdvmc_gpu_handle* h, const double* O, const double* e_r, const double* e_i, int64_t M,
                               double* S, double* f_r, double* f_i, double* o);
/* The proposal stream: particle, displacement[3], log(u) for (walker, step) -- lets tests replay a chain. */
int tdvmc_gpu_proposals(tdvmc_gpu_handle* h, int32_t global_walker, int64_t first_step, int32_t n, int32_t* particle,
                        double* disp, double* log_u);

/* ---- additional observables: g(r) and S(k) of the bulk spline systems ----
 * Replaces IPhysicalSystem::CalculateAdditionalSystemProperties (BosonsBulk.cpp:474-520,
 * NUBosonsBulkPB.cpp:597-639) and the driver loop around it (src/TDVMC.cpp:1332-1388, 1438-1444).
 * The description is data the reference's InitSystem() builds (BosonsBulk.cpp:124-153). */
typedef struct tdvmc_observable_desc
{
    int32_t gr_count;          /* pairDistribution.grid.count (Grid.cpp:21; 0: no g(r)) */
    int32_t n_shells;          /* numOfkValues (BosonsBulk.cpp:55; 0: no S(k)) */
    double gr_spacing;         /* pairDistribution.grid.spacing */
    double gr_max;             /* pairDistribution.grid.max */
    double gr_weight;          /* per pair: DIM/(N-1) (BosonsBulk.cpp:481) or 1 (NUBosonsBulkPB.cpp:611) */
    const double* gr_scaling;  /* [gr_count] scalingGrid (ObservableVsOnGridWithScaling.cpp:19-44) */
    const int32_t* shell_ptr;  /* [n_shells+1] first wave vector of every shell */
    const double* kvec;        /* [shell_ptr[n_shells]][3] kValues, already times 2 pi / LBOX (BosonsBulk.cpp:127-137) */
} tdvmc_observable_desc;
/* sys->CalculateAdditionalSystemProperties(R, ...) for n_cfg caller-given configurations R[n_cfg][N][3]:
 * gr[n_cfg][gr_count] (pairDistribution values), sk[n_cfg][n_shells] (structureFactor values). */
int tdvmc_gpu_observables_fixed(tdvmc_gpu_handle* h, const tdvmc_observable_desc* od, const double* R, int32_t n_cfg,
                                double* gr, double* sk);
/* ParallelCalculateAdditionalSystemProperties (src/TDVMC.cpp:1438-1444): n_init steps
 * (MC_NADDITIONALINITIALIZATIONSTEPS), then n_samples x (n_therm steps + observables) per resident walker;
 * gr[gr_count], sk[n_shells] = mean over samples, walkers and ranks (additionalObservablesMean after
 * MPIMethods::ReduceToAverage); identical on every rank. */
int tdvmc_gpu_sample_observables(tdvmc_gpu_handle* h, const tdvmc_observable_desc* od, int32_t n_samples, int32_t n_therm,
                                 int32_t n_init, double* gr, double* sk);

/* ---- additional observables of the three-particle mixture cluster (config/He4He4Na.config's whole workload) ----
 * Replaces BosonMixtureCluster::CalculateAdditionalSystemProperties (BosonMixtureCluster.cpp:680-741) and the driver
 * loop around it (src/TDVMC.cpp:1332-1388, 1438-1444).  Grids as InitSystem builds them (BosonMixtureCluster.cpp:327-340);
 * the reference hard-codes three particles here, so the handle must hold a three-particle mixture. */
typedef struct tdvmc_cluster_observable_desc
{
    int32_t n_angle, n_density, n_distance;   /* grid.count of angularDistribution, densityFromCOM, particleDistances */
    int32_t reserved;
    double angle_spacing;                     /* 1.0 degree */
    double density_spacing, density_max;      /* 0.1, 80 */
    double distance_spacing, distance_max;    /* 0.5, 80 */
    const double* density_scaling;            /* [n_density] densityFromCOM.scalingGrid */
} tdvmc_cluster_observable_desc;
/* For n_cfg given configurations R[n_cfg][3][3]: r2[n_cfg], angle[n_cfg][3][n_angle] (1-2-3, 1-3-2, 2-1-3),
 * density[n_cfg][3][n_density], distance[n_cfg][3][n_distance] (pairs 1-2, 1-3, 2-3). */
int tdvmc_gpu_cluster_observables_fixed(tdvmc_gpu_handle* h, const tdvmc_cluster_observable_desc* od, const double* R,
                                        int32_t n_cfg, double* r2, double* angle, double* density, double* distance);
/* ParallelCalculateAdditionalSystemProperties for the cluster: n_init steps, then n_samples x (n_therm steps +
 * observables) per walker; outputs are the means over samples, walkers and ranks (additionalObservablesMean). */
int tdvmc_gpu_sample_cluster_observables(tdvmc_gpu_handle* h, const tdvmc_cluster_observable_desc* od, int32_t n_samples,
                                         int32_t n_therm, int32_t n_init, double* r2, double* angle, double* density,
                                         double* distance);

/* ---- measurement hooks ---- */
enum tdvmc_kernel_id
{
    TDVMC_KERNEL_SWEEP = 0,
    TDVMC_KERNEL_EVALUATE = 1,
    TDVMC_KERNEL_ACCUMULATE = 2,
    TDVMC_KERNEL_TABLES = 3,
    TDVMC_KERNEL_CONTRACT = 4,
    TDVMC_KERNEL_OTHER = 5,
    TDVMC_KERNEL_SOLVE = 6,
    TDVMC_KERNEL_COUNT = 7
};
/* Per-kernel CUDA-event timing on the library's stream. enable=1 starts, stats are cumulative since the last reset. */
int tdvmc_gpu_profile(tdvmc_gpu_handle* h, int32_t enable, int32_t reset);
int tdvmc_gpu_kernel_stats(tdvmc_gpu_handle* h, int32_t kernel_id, int64_t* launches, double* total_ms);
int tdvmc_gpu_synchronize(tdvmc_gpu_handle* h);
/* CUDA-event stopwatch on the library's stream (the stream every kernel and copy of this handle is issued on). */
int tdvmc_gpu_timer_start(tdvmc_gpu_handle* h);
int tdvmc_gpu_timer_stop(tdvmc_gpu_handle* h, double* elapsed_ms); /* records, synchronises, returns the elapsed time */
/* Number of kernels of this library launched on the handle since creation. */
int tdvmc_gpu_launch_count(tdvmc_gpu_handle* h, int64_t* n);
/* Evict the L2 cache: overwrite a scratch buffer of n_bytes (>= L2 size) on the library's stream. */
int tdvmc_gpu_flush_l2(tdvmc_gpu_handle* h, int64_t n_bytes);
/* K3/K4 exhibits on the resident walkers (reference table semantics, BosonsBulk.cpp:220-336 and :349-458):
 * materialise the sD/sD2 tables of every local walker and contract them again.  Used by bench.py. */
int tdvmc_gpu_tables_resident(tdvmc_gpu_handle* h, int32_t n_walkers);
int tdvmc_gpu_contract_resident(tdvmc_gpu_handle* h, int32_t n_walkers, double* e_r, double* e_i);
/* Walkers the sweep kernel keeps resident per SM (one warp each) and the SM count: an ensemble that is a
 * multiple of per_sm * sm_count fills the machine without a partial last wave. */
int tdvmc_gpu_resident_walkers(tdvmc_gpu_handle* h, int32_t* per_sm, int32_t* sm_count);
/* Device microbenchmarks for the roofline denominators the driver does not provide (FP64). */
int tdvmc_gpu_measure_fp64_peak(tdvmc_gpu_handle* h, double* dfma_tflops, double* dmma_tflops);

#ifdef __cplusplus
}
#endif
#endif
