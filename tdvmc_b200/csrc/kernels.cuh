// Kernel argument blocks and host-side launchers of the walker-ensemble library.
#pragma once

#include "common.cuh"

namespace tdvmc
{

#ifndef TDVMC_SWEEP_SMALL_THREADS
#define TDVMC_SWEEP_SMALL_THREADS 704
#endif
constexpr int kSweepMaxThreads = TDVMC_SWEEP_SMALL_THREADS;      // packed small systems (8 / 16 lanes per walker): up to 22 warps, 93 registers
#ifndef TDVMC_SWEEP_THREADS
#define TDVMC_SWEEP_THREADS 640
#endif
constexpr int kSweepMaxThreadsWarp = TDVMC_SWEEP_THREADS;  // one warp per walker: up to 20 warps per block, which lets the kernel take 94 registers
constexpr int kSweepMinBlocks = 1;

// ---- K1: Metropolis sweep (sweep.cu) ----
struct SweepArgs
{
    SysDev s;
    double* pos;                  // [W][3][Np]
    unsigned long long* accepted; // [W]
    int W;                        // local walkers
    int first_walker;             // global id of local walker 0
    int wpb;                      // warps per block (one walker per warp; 2 or 4 for systems of <= 16 / <= 8 particles)
    int npp;                      // padded row length of the shared-memory position arrays
    int pos_offset;               // byte offset of the position arrays in dynamic shared memory
    uint64_t seed;
    uint64_t first_step;          // per-walker step counter at launch (same for every walker)
    long long n_steps;
    double mc_step;
};
cudaError_t launch_sweep(SweepArgs a, cudaStream_t st);
// several warps per walker for small ensembles / large systems (sweep_split_kernel); warps from sweep_split_warps (> 1)
int sweep_split_warps(const SysDev& s, int W, int sm_count, int resident_per_sm);
cudaError_t launch_sweep_split(SweepArgs a, int warps, int sm_count, int smem_optin, cudaStream_t st, int copies = 8);
bool sweep_large_fits(const SysDev& s, int npp, int smem_optin); // one walker per block, 8 warps, 4 table replicas
// ensembles that are not a whole number of waves: walkers time-share the resident warps (sweep_queue_kernel)
bool sweep_queue_wanted(const SysDev& s, int W, int sm_count, int resident_per_sm, long long n_steps);
cudaError_t launch_sweep_queue(SweepArgs a, int sm_count, int smem_optin, cudaStream_t st);
int sweep_blocks_per_sm(const SysDev& s, int wpb, int npp);
int sweep_walkers_per_warp(const SysDev& s);
int sweep_max_threads(const SysDev& s);

// ---- K2+K3+K4: fused evaluation of one configuration per block (evaluate.cu) ----
struct EvalArgs
{
    SysDev s;
    const double* pos;      // [n_cfg][3][Np]
    int n_cfg;
    // sample row of configuration c: row0 + c * row_stride
    double* A;              // [rows][lda]: O_0..O_{P-1}, E^R, E^I, 1, 0...
    int lda;
    long long row0, row_stride;
    double* other;          // [rows][n_other]
    double* exponent;       // [rows] or null
    double* drift_r;        // [n_cfg][N][3] or null
    double* drift_i;
    double* ss_out;         // [n_cfg][K] or null
    double* outer_out;      // [n_cfg] or null
    double* scratch;        // large systems only: [blocks][9][NT*32] positions + forces of the configuration in flight
    int ucopies;            // set by launch_evaluate: shared-memory replicas of the u~ table (8 where they fit)
};
cudaError_t launch_evaluate(const EvalArgs& a, cudaStream_t st);
int evaluate_blocks_per_sm(const SysDev& s);   // 2 (12-warp blocks) or 1 (24-warp blocks, large N)
size_t evaluate_smem_bytes(const SysDev& s);   // dynamic shared memory of one evaluation block
size_t evaluate_scratch_doubles(const SysDev& s, int sm_count); // 0 unless the system needs the global-memory variant
cudaError_t launch_evaluate_he(const EvalArgs& a, cudaStream_t st);  // HeBulk, HeDrop (evaluate_he.cu)

// single-particle move ratios for scripted moves of one configuration (quotient_fixed)
struct QuotientArgs
{
    SysDev s;
    const double* pos;   // [3][Np]
    const double* moves; // [n][4]
    int n_moves;
    double* delta;       // [n] exponentNew - exponent
};
cudaError_t launch_quotient(const QuotientArgs& a, cudaStream_t st);

// ---- BosonMixtureCluster: thread-per-walker kernels (mixture.cu) ----
cudaError_t launch_evaluate_mix(const EvalArgs& a, cudaStream_t st); // BosonMixtureCluster (mixture.cu)
cudaError_t launch_sweep_mix(const SweepArgs& a, cudaStream_t st);
cudaError_t launch_quotient_mix(const QuotientArgs& a, cudaStream_t st);
cudaError_t launch_com_mix(const SysDev& s, double* pos, int W, cudaStream_t st);

// ---- NUBosonsBulkPBBoxAndRadial: radial + box spline bases (boxradial.cu) ----
cudaError_t launch_evaluate_br(const EvalArgs& a, cudaStream_t st);
cudaError_t launch_sweep_br(SweepArgs a, cudaStream_t st);
cudaError_t launch_quotient_br(const QuotientArgs& a, cudaStream_t st);
int sweep_br_fits(const SysDev& s, int wpb, int npp, int smem_optin);

// ---- InhContactBosons: one-dimensional, thread-per-walker kernels (inhcontact.cu) ----
cudaError_t launch_evaluate_inh(const EvalArgs& a, cudaStream_t st);
cudaError_t launch_sweep_inh(const SweepArgs& a, cudaStream_t st);
cudaError_t launch_quotient_inh(const QuotientArgs& a, cudaStream_t st);

// ---- K3 (table form) and K4 (contraction from tables): reference semantics (tables.cu) ----
struct TableArgs
{
    SysDev s;
    const double* pos;   // [n_cfg][3][Np]
    int n_cfg;
    double* T;           // [n_cfg][N][K][4] = {sD_x, sD_y, sD_z, sD2}
    double* v_int;       // [n_cfg]
};
cudaError_t launch_tables(const TableArgs& a, cudaStream_t st);

struct ContractArgs
{
    SysDev s;
    const double* T;     // [n_cfg][N][K][4]
    const double* v_int; // [n_cfg]
    int n_cfg;
    double* e_r;         // [n_cfg]
    double* e_i;
    double* sums;        // [n_cfg][5] = R1, I1, R2, I2, R1I1 (or null)
};
cudaError_t launch_contract(const ContractArgs& a, cudaStream_t st);

// ---- K5: S / F accumulation, FP64 tensor-core SYRK on the augmented sample matrix (accumulate.cu) ----
constexpr int kAccChunkRows = 32;  // samples per TMA stage
constexpr int kAccMaxTiles = 26;   // 8x8 output tiles per dimension -> P + 3 <= 208
struct AccArgs
{
    const double* A;      // [rows_padded][lda], rows beyond M are zero
    int lda;              // row stride in doubles, lda % 16 == 4
    int ncols;            // P + 3 used columns
    long long n_chunks;   // rows_padded / kAccChunkRows
    int n_cta;
    double* partial;      // [n_cta][ldc][ldc], ldc = 8 * ceil(ncols / 8)
    int ldc;
};
cudaError_t launch_accumulate(const AccArgs& a, cudaStream_t st);

struct AccFinishArgs
{
    const double* partial;
    int n_cta, ldc, P;
    const double* other;  // [M][n_other]
    long long M;
    int n_other;
    const unsigned long long* accepted; // [W]
    int W;
    double n_trials;      // proposals since the counters were cleared, this rank
    double* est;          // packed: S[P*P] | F_R[P] | F_I[P] | O[P] | E_R | E_I | other[n_other] | acc | trials | samples
};
cudaError_t launch_acc_finish(const AccFinishArgs& a, cudaStream_t st);

// ---- parameter linear solve on the device (solve.cu) ----
struct SolveArgs
{
    const double* est;     // packed estimator SUMS: S[P*P] | F_R[P] | F_I[P] | O[P] | E_R | E_I | other | acc | trials | samples
    int cnt_offset;        // index of `acc` in est
    int P;
    int imaginary_time;    // IMAGINARY_TIME: 0 real time, 1 imaginary time, -1 time rotation
    int use_preconditioning;
    double regularization; // 0.001 in the reference (src/TDVMC.cpp:1737)
    double min_scaling;    // 0: reference behaviour
    double* L_global;      // [P (P + 1) / 2] scratch for P too large for shared memory
    int L_in_smem;         // set by the launcher
    int force_global;      // tests: exercise the global-memory variant at small P
    double* out;           // uDotR[P] | uDotI[P] | phiDotR | phiDotI | not-positive-definite flag | <E^R> | <E^I>
};
cudaError_t launch_solve(SolveArgs a, int smem_optin, cudaStream_t st);
cudaError_t launch_solve_qr(SolveArgs a, cudaStream_t st); // LINEAR_EQUATION_SOLVER_TYPE = 1; a.L_global = [P * P] scratch

// ---- small utilities (util.cu) ----
cudaError_t launch_wrap(const SysDev& s, double* pos, int W, cudaStream_t st);
// ---- additional observables g(r), S(k) (observables.cu) ----
struct ObsArgs
{
    SysDev s;
    const double* pos;            // [n_cfg][3][Np]
    int n_cfg;
    int accumulate;               // 1: add into the configuration's rows, 0: overwrite
    int gr_count;
    double gr_spacing, gr_max;
    int n_shells, n_kvec;
    const int* shell_ptr;         // [n_shells + 1]
    const double* kvec;           // [n_kvec][3]
    unsigned long long* gr_rows;  // [n_cfg][gr_count] pair counts
    double* sk_rows;              // [n_cfg][n_shells]
};
cudaError_t launch_observables(const ObsArgs& a, cudaStream_t st);
cudaError_t launch_obs_reduce(const unsigned long long* gr_rows, const double* sk_rows, int n_rows, int gr_count, int n_shells,
                              double* out, cudaStream_t st);

// ---- three-particle cluster observables (observables.cu) ----
struct ClusterObsArgs
{
    SysDev s;
    const double* pos; // [n_cfg][3][Np]
    int n_cfg;
    int accumulate;    // r2_rows: add (1) or overwrite (0)
    int per_cfg_hist;  // 1: one configuration per block, hist[n_cfg][nh] overwritten; 0: hist[nh] accumulated over all
    int n_angle, n_density, n_distance;
    double angle_spacing, density_spacing, density_max, distance_spacing, distance_max;
    double* r2_rows;          // [n_cfg]
    unsigned long long* hist; // counts
};
cudaError_t launch_cluster_observables(const ClusterObsArgs& a, cudaStream_t st);
cudaError_t launch_sum_rows(const double* rows, int n, double* out, cudaStream_t st);

cudaError_t launch_min_image(const SysDev& s, double L, const double* a, const double* b, int n, double* norm, double* disp,
                             cudaStream_t st);
cudaError_t launch_proposals(uint64_t seed, uint32_t walker, uint64_t first_step, int n, int n_particles, double mc_step,
                             int* particle, double* disp, double* log_u, cudaStream_t st);
cudaError_t launch_transpose_in(const double* aos, double* soa, int n_cfg, int N, int Np, cudaStream_t st);  // [c][N][3] -> [c][3][Np]
cudaError_t launch_transpose_out(const double* soa, double* aos, int n_cfg, int N, int Np, cudaStream_t st);
cudaError_t launch_fill_rows(double* A, int lda, int P, const double* O, const double* e_r, const double* e_i, long long M,
                             cudaStream_t st);
cudaError_t measure_fp64(double* dfma_tflops, double* dmma_tflops, cudaStream_t st);

} // namespace tdvmc
