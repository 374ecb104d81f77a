// Shared device-side definitions of the walker-ensemble library (sm_100a).
//
// The whole library is compiled with -fmad=false: every a*b+c written with operators is a
// separate IEEE multiply and add, in the association order of the expression, exactly like the
// reference compiled for baseline x86-64.  Fused operations appear only where fma() is spelled out.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace tdvmc
{

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int kRecStride = 18;   // doubles per knot-interval record of the evaluation table (bank spreading)

// Read-only description of the system as the kernels see it.  Pointers are device memory.
struct SysDev
{
    int N;            // particles
    int Np;           // padded particle count (SoA row length in HBM)
    int P;            // parameters
    int K;            // basis splines
    int pair_rule;    // TDVMC_PAIR_RULE_*
    int tail_param;
    int n_other;
    int first_bin;    // first knot interval a distance can fall in (knots[first_bin] == 0)
    int nbins;        // K - first_bin
    int ncell;        // cells of the uniform bin lookup grid
    int uniform;      // != 0: knots are a uniform grid, interval index = floor(r / h)
    int kind;         // TDVMC_SYSTEM_*: 0 spline table (BosonsBulk, NUBosonsBulkPB), 1 HeBulk, 2 HeDrop, 3 mixture, 4 box + radial
    int periodic;     // 1: minimum image in the box, 0: open boundary (HeDrop)
    int n_short;      // He family: splines on the short grid (HeDrop: 70; HeBulk: all)
    int potential;    // He family: 0 Aziz HFD-B(He) (HeBulk.cpp:187-195), 1 Lennard-Jones sigma=4 eps=3.56 (HeDrop.cpp:389-394)
    int rho_bins;     // density-profile bins carried in other[] after g(r) (HeDrop: 200)
    int use_phi;      // wf = exp(exponent + phiR) (HeDrop.cpp:783) or exp(exponent) (HeBulk.cpp:500)
    int n_ext;        // columns of the parameter map: K spline sums + analytic extras
    int gr_bins;      // g(r) bins carried in other[] (HeBulk: 100)
    double L, Linv, Lhalf;   // LBOX, 1/LBOX, LBOX/2 (src/TDVMC.cpp:535-536)
    double rmax;             // maxDistance = knots[K]
    double hbar;             // HBAR2_2M
    double pot_a, pot_b;     // square well, time switch applied
    double phiR;
    double inv_cell;         // ncell / rmax
    double h, inv_h;         // uniform knot spacing
    double u_tail;           // uR[tail_param]
    double r0;               // start of the uniform spline grid (HeBulk: rijSplit; 0 otherwise)
    double core_m;           // McMillan exponent (HeBulk: -5)
    double u_core;           // u~ of the McMillan column
    double g0R, g0I;         // sum_p u_p grad_const[p]: the literal gradient constant (HeBulk.cpp:351)
    double r_split2;         // He family: start of the long grid (HeDrop: rijSplineSplit), else huge
    double r_tail;           // He family: const + linear tails from here on (HeDrop: rijTail), else huge
    double h_large;          // spacing of the long grid
    double gr_max;           // g(r) / density-profile range
    double u_const, u_lin;   // u~ of the const and linear tail columns
    const double* knots;     // [K+4]
    const double* rec;       // [nbins][kRecStride]: piece p of spline (bin-p) at [p*4 + c]
    const double* cub;       // sweep table, 3 planes of (nbins+1) double2: [c0,c1] | [c2,c3] | [t_lo,t_hi];
                             // u(r) = c0 + c1 s + c2 s^2 + c3 s^3 on the interval, s = r - t_lo
    const unsigned short* lut; // [ncell] -> interval index guess
    const int* map_ptr;
    const int* map_col;
    const double* map_val;
    const double* uR;        // [P]
    const double* uI;        // [P]
    const double* utR;       // [n_ext]  parameters in spline space, u~_k = sum_p u_p M[p][k]
    const double* utI;       // [n_ext]
    const double* map_const; // [P] constant part of O_p
    // BosonMixtureCluster (kind 3): per-pair-type bases, per-particle species data
    int n_types;               // pair types T
    int spline_order;          // 3 (BosonMixtureCluster) or 4 (BosonMixtureCluster_4thorder)
    const int* pair_type;      // [N][N] correlationTypes
    const double* hbar_n;      // [N] hbar^2/2m of each particle's species
    const double* mass_n;      // [N]
    const double* t_knots;     // [T][K+order+1]
    const double* t_weights;   // [T][K][order+1][order+1] monomial spline tables (SplineFactory::GetWeights3 / GetWeights4)
    const double* t_mcm;       // [T] McMillan exponents
    const int* t_pot;          // [T] pair potential ids
    const double* t_cub;       // [T][K-2*order+3][order+3] sweep polynomials {c0..c_order, t_lo, t_hi} per knot interval
    // NUBosonsBulkPBBoxAndRadial (kind 4): ext = [ssRad (K) | ss (K)] on one knot vector; cub holds the radial planes
    // followed by the box planes
    const double* ugR;         // [n_ext] u~ as the DRIFT uses it: the last radial spline's parameter multiplies the box
    const double* ugI;         //         table (NUBosonsBulkPBBoxAndRadial.cpp:493-497); the Laplacian uses utR / utI
    const double* gr_vol;      // [gr_bins] g(r) shell volumes (:149-169)
    double gr_spacing;         // grNodePointSpacing (:147)
    // InhContactBosons (kind 5, one-dimensional): n_short = K1 splines of the single-particle function, then K - K1 of the
    // pair correlation; knots [K1+4 | K2+4]; rec = the raw spline table [K][4][4]; cub = [K1-3 | K2-3] interval records
    double gamma;              // contact strength (InhContactBosons.cpp:25-29); pot_a / pot_b = square-well range / strength
    double ext_k, ext_v0;      // lattice potential k^2 V0 sin^2(k x), SYSTEM_PARAMS[2], [3] (:448-467)
    double exp_const;          // -2 gamma h_pc: coefficient of ss_pc[0] in the exponent (:764)
    // DIM of the spline-table / box+radial systems (1, 2 or 3; unused coordinates stay zero), else 3.  Kept at the END of the
    // struct: inserting fields further up shifts L / Linv / Lhalf off their 16-byte pairs in the constant bank and costs the
    // sweep 1.2 % (13.63 -> 13.80 ms per launch, measured on the same GPU)
    double dm1;                // DIM - 1: secondDerivativeFactor of BosonsBulk.cpp:319, NUBosonsBulkPB.cpp:392
    int dim;
    // uniform knots (BosonsBulk.cpp:61-67): a distance whose r / h lies further than bin_guard from an integer is in
    // interval first_bin + floor(r / h) whatever the rounding of the stored knots (bin_guard = twice their largest
    // deviation from the exact grid, in units of h, + 1e-10); 0: not available, the knots are always consulted
    double bin_guard;
};

// ---- minimum image -------------------------------------------------------------------------
// GetCoordinateNIC (src/Utils.cpp:266-281): round half away from zero through an int cast, and a
// 1e-10 nudge when the result lands exactly on +-L/2.  Same expressions, same order, no FMA.
__device__ __forceinline__ double nic_exact(double r, double L, double Linv, double Lhalf)
{
    int k = (int)(r * Linv + ((r >= 0.0) ? 0.5 : -0.5));
    double result = r - k * L;
    if (result == Lhalf) result -= 1e-10;
    else if (result == -Lhalf) result += 1e-10;
    return result;
}

// VectorDisplacementNIC_3D (src/Utils.cpp:329-338, 368-374): a - b per coordinate, then the norm
// sqrt(x*x + y*y + z*z) in that association (src/Utils.cpp:174-179, 207-212).
__device__ __forceinline__ double disp_exact(const SysDev& s, double ax, double ay, double az, double bx, double by,
                                             double bz, double& vx, double& vy, double& vz)
{
    vx = nic_exact(ax - bx, s.L, s.Linv, s.Lhalf);
    vy = nic_exact(ay - by, s.L, s.Linv, s.Lhalf);
    vz = nic_exact(az - bz, s.L, s.Linv, s.Lhalf);
    return sqrt(vx * vx + vy * vy + vz * vz);
}

// std::lower_bound(nodes, r) - 1  <=>  knots[bin] < r <= knots[bin+1]  (BosonsBulk.cpp:197-198).
// A uniform lookup grid gives the starting guess, two short loops make it exact.
__device__ __forceinline__ int find_bin_exact(const SysDev& s, const double* knots, const unsigned short* lut, double r)
{
    int c = (int)(r * s.inv_cell);
    c = max(0, min(c, s.ncell - 1));
    int b = lut[c];
    while (b < s.K - 1 && r > knots[b + 1]) b++;
    while (b > s.first_bin && !(knots[b] < r)) b--;
    return b;
}

// Uniform knots: the interval from r / h alone when r is not within bin_guard (in units of h) of a knot - no table
// look-up, no knot loads; the rare distance next to a knot takes the exact search.  Same result as find_bin_exact.
__device__ __forceinline__ int find_bin_uniform(const SysDev& s, const double* knots, const unsigned short* lut, double r)
{
    const double magic = 6755399441055744.0; // 1.5 * 2^52: x + magic carries round-to-nearest(x) in its low word
    const double x = r * s.inv_h;
    const double y = x + magic;
    const int j = __double2loint(y);
    const double d = x - (y - magic);        // x - nearest integer, exact
    if (fabs(d) > s.bin_guard) return s.first_bin + j - (d < 0.0 ? 1 : 0);
    return find_bin_exact(s, knots, lut, r);
}

// ---- warp / block reductions ------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ int warp_sum_int(int v)
{
    return __reduce_add_sync(FULL_MASK, v);
}

// ---- basis-sum histogram (evaluate.cu, evaluate_he.cu) ---------------------------------------------
// Conflict-free warp-cooperative  hist[bin - p] += v[p], p = 0..3  for the lanes with active == true.
// Must be called by all 32 lanes.  Lanes sharing a bin are serialised by their rank within the group;
// the four pieces are separated by __syncwarp() because bin-p of one lane aliases bin'-p' of another.
__device__ __forceinline__ void warp_hist_add4(double* hist, int bin, bool active, const double (&v)[4], int lane)
{
    const unsigned amask = __ballot_sync(FULL_MASK, active);
    if (amask == 0u) return;
    const int key = active ? bin : (-1 - lane);
    const unsigned peers = __match_any_sync(FULL_MASK, key);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    for (int round = 0;; round++)
    {
        const bool mine = active && (rank == round);
        if (__ballot_sync(FULL_MASK, mine) == 0u) break;
#pragma unroll
        for (int p = 0; p < 4; p++)
        {
            if (mine) hist[bin - p] += v[p];
            __syncwarp();
        }
    }
}

// The same with ballot and match.any issued by the caller as soon as the interval is known, so that the long latency of
// MATCH.ANY passes behind the spline arithmetic instead of in front of the first read-modify-write.
__device__ __forceinline__ void warp_hist_add4_ranked(double* hist, int bin, bool active, const double (&v)[4], int lane,
                                                      unsigned amask, unsigned peers)
{
    if (amask == 0u) return;
    const int rank = __popc(peers & ((1u << lane) - 1u));
    for (int round = 0;; round++)
    {
        const bool mine = active && (rank == round);
        if (__ballot_sync(FULL_MASK, mine) == 0u) break;
#pragma unroll
        for (int p = 0; p < 4; p++)
        {
            if (mine) hist[bin - p] += v[p];
            __syncwarp();
        }
    }
}

// ---- Philox4x32-10 proposal stream (same definition as oracle_proposal in oracle/tdvmc_oracle.c) ----
struct Philox4
{
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1)
{
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int round = 0; round < 10; round++)
    {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 r = { c0, c1, c2, c3 };
    return r;
}

__host__ __device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) // (0, 1]
{
    uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)x + 1.0) * (1.0 / 9007199254740992.0);
}

struct Proposal
{
    int particle;
    double dx, dy, dz;
    double log_u;
};

// Replaces randomParticleIndex() + DIM x randomNormal(MC_STEP) + random01() of DoMetropolisStep
// (src/TDVMC.cpp:870-875, 900) by a counter-based stream: a pure function of (seed, walker, step).
__device__ __forceinline__ Proposal make_proposal(uint64_t seed, uint32_t walker, uint64_t step, int n_particles,
                                                  double mc_step)
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t s0 = (uint32_t)step, s1 = (uint32_t)(step >> 32);
    Philox4 a = philox4x32_10(s0, s1, walker, 0u, k0, k1);
    Philox4 b = philox4x32_10(s0, s1, walker, 1u, k0, k1);
    Philox4 c = philox4x32_10(s0, s1, walker, 2u, k0, k1);
    Proposal p;
    p.particle = (int)(((uint64_t)a.x * (uint64_t)n_particles) >> 32);
    p.log_u = log(u53(a.y, a.z));
    const double two_pi = 6.283185307179586476925286766559;
    double rad0 = sqrt(-2.0 * log(u53(a.w, b.x)));
    double ang0 = two_pi * (u53(b.y, b.z) - 1.0 / 9007199254740992.0);
    double rad1 = sqrt(-2.0 * log(u53(b.w, c.x)));
    double ang1 = two_pi * (u53(c.y, c.z) - 1.0 / 9007199254740992.0);
    double s, co;
    sincos(ang0, &s, &co);
    p.dx = rad0 * co * mc_step;
    p.dy = rad0 * s * mc_step;
    p.dz = rad1 * cos(ang1) * mc_step;
    return p;
}

} // namespace tdvmc
