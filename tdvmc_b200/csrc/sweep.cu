// K1 — Metropolis sweep: one warp owns one walker.
//
// Replaces DoMetropolisStep + CalculateWFChange/Quotient + AcceptMove of the reference
// (src/TDVMC.cpp:858-916, BosonsBulk.cpp:553-665, NUBosonsBulkPB.cpp:672-773).
//
// The reference rebuilds two K-vectors of basis sums per proposal and takes a P-term dot product
// with uR.  Only the difference of exponents enters the acceptance test, and the exponent is
// linear in the pair terms, so here the parameters are contracted once per parameter update into
// a per-knot-interval cubic u(r) = sum_k u~_k B_k(r) (local coordinate r - t_lo, built in extended
// precision on the host from the caller's spline table), and a proposal costs 2(N-1) distance +
// cubic evaluations and one warp reduction:
//     delta = sum_{i != p} u(|r_new - r_i|) - u(|r_old - r_i|),   accept iff log(U) <= 2 delta.
//
// Positions stay in shared memory for the whole launch (structure of arrays, one row per
// coordinate); HBM is touched once on entry and once on exit.  ncu (profiles/r01a_sweep_ncu.txt)
// showed the first version bound by shared-memory wavefronts (83 %) ahead of the FP64 pipe (55 %),
// so the table is split into 16-byte planes (a random record index then spreads over all eight
// 16-byte slots of a bank row), uniform knots need no t_lo load, positions are kept wrapped into
// the first cell so the minimum image is min(|d|, L - |d|) (3 FP64 ops per coordinate instead of 4),
// and the square root is the hardware seed plus ONE Newton step (sweep_math.cuh: three FP64 instructions, relative error
// <= 1.4e-12 - irrelevant for sampling, and a deterministic function of the distance; 22 FP64 instructions per pair in all.
// Round 1's third-order step, five instructions and ~2 ulp, gave 781 against 841 M walker-steps/s at 4096 walkers, with
// bit-identical chains over 82 M proposals, profiles/ab_sweep_sqrt.py).
// Not kept: 64-bit fixed-point coordinates (the two's-complement difference IS the minimum image, which
// moves 9 of 26 FP64 instructions per pair to the integer pipe and three I2F.F64.S64).  Measured equal
// (14.01 vs 14.07 ms per launch): the conversions issue at 14 lanes/clk/SM (profiles/microbench/
// conv_throughput.cu) and with five warps per scheduler they serialise with the FP64 work instead of
// overlapping it (ncu: FP64 53 % + conversion unit 49 %).
// Also not kept: deciding the fold on the integer pipe (64-bit compare of |d|'s bits with those of L/2, then ONE FP64
// add |d| - (wrap ? L : 0)): 21 instead of 24 FP64 instructions per pair but 48 instead of 36 instructions in all,
// and the launch went from 13.7 to 15.4 ms - with five warps per scheduler the issue slots matter as much as the
// FP64 pipe.
// A second capture (profiles/r01c_sweep_ncu.txt) still had 80 % shared-memory wavefronts: a random
// LDS.128 is served quarter-warp by quarter-warp, and two of eight lanes hitting the same 16-byte
// slot of a bank row with different records serialise.  The coefficient planes are therefore
// replicated eight times, copy c in slot c of every 128-byte row, and lane l reads copy l % 8: each
// quarter-warp covers the eight slots exactly once, whatever the record indices are.
#include "kernels.cuh"
#include "sweep_math.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace tdvmc
{

#ifndef TDVMC_CUB_COPIES
#define TDVMC_CUB_COPIES 8
#endif
constexpr int kCubCopies = TDVMC_CUB_COPIES;  // shared-memory replicas of the coefficient planes (one per 16-byte slot)

// sum over the GROUP lanes of a walker (all lanes of the warp take part)
template <int GROUP>
__device__ __forceinline__ double group_sum(double v)
{
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// lanes per walker for a system of N particles
static inline int sweep_group(const SysDev& s)
{
    if (s.dim < 3 && s.kind == 0) return 32; // the low-dimensional instance exists for whole-warp walkers only
    return s.N <= 8 ? 8 : (s.N <= 16 ? 16 : 32);
}

// pair term of the exponent at distance r, with the system's cut rule
template <bool UNIFORM, bool REFLECT, int STRIDE, bool HE = false>
__device__ __forceinline__ double pair_u(const SysDev& s, const double2* __restrict__ c01p, const double2* __restrict__ c23p,
                                         const double2* __restrict__ ttp, const unsigned short* __restrict__ lut, double r)
{
    // Beyond r_max every pair contributes the constant u_tail (BosonsBulk.cpp:207-210, 532-534): record
    // nbins of the table holds exactly that constant, so the cut needs no compare/select.  (At r == r_max
    // the reference's '<=' (:571) / '<' (:593) pick the spline or the tail; a null set for sampling.)
    if (REFLECT) r = (r < s.rmax) ? r : 2.0 * s.rmax - r; // NUBosonsBulkPB.cpp:689-692
    if (HE)
    {
        // HeBulk: McMillan core u~_mc r^m below rijSplit (HeBulk.cpp:471-474); the uniform grid starts at rijSplit
        if (r < s.r0) return s.u_core * pow(r, s.core_m);
        if (r >= s.r_tail) return fma(s.u_lin, r, s.u_const); // const + linear tails, HeDrop.cpp:732-737
        r -= s.r0;
    }
    // floor(r * inv) via the rounding constant, added with round-down in the same FMA; the integer sits in the low word
    const double y = __fma_rd(r, UNIFORM ? s.inv_h : s.inv_cell, kMagic);
    const int c = __double2loint(y);
    double t;
    int j;
    if (UNIFORM)
    {
        j = max(0, min(c, s.nbins));          // record nbins: constant tail
        t = fma(-(y - kMagic), s.h, r);       // r - floor(r/h) h (multiplies zero coefficients when j was clamped)
        // (the index converted back on the conversion unit, (double)j, instead of this FP64 add: 838 against 841 M walker-steps/s)
    }
    else
    {
        j = (int)lut[max(0, min(c, s.ncell - 1))] - s.first_bin;
        double2 tt = ttp[j];
        while (r > tt.y)
        {
            j++;
            tt = ttp[j];
        }
        t = r - tt.x;
    }
    const double2 c01 = c01p[j * STRIDE]; // STRIDE = 8: replicated planes, the caller's pointers select this lane's copy
    const double2 c23 = c23p[j * STRIDE];
    const double v = fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
    return v;
}

// GROUP = lanes per walker (32, or 16 / 8 for systems of at most 16 / 8 particles, where a whole warp per walker would
// leave most lanes without a partner: HeDrop's six atoms run four walkers per warp).
// LOWDIM: BosonsBulk / NUBosonsBulkPB with DIM = 1 or 2 (unused coordinates zero, only DIM Gaussian components move a
// particle, src/TDVMC.cpp:872-875) - a separate instance, because even two predicated moves per proposal batch perturb the
// register allocation of the three-dimensional kernel (measured: 651 vs 664 M walker-steps/s on the same GPU).
template <bool UNIFORM, bool REFLECT, int UNROLL, bool HE = false, bool OPEN = false, int GROUP = 32, bool LOWDIM = false>
__global__ void __launch_bounds__(GROUP == 32 ? kSweepMaxThreadsWarp : kSweepMaxThreads, kSweepMinBlocks) sweep_kernel(SweepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SysDev& s = a.s;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int WPW = 32 / GROUP;          // walkers per warp
    const int gl = lane & (GROUP - 1);       // lane within the walker's group
    const int grp = lane / GROUP;
    const int Npp = a.npp;
    const int nrec = s.nbins + 1;

    // shared memory: c01 plane [nrec][8 copies] | c23 plane [nrec][8 copies] | (non-uniform: t plane [nrec] | lut) | positions
    double2* c01s = reinterpret_cast<double2*>(smem_raw);
    double2* c23s = c01s + (size_t)nrec * kCubCopies;
    double2* tts = c23s + (size_t)nrec * kCubCopies;
    unsigned short* lut = reinterpret_cast<unsigned short*>(tts + (UNIFORM ? 0 : nrec));
    double* pos_base = reinterpret_cast<double*>(smem_raw + a.pos_offset);
    {
        const double2* g01 = reinterpret_cast<const double2*>(s.cub);
        const double2* g23 = g01 + nrec;
        const double2* gtt = g23 + nrec;
        for (int i = threadIdx.x; i < nrec * kCubCopies; i += blockDim.x)
        {
            c01s[i] = g01[i / kCubCopies];
            c23s[i] = g23[i / kCubCopies];
        }
        if (!UNIFORM)
        {
            for (int i = threadIdx.x; i < nrec; i += blockDim.x) tts[i] = gtt[i];
            for (int i = threadIdx.x; i < s.ncell; i += blockDim.x) lut[i] = s.lut[i];
        }
    }
    const double2* c01p = c01s + (lane & (kCubCopies - 1));
    const double2* c23p = c23s + (lane & (kCubCopies - 1));
    const double2* ttp = tts;

    const int w = (blockIdx.x * a.wpb + warp) * WPW + grp; // local walker
    double* px = pos_base + (size_t)(warp * WPW + grp) * 3 * Npp;
    double* py = px + Npp;
    double* pz = py + Npp;
    const bool have = w < a.W;
    double* gpos = a.pos + (size_t)(have ? w : 0) * 3 * s.Np; // a group without a walker idles on a copy of walker 0
    const double L = s.L, Linv = s.Linv, Lhalf = s.Lhalf;
    {
        for (int i = gl; i < s.N; i += GROUP) // positions live wrapped into the first cell during the sweep
        {
            px[i] = OPEN ? gpos[i] : wrap_fast(gpos[i], L, Linv);
            py[i] = OPEN ? gpos[s.Np + i] : wrap_fast(gpos[s.Np + i], L, Linv);
            pz[i] = OPEN ? gpos[2 * s.Np + i] : wrap_fast(gpos[2 * s.Np + i], L, Linv);
        }
    }
    __syncthreads();
    if (GROUP == 32 && !have) return; // (smaller groups stay: the warp's shuffles need every lane)

    const uint32_t gw = (uint32_t)(a.first_walker + w);
    const int N = s.N;
    unsigned long long n_acc = 0;

    for (long long t0 = 0; t0 < a.n_steps; t0 += GROUP)
    {
        // every lane draws the proposal of one of its walker's next GROUP steps
        Proposal mine;
        mine.particle = 0;
        mine.dx = mine.dy = mine.dz = 0.0;
        mine.log_u = 0.0;
        if (t0 + gl < a.n_steps) mine = make_proposal(a.seed, gw, a.first_step + (uint64_t)(t0 + gl), N, a.mc_step);
        if (LOWDIM)
        {
            if (s.dim < 3) mine.dz = 0.0;
            if (s.dim < 2) mine.dy = 0.0;
        }
        const int nsub = (int)min((long long)GROUP, a.n_steps - t0);

        for (int sidx = 0; sidx < nsub; sidx++)
        {
            const int p = __shfl_sync(FULL_MASK, mine.particle, sidx, GROUP);
            const double ddx = __shfl_sync(FULL_MASK, mine.dx, sidx, GROUP);
            const double ddy = __shfl_sync(FULL_MASK, mine.dy, sidx, GROUP);
            const double ddz = __shfl_sync(FULL_MASK, mine.dz, sidx, GROUP);
            const double log_u = __shfl_sync(FULL_MASK, mine.log_u, sidx, GROUP);

            const double ox = px[p], oy = py[p], oz = pz[p];
            const double nx = OPEN ? ox + ddx : wrap_fast(ox + ddx, L, Linv); // src/TDVMC.cpp:872-875, kept in the first cell
            const double ny = OPEN ? oy + ddy : wrap_fast(oy + ddy, L, Linv);
            const double nz = OPEN ? oz + ddz : wrap_fast(oz + ddz, L, Linv);

            double delta = 0.0;
#pragma unroll UNROLL
            for (int i = gl; i < N; i += GROUP)
            {
                const double xi = px[i], yi = py[i], zi = pz[i];
                const double r_old = sqrt_fast(dist2<OPEN>(xi - ox, yi - oy, zi - oz, Lhalf));
                const double r_new = sqrt_fast(dist2<OPEN>(xi - nx, yi - ny, zi - nz, Lhalf));
                const double u_old = pair_u<UNIFORM, REFLECT, kCubCopies, HE>(s, c01p, c23p, ttp, lut, r_old);
                const double u_new = pair_u<UNIFORM, REFLECT, kCubCopies, HE>(s, c01p, c23p, ttp, lut, r_new);
                const double d = u_new - u_old;
                if (i != p) delta += d;
            }
            delta = group_sum<GROUP>(delta);

            // quotient = exp(2 delta) must be finite and >= U (src/TDVMC.cpp:886-913), in the log domain
            const double two_delta = 2.0 * delta;
            const bool accept = (two_delta >= log_u) && (two_delta <= 709.782712893384);
            __syncwarp(); // every lane has read the old positions before lane 0 overwrites one
            if (accept)
            {
                if (gl == 0)
                {
                    px[p] = nx;
                    py[p] = ny;
                    pz[p] = nz;
                }
                n_acc++;
            }
            __syncwarp();
        }
    }

    if (!have) return;
    for (int i = gl; i < N; i += GROUP)
    {
        gpos[i] = px[i];
        gpos[s.Np + i] = py[i];
        gpos[2 * s.Np + i] = pz[i];
    }
    if (gl == 0) a.accepted[w] += n_acc;
}

// ---------------------------------------------------------------------------------------------------------------
// Small ensembles / large systems: WARPS warps share one walker (r02).
//
// With one warp per walker an ensemble of fewer than ~20 x 148 walkers leaves FP64 issue slots empty - the reference's own
// production runs are O(256) chains (src/MPIMethods.h:334-347) - and a walker of N = 1728 particles fills 41 KB of shared
// memory, so only four warps fit on an SM.  Here the partners of the moved particle are dealt out to WARPS x 32 lanes; every
// warp draws the same proposals (counter-based stream: redundant, no exchange), reduces its part of the exponent change
// by shuffle, the parts meet in shared memory at a named barrier of the walker's warps, and every warp takes the same
// accept decision from the same fixed-order sum.  A second barrier orders the position write of the leading warp before
// the next proposal reads it.  Same proposal stream, same accept rule as sweep_kernel: the chains coincide up to the
// summation order of the exponent change.
template <bool UNIFORM, bool REFLECT, int WARPS, int COPIES = kCubCopies>
__global__ void __launch_bounds__(kSweepMaxThreadsWarp, 1) sweep_split_kernel(SweepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SysDev& s = a.s;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int slot = warp / WARPS;            // walker of this block
    const int wsub = warp % WARPS;            // this warp's share of the walker
    const int gl = wsub * 32 + lane;          // lane within the walker's WARPS x 32 lanes
    constexpr int STRIDE = WARPS * 32;
    const int Npp = a.npp;
    const int nrec = s.nbins + 1;

    double2* c01s = reinterpret_cast<double2*>(smem_raw);
    double2* c23s = c01s + (size_t)nrec * COPIES;
    double2* tts = c23s + (size_t)nrec * COPIES;
    unsigned short* lut = reinterpret_cast<unsigned short*>(tts + (UNIFORM ? 0 : nrec));
    double* pos_base = reinterpret_cast<double*>(smem_raw + a.pos_offset);
    double* red_base = pos_base + (size_t)a.wpb * 3 * Npp; // [slots][WARPS] partial exponent changes
    {
        const double2* g01 = reinterpret_cast<const double2*>(s.cub);
        const double2* g23 = g01 + nrec;
        const double2* gtt = g23 + nrec;
        for (int i = threadIdx.x; i < nrec * COPIES; i += blockDim.x)
        {
            c01s[i] = g01[i / COPIES];
            c23s[i] = g23[i / COPIES];
        }
        if (!UNIFORM)
        {
            for (int i = threadIdx.x; i < nrec; i += blockDim.x) tts[i] = gtt[i];
            for (int i = threadIdx.x; i < s.ncell; i += blockDim.x) lut[i] = s.lut[i];
        }
    }
    const double2* c01p = c01s + (lane & (COPIES - 1));
    const double2* c23p = c23s + (lane & (COPIES - 1));
    const double2* ttp = tts;

    const int w = blockIdx.x * a.wpb + slot; // local walker (a.wpb = walkers per block here)
    double* px = pos_base + (size_t)slot * 3 * Npp;
    double* py = px + Npp;
    double* pz = py + Npp;
    double* red = red_base + slot * WARPS;
    const bool have = w < a.W;
    double* gpos = a.pos + (size_t)(have ? w : 0) * 3 * s.Np;
    const double L = s.L, Linv = s.Linv, Lhalf = s.Lhalf;
    for (int i = gl; i < s.N; i += STRIDE)
    {
        px[i] = wrap_fast(gpos[i], L, Linv);
        py[i] = wrap_fast(gpos[s.Np + i], L, Linv);
        pz[i] = wrap_fast(gpos[2 * s.Np + i], L, Linv);
    }
    __syncthreads();
    if (!have) return; // all warps of the slot leave together

    const uint32_t gw = (uint32_t)(a.first_walker + w);
    const int N = s.N;
    const int bar_id = 1 + slot; // named barrier of this walker's warps (0 is __syncthreads)
    unsigned long long n_acc = 0;

    for (long long t0 = 0; t0 < a.n_steps; t0 += 32)
    {
        Proposal mine;
        mine.particle = 0;
        mine.dx = mine.dy = mine.dz = 0.0;
        mine.log_u = 0.0;
        if (t0 + lane < a.n_steps) mine = make_proposal(a.seed, gw, a.first_step + (uint64_t)(t0 + lane), N, a.mc_step);
        const int nsub = (int)min(32ll, a.n_steps - t0);

        for (int sidx = 0; sidx < nsub; sidx++)
        {
            const int p = __shfl_sync(FULL_MASK, mine.particle, sidx);
            const double ddx = __shfl_sync(FULL_MASK, mine.dx, sidx);
            const double ddy = __shfl_sync(FULL_MASK, mine.dy, sidx);
            const double ddz = __shfl_sync(FULL_MASK, mine.dz, sidx);
            const double log_u = __shfl_sync(FULL_MASK, mine.log_u, sidx);

            const double ox = px[p], oy = py[p], oz = pz[p];
            const double nx = wrap_fast(ox + ddx, L, Linv);
            const double ny = wrap_fast(oy + ddy, L, Linv);
            const double nz = wrap_fast(oz + ddz, L, Linv);

            double delta = 0.0;
#pragma unroll 2
            for (int i = gl; i < N; i += STRIDE)
            {
                const double xi = px[i], yi = py[i], zi = pz[i];
                const double r_old = sqrt_fast(dist2<false>(xi - ox, yi - oy, zi - oz, Lhalf));
                const double r_new = sqrt_fast(dist2<false>(xi - nx, yi - ny, zi - nz, Lhalf));
                const double u_old = pair_u<UNIFORM, REFLECT, COPIES>(s, c01p, c23p, ttp, lut, r_old);
                const double u_new = pair_u<UNIFORM, REFLECT, COPIES>(s, c01p, c23p, ttp, lut, r_new);
                const double d = u_new - u_old;
                if (i != p) delta += d;
            }
            delta = group_sum<32>(delta);
            if (lane == 0) red[wsub] = delta;
            // all warps of the walker: parts published, old positions read
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(STRIDE) : "memory");
            double tot = 0.0;
#pragma unroll
            for (int q = 0; q < WARPS; q++) tot += red[q];

            const double two_delta = 2.0 * tot;
            const bool accept = (two_delta >= log_u) && (two_delta <= 709.782712893384);
            if (accept)
            {
                if (gl == 0)
                {
                    px[p] = nx;
                    py[p] = ny;
                    pz[p] = nz;
                }
                n_acc++;
            }
            // new position visible, sums consumed, before the next proposal
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(STRIDE) : "memory");
        }
    }

    for (int i = gl; i < N; i += STRIDE)
    {
        gpos[i] = px[i];
        gpos[s.Np + i] = py[i];
        gpos[2 * s.Np + i] = pz[i];
    }
    if (gl == 0) a.accepted[w] += n_acc;
}

// warps per walker for an ensemble of W walkers on sm_count SMs: enough to reach ~16 warps per SM, at least two partners
// per lane, 1 (= sweep_kernel) when the ensemble fills the machine by itself
int sweep_split_warps(const SysDev& s, int W, int sm_count, int resident_per_sm)
{
    if (s.kind != 0 || s.dim != 3 || W <= 0) return 1;
    if (const char* e = getenv("TDVMC_SWEEP_SPLIT")) // tuning knob: 1 = never split
    {
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) return v;
    }
    const int wps = (W + sm_count - 1) / sm_count; // walkers per SM
    const int target = resident_per_sm > 0 ? (resident_per_sm < 16 ? 16 : resident_per_sm) : 16;
    int warps = 1;
    while (warps < 8 && wps * warps * 2 <= target && 64 * warps * 2 <= s.N) warps *= 2;
    return warps;
}

static size_t sweep_split_smem_bytes(const SysDev& s, int spb, int warps, int npp, int copies, size_t* pos_offset)
{
    const size_t nrec = (size_t)s.nbins + 1;
    size_t off = nrec * copies * 2 * sizeof(double2);
    if (!s.uniform) off += nrec * sizeof(double2) + (size_t)s.ncell * sizeof(unsigned short);
    off = (off + 15) & ~(size_t)15;
    *pos_offset = off;
    return off + (size_t)spb * 3 * npp * sizeof(double) + (size_t)spb * warps * sizeof(double);
}

bool sweep_large_fits(const SysDev& s, int npp, int smem_optin)
{
    size_t off;
    return sweep_split_smem_bytes(s, 1, 8, npp, 4, &off) <= (size_t)smem_optin;
}

template <int WARPS, int COPIES>
static const void* sweep_split_fn(const SysDev& s)
{
    const bool refl = s.pair_rule == 1;
    return s.uniform ? (refl ? (const void*)sweep_split_kernel<true, true, WARPS, COPIES> : (const void*)sweep_split_kernel<true, false, WARPS, COPIES>)
                     : (refl ? (const void*)sweep_split_kernel<false, true, WARPS, COPIES> : (const void*)sweep_split_kernel<false, false, WARPS, COPIES>);
}

// a.wpb is ignored; walkers per block follow from the ensemble size, the thread budget, the named barriers and shared memory.
// copies = 4: the large-system variant (eight warps per walker only)
cudaError_t launch_sweep_split(SweepArgs a, int warps, int sm_count, int smem_optin, cudaStream_t st, int copies)
{
    if (copies == 4) warps = 8;
    const int wps = (a.W + sm_count - 1) / sm_count;
    int spb = wps < 1 ? 1 : wps;
    spb = std::min(spb, std::min(kSweepMaxThreadsWarp / (32 * warps), 15));
    size_t pos_off = 0, smem = 0;
    for (; spb >= 1; spb--)
    {
        smem = sweep_split_smem_bytes(a.s, spb, warps, a.npp, copies, &pos_off);
        if (smem <= (size_t)smem_optin) break;
    }
    if (spb < 1) return cudaErrorInvalidConfiguration;
    a.wpb = spb;
    a.pos_offset = (int)pos_off;
    const void* fn = copies == 4 ? sweep_split_fn<8, 4>(a.s)
                     : warps == 2 ? sweep_split_fn<2, kCubCopies>(a.s) : (warps == 4 ? sweep_split_fn<4, kCubCopies>(a.s) : sweep_split_fn<8, kCubCopies>(a.s));
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    void* args[] = { &a };
    return cudaLaunchKernel(fn, dim3((a.W + spb - 1) / spb), dim3(spb * warps * 32), args, smem, st);
}

// ---------------------------------------------------------------------------------------------------------------
// Ensembles that are not a whole number of waves: walkers time-share the resident warps (r02).
//
// sweep_kernel keeps one walker per warp for the whole launch, so an ensemble of 4096 walkers on 148 SMs x 20 resident
// warps (1.38 waves - the size SURVEY.md 8(d) specifies) takes the time of two waves.  Here every SM gets an equal share of
// the walkers (27 or 28 at W = 4096) and its 20 warps draw work units (walker, chunk of ~n_steps / 10 steps) from a
// shared-memory counter, chunk-major: all warps stay busy until the last round, the launch takes ~(walkers per SM / 20)
// instead of ceil(.) wave times.  A walker's state travels through HBM between its chunks (16.5 KB per unit - nothing
// against ~10^7 FP64 instructions per unit); a per-walker flag orders chunk c + 1 after chunk c.  The proposal stream is a
// function of (seed, walker, step), so the chains are those of sweep_kernel.
template <bool UNIFORM, bool REFLECT>
__global__ void __launch_bounds__(kSweepMaxThreadsWarp, 1) sweep_queue_kernel(SweepArgs a, int chunk_steps)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SysDev& s = a.s;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int Npp = a.npp;
    const int nrec = s.nbins + 1;

    double2* c01s = reinterpret_cast<double2*>(smem_raw);
    double2* c23s = c01s + (size_t)nrec * kCubCopies;
    double2* tts = c23s + (size_t)nrec * kCubCopies;
    unsigned short* lut = reinterpret_cast<unsigned short*>(tts + (UNIFORM ? 0 : nrec));
    double* pos_base = reinterpret_cast<double*>(smem_raw + a.pos_offset);
    int* s_next = reinterpret_cast<int*>(pos_base + (size_t)a.wpb * 3 * Npp); // unit counter, then done[walkers of this block]
    int* s_done = s_next + 1; // per-walker count of finished chunks: written and polled with atomics only (one lane each)

    const int w_begin = (int)((long long)blockIdx.x * a.W / gridDim.x);
    const int w_end = (int)((long long)(blockIdx.x + 1) * a.W / gridDim.x);
    const int nW = w_end - w_begin;
    {
        const double2* g01 = reinterpret_cast<const double2*>(s.cub);
        const double2* g23 = g01 + nrec;
        const double2* gtt = g23 + nrec;
        for (int i = threadIdx.x; i < nrec * kCubCopies; i += blockDim.x)
        {
            c01s[i] = g01[i / kCubCopies];
            c23s[i] = g23[i / kCubCopies];
        }
        if (!UNIFORM)
        {
            for (int i = threadIdx.x; i < nrec; i += blockDim.x) tts[i] = gtt[i];
            for (int i = threadIdx.x; i < s.ncell; i += blockDim.x) lut[i] = s.lut[i];
        }
        if (threadIdx.x == 0) *s_next = 0;
        for (int i = threadIdx.x; i < nW; i += blockDim.x) s_done[i] = 0;
    }
    __syncthreads();
    const double2* c01p = c01s + (lane & (kCubCopies - 1));
    const double2* c23p = c23s + (lane & (kCubCopies - 1));
    const double2* ttp = tts;
    double* px = pos_base + (size_t)warp * 3 * Npp;
    double* py = px + Npp;
    double* pz = py + Npp;
    const double L = s.L, Linv = s.Linv, Lhalf = s.Lhalf;
    const int N = s.N;
    const int n_chunks = (int)((a.n_steps + chunk_steps - 1) / chunk_steps);
    const int units = nW * n_chunks;

    for (;;)
    {
        int unit = 0;
        if (lane == 0) unit = atomicAdd(s_next, 1);
        unit = __shfl_sync(FULL_MASK, unit, 0);
        if (unit >= units) break;
        const int c = unit / nW, wl = unit - c * nW;
        const int w = w_begin + wl;
        if (lane == 0)
            while (atomicAdd(s_done + wl, 0) < c) __nanosleep(200); // the walker's previous chunk (a smaller unit: no deadlock)
        __syncwarp();
        __threadfence_block();
        double* gpos = a.pos + (size_t)w * 3 * s.Np;
        for (int i = lane; i < N; i += 32)
        {
            px[i] = wrap_fast(__ldcg(gpos + i), L, Linv);
            py[i] = wrap_fast(__ldcg(gpos + s.Np + i), L, Linv);
            pz[i] = wrap_fast(__ldcg(gpos + 2 * s.Np + i), L, Linv);
        }
        __syncwarp();

        const uint32_t gw = (uint32_t)(a.first_walker + w);
        const long long t_begin = (long long)c * chunk_steps;
        const long long t_end = min(a.n_steps, t_begin + chunk_steps);
        unsigned long long n_acc = 0;
        for (long long t0 = t_begin; t0 < t_end; t0 += 32)
        {
            Proposal mine;
            mine.particle = 0;
            mine.dx = mine.dy = mine.dz = 0.0;
            mine.log_u = 0.0;
            if (t0 + lane < t_end) mine = make_proposal(a.seed, gw, a.first_step + (uint64_t)(t0 + lane), N, a.mc_step);
            const int nsub = (int)min(32ll, t_end - t0);
            for (int sidx = 0; sidx < nsub; sidx++)
            {
                const int p = __shfl_sync(FULL_MASK, mine.particle, sidx);
                const double ddx = __shfl_sync(FULL_MASK, mine.dx, sidx);
                const double ddy = __shfl_sync(FULL_MASK, mine.dy, sidx);
                const double ddz = __shfl_sync(FULL_MASK, mine.dz, sidx);
                const double log_u = __shfl_sync(FULL_MASK, mine.log_u, sidx);
                const double ox = px[p], oy = py[p], oz = pz[p];
                const double nx = wrap_fast(ox + ddx, L, Linv);
                const double ny = wrap_fast(oy + ddy, L, Linv);
                const double nz = wrap_fast(oz + ddz, L, Linv);
                double delta = 0.0;
#pragma unroll 2
                for (int i = lane; i < N; i += 32)
                {
                    const double xi = px[i], yi = py[i], zi = pz[i];
                    const double r_old = sqrt_fast(dist2<false>(xi - ox, yi - oy, zi - oz, Lhalf));
                    const double r_new = sqrt_fast(dist2<false>(xi - nx, yi - ny, zi - nz, Lhalf));
                    const double u_old = pair_u<UNIFORM, REFLECT, kCubCopies>(s, c01p, c23p, ttp, lut, r_old);
                    const double u_new = pair_u<UNIFORM, REFLECT, kCubCopies>(s, c01p, c23p, ttp, lut, r_new);
                    const double d = u_new - u_old;
                    if (i != p) delta += d;
                }
                delta = group_sum<32>(delta);
                const double two_delta = 2.0 * delta;
                const bool accept = (two_delta >= log_u) && (two_delta <= 709.782712893384);
                __syncwarp();
                if (accept)
                {
                    if (lane == 0)
                    {
                        px[p] = nx;
                        py[p] = ny;
                        pz[p] = nz;
                    }
                    n_acc++;
                }
                __syncwarp();
            }
        }
        for (int i = lane; i < N; i += 32)
        {
            __stcg(gpos + i, px[i]);
            __stcg(gpos + s.Np + i, py[i]);
            __stcg(gpos + 2 * s.Np + i, pz[i]);
        }
        if (lane == 0) atomicAdd(a.accepted + w, n_acc); // (successive chunks of a walker run on different warps)
        __threadfence_block();
        __syncwarp();
        if (lane == 0) atomicExch(s_done + wl, c + 1);
    }
}

// Use the time-shared launch?  Only for the three-dimensional spline-table systems, when the ensemble is more than one
// wave and the last wave of a plain launch would leave more than ~8 % of the machine idle.
bool sweep_queue_wanted(const SysDev& s, int W, int sm_count, int resident_per_sm, long long n_steps)
{
    if (s.kind != 0 || s.dim != 3 || resident_per_sm <= 0 || n_steps < 320) return false;
    int knob = 1;
    if (const char* e = getenv("TDVMC_SWEEP_QUEUE")) knob = atoi(e); // tuning knob: 0 = never, 2 = whenever possible (tests)
    if (knob == 0) return false;
    if ((W + sm_count - 1) / sm_count > 4096) return false; // (the flag array lives in shared memory)
    if (knob == 2) return true;
    const double waves = (double)W / ((double)sm_count * resident_per_sm);
    if (waves <= 1.0) return false;
    return waves / std::ceil(waves) < 0.92;
}

cudaError_t launch_sweep_queue(SweepArgs a, int sm_count, int smem_optin, cudaStream_t st)
{
    const size_t nrec = (size_t)a.s.nbins + 1;
    size_t off = nrec * kCubCopies * 2 * sizeof(double2);
    if (!a.s.uniform) off += nrec * sizeof(double2) + (size_t)a.s.ncell * sizeof(unsigned short);
    off = (off + 15) & ~(size_t)15;
    a.pos_offset = (int)off;
    const int per_sm = (a.W + sm_count - 1) / sm_count;
    const size_t smem = off + (size_t)a.wpb * 3 * a.npp * sizeof(double) + (size_t)(per_sm + 2) * sizeof(int);
    if (smem > (size_t)smem_optin) return cudaErrorInvalidConfiguration;
    int chunk = (int)(a.n_steps / 10);
    chunk = std::max(32, (chunk + 31) / 32 * 32);
    const void* fn = a.s.uniform ? (a.s.pair_rule == 1 ? (const void*)sweep_queue_kernel<true, true> : (const void*)sweep_queue_kernel<true, false>)
                                 : (a.s.pair_rule == 1 ? (const void*)sweep_queue_kernel<false, true> : (const void*)sweep_queue_kernel<false, false>);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    void* args[] = { &a, &chunk };
    return cudaLaunchKernel(fn, dim3(sm_count), dim3(a.wpb * 32), args, smem, st);
}

// exponentNew - exponent for scripted moves of one configuration: the ratio evaluator of the sweep,
// one warp per move, tables read straight from global memory (parity entry point, not a hot path)
template <bool UNIFORM, bool REFLECT, bool HE = false, bool OPEN = false>
__global__ void quotient_kernel(QuotientArgs a)
{
    const SysDev& s = a.s;
    const int lane = threadIdx.x & 31;
    const int mv = blockIdx.x;
    const int nrec = s.nbins + 1;
    const double2* c01p = reinterpret_cast<const double2*>(s.cub);
    const double2* c23p = c01p + nrec;
    const double2* ttp = c23p + nrec;
    const double* px = a.pos;
    const double* py = a.pos + s.Np;
    const double* pz = a.pos + 2 * s.Np;
    const double L = s.L, Linv = s.Linv, Lhalf = s.Lhalf;
    const int p = (int)a.moves[mv * 4];
    auto W = [&](double x) { return OPEN ? x : wrap_fast(x, L, Linv); };
    const double nx = W(a.moves[mv * 4 + 1]), ny = W(a.moves[mv * 4 + 2]), nz = W(a.moves[mv * 4 + 3]);
    const double ox = W(px[p]), oy = W(py[p]), oz = W(pz[p]);
    double delta = 0.0;
    for (int i = lane; i < s.N; i += 32)
    {
        const double xi = W(px[i]), yi = W(py[i]), zi = W(pz[i]);
        const double r_old = sqrt_fast(dist2<OPEN>(xi - ox, yi - oy, zi - oz, Lhalf));
        const double r_new = sqrt_fast(dist2<OPEN>(xi - nx, yi - ny, zi - nz, Lhalf));
        const double d = pair_u<UNIFORM, REFLECT, 1, HE>(s, c01p, c23p, ttp, s.lut, r_new) -
                         pair_u<UNIFORM, REFLECT, 1, HE>(s, c01p, c23p, ttp, s.lut, r_old);
        if (i != p) delta += d;
    }
    delta = warp_sum(delta);
    if (lane == 0) a.delta[mv] = delta;
}

cudaError_t launch_quotient(const QuotientArgs& a, cudaStream_t st)
{
    if (a.n_moves <= 0) return cudaSuccess;
    const bool refl = a.s.pair_rule == 1;
    if (a.s.kind == 1)
    {
        quotient_kernel<true, false, true><<<a.n_moves, 32, 0, st>>>(a);
    }
    else if (a.s.kind == 2)
    {
        quotient_kernel<false, false, true, true><<<a.n_moves, 32, 0, st>>>(a);
    }
    else if (a.s.uniform)
    {
        if (refl) quotient_kernel<true, true><<<a.n_moves, 32, 0, st>>>(a);
        else quotient_kernel<true, false><<<a.n_moves, 32, 0, st>>>(a);
    }
    else
    {
        if (refl) quotient_kernel<false, true><<<a.n_moves, 32, 0, st>>>(a);
        else quotient_kernel<false, false><<<a.n_moves, 32, 0, st>>>(a);
    }
    return cudaGetLastError();
}

size_t sweep_smem_bytes(const SysDev& s, int wpb, int npp, size_t* pos_offset)
{
    const size_t nrec = (size_t)s.nbins + 1;
    size_t off = nrec * kCubCopies * 2 * sizeof(double2);
    if (!s.uniform) off += nrec * sizeof(double2) + (size_t)s.ncell * sizeof(unsigned short);
    off = (off + 15) & ~(size_t)15;
    *pos_offset = off;
    return off + (size_t)wpb * (32 / sweep_group(s)) * 3 * npp * sizeof(double);
}

static int sweep_unroll()
{
    static int u = -1;
    if (u < 0)
    {
        const char* e = getenv("TDVMC_SWEEP_UNROLL"); // tuning knob, default 2
        u = e ? atoi(e) : 2;
        if (u != 1 && u != 2 && u != 4) u = 2;
    }
    return u;
}

// the kernel instance for a system: (uniform knots, reflection rule, He family, open boundary) x unroll x lanes per walker
template <int UNROLL, int GROUP>
static const void* sweep_fn_ug(const SysDev& s)
{
    const bool refl = s.pair_rule == 1;
    if (s.kind == 1) return (const void*)sweep_kernel<true, false, UNROLL, true, false, GROUP>;
    if (s.kind == 2) return (const void*)sweep_kernel<false, false, UNROLL, true, true, GROUP>;
    if (s.dim < 3 && GROUP == 32 && UNROLL == 2) // (low-dimensional systems always run the default instance shape)
        return s.uniform ? (refl ? (const void*)sweep_kernel<true, true, 2, false, false, 32, true>
                                 : (const void*)sweep_kernel<true, false, 2, false, false, 32, true>)
                         : (refl ? (const void*)sweep_kernel<false, true, 2, false, false, 32, true>
                                 : (const void*)sweep_kernel<false, false, 2, false, false, 32, true>);
    return s.uniform ? (refl ? (const void*)sweep_kernel<true, true, UNROLL, false, false, GROUP>
                             : (const void*)sweep_kernel<true, false, UNROLL, false, false, GROUP>)
                     : (refl ? (const void*)sweep_kernel<false, true, UNROLL, false, false, GROUP>
                             : (const void*)sweep_kernel<false, false, UNROLL, false, false, GROUP>);
}

static const void* sweep_fn(const SysDev& s)
{
    if (s.dim < 3 && s.kind == 0) return sweep_fn_ug<2, 32>(s);
    const int g = sweep_group(s);
    if (g == 8) return sweep_fn_ug<1, 8>(s);   // one partner per lane at most: nothing to unroll
    if (g == 16) return sweep_fn_ug<1, 16>(s);
    const int u = sweep_unroll();
    return u == 1 ? sweep_fn_ug<1, 32>(s) : (u == 4 ? sweep_fn_ug<4, 32>(s) : sweep_fn_ug<2, 32>(s));
}

cudaError_t launch_sweep(SweepArgs a, cudaStream_t st)
{
    size_t pos_off;
    size_t smem = sweep_smem_bytes(a.s, a.wpb, a.npp, &pos_off);
    a.pos_offset = (int)pos_off;
    const int per_block = a.wpb * (32 / sweep_group(a.s)); // walkers per block
    int grid = (a.W + per_block - 1) / per_block;
    int threads = a.wpb * 32;
    const void* fn = sweep_fn(a.s);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    void* args[] = { &a };
    return cudaLaunchKernel(fn, dim3(grid), dim3(threads), args, smem, st);
}

int sweep_walkers_per_warp(const SysDev& s)
{
    return 32 / sweep_group(s);
}

int sweep_max_threads(const SysDev& s)
{
    return sweep_group(s) == 32 ? kSweepMaxThreadsWarp : kSweepMaxThreads;
}

int sweep_blocks_per_sm(const SysDev& s, int wpb, int npp)
{
    size_t pos_off;
    size_t smem = sweep_smem_bytes(s, wpb, npp, &pos_off);
    int nb = 0;
    const void* fn = sweep_fn(s);
    static int optin = -1;
    if (optin < 0)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) optin = 48 * 1024;
    }
    if (smem > (size_t)optin) return 0;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, wpb * 32, smem) != cudaSuccess) return 0;
    return nb;
}

} // namespace tdvmc
