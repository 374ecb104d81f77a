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
// Positions stay in shared memory for the whole launch (structure of arrays, one row per
// coordinate), HBM is touched once on entry and once on exit.
#include "kernels.cuh"

namespace tdvmc
{

constexpr double kMagic = 6755399441055744.0; // 2^52 + 2^51: (x + kMagic) - kMagic rounds x to nearest

// minimum image of one coordinate difference; ties at exactly +-L/2 are irrelevant for sampling
__device__ __forceinline__ double mi_fast(double d, double L, double Linv)
{
    double k = fma(d, Linv, kMagic) - kMagic;
    return fma(-k, L, d);
}

// pair term of the exponent at distance r, with the system's cut rule
template <bool UNIFORM, bool REFLECT>
__device__ __forceinline__ double pair_u(const SysDev& s, const double* __restrict__ cub,
                                         const unsigned short* __restrict__ lut, double r)
{
    bool inside;
    if (REFLECT)
    {
        if (!(r < s.rmax)) r = 2.0 * s.rmax - r; // NUBosonsBulkPB.cpp:689-692
        inside = r < s.rmax;
    }
    else
    {
        inside = r <= s.rmax; // BosonsBulk.cpp:571 (the strict '<' of :593 differs on a null set)
    }
    // floor(r * inv) via the rounding constant; the integer sits in the low word
    double y = fma(r, UNIFORM ? s.inv_h : s.inv_cell, -0.5) + kMagic;
    int c = __double2loint(y);
    int j;
    if (UNIFORM)
    {
        j = max(0, min(c, s.nbins - 1));
    }
    else
    {
        c = max(0, min(c, s.ncell - 1));
        j = (int)lut[c] - s.first_bin;
    }
    const double* q = cub + j * kCubStride;
    double2 c01 = *reinterpret_cast<const double2*>(q);
    double2 c23 = *reinterpret_cast<const double2*>(q + 2);
    double2 tt = *reinterpret_cast<const double2*>(q + 4);
    if (!UNIFORM)
    {
        while (r > tt.y && j < s.nbins - 1)
        {
            j++;
            q += kCubStride;
            c01 = *reinterpret_cast<const double2*>(q);
            c23 = *reinterpret_cast<const double2*>(q + 2);
            tt = *reinterpret_cast<const double2*>(q + 4);
        }
    }
    double t = r - tt.x;
    double v = fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
    return inside ? v : s.u_tail;
}

template <bool UNIFORM, bool REFLECT>
__global__ void __launch_bounds__(kSweepMaxThreads) sweep_kernel(SweepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SysDev& s = a.s;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int Npp = a.npp;

    double* cub = reinterpret_cast<double*>(smem_raw);
    unsigned short* lut = reinterpret_cast<unsigned short*>(cub + s.nbins * kCubStride);
    double* pos_base = reinterpret_cast<double*>(smem_raw + a.pos_offset);

    for (int i = threadIdx.x; i < s.nbins * kCubStride; i += blockDim.x) cub[i] = s.cub[i];
    if (!UNIFORM)
        for (int i = threadIdx.x; i < s.ncell; i += blockDim.x) lut[i] = s.lut[i];

    const int w = blockIdx.x * a.wpb + warp; // local walker
    double* px = pos_base + (size_t)warp * 3 * Npp;
    double* py = px + Npp;
    double* pz = py + Npp;
    const bool have = w < a.W;
    double* gpos = a.pos + (size_t)(have ? w : 0) * 3 * s.Np;
    if (have)
    {
        for (int i = lane; i < s.N; i += 32)
        {
            px[i] = gpos[i];
            py[i] = gpos[s.Np + i];
            pz[i] = gpos[2 * s.Np + i];
        }
    }
    __syncthreads();
    if (!have) return;

    const uint32_t gw = (uint32_t)(a.first_walker + w);
    const int N = s.N;
    const double L = s.L, Linv = s.Linv;
    unsigned long long n_acc = 0;

    for (long long t0 = 0; t0 < a.n_steps; t0 += 32)
    {
        // every lane draws the proposal of one of the next 32 steps
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
            const double nx = ox + ddx, ny = oy + ddy, nz = oz + ddz; // src/TDVMC.cpp:872-875

            double delta = 0.0;
#pragma unroll 2
            for (int i = lane; i < N; i += 32)
            {
                const double xi = px[i], yi = py[i], zi = pz[i];
                double ax = mi_fast(xi - ox, L, Linv);
                double ay = mi_fast(yi - oy, L, Linv);
                double az = mi_fast(zi - oz, L, Linv);
                double bx = mi_fast(xi - nx, L, Linv);
                double by = mi_fast(yi - ny, L, Linv);
                double bz = mi_fast(zi - nz, L, Linv);
                double r_old = sqrt(fma(az, az, fma(ay, ay, ax * ax)));
                double r_new = sqrt(fma(bz, bz, fma(by, by, bx * bx)));
                double u_old = pair_u<UNIFORM, REFLECT>(s, cub, lut, r_old);
                double u_new = pair_u<UNIFORM, REFLECT>(s, cub, lut, r_new);
                double d = u_new - u_old;
                delta += (i == p) ? 0.0 : d;
            }
            delta = warp_sum(delta);

            // quotient = exp(2 delta) must be finite and >= U (src/TDVMC.cpp:886-913), in the log domain
            const double two_delta = 2.0 * delta;
            const bool accept = (two_delta >= log_u) && (two_delta <= 709.782712893384);
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
        gpos[i] = px[i];
        gpos[s.Np + i] = py[i];
        gpos[2 * s.Np + i] = pz[i];
    }
    if (lane == 0) a.accepted[w] += n_acc;
}

// exponentNew - exponent for scripted moves of one configuration: the ratio evaluator of the sweep,
// one warp per move, tables read straight from global memory (parity entry point, not a hot path)
template <bool UNIFORM, bool REFLECT>
__global__ void quotient_kernel(QuotientArgs a)
{
    const SysDev& s = a.s;
    const int lane = threadIdx.x & 31;
    const int mv = blockIdx.x;
    const double* px = a.pos;
    const double* py = a.pos + s.Np;
    const double* pz = a.pos + 2 * s.Np;
    const int p = (int)a.moves[mv * 4];
    const double nx = a.moves[mv * 4 + 1], ny = a.moves[mv * 4 + 2], nz = a.moves[mv * 4 + 3];
    const double ox = px[p], oy = py[p], oz = pz[p];
    double delta = 0.0;
    for (int i = lane; i < s.N; i += 32)
    {
        const double xi = px[i], yi = py[i], zi = pz[i];
        double ax = mi_fast(xi - ox, s.L, s.Linv), ay = mi_fast(yi - oy, s.L, s.Linv), az = mi_fast(zi - oz, s.L, s.Linv);
        double bx = mi_fast(xi - nx, s.L, s.Linv), by = mi_fast(yi - ny, s.L, s.Linv), bz = mi_fast(zi - nz, s.L, s.Linv);
        double r_old = sqrt(fma(az, az, fma(ay, ay, ax * ax)));
        double r_new = sqrt(fma(bz, bz, fma(by, by, bx * bx)));
        double d = pair_u<UNIFORM, REFLECT>(s, s.cub, s.lut, r_new) - pair_u<UNIFORM, REFLECT>(s, s.cub, s.lut, r_old);
        delta += (i == p) ? 0.0 : d;
    }
    delta = warp_sum(delta);
    if (lane == 0) a.delta[mv] = delta;
}

cudaError_t launch_quotient(const QuotientArgs& a, cudaStream_t st)
{
    if (a.n_moves <= 0) return cudaSuccess;
    const bool refl = a.s.pair_rule == 1;
    if (a.s.uniform)
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
    size_t off = (size_t)s.nbins * kCubStride * sizeof(double);
    if (!s.uniform) off += (size_t)s.ncell * sizeof(unsigned short);
    off = (off + 15) & ~(size_t)15;
    *pos_offset = off;
    return off + (size_t)wpb * 3 * npp * sizeof(double);
}

template <bool U, bool R>
static cudaError_t launch_one(const SweepArgs& a, int grid, int threads, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(sweep_kernel<U, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    sweep_kernel<U, R><<<grid, threads, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_sweep(SweepArgs a, cudaStream_t st)
{
    size_t pos_off;
    size_t smem = sweep_smem_bytes(a.s, a.wpb, a.npp, &pos_off);
    a.pos_offset = (int)pos_off;
    int grid = (a.W + a.wpb - 1) / a.wpb;
    int threads = a.wpb * 32;
    bool refl = a.s.pair_rule == 1;
    if (a.s.uniform) return refl ? launch_one<true, true>(a, grid, threads, smem, st) : launch_one<true, false>(a, grid, threads, smem, st);
    return refl ? launch_one<false, true>(a, grid, threads, smem, st) : launch_one<false, false>(a, grid, threads, smem, st);
}

int sweep_blocks_per_sm(const SysDev& s, int wpb, int npp)
{
    size_t pos_off;
    size_t smem = sweep_smem_bytes(s, wpb, npp, &pos_off);
    int nb = 0;
    bool refl = s.pair_rule == 1;
    const void* fn = s.uniform ? (refl ? (const void*)sweep_kernel<true, true> : (const void*)sweep_kernel<true, false>)
                               : (refl ? (const void*)sweep_kernel<false, true> : (const void*)sweep_kernel<false, false>);
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, wpb * 32, smem) != cudaSuccess) return 0;
    return nb;
}

} // namespace tdvmc
