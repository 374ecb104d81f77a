// NUBosonsBulkPBBoxAndRadial (src/PhysicalSystems/NUBosonsBulkPBBoxAndRadial.cpp) on the device: periodic bosons whose
// Jastrow exponent has a radial spline basis in r_ij (inside maxDistanceRad) AND a "box" spline basis evaluated at
// |x_ij|, |y_ij|, |z_ij| of the minimum-image displacement, both on the same (non-uniform) knots; Gauss pair potential;
// g(r) histogram carried in otherExpectationValues.  Extended sums ext = [ssRad_0..K-1 | ss_0..K-1].
//
//   evaluate_br_kernel   CalculateLocalOperators :213-262, CalculateOtherLocalOperators :264-425,
//                        CalculateExpectationValues :438-580 fused: ONE WARP per configuration, lane n owns particle n
//                        (drift in registers, no atomics on forces), the four tables of the reference never exist
//   sweep_br_kernel      DoMetropolisStep (src/TDVMC.cpp:858-916) + CalculateWFChange/Quotient :632-747 + AcceptMove
//                        :749-755: one warp per walker, positions in shared memory, both bases pre-contracted with the
//                        parameters into per-interval cubics (as sweep.cu does for the radial systems)
//   quotient_br_kernel   the sweep's ratio evaluator on scripted moves (parity entry point)
//
// The per-pair expressions are the reference's, in its order (compiled with -fmad=false); only the order of the sums
// over partners and basis functions differs.  One deviation of the reference is reproduced, not corrected: the drift
// contracts the last radial spline's parameter with the BOX table (:493-497) - s.ugR / s.ugI carry that.
// (A displacement component of exactly zero crashes the reference - lower_bound lands on bin 2 and it reads
// splineWeights[-1], :247-256; here it evaluates the first interval.)
#include "kernels.cuh"
#include "sweep_math.cuh"

namespace tdvmc
{

constexpr int kBrWarps = 4; // configurations (warps) per block of the evaluation kernel

// the four overlapping pieces at x on interval `bin`: value sums (optional), first / second derivative contracted with u~
struct BrAcc
{
    double gR, gI, lR, lI;
};

struct BrEvalTables // block-shared copies in shared memory
{
    const double* knots;
    const double* rec;
    const unsigned short* lut;
};

template <bool RADIAL>
__device__ __forceinline__ void br_pieces(const SysDev& s, const BrEvalTables& tb, double x, double f2, bool values,
                                          double* __restrict__ ext, const double* __restrict__ ugR,
                                          const double* __restrict__ ugI, const double* __restrict__ ulR,
                                          const double* __restrict__ ulI, BrAcc& acc)
{
    const int bin = find_bin_exact(s, tb.knots, tb.lut, x);
    const double* rec = tb.rec + (size_t)(bin - s.first_bin) * kRecStride;
    const double x2 = x * x, x3 = x2 * x;
#pragma unroll
    for (int p = 0; p < 4; p++)
    {
        const double* q = rec + p * 4;
        const double w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
        const double d1 = w1 + 2.0 * w2 * x + 3.0 * w3 * x2; // :356, :393
        const double d2 = 2.0 * w2 + 6.0 * w3 * x;           // :358, :395
        const int k = bin - p;
        acc.gR = fma(ugR[k], d1, acc.gR);
        acc.gI = fma(ugI[k], d1, acc.gI);
        const double l = RADIAL ? d2 + f2 * d1 : d2;         // :373 (secondDerivativeFactor / rni * tmp1), :411
        acc.lR = fma(ulR[k], l, acc.lR);
        acc.lI = fma(ulI[k], l, acc.lI);
        const double v = w0 + w1 * x + w2 * x2 + w3 * x3;    // :240, :255 (computed by every lane: no divergent path)
        if (values) atomicAdd(&ext[k], v);
    }
}

__global__ void __launch_bounds__(kBrWarps * 32) evaluate_br_kernel(EvalArgs a)
{
    extern __shared__ __align__(16) double br_sm[];
    const SysDev& s = a.s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cfg = blockIdx.x * kBrWarps + warp;
    const int N = s.N, K = s.K, P = s.P, NE = 2 * K, NG = s.gr_bins;
    // block-shared tables (first capture: the L1 path of the global-memory tables was the limit, 86 %):
    // knots [K+4 (+pad)] | rec [nbins][18] | u~ drift R, I and plain R, I [4][2K] | lut; then per warp:
    // positions [3][Np] | ext [2K] | g(r) counts [NG] (as 32-bit integers)
    const int nk = (K + 4 + 1) & ~1;
    double* sk = br_sm;
    double* srec = sk + nk;
    double* su = srec + (size_t)s.nbins * kRecStride;
    unsigned short* slut = reinterpret_cast<unsigned short*>(su + 4 * NE);
    double* warp_base = su + 4 * NE + ((s.ncell + 3) / 4);
    for (int i = threadIdx.x; i < K + 4; i += blockDim.x) sk[i] = s.knots[i];
    for (int i = threadIdx.x; i < s.nbins * kRecStride; i += blockDim.x) srec[i] = s.rec[i];
    for (int i = threadIdx.x; i < NE; i += blockDim.x)
    {
        su[i] = s.ugR[i];
        su[NE + i] = s.ugI[i];
        su[2 * NE + i] = s.utR[i];
        su[3 * NE + i] = s.utI[i];
    }
    for (int i = threadIdx.x; i < s.ncell; i += blockDim.x) slut[i] = s.lut[i];
    const int per_warp = 3 * s.Np + NE + (NG + 1) / 2;
    double* px = warp_base + (size_t)warp * per_warp;
    double* py = px + s.Np;
    double* pz = py + s.Np;
    double* ext = pz + s.Np;
    unsigned* grc = reinterpret_cast<unsigned*>(ext + NE);
    __syncthreads();
    if (cfg >= a.n_cfg) return; // whole warps leave together; no block-wide barrier below

    const double* gpos = a.pos + (size_t)cfg * 3 * s.Np;
    for (int i = lane; i < N; i += 32)
    {
        px[i] = gpos[i];
        py[i] = gpos[s.Np + i];
        pz[i] = gpos[2 * s.Np + i];
    }
    for (int k = lane; k < NE; k += 32) ext[k] = 0.0;
    for (int k = lane; k < NG; k += 32) grc[k] = 0u;
    __syncwarp();

    BrEvalTables tb;
    tb.knots = sk;
    tb.rec = srec;
    tb.lut = slut;
    const double* ugRr = su;
    const double* ugIr = su + NE;
    const double* ulRr = su + 2 * NE;
    const double* ulIr = su + 3 * NE;
    double potential = 0.0, R1 = 0.0, I1 = 0.0, R1I1 = 0.0, R2 = 0.0, I2 = 0.0;
    for (int n = lane; n < N; n += 32)
    {
        const double xn = px[n], yn = py[n], zn = pz[n];
        double fRx = 0, fRy = 0, fRz = 0, fIx = 0, fIy = 0, fIz = 0, lR = 0, lI = 0;
        for (int i = 0; i < N; i++)
        {
            if (i == n) continue;
            double vx, vy, vz;
            const double r = disp_exact(s, xn, yn, zn, px[i], py[i], pz[i], vx, vy, vz);
            const bool lower = i < n;
            if (lower && r < s.gr_max) // :312-318
            {
                const double grBinInterval = r / s.gr_spacing;
                atomicAdd(&grc[(int)grBinInterval], 1u);
            }
            {
                // radial basis inside maxDistanceRad (:346-376), Gauss potential for i < n (:326-328).  Every lane runs
                // the same instructions; a lane whose pair is outside evaluates the last interval and is masked out.
                const bool inside = r < s.rmax;
                const double ra = r / s.pot_a;
                const double pot = s.pot_b * exp(-(ra * ra) / 2.0);
                if (lower && inside) potential += pot;
                BrAcc acc = { 0.0, 0.0, 0.0, 0.0 };
                br_pieces<true>(s, tb, inside ? r : s.rmax, s.dm1 / r, lower && inside, ext, ugRr, ugIr, ulRr, ulIr, acc);
                const double ex = vx / r, ey = vy / r, ez = vz / r; // :361-364
                const double gR = inside ? acc.gR : 0.0, gI = inside ? acc.gI : 0.0;
                fRx = fma(gR, ex, fRx); fRy = fma(gR, ey, fRy); fRz = fma(gR, ez, fRz);
                fIx = fma(gI, ex, fIx); fIy = fma(gI, ey, fIy); fIz = fma(gI, ez, fIz);
                lR += inside ? acc.lR : 0.0;
                lI += inside ? acc.lI : 0.0;
            }
            // box basis, one coordinate at a time (:378-416); sign = vecrni[a] < 0 ? -1 : 1
            {
                BrAcc acc = { 0.0, 0.0, 0.0, 0.0 };
                br_pieces<false>(s, tb, fabs(vx), 0.0, lower, ext + K, ugRr + K, ugIr + K, ulRr + K, ulIr + K, acc);
                const double sg = vx < 0 ? -1.0 : 1.0;
                fRx = fma(acc.gR, sg, fRx);
                fIx = fma(acc.gI, sg, fIx);
                lR += acc.lR;
                lI += acc.lI;
            }
            if (s.dim > 1) // (the reference loops a < DIM, :378)
            {
                BrAcc acc = { 0.0, 0.0, 0.0, 0.0 };
                br_pieces<false>(s, tb, fabs(vy), 0.0, lower, ext + K, ugRr + K, ugIr + K, ulRr + K, ulIr + K, acc);
                const double sg = vy < 0 ? -1.0 : 1.0;
                fRy = fma(acc.gR, sg, fRy);
                fIy = fma(acc.gI, sg, fIy);
                lR += acc.lR;
                lI += acc.lI;
            }
            if (s.dim > 2)
            {
                BrAcc acc = { 0.0, 0.0, 0.0, 0.0 };
                br_pieces<false>(s, tb, fabs(vz), 0.0, lower, ext + K, ugRr + K, ugIr + K, ulRr + K, ulIr + K, acc);
                const double sg = vz < 0 ? -1.0 : 1.0;
                fRz = fma(acc.gR, sg, fRz);
                fIz = fma(acc.gI, sg, fIz);
                lR += acc.lR;
                lI += acc.lI;
            }
        }
        R1I1 += 2.0 * (fRx * fIx + fRy * fIy + fRz * fIz); // :525-527
        R1 += fRx * fRx + fRy * fRy + fRz * fRz;
        I1 += fIx * fIx + fIy * fIy + fIz * fIz;
        R2 += lR;
        I2 += lI;
        if (a.drift_r)
        {
            double* d = a.drift_r + ((size_t)cfg * N + n) * 3;
            d[0] = fRx; d[1] = fRy; d[2] = fRz;
        }
        if (a.drift_i)
        {
            double* d = a.drift_i + ((size_t)cfg * N + n) * 3;
            d[0] = fIx; d[1] = fIy; d[2] = fIz;
        }
    }
    potential = warp_sum(potential);
    R1 = warp_sum(R1);
    I1 = warp_sum(I1);
    R1I1 = warp_sum(R1I1);
    R2 = warp_sum(R2);
    I2 = warp_sum(I2);
    __syncwarp(); // ext and the g(r) counts are complete

    const long long row = a.row0 + (long long)cfg * a.row_stride;
    double* Arow = a.A + (size_t)row * a.lda;
    double exponent = 0.0;
    for (int p = lane; p < P; p += 32) // RefreshLocalOperators :193-211 through the CSR map
    {
        double o = 0.0;
        for (int j = s.map_ptr[p]; j < s.map_ptr[p + 1]; j++) o += ext[s.map_col[j]] * s.map_val[j];
        Arow[p] = o;
        exponent = fma(s.uR[p], o, exponent); // :614-620
    }
    exponent = warp_sum(exponent);
    const double kineticR = -(R1 - I1 + R2); // :530-531
    const double kineticI = -(R1I1 + I2);
    double* o = a.other + (size_t)row * s.n_other;
    if (lane == 0)
    {
        Arow[P] = kineticR + potential;      // :535-536 (external and complex potentials are zero)
        Arow[P + 1] = kineticI;
        Arow[P + 2] = 1.0;
        o[0] = kineticR;                     // :574-576
        o[1] = potential;
        o[2] = exp(exponent + s.phiR);
        if (a.exponent) a.exponent[row] = exponent;
        if (a.outer_out) a.outer_out[cfg] = 0.0;
    }
    for (int b = lane; b < NG; b += 32) o[3 + b] = (double)grc[b] * (1.0 / s.gr_vol[b]); // grBins[b] += 1 / volume, :317
    if (a.ss_out)
        for (int k = lane; k < NE; k += 32) a.ss_out[(size_t)cfg * NE + k] = ext[k];
}

cudaError_t launch_evaluate_br(const EvalArgs& a, cudaStream_t st)
{
    if (a.n_cfg <= 0) return cudaSuccess;
    const size_t per_warp = (size_t)3 * a.s.Np + 2 * a.s.K + (a.s.gr_bins + 1) / 2;
    const size_t shared = (size_t)((a.s.K + 4 + 1) & ~1) + (size_t)a.s.nbins * kRecStride + 8 * (size_t)a.s.K + (a.s.ncell + 3) / 4;
    const size_t smem = (shared + per_warp * kBrWarps) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(evaluate_br_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    evaluate_br_kernel<<<(a.n_cfg + kBrWarps - 1) / kBrWarps, kBrWarps * 32, smem, st>>>(a);
    return cudaGetLastError();
}

// ---- sweep ----------------------------------------------------------------------------------------------------------
// u(x) on the interval that holds x: planes [c0,c1] | [c2,c3] | [t_lo,t_hi] of nrec double2 each (record nrec - 1 of the
// radial set is the zero tail beyond maxDistanceRad); lut gives the starting interval, the loop only walks upwards
template <int STRIDE>
__device__ __forceinline__ double br_cubic(const double2* __restrict__ c01p, const double2* __restrict__ c23p,
                                           const double2* __restrict__ ttp, const unsigned short* __restrict__ lut, int ncell,
                                           double inv_cell, int first_bin, double x)
{
    // floor(x * inv_cell) through the rounding constant (round-down FMA, integer in the low word): no F2I conversion,
    // which issues at a quarter of the FP64 rate (profiles/microbench/conv_throughput.cu)
    const int c = __double2loint(__fma_rd(x, inv_cell, kMagic));
    int j = (int)lut[max(0, min(c, ncell - 1))] - first_bin;
    double2 tt = ttp[j * STRIDE];
    while (x > tt.y)
    {
        j++;
        tt = ttp[j * STRIDE];
    }
    const double t = x - tt.x;
    const double2 c01 = c01p[j * STRIDE];
    const double2 c23 = c23p[j * STRIDE];
    return fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
}

struct BrTables
{
    const double2 *r01, *r23, *rtt, *b01, *b23, *btt;
    const unsigned short* lut;
    int ncell, first_bin;
    double inv_cell;
};

// pair term of the exponent between two points wrapped into the first cell; STRIDE = 8: shared-memory planes replicated
// eight times, copy c in 16-byte slot c of every 128-byte row, the pointers in t select this lane's copy (a random
// LDS.128 is served quarter-warp by quarter-warp: each quarter then covers the eight slots once - the cure of sweep.cu;
// first capture of this kernel: 87 % shared-memory wavefronts, 37 % of them bank conflicts)
template <int STRIDE>
__device__ __forceinline__ double br_pair_u(const BrTables& t, double dx, double dy, double dz, double Lhalf)
{
    const double mx = Lhalf - fabs(fabs(dx) - Lhalf); // |minimum-image component|
    const double my = Lhalf - fabs(fabs(dy) - Lhalf);
    const double mz = Lhalf - fabs(fabs(dz) - Lhalf);
    const double r = sqrt_fast(fma(mz, mz, fma(my, my, mx * mx)));
    double u = br_cubic<STRIDE>(t.r01, t.r23, t.rtt, t.lut, t.ncell, t.inv_cell, t.first_bin, r); // zero beyond maxDistanceRad
    u += br_cubic<STRIDE>(t.b01, t.b23, t.btt, t.lut, t.ncell, t.inv_cell, t.first_bin, mx);
    u += br_cubic<STRIDE>(t.b01, t.b23, t.btt, t.lut, t.ncell, t.inv_cell, t.first_bin, my);
    u += br_cubic<STRIDE>(t.b01, t.b23, t.btt, t.lut, t.ncell, t.inv_cell, t.first_bin, mz);
    return u;
}

__device__ __forceinline__ BrTables br_tables_global(const SysDev& s)
{
    const int nrec = s.nbins + 1;
    BrTables t;
    t.r01 = reinterpret_cast<const double2*>(s.cub);
    t.r23 = t.r01 + nrec;
    t.rtt = t.r23 + nrec;
    t.b01 = t.rtt + nrec;
    t.b23 = t.b01 + nrec;
    t.btt = t.b23 + nrec;
    t.lut = s.lut;
    t.ncell = s.ncell;
    t.first_bin = s.first_bin;
    t.inv_cell = s.inv_cell;
    return t;
}

constexpr int kBrCopies = 8;

__global__ void __launch_bounds__(256, 3) sweep_br_kernel(SweepArgs a)
{
    extern __shared__ __align__(16) unsigned char br_raw[];
    const SysDev& s = a.s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nrec = s.nbins + 1, Npp = a.npp;
    // shared memory: six planes of nrec x 8 copies double2 | lut | positions
    double2* planes = reinterpret_cast<double2*>(br_raw);
    unsigned short* lut = reinterpret_cast<unsigned short*>(planes + 6 * (size_t)nrec * kBrCopies);
    double* pos_base = reinterpret_cast<double*>(br_raw + a.pos_offset);
    {
        const double2* g = reinterpret_cast<const double2*>(s.cub);
        for (int i = threadIdx.x; i < 6 * nrec * kBrCopies; i += blockDim.x) planes[i] = g[i / kBrCopies];
        for (int i = threadIdx.x; i < s.ncell; i += blockDim.x) lut[i] = s.lut[i];
    }
    const int copy = lane & (kBrCopies - 1);
    const size_t plane = (size_t)nrec * kBrCopies;
    BrTables t;
    t.r01 = planes + copy;
    t.r23 = planes + plane + copy;
    t.rtt = planes + 2 * plane + copy;
    t.b01 = planes + 3 * plane + copy;
    t.b23 = planes + 4 * plane + copy;
    t.btt = planes + 5 * plane + copy;
    t.lut = lut;
    t.ncell = s.ncell;
    t.first_bin = s.first_bin;
    t.inv_cell = s.inv_cell;

    const int w = blockIdx.x * a.wpb + warp;
    double* px = pos_base + (size_t)warp * 3 * Npp;
    double* py = px + Npp;
    double* pz = py + Npp;
    const bool have = w < a.W;
    double* gpos = a.pos + (size_t)(have ? w : 0) * 3 * s.Np;
    const double L = s.L, Linv = s.Linv, Lhalf = s.Lhalf;
    const int N = s.N;
    for (int i = lane; i < N; i += 32) // positions live wrapped into the first cell during the sweep
    {
        px[i] = wrap_fast(gpos[i], L, Linv);
        py[i] = wrap_fast(gpos[s.Np + i], L, Linv);
        pz[i] = wrap_fast(gpos[2 * s.Np + i], L, Linv);
    }
    __syncthreads();
    if (!have) return;

    const uint32_t gw = (uint32_t)(a.first_walker + w);
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
            const double ddy_ = s.dim > 1 ? ddy : 0.0, ddz_ = s.dim > 2 ? ddz : 0.0; // DIM coordinates move; an unused
            // coordinate contributes the same box-spline constant u_box(0) to the old and the new exponent
            const double nx = wrap_fast(ox + ddx, L, Linv); // src/TDVMC.cpp:872-875, kept in the first cell
            const double ny = wrap_fast(oy + ddy_, L, Linv);
            const double nz = wrap_fast(oz + ddz_, L, Linv);
            double delta = 0.0;
            for (int i = lane; i < N; i += 32)
            {
                if (i == p) continue;
                const double xi = px[i], yi = py[i], zi = pz[i];
                delta += br_pair_u<kBrCopies>(t, xi - nx, yi - ny, zi - nz, Lhalf) - br_pair_u<kBrCopies>(t, xi - ox, yi - oy, zi - oz, Lhalf);
            }
            delta = warp_sum(delta);
            const double two_delta = 2.0 * delta; // quotient = exp(2 delta) finite and >= U (src/TDVMC.cpp:886-913)
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
        gpos[i] = px[i];
        gpos[s.Np + i] = py[i];
        gpos[2 * s.Np + i] = pz[i];
    }
    if (lane == 0) a.accepted[w] += n_acc;
}

static size_t sweep_br_smem(const SysDev& s, int wpb, int npp, size_t* pos_offset)
{
    size_t off = (size_t)6 * (s.nbins + 1) * kBrCopies * sizeof(double2) + (size_t)s.ncell * sizeof(unsigned short);
    off = (off + 15) & ~(size_t)15;
    *pos_offset = off;
    return off + (size_t)wpb * 3 * npp * sizeof(double);
}

cudaError_t launch_sweep_br(SweepArgs a, cudaStream_t st)
{
    size_t pos_off;
    const size_t smem = sweep_br_smem(a.s, a.wpb, a.npp, &pos_off);
    a.pos_offset = (int)pos_off;
    cudaError_t e = cudaFuncSetAttribute(sweep_br_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (a.W + a.wpb - 1) / a.wpb;
    sweep_br_kernel<<<grid, a.wpb * 32, smem, st>>>(a);
    return cudaGetLastError();
}

int sweep_br_fits(const SysDev& s, int wpb, int npp, int smem_optin)
{
    size_t pos_off;
    return sweep_br_smem(s, wpb, npp, &pos_off) <= (size_t)smem_optin ? 1 : 0;
}

// exponentNew - exponent for scripted moves of one configuration, one warp per move (tables in global memory)
__global__ void quotient_br_kernel(QuotientArgs a)
{
    const SysDev& s = a.s;
    const int lane = threadIdx.x & 31;
    const int mv = blockIdx.x;
    const BrTables t = br_tables_global(s);
    const double* px = a.pos;
    const double* py = a.pos + s.Np;
    const double* pz = a.pos + 2 * s.Np;
    const double L = s.L, Linv = s.Linv, Lhalf = s.Lhalf;
    const int p = (int)a.moves[mv * 4];
    const double nx = wrap_fast(a.moves[mv * 4 + 1], L, Linv), ny = wrap_fast(a.moves[mv * 4 + 2], L, Linv),
                 nz = wrap_fast(a.moves[mv * 4 + 3], L, Linv);
    const double ox = wrap_fast(px[p], L, Linv), oy = wrap_fast(py[p], L, Linv), oz = wrap_fast(pz[p], L, Linv);
    double delta = 0.0;
    for (int i = lane; i < s.N; i += 32)
    {
        if (i == p) continue;
        const double xi = wrap_fast(px[i], L, Linv), yi = wrap_fast(py[i], L, Linv), zi = wrap_fast(pz[i], L, Linv);
        delta += br_pair_u<1>(t, xi - nx, yi - ny, zi - nz, Lhalf) - br_pair_u<1>(t, xi - ox, yi - oy, zi - oz, Lhalf);
    }
    delta = warp_sum(delta);
    if (lane == 0) a.delta[mv] = delta;
}

cudaError_t launch_quotient_br(const QuotientArgs& a, cudaStream_t st)
{
    if (a.n_moves <= 0) return cudaSuccess;
    quotient_br_kernel<<<a.n_moves, 32, 0, st>>>(a);
    return cudaGetLastError();
}

} // namespace tdvmc
