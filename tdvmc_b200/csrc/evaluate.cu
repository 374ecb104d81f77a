// K2 + K3 + K4 fused — evaluate one configuration per thread block, nothing materialised.
//
// Replaces, per sample, CalculateLocalOperators / RefreshLocalOperators (BosonsBulk.cpp:158-218),
// CalculateOtherLocalOperators (:220-336) and CalculateExpectationValues (:349-458) (and their
// NUBosonsBulkPB twins, NUBosonsBulkPB.cpp:219-277, 279-414, 427-560).
//
// The reference first fills sD[k][n][a] and sD2[k][n] (2.2 MB per sample at N=343, K=203) and then
// contracts them with u.  Both steps are linear, so the contraction is done on the fly:
//     F_n   = sum_i  e_ni  * sum_p u~[bin-p] B'_{bin-p}(r_ni)
//     lap_n = sum_i          sum_p u~[bin-p] (B''_{bin-p}(r_ni) + (D-1)/r_ni B'_{bin-p}(r_ni))
// with u~ = M^T u (boundary-condition map applied to the parameters).  Per-term arithmetic keeps the
// reference's expressions and association (the monomial table in absolute r is ill-conditioned, so
// a different evaluation order would differ from the reference at the 1e-10 level); only the order
// of the sums differs.
//
// Every unordered pair is visited ONCE.  The pair matrix is cut into 32x32 tiles; in "shift" s, warp I
// takes tile (I, (I+s) mod NT), so the tiles of one shift have distinct row blocks and distinct column
// blocks.  Inside a tile lane l owns row particle 32I+l and meets column particle 32J+(l+k)%32 in step
// k: the row force stays in the lane's registers, the column force accumulators ROTATE through the warp
// by shuffle (lane l receives lane l+1's), so no lane ever adds into another lane's data and nothing
// needs an atomic.  Tiles are flushed into the block's force arrays columns first, rows second, a
// barrier after each.
//
// The basis sums ss[k] (needed for O_k) are a histogram over knot intervals; each warp keeps a private copy and adds to
// it without atomics: lanes whose pair falls into the same interval are ranked with match.any and take turns, so every
// round touches distinct addresses (f64 shared-memory atomics are CAS loops on sm_100a, 64-bit integer ones too).
//
// r02, what bounds the kernel and what was tried (profiles/r02_evaluate_*.txt): everything a step moves between lanes or
// to and from shared memory - table records, u~, the histogram read-modify-writes, positions, the twelve shuffles of the
// column rotation - goes through ONE 128 B/clk crossbar per SM; at N = 343 that is ~140 wavefronts per 32-pair step against
// 80 clocks of FP64 work.  (1) ROT = true: lane l starts with spline piece (l >> 1) & 3 and reads copy l & 1 of a table
// with 128-byte records, the second copy shifted by 16 bytes; the 16-byte bank group of every LDS.128 is then
// (2 piece + half + copy) mod 8 whatever the interval, i.e. conflict-free by construction (ncu: 3.75 wavefronts per
// LDS.128 = one per quarter-warp, against 10 with the 144-byte records).  Used where every lane of a step is inside the
// cut (the reflection rule of NUBosonsBulkPB); at BosonsBulk's cut r <= L/2 the selects that bring the values back
// to piece order for the histogram cost what the replays saved.  (2) A piece-parallel variant - pairs inside the cut
// compacted by ballot, four lanes per pair, one spline piece each, conflict-free table reads, results summed and handed
// back by shuffle - ran the FP64 work of a step in 137 instead of 160 instructions but needed 48 shuffles per step (each
// two crossbar wavefronts) and a match.any per round of eight pairs: 4.94 against 3.64 ms per 2960 configurations.  Not kept.
#include "kernels.cuh"

#include <cstdlib>

namespace tdvmc
{

struct SmemCarver
{
    unsigned char* base;
    size_t off;
    __host__ __device__ double* take(size_t n_doubles)
    {
        double* p = reinterpret_cast<double*>(base + off);
        off += ((n_doubles + 1) & ~(size_t)1) * sizeof(double);
        return p;
    }
};

// 1/r to full precision without the IEEE division sequence: hardware seed + two Newton steps.
// Enters only the well-conditioned factors (unit vector, (D-1)/r), never the spline argument.
__device__ __forceinline__ double rcp_refined(double r)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r));
    double e = fma(-r, y, 1.0);
    y = fma(y, e, y);
    e = fma(-r, y, 1.0);
    y = fma(y, e, y);
    return y;
}

struct EvalSmem2
{
    double* knots;
    double* wtab;  // two copies of [nbins][4 pieces][4 coefficients]; the second starts 16 bytes past a multiple of 128
    double2* ut;   // (u~R_k, u~I_k), `ucopies` interleaved replicas: entry k of replica c at ut[k * ucopies + c]
    int ucopies;   // 8 where it fits (lane l reads replica l & 7: the eight lanes of a quarter-warp never collide), else 1
    double* px;
    double* py;
    double* pz;
    double* hist;  // [nwarps][K]
    double* frc;   // [6][NT*32]
    double* sstot;
    double* red;
    unsigned short* lut;
};


__host__ __device__ inline size_t eval_smem_layout2(const SysDev& s, int nwarps, bool rot, bool gmem, EvalSmem2* out, unsigned char* base,
                                                    int ucopies = 1)
{
    SmemCarver c = { base, 0 };
    const int Npad = ((s.N + 31) >> 5) << 5;
    EvalSmem2 m;
    // rot: two copies of 128-byte records on 128-byte phases 0 and 16; else one copy of the 144-byte records
    m.wtab = c.take(rot ? (size_t)2 * s.nbins * 16 + 2 : (size_t)s.nbins * kRecStride);
    m.knots = c.take(s.K + 4);
    m.ucopies = ucopies;
    m.ut = reinterpret_cast<double2*>(c.take((size_t)2 * s.K * ucopies));
    // gmem: positions and forces of the configuration live in a per-block slab of global memory (L2) instead - systems too
    // large for one SM's shared memory (N = 8000 of config/BosonsBulk3D.config as shipped: 576 KB)
    m.px = gmem ? nullptr : c.take(Npad);
    m.py = gmem ? nullptr : c.take(Npad);
    m.pz = gmem ? nullptr : c.take(Npad);
    m.hist = c.take((size_t)nwarps * s.K);
    m.frc = gmem ? nullptr : c.take((size_t)6 * Npad);
    m.sstot = c.take(s.K);
    m.red = c.take((size_t)nwarps * 8);
    m.lut = reinterpret_cast<unsigned short*>(base + c.off);
    c.off += ((size_t)s.ncell * sizeof(unsigned short) + 15) & ~(size_t)15;
    if (out) *out = m;
    return c.off;
}

// Everything after the pair loop: Laplacian factor, |F|^2 sums, drift output, fixed-order block reduction, the histogram
// folded into O_k through the boundary-condition map, local energy and the otherExpectationValues row.
__device__ __forceinline__ void eval_finish(const EvalArgs& a, const SysDev& s, const EvalSmem2& m, int cfg, double lapR, double lapI,
                                            int vcount, int outer)
{
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int N = s.N, K = s.K, P = s.P;
    const int NP32 = ((N + 31) >> 5) * 32;
    const double* fRx = m.frc;
    const double* fRy = fRx + NP32;
    const double* fRz = fRy + NP32;
    const double* fIx = fRz + NP32;
    const double* fIy = fIx + NP32;
    const double* fIz = fIy + NP32;
    lapR *= 2.0; // each pair enters the Laplacian of both partners with the same value
    lapI *= 2.0;

    double R1 = 0.0, I1 = 0.0, RI = 0.0;
    for (int n = tid; n < N; n += blockDim.x)
    {
        const double ax = fRx[n], ay = fRy[n], az = fRz[n], bx = fIx[n], by = fIy[n], bz = fIz[n];
        R1 += ax * ax + ay * ay + az * az;          // VectorNorm2, BosonsBulk.cpp:418-419
        I1 += bx * bx + by * by + bz * bz;
        RI += ax * bx + ay * by + az * bz;          // kineticSumR1I1 / 2, BosonsBulk.cpp:417
        if (a.drift_r)
        {
            double* d = a.drift_r + ((size_t)cfg * N + n) * 3;
            d[0] = ax; d[1] = ay; d[2] = az;
        }
        if (a.drift_i)
        {
            double* d = a.drift_i + ((size_t)cfg * N + n) * 3;
            d[0] = bx; d[1] = by; d[2] = bz;
        }
    }

    // block reduction (fixed order -> deterministic)
    R1 = warp_sum(R1);
    I1 = warp_sum(I1);
    RI = warp_sum(RI);
    lapR = warp_sum(lapR);
    lapI = warp_sum(lapI);
    vcount = warp_sum_int(vcount);
    outer = warp_sum_int(outer);
    if (lane == 0)
    {
        double* r = m.red + warp * 8;
        r[0] = R1; r[1] = I1; r[2] = RI; r[3] = lapR; r[4] = lapI; r[5] = (double)vcount; r[6] = (double)outer;
    }
    __syncthreads();

    for (int k = tid; k < K; k += blockDim.x)
    {
        double t = 0.0;
        for (int w = 0; w < nwarps; w++) t += m.hist[(size_t)w * K + k];
        m.sstot[k] = t;
        if (a.ss_out) a.ss_out[(size_t)cfg * K + k] = t;
    }
    __syncthreads();

    const long long row = a.row0 + (long long)cfg * a.row_stride;
    double* Arow = a.A + (size_t)row * a.lda;
    double epart = 0.0;
    for (int p = tid; p < P; p += blockDim.x)
    {
        double o = 0.0;
        for (int j = s.map_ptr[p]; j < s.map_ptr[p + 1]; j++) o += s.map_val[j] * m.sstot[s.map_col[j]]; // BosonsBulk.cpp:158-177
        Arow[p] = o;
        epart = fma(s.uR[p], o, epart); // BosonsBulk.cpp:526-529
    }
    epart = warp_sum(epart);
    if (lane == 0) m.red[warp * 8 + 7] = epart;
    __syncthreads();

    if (tid == 0)
    {
        double t[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        for (int w = 0; w < nwarps; w++)
            for (int q = 0; q < 8; q++) t[q] += m.red[w * 8 + q];
        const double outer_sum = t[6];
        const double v_int = s.pot_b * t[5];
        const double exponent = t[7] + s.uR[s.tail_param] * outer_sum; // BosonsBulk.cpp:532-536
        const double kR1 = t[0], kI1 = t[1], kRI = 2.0 * t[2], kR2 = t[3], kI2 = t[4];
        const double kin_r = -(kR1 - kI1 + kR2) * s.hbar; // BosonsBulk.cpp:422
        const double kin_i = -(kRI + kI2) * s.hbar;       // BosonsBulk.cpp:423
        const double e_r = kin_r + v_int;                 // :425, external potential is zero (:344-347)
        const double e_i = kin_i;
        Arow[P] = e_r;
        Arow[P + 1] = e_i;
        Arow[P + 2] = 1.0;
        double* o = a.other + (size_t)row * s.n_other;   // BosonsBulk.cpp:449-457
        o[0] = kin_r;
        o[1] = v_int;
        o[2] = exp(exponent + s.phiR);
        o[3] = exponent;
        o[4] = kR1;
        o[5] = kI1;
        o[6] = kR2;
        o[7] = kI2;
        o[8] = kRI;
        if (a.exponent) a.exponent[row] = exponent;
        if (a.outer_out) a.outer_out[cfg] = outer_sum;
    }
}

#ifdef TDVMC_EVAL_KO
// measurement builds only (profiles/ab_evaluate_variants.py knockout): parts of a step switched off to read their cost
__device__ int g_eval_ko;
#define KO(bit) (ko & (bit))
#else
#define KO(bit) false
#endif

// UNI: interval index of a uniform knot vector without the look-up table and the two knot loads (find_bin_uniform):
// 4.99 -> 4.73 ms per 4096 configurations at N = 343, bit-identical results.
template <bool REFLECT, bool WIDE, bool ROT, bool GMEM = false, bool UNI = false>
__global__ void __launch_bounds__(WIDE ? 768 : 384, WIDE ? 1 : 2) evaluate_kernel(EvalArgs a)
{
#ifdef TDVMC_EVAL_KO
    const int ko = g_eval_ko;
#endif
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const SysDev& s = a.s;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int N = s.N, K = s.K;
    const int NT = (N + 31) >> 5; // tiles of 32 particles

    EvalSmem2 m;
    eval_smem_layout2(s, nwarps, ROT, GMEM, &m, smem_raw, a.ucopies);
    if (GMEM)
    {
        double* slab = a.scratch + (size_t)blockIdx.x * 9 * NT * 32; // [px | py | pz | 6 force rows], padded to whole tiles
        m.px = slab;
        m.py = slab + NT * 32;
        m.pz = slab + 2 * NT * 32;
        m.frc = slab + 3 * NT * 32;
    }

    const int ntab = s.nbins * 16;
    double* wtab1 = m.wtab + ntab + 2;
    if (ROT)
    {
        for (int i = tid; i < ntab; i += blockDim.x)
        {
            const double v = s.rec[(size_t)(i >> 4) * kRecStride + (i & 15)];
            m.wtab[i] = v;
            wtab1[i] = v;
        }
    }
    else
    {
        for (int i = tid; i < s.nbins * kRecStride; i += blockDim.x) m.wtab[i] = s.rec[i];
    }
    for (int i = tid; i < K + 4; i += blockDim.x) m.knots[i] = s.knots[i];
    for (int i = tid; i < K; i += blockDim.x)
    {
        const double2 u = make_double2(s.utR[i], s.utI[i]);
        for (int c = 0; c < m.ucopies; c++) m.ut[(size_t)i * m.ucopies + c] = u;
    }
    for (int i = tid; i < s.ncell; i += blockDim.x) m.lut[i] = s.lut[i];
    // GMEM: one block per SM walks the configurations (its slab is reused); otherwise one block per configuration
    int cfg = blockIdx.x;
    do
    {
    const double* gpos = a.pos + (size_t)cfg * 3 * s.Np;
    for (int i = tid; i < NT * 32; i += blockDim.x)
    {
        const bool v = i < N;
        m.px[i] = v ? gpos[i] : 0.0;
        m.py[i] = v ? gpos[s.Np + i] : 0.0;
        m.pz[i] = v ? gpos[2 * s.Np + i] : 0.0;
    }
    for (int i = tid; i < 6 * NT * 32; i += blockDim.x) m.frc[i] = 0.0;
    for (int i = tid; i < nwarps * K; i += blockDim.x) m.hist[i] = 0.0;
    __syncthreads();

    double* hist = m.hist + (size_t)warp * K;
    // lane l starts with spline piece (l >> 1) & 3 and reads table copy l & 1: with 128-byte records the 16-byte bank
    // group of a coefficient pair is (2 piece + half + copy) mod 8 whatever the interval, so the eight lanes of a
    // quarter-warp hit eight distinct groups in every one of the eight LDS.128 of a step
    const int p0 = ROT ? (lane >> 1) & 3 : 0;
    const int wstride = ROT ? 16 : kRecStride;
    const double* wmine = ((ROT && (lane & 1)) ? wtab1 : m.wtab) - (size_t)s.first_bin * wstride;
    const int ucop = m.ucopies;
    const double2* umine = m.ut + (lane & (ucop - 1));
    const double rmax = s.rmax;
    const double pot_a = s.pot_a;
    const int NP32 = NT * 32;
    double* fRx = m.frc;
    double* fRy = fRx + NP32;
    double* fRz = fRy + NP32;
    double* fIx = fRz + NP32;
    double* fIy = fIx + NP32;
    double* fIz = fIy + NP32;

    double lapR = 0.0, lapI = 0.0;   // sum u~ B'' (piece lanes) + (D-1)/r sum u~ B' (owner lanes)
    int vcount = 0, outer = 0;

    const int half = NT >> 1;
    const bool nt_even = (NT & 1) == 0;
    for (int sft = 0; sft <= half; sft++)
    {
        // NT even: the shift NT/2 pairs block I with I+NT/2 from both sides; warps I and I+NT/2 share that tile
        // (rotations 0..15 and 16..31) instead of one of them idling
        const bool shared_shift = nt_even && sft == half && sft != 0;
        for (int base = 0; base < NT; base += nwarps)
        {
            const int Iw = base + warp;
            const bool tile = Iw < NT;
            const bool second = shared_shift && Iw >= half;
            const int I = second ? Iw - half : Iw;
            int J = I + sft;
            if (J >= NT) J -= NT;
            // diagonal tiles: rotations 1..16 (16 by the lower half only) cover each pair of the tile once
            const int kfirst = sft == 0 ? 1 : (second ? 16 : 0);
            const int klast = sft == 0 ? 16 : (shared_shift && !second ? 15 : 31);
            const int n = 32 * I + lane;
            double rRx = 0.0, rRy = 0.0, rRz = 0.0, rIx = 0.0, rIy = 0.0, rIz = 0.0; // force on the row particle
            double cRx = 0.0, cRy = 0.0, cRz = 0.0, cIx = 0.0, cIy = 0.0, cIz = 0.0; // rotating column accumulators
            if (tile)
            {
                const bool vn = n < N;
                const double xn = m.px[n], yn = m.py[n], zn = m.pz[n];
                for (int k = kfirst; k <= klast; k++)
                {
                    if (k > kfirst && !KO(16))
                    {
                        const int src = (lane + 1) & 31;
                        cRx = __shfl_sync(FULL_MASK, cRx, src);
                        cRy = __shfl_sync(FULL_MASK, cRy, src);
                        cRz = __shfl_sync(FULL_MASK, cRz, src);
                        cIx = __shfl_sync(FULL_MASK, cIx, src);
                        cIy = __shfl_sync(FULL_MASK, cIy, src);
                        cIz = __shfl_sync(FULL_MASK, cIz, src);
                    }
                    const int i = 32 * J + ((lane + k) & 31);
                    double vx, vy, vz;
                    double r = disp_exact(s, xn, yn, zn, m.px[i], m.py[i], m.pz[i], vx, vy, vz); // R[n] - R[i]
                    const bool pair = vn && (i < N) && (sft != 0 || k < 16 || lane < 16);
                    bool inside;
                    if (REFLECT)
                    {
                        if (!(r < rmax)) r = 2 * rmax - r; // NUBosonsBulkPB.cpp:250-253, 331-335
                        inside = r < rmax;
                    }
                    else
                    {
                        inside = r <= rmax; // BosonsBulk.cpp:195, 257
                    }
                    const bool act = pair && inside;
                    if (pair && !inside) outer++;      // BosonsBulk.cpp:207-210
                    if (act && (r < pot_a)) vcount++;  // BosonsBulk.cpp:268-271

                    int bin = 0;
                    double val[4] = { 0.0, 0.0, 0.0, 0.0 };
                    // the interval first, and the match.any that ranks the lanes sharing one (warp_hist_add4_ranked) right away:
                    // its latency - 10 % of the kernel's stall samples when it sat in front of the histogram update -
                    // passes behind the spline arithmetic (4.73 -> 4.65 ms per 4096 configurations, bit-identical)
                    if (act) bin = UNI ? find_bin_uniform(s, m.knots, m.lut, r) : find_bin_exact(s, m.knots, m.lut, r);
                    const unsigned amask = __ballot_sync(FULL_MASK, act);
                    unsigned peers = 0u;
                    if (amask != 0u) peers = __match_any_sync(FULL_MASK, act ? bin : (-1 - lane)); // (amask is warp-uniform)
                    if (act)
                    {
                        const double* w = wmine + (size_t)(KO(4) ? s.first_bin : bin) * wstride;
                        const double r2 = r * r;
                        const double rinv = rcp_refined(r);
                        const double f2 = s.dm1 * rinv; // (DIM - 1) / r, BosonsBulk.cpp:319-322
                        double gR = 0.0, gI = 0.0;
                        double vq[4];
#pragma unroll
                        for (int q = 0; q < 4; q++)
                        {
                            const int p = ROT ? (p0 + q) & 3 : q;
                            const double2 w01 = *reinterpret_cast<const double2*>(w + p * 4);
                            const double2 w23 = *reinterpret_cast<const double2*>(w + p * 4 + 2);
                            const double d1 = w01.y + 2.0 * w23.x * r + 3.0 * w23.y * r2; // BosonsBulk.cpp:299
                            const double d2 = 2.0 * w23.x + 6.0 * w23.y * r;              // BosonsBulk.cpp:301
                            const double2 uk = KO(2) ? make_double2(0.5 + p, 0.25) : umine[(bin - p) * ucop];
                            const double uRk = uk.x, uIk = uk.y;
                            const double t2 = d2 + f2 * d1;
                            gR = fma(uRk, d1, gR);
                            gI = fma(uIk, d1, gI);
                            lapR = fma(uRk, t2, lapR);
                            lapI = fma(uIk, t2, lapI);
                            vq[q] = w01.x + w01.y * r + w23.x * r2 + w23.y * (r2 * r); // BosonsBulk.cpp:204
                        }
                        // back to piece order for the histogram: val[p] = vq[(p - p0) & 3]
                        if (ROT)
                        {
                            const bool s1 = p0 & 1, s2 = p0 & 2;
                            const double a0 = s1 ? vq[3] : vq[0], a1 = s1 ? vq[0] : vq[1], a2 = s1 ? vq[1] : vq[2], a3 = s1 ? vq[2] : vq[3];
                            val[0] = s2 ? a2 : a0;
                            val[1] = s2 ? a3 : a1;
                            val[2] = s2 ? a0 : a2;
                            val[3] = s2 ? a1 : a3;
                        }
                        else
                        {
                            val[0] = vq[0];
                            val[1] = vq[1];
                            val[2] = vq[2];
                            val[3] = vq[3];
                        }
                        const double ex = vx * rinv, ey = vy * rinv, ez = vz * rinv; // unreflected vec / (reflected) r
                        rRx = fma(gR, ex, rRx);
                        rRy = fma(gR, ey, rRy);
                        rRz = fma(gR, ez, rRz);
                        rIx = fma(gI, ex, rIx);
                        rIy = fma(gI, ey, rIy);
                        rIz = fma(gI, ez, rIz);
                        cRx = fma(-gR, ex, cRx); // the partner sees the opposite unit vector
                        cRy = fma(-gR, ey, cRy);
                        cRz = fma(-gR, ez, cRz);
                        cIx = fma(-gI, ex, cIx);
                        cIy = fma(-gI, ey, cIy);
                        cIz = fma(-gI, ez, cIz);
                    }
                    if (!KO(1)) warp_hist_add4_ranked(hist, bin, act, val, lane, amask, peers);
                }
            }
            // flush: columns first, rows second; in the shared shift the two warps of a tile take turns
            const int ic = 32 * J + ((lane + klast) & 31); // the column whose accumulator ended up in this lane
            for (int role = 0; role < (shared_shift ? 2 : 1); role++)
            {
                if (tile && (int)second == role)
                {
                    fRx[ic] += cRx;
                    fRy[ic] += cRy;
                    fRz[ic] += cRz;
                    fIx[ic] += cIx;
                    fIy[ic] += cIy;
                    fIz[ic] += cIz;
                }
                __syncthreads();
            }
            for (int role = 0; role < (shared_shift ? 2 : 1); role++)
            {
                if (tile && (int)second == role)
                {
                    fRx[n] += rRx;
                    fRy[n] += rRy;
                    fRz[n] += rRz;
                    fIx[n] += rIx;
                    fIy[n] += rIy;
                    fIz[n] += rIz;
                }
                __syncthreads();
            }
        }
    }
    eval_finish(a, s, m, cfg, lapR, lapI, vcount, outer);
    if (!GMEM) break;  // (compile-time: the one-block-per-configuration kernels have no loop)
    __syncthreads();   // the slab and the histograms are reused by the next configuration
    cfg += gridDim.x;
    } while (cfg < a.n_cfg);
}

// Rotated piece order (two table copies, conflict-free LDS.128) where nearly every lane of a step is inside the cut - the
// reflection rule of NUBosonsBulkPB: 22.8 against 25.1 ms per 512 configurations at N = 1728; at BosonsBulk's cut r <= L/2
// half the lanes idle, the records collide less, and the extra selects cost more than the replays saved (3.80 against
// 3.64 ms per 2960 configurations at N = 343).
constexpr size_t kEvalSmemLimit = (size_t)227 * 1024; // opt-in shared memory per block on sm_100a

static int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// tuning knob (profiles/ab_evaluate_variants.py, tests): TDVMC_EVAL_UNIBIN = 0 takes the exact search everywhere
constexpr int kEvalUniBinDefault = 1;

static bool evaluate_rotated(const SysDev& s) { return s.pair_rule == 1; }

// positions + forces in global memory: when even one 24-warp block per SM does not fit shared memory
static bool evaluate_gmem(const SysDev& s)
{
    return eval_smem_layout2(s, 24, evaluate_rotated(s), false, nullptr, nullptr) + 1024 > kEvalSmemLimit;
}

static size_t eval_layout_bytes(const SysDev& s, int nwarps)
{
    return eval_smem_layout2(s, nwarps, evaluate_rotated(s), evaluate_gmem(s), nullptr, nullptr);
}

// two blocks of <= 12 warps per SM when they fit, else one block of <= 24 warps
static bool evaluate_wide(const SysDev& s)
{
    if (evaluate_gmem(s)) return true;
    int nt = (s.N + 31) / 32;
    if (nt <= 12) return false;
    return 2 * (eval_layout_bytes(s, 12) + 1024) > kEvalSmemLimit;
}

int evaluate_blocks_per_sm(const SysDev& s) { return evaluate_wide(s) ? 1 : 2; }

// Warps per block: one tile per warp and round, so ceil(NT / warps) rounds per shift with the last one partly idle (NT = 54
// tiles at N = 1728 on 24 warps: rounds of 24, 24 and 6 tiles, a quarter of the warp-time waiting at the flush barriers).  Take
// the warp count in the upper third of the allowed range that wastes least (N = 1728: 18 warps, three full rounds).
int evaluate_threads(const SysDev& s)
{
    const int nt = (s.N + 31) / 32;
    const int cap = evaluate_wide(s) ? 24 : 12;
    if (nt <= cap) return nt * 32;
    int best = cap, best_cost = ((nt + cap - 1) / cap) * cap;
    if (const char* e = getenv("TDVMC_EVAL_WARPS")) // tuning knob
    {
        const int w = atoi(e);
        if (w >= 1 && w <= cap) return w * 32;
    }
    for (int w = cap - 1; w >= (2 * cap) / 3; w--)
    {
        const int cost = ((nt + w - 1) / w) * w; // warp-rounds per shift
        if (cost < best_cost)
        {
            best_cost = cost;
            best = w;
        }
    }
    return best * 32;
}

// Replicas of the u~ table: eight (lane l reads replica l & 7, so the lanes of a quarter-warp never collide in an LDS.128)
// when the block still fits - twice per SM for the 12-warp blocks - else one
static int evaluate_ucopies(const SysDev& s)
{
    const int want = env_int("TDVMC_EVAL_UCOPIES", 8) >= 8 ? 8 : 1; // tuning knob
    if (want == 1) return 1;
    const size_t bytes = eval_smem_layout2(s, evaluate_threads(s) / 32, evaluate_rotated(s), evaluate_gmem(s), nullptr, nullptr, 8) + 1024;
    return evaluate_blocks_per_sm(s) * bytes <= kEvalSmemLimit ? 8 : 1;
}

size_t evaluate_smem_bytes(const SysDev& s)
{
    return eval_smem_layout2(s, evaluate_threads(s) / 32, evaluate_rotated(s), evaluate_gmem(s), nullptr, nullptr, evaluate_ucopies(s));
}

size_t evaluate_scratch_doubles(const SysDev& s, int sm_count)
{
    if (!evaluate_gmem(s)) return 0;
    return (size_t)sm_count * 9 * (size_t)(((s.N + 31) / 32) * 32);
}

template <typename KernelT>
static cudaError_t launch_eval_kernel(KernelT kernel, const EvalArgs& a, int threads, size_t smem, int grid, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    kernel<<<grid, threads, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_evaluate(const EvalArgs& a_in, cudaStream_t st)
{
    if (a_in.n_cfg <= 0) return cudaSuccess;
    EvalArgs a = a_in;
    a.ucopies = evaluate_ucopies(a.s);
    const int threads = evaluate_threads(a.s);
    const size_t smem = evaluate_smem_bytes(a.s);
    const bool refl = a.s.pair_rule == 1;
    // uniform knots (BosonsBulk's own grid; config/NUBosonsBulkPB3D.config's NURBS_GRID is one too): the interval index needs no table
    const bool uni = a.s.uniform && a.s.bin_guard > 0.0 && env_int("TDVMC_EVAL_UNIBIN", kEvalUniBinDefault) != 0;
#ifdef TDVMC_EVAL_KO
    {
        const int ko = env_int("TDVMC_EVAL_KO", 0);
        cudaMemcpyToSymbolAsync(g_eval_ko, &ko, sizeof(int), 0, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
    }
#endif
    if (evaluate_gmem(a.s))
    {
        if (!a.scratch) return cudaErrorInvalidValue;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int grid = a.n_cfg < sms ? a.n_cfg : sms; // one block per SM walks the configurations
        if (refl)
            return uni ? launch_eval_kernel(evaluate_kernel<true, true, true, true, true>, a, threads, smem, grid, st)
                       : launch_eval_kernel(evaluate_kernel<true, true, true, true>, a, threads, smem, grid, st);
        return uni ? launch_eval_kernel(evaluate_kernel<false, true, false, true, true>, a, threads, smem, grid, st)
                   : launch_eval_kernel(evaluate_kernel<false, true, false, true>, a, threads, smem, grid, st);
    }
    if (evaluate_wide(a.s))
    {
        if (refl)
            return uni ? launch_eval_kernel(evaluate_kernel<true, true, true, false, true>, a, threads, smem, a.n_cfg, st)
                       : launch_eval_kernel(evaluate_kernel<true, true, true>, a, threads, smem, a.n_cfg, st);
        return uni ? launch_eval_kernel(evaluate_kernel<false, true, false, false, true>, a, threads, smem, a.n_cfg, st)
                   : launch_eval_kernel(evaluate_kernel<false, true, false>, a, threads, smem, a.n_cfg, st);
    }
    if (refl)
        return uni ? launch_eval_kernel(evaluate_kernel<true, false, true, false, true>, a, threads, smem, a.n_cfg, st)
                   : launch_eval_kernel(evaluate_kernel<true, false, true>, a, threads, smem, a.n_cfg, st);
    return uni ? launch_eval_kernel(evaluate_kernel<false, false, false, false, true>, a, threads, smem, a.n_cfg, st)
               : launch_eval_kernel(evaluate_kernel<false, false, false>, a, threads, smem, a.n_cfg, st);
}

} // namespace tdvmc
