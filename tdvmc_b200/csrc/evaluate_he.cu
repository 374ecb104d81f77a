// Fused evaluation for the He family: HeBulk (src/PhysicalSystems/HeBulk.cpp) and HeDrop
// (src/PhysicalSystems/HeDrop.cpp).  McMillan core r^m below rijSplit, uniform cubic B-splines written in
// the local coordinate res = (r - r0)/h - bin on one (HeBulk) or two (HeDrop: h = 0.1 then 0.5) grids,
// constant + linear tails beyond rijTail (HeDrop), Aziz HFD-B(He) or Lennard-Jones potential inline,
// g(r) -- and for HeDrop the density profile around the centre of mass -- carried in otherExpectationValues.
//
// Same structure as evaluate.cu (one block per configuration, thread n owns particle n, contraction with
// u~ = M^T u on the fly, per-warp conflict-free histogram for the value sums).  The per-pair expressions are
// the reference's (HeBulk.cpp:268-302, 471-486; HeDrop.cpp:399-468, 728-761), so each term agrees with the
// reference to the last bits (libm pow/exp differ from glibc by <= 1-2 ulp).
// Extended basis sums: ext = [ss_0 .. ss_{K-1} | mcMillanSum | constSum | linearSum].
#include "kernels.cuh"

#include <cstdlib>

namespace tdvmc
{

__device__ __forceinline__ double he_pair_potential(const SysDev& s, double r)
{
    if (s.potential == 0)
    {
        // Aziz HFD-B(He), HeBulk.cpp:187-195, 251-261
        const double e = 10.948, rm = 2.963, aa = 184431.01, alpha = 10.43329537, beta = -2.27965105, dd = 1.4826,
                     c6 = 1.36745214, c8 = 0.42123807, c10 = 0.17473318;
        const double x = r / rm;
        const double x2 = x * x;
        const double xm2 = 1.0 / x2;
        const double xm6 = xm2 * xm2 * xm2;
        double F = 1;
        if (x < dd)
        {
            const double q = dd / x - 1;
            F = exp(-(q * q));
        }
        return e * (aa * exp(-alpha * x + beta * x2) - F * xm6 * (c6 + xm2 * (c8 + xm2 * c10)));
    }
    // Lennard-Jones, HeDrop.cpp:389-394
    const double sigma = 4.0, eps = 3.56;
    const double q = sigma / r;
    const double q2 = q * q;
    const double s6 = q2 * q2 * q2;
    return 4.0 * eps * s6 * (s6 - 1.0);
}

__global__ void __launch_bounds__(256, 4) evaluate_he_kernel(EvalArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SysDev& s = a.s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int cfg = blockIdx.x;
    const int N = s.N, K = s.K, P = s.P, G = s.gr_bins, NR = s.rho_bins, NE = s.n_ext;
    const int MC = K, CO = K + 1, LI = K + 2;

    double* utR = reinterpret_cast<double*>(smem_raw);
    double* utI = utR + NE;
    double* px = utI + NE;
    double* py = px + N;
    double* pz = py + N;
    double* hist = pz + N;                    // [nwarps][K]
    double* ext = hist + (size_t)nwarps * K;  // [NE]
    double* red = ext + NE;                   // [nwarps][12]
    double* com = red + (size_t)nwarps * 12;  // [4]
    int* gr = reinterpret_cast<int*>(com + 4); // [G + NR]

    for (int i = tid; i < NE; i += blockDim.x)
    {
        utR[i] = s.utR[i];
        utI[i] = s.utI[i];
    }
    const double* gpos = a.pos + (size_t)cfg * 3 * s.Np;
    for (int i = tid; i < N; i += blockDim.x)
    {
        px[i] = gpos[i];
        py[i] = gpos[s.Np + i];
        pz[i] = gpos[2 * s.Np + i];
    }
    for (int i = tid; i < nwarps * K; i += blockDim.x) hist[i] = 0.0;
    for (int i = tid; i < G + NR; i += blockDim.x) gr[i] = 0;
    __syncthreads();
    if (tid == 0 && NR > 0) // GetCenterOfMass, HeDrop.cpp:255-270
    {
        double cx = 0, cy = 0, cz = 0;
        for (int i = 0; i < N; i++)
        {
            cx += px[i];
            cy += py[i];
            cz += pz[i];
        }
        com[0] = cx / (double)N;
        com[1] = cy / (double)N;
        com[2] = cz / (double)N;
    }
    __syncthreads();

    double* myhist = hist + (size_t)warp * K;
    const double rs = s.r0, rmax = s.rmax, r2s = s.r_split2, rt = s.r_tail, m = s.core_m;
    const double gr_spacing = s.gr_max / (double)G;
    const double ucR = utR[MC], ucI = utI[MC], ulR = utR[LI], ulI = utI[LI];

    double R1 = 0.0, I1 = 0.0, RI = 0.0, lapR = 0.0, lapI = 0.0, pot = 0.0, mcm = 0.0, csum = 0.0, lsum = 0.0;

    for (int n0 = warp * 32; n0 < N; n0 += blockDim.x)
    {
        const int n = n0 + lane;
        const bool valid = n < N;
        const double xn = valid ? px[n] : 0.0, yn = valid ? py[n] : 0.0, zn = valid ? pz[n] : 0.0;
        double fRx = 0.0, fRy = 0.0, fRz = 0.0, fIx = 0.0, fIy = 0.0, fIz = 0.0;
        for (int i = 0; i < N; i++)
        {
            double vx, vy, vz, r;
            if (s.periodic)
            {
                r = disp_exact(s, xn, yn, zn, px[i], py[i], pz[i], vx, vy, vz);
            }
            else
            {
                vx = xn - px[i]; // VectorDisplacement, Utils.cpp:253-263
                vy = yn - py[i];
                vz = zn - pz[i];
                r = sqrt(vx * vx + vy * vy + vz * vz);
            }
            const bool act = valid && (i != n) && (r < rmax); // HeBulk.cpp:232 (HeDrop: no cut)
            const bool lower = act && (i < n);
            int bin = 0;
            double val[4] = { 0.0, 0.0, 0.0, 0.0 };
            bool spline_val = false;
            if (lower) pot += he_pair_potential(s, r);
            if (valid && (i < n) && (r < s.gr_max)) atomicAdd(&gr[min((int)floor(r / gr_spacing), G - 1)], 1); // g(r) counts
            if (act)
            {
                if (r < rs)
                {
                    // McMillan core, HeBulk.cpp:268-276 (m = -5), HeDrop.cpp:402-410
                    const double rp = pow(r, m - 2.0);
                    const double g = m * rp;
                    const double l2 = m * (m + 1.0) * rp;
                    fRx = fma(ucR * g, vx, fRx); fRy = fma(ucR * g, vy, fRy); fRz = fma(ucR * g, vz, fRz);
                    fIx = fma(ucI * g, vx, fIx); fIy = fma(ucI * g, vy, fIy); fIz = fma(ucI * g, vz, fIz);
                    lapR = fma(ucR, l2, lapR);
                    lapI = fma(ucI, l2, lapI);
                    if (lower) mcm += pow(r, m);
                }
                else if (r >= rt)
                {
                    // constant + linear tails, HeDrop.cpp:411-426, 732-737
                    const double ex = vx / r, ey = vy / r, ez = vz / r;
                    fRx = fma(ulR, ex, fRx); fRy = fma(ulR, ey, fRy); fRz = fma(ulR, ez, fRz);
                    fIx = fma(ulI, ex, fIx); fIy = fma(ulI, ey, fIy); fIz = fma(ulI, ez, fIz);
                    lapR = fma(ulR, 2.0 / r, lapR);
                    lapI = fma(ulI, 2.0 / r, lapI);
                    if (lower)
                    {
                        csum += 1.0;
                        lsum += r;
                    }
                }
                else
                {
                    double interval, nps; // HeBulk.cpp:279-282, HeDrop.cpp:431-446
                    if (r < r2s)
                    {
                        nps = s.h;
                        interval = (r - rs) / nps;
                        bin = (int)floor(interval);
                    }
                    else
                    {
                        nps = s.h_large;
                        interval = (r - r2s) / nps;
                        bin = (int)floor(interval) + s.n_short;
                    }
                    const double res = interval - floor(interval);
                    const double res2 = res * res;
                    const double nps2 = nps * nps;
                    double tmp[4], l2[4];
                    tmp[0] = -1.0 / 2.0 * (1.0 - 2.0 * res + res2); // HeBulk.cpp:284-287
                    tmp[1] = 1.0 / 6.0 * (-12.0 * res + 9.0 * res2);
                    tmp[2] = 1.0 / 6.0 * (3.0 + 6.0 * res - 9.0 * res2);
                    tmp[3] = 1.0 / 2.0 * res2;
                    const double f2 = 2.0 / (nps * r);
                    l2[0] = 1.0 / nps2 * (1.0 - res) + f2 * tmp[0]; // HeBulk.cpp:299-302
                    l2[1] = 1.0 / nps2 * (1.0 / 6.0 * (-12.0 + 18.0 * res)) + f2 * tmp[1];
                    l2[2] = 1.0 / nps2 * (1.0 / 6.0 * (6.0 - 18.0 * res)) + f2 * tmp[2];
                    l2[3] = 1.0 / nps2 * (res) + f2 * tmp[3];
                    const double ex = vx / r, ey = vy / r, ez = vz / r;
                    double gR = 0.0, gI = 0.0;
#pragma unroll
                    for (int b = 0; b < 4; b++)
                    {
                        const double uRk = utR[bin + b], uIk = utI[bin + b];
                        gR = fma(uRk, tmp[b], gR);
                        gI = fma(uIk, tmp[b], gI);
                        lapR = fma(uRk, l2[b], lapR);
                        lapI = fma(uIk, l2[b], lapI);
                    }
                    gR = gR / nps;
                    gI = gI / nps;
                    fRx = fma(gR, ex, fRx); fRy = fma(gR, ey, fRy); fRz = fma(gR, ez, fRz);
                    fIx = fma(gI, ex, fIx); fIy = fma(gI, ey, fIy); fIz = fma(gI, ez, fIz);
                    if (lower)
                    {
                        const double res3 = res2 * res; // pow(res, 3)
                        // values for splines bin .. bin+3 (HeBulk.cpp:483-486), stored for hist[bin + 3 - p]
                        val[3] = -1.0 / 6.0 * (-1.0 + 3.0 * res - 3.0 * res2 + res3);
                        val[2] = 1.0 / 6.0 * (4.0 - 6.0 * res2 + 3.0 * res3);
                        val[1] = 1.0 / 6.0 * (1.0 + 3.0 * res + 3.0 * res2 - 3.0 * res3);
                        val[0] = 1.0 / 6.0 * res3;
                        spline_val = true;
                    }
                }
            }
            warp_hist_add4(myhist, bin + 3, spline_val, val, lane);
        }
        if (valid)
        {
            if (NR > 0) // density profile around the centre of mass, HeDrop.cpp:482-491
            {
                const double d0 = xn - com[0], d1 = yn - com[1], d2 = zn - com[2];
                const double rr = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                if (rr < s.gr_max) atomicAdd(&gr[G + min((int)floor(rr / gr_spacing), NR - 1)], 1);
            }
            fRx += s.g0R; fRy += s.g0R; fRz += s.g0R; // the literal 1 of the last parameter, HeBulk.cpp:351
            fIx += s.g0I; fIy += s.g0I; fIz += s.g0I;
            R1 += fRx * fRx + fRy * fRy + fRz * fRz;
            I1 += fIx * fIx + fIy * fIy + fIz * fIz;
            RI += fRx * fIx + fRy * fIy + fRz * fIz;
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
    }

    R1 = warp_sum(R1); I1 = warp_sum(I1); RI = warp_sum(RI);
    lapR = warp_sum(lapR); lapI = warp_sum(lapI); pot = warp_sum(pot);
    mcm = warp_sum(mcm); csum = warp_sum(csum); lsum = warp_sum(lsum);
    if (lane == 0)
    {
        double* r = red + warp * 12;
        r[0] = R1; r[1] = I1; r[2] = RI; r[3] = lapR; r[4] = lapI; r[5] = pot; r[6] = mcm; r[7] = csum; r[8] = lsum;
    }
    __syncthreads();
    for (int k = tid; k < K; k += blockDim.x)
    {
        double t = 0.0;
        for (int w = 0; w < nwarps; w++) t += hist[(size_t)w * K + k];
        ext[k] = t;
    }
    if (tid == 0)
    {
        double t6 = 0.0, t7 = 0.0, t8 = 0.0;
        for (int w = 0; w < nwarps; w++)
        {
            t6 += red[w * 12 + 6];
            t7 += red[w * 12 + 7];
            t8 += red[w * 12 + 8];
        }
        ext[MC] = t6;
        ext[CO] = t7;
        ext[LI] = t8;
    }
    __syncthreads();
    if (a.ss_out)
        for (int k = tid; k < NE; k += blockDim.x) a.ss_out[(size_t)cfg * NE + k] = ext[k];

    const long long row = a.row0 + (long long)cfg * a.row_stride;
    double* Arow = a.A + (size_t)row * a.lda;
    double epart = 0.0;
    for (int p = tid; p < P; p += blockDim.x)
    {
        double o = s.map_const[p]; // HeBulk.cpp:376-383, HeDrop.cpp:609-626
        for (int j = s.map_ptr[p]; j < s.map_ptr[p + 1]; j++) o += s.map_val[j] * ext[s.map_col[j]];
        Arow[p] = o;
        epart = fma(s.uR[p], o, epart); // HeBulk.cpp:491-498
    }
    epart = warp_sum(epart);
    if (lane == 0) red[warp * 12 + 9] = epart;
    __syncthreads();
    double* orow = a.other + (size_t)row * s.n_other;
    for (int b = tid; b < G + NR; b += blockDim.x)
    {
        // 1 / grBinVolumes[b], HeBulk.cpp:133-143 (the density profile uses the same shell volumes, HeDrop.cpp:490)
        const int bb = b < G ? b : b - G;
        const double r1 = gr_spacing * (bb + 1), r0 = gr_spacing * bb;
        double vol = 4.0 * M_PI * (r1 * r1 * r1) / 3.0;
        if (bb > 0) vol = vol - 4.0 * M_PI * (r0 * r0 * r0) / 3.0;
        orow[3 + b] = (double)gr[b] * (1.0 / vol);
    }
    if (tid == 0)
    {
        double t[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        for (int w = 0; w < nwarps; w++)
            for (int q = 0; q < 10; q++) t[q] += red[w * 12 + q];
        const double exponent = t[9];
        const double kRI = 2.0 * t[2];
        const double kin_r = -s.hbar * (t[0] - t[1] + t[3]); // HeBulk.cpp:363-364
        const double kin_i = -s.hbar * (kRI + t[4]);
        const double e_r = kin_r + t[5];
        Arow[P] = e_r;
        Arow[P + 1] = kin_i;
        Arow[P + 2] = 1.0;
        orow[0] = kin_r; // HeBulk.cpp:395-397
        orow[1] = t[5];
        orow[2] = s.use_phi ? exp(exponent + s.phiR) : exp(exponent);
        if (a.exponent) a.exponent[row] = exponent;
        if (a.outer_out) a.outer_out[cfg] = 0.0;
    }
}

// ---- tile version: every unordered pair once (HeBulk's 64 atoms; see evaluate.cu for the scheme) -----------------
// 32x32 tiles of the pair matrix, shift s pairs row block I with column block (I+s) mod NT, row forces in registers,
// column-force accumulators rotating through the warp by shuffle.  With an even number of blocks the last shift pairs
// I with I+NT/2 from both sides; here the two warps share that tile (rotations 0..15 and 16..31) instead of one
// idling, which is what makes NT = 2 (N = 64) twice as fast as walking all partners.  Divisions that only scale
// well-conditioned factors go through one refined reciprocal.
__device__ __forceinline__ double he_rcp(double r)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r));
    double e = fma(-r, y, 1.0);
    y = fma(y, e, y);
    e = fma(-r, y, 1.0);
    y = fma(y, e, y);
    return y;
}

__global__ void __launch_bounds__(256, 4) evaluate_he_tile_kernel(EvalArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SysDev& s = a.s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int cfg = blockIdx.x;
    const int N = s.N, K = s.K, P = s.P, G = s.gr_bins, NR = s.rho_bins, NE = s.n_ext;
    const int MC = K, CO = K + 1, LI = K + 2;
    const int NT = (N + 31) >> 5, NP32 = NT * 32;

    double* utR = reinterpret_cast<double*>(smem_raw);
    double* utI = utR + NE;
    double* px = utI + NE;
    double* py = px + NP32;
    double* pz = py + NP32;
    double* hist = pz + NP32;                 // [nwarps][K]
    double* ext = hist + (size_t)nwarps * K;  // [NE]
    double* red = ext + NE;                   // [nwarps][12]
    double* com = red + (size_t)nwarps * 12;  // [4]
    double* frc = com + 4;                    // [6][NP32]
    int* gr = reinterpret_cast<int*>(frc + (size_t)6 * NP32); // [G + NR]

    for (int i = tid; i < NE; i += blockDim.x)
    {
        utR[i] = s.utR[i];
        utI[i] = s.utI[i];
    }
    const double* gpos = a.pos + (size_t)cfg * 3 * s.Np;
    for (int i = tid; i < NP32; i += blockDim.x)
    {
        const bool v = i < N;
        px[i] = v ? gpos[i] : 0.0;
        py[i] = v ? gpos[s.Np + i] : 0.0;
        pz[i] = v ? gpos[2 * s.Np + i] : 0.0;
    }
    for (int i = tid; i < 6 * NP32; i += blockDim.x) frc[i] = 0.0;
    for (int i = tid; i < nwarps * K; i += blockDim.x) hist[i] = 0.0;
    for (int i = tid; i < G + NR; i += blockDim.x) gr[i] = 0;
    __syncthreads();
    if (tid == 0 && NR > 0) // GetCenterOfMass, HeDrop.cpp:255-270
    {
        double cx = 0, cy = 0, cz = 0;
        for (int i = 0; i < N; i++)
        {
            cx += px[i];
            cy += py[i];
            cz += pz[i];
        }
        com[0] = cx / (double)N;
        com[1] = cy / (double)N;
        com[2] = cz / (double)N;
    }
    __syncthreads();

    double* myhist = hist + (size_t)warp * K;
    double* fRx = frc;
    double* fRy = fRx + NP32;
    double* fRz = fRy + NP32;
    double* fIx = fRz + NP32;
    double* fIy = fIx + NP32;
    double* fIz = fIy + NP32;
    const double rs = s.r0, rmax = s.rmax, r2s = s.r_split2, rt = s.r_tail, m = s.core_m;
    const double gr_spacing = s.gr_max / (double)G;
    const double ucR = utR[MC], ucI = utI[MC], ulR = utR[LI], ulI = utI[LI];
    const double inv_h = 1.0 / s.h, inv_hl = 1.0 / s.h_large;

    double R1 = 0.0, I1 = 0.0, RI = 0.0, lapR = 0.0, lapI = 0.0, pot = 0.0, mcm = 0.0, csum = 0.0, lsum = 0.0;

    const int half = NT >> 1;
    const bool nt_even = (NT & 1) == 0;
    for (int sft = 0; sft <= half; sft++)
    {
        const bool shared_shift = nt_even && sft == half && sft != 0; // tile (I, I+half) split between warps I and I+half
        for (int base = 0; base < NT; base += nwarps)
        {
            const int Iw = base + warp;
            const bool tile = Iw < NT;
            const bool second = shared_shift && Iw >= half; // this warp takes rotations 16..31 of the lower warp's tile
            const int I = second ? Iw - half : Iw;
            int J = I + sft;
            if (J >= NT) J -= NT;
            const int kfirst = sft == 0 ? 1 : (shared_shift ? (second ? 16 : 0) : 0);
            const int klast = sft == 0 ? 16 : (shared_shift ? (second ? 31 : 15) : 31);
            const int n = 32 * I + lane;
            double rRx = 0.0, rRy = 0.0, rRz = 0.0, rIx = 0.0, rIy = 0.0, rIz = 0.0;
            double cRx = 0.0, cRy = 0.0, cRz = 0.0, cIx = 0.0, cIy = 0.0, cIz = 0.0;
            if (tile)
            {
                const bool vn = n < N;
                const double xn = px[n], yn = py[n], zn = pz[n];
                for (int k = kfirst; k <= klast; k++)
                {
                    if (k > kfirst)
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
                    double vx, vy, vz, r;
                    if (s.periodic)
                    {
                        r = disp_exact(s, xn, yn, zn, px[i], py[i], pz[i], vx, vy, vz);
                    }
                    else
                    {
                        vx = xn - px[i]; // VectorDisplacement, Utils.cpp:253-263
                        vy = yn - py[i];
                        vz = zn - pz[i];
                        r = sqrt(vx * vx + vy * vy + vz * vz);
                    }
                    const bool pair = vn && (i < N) && (sft != 0 || k < 16 || lane < 16);
                    const bool act = pair && (r < rmax); // HeBulk.cpp:232 (HeDrop: no cut)
                    int bin = 0;
                    double val[4] = { 0.0, 0.0, 0.0, 0.0 };
                    bool spline_val = false;
                    if (act) pot += he_pair_potential(s, r);
                    if (pair && (r < s.gr_max)) atomicAdd(&gr[min((int)floor(r / gr_spacing), G - 1)], 1); // g(r) counts
                    if (act)
                    {
                        double gxR, gyR, gzR, gxI, gyI, gzI; // gradient contribution on the row particle
                        if (r < rs)
                        {
                            // McMillan core, HeBulk.cpp:268-276 (m = -5), HeDrop.cpp:402-410
                            const double rp = pow(r, m - 2.0);
                            const double g = m * rp;
                            const double l2 = m * (m + 1.0) * rp;
                            gxR = ucR * g * vx; gyR = ucR * g * vy; gzR = ucR * g * vz;
                            gxI = ucI * g * vx; gyI = ucI * g * vy; gzI = ucI * g * vz;
                            lapR = fma(ucR, l2, lapR);
                            lapI = fma(ucI, l2, lapI);
                            mcm += pow(r, m);
                        }
                        else if (r >= rt)
                        {
                            // constant + linear tails, HeDrop.cpp:411-426, 732-737
                            const double rinv = he_rcp(r);
                            const double ex = vx * rinv, ey = vy * rinv, ez = vz * rinv;
                            gxR = ulR * ex; gyR = ulR * ey; gzR = ulR * ez;
                            gxI = ulI * ex; gyI = ulI * ey; gzI = ulI * ez;
                            lapR = fma(ulR, 2.0 * rinv, lapR);
                            lapI = fma(ulI, 2.0 * rinv, lapI);
                            csum += 1.0;
                            lsum += r;
                        }
                        else
                        {
                            double interval, inps; // HeBulk.cpp:279-282, HeDrop.cpp:431-446
                            if (r < r2s)
                            {
                                inps = inv_h;
                                interval = (r - rs) / s.h;
                                bin = (int)floor(interval);
                            }
                            else
                            {
                                inps = inv_hl;
                                interval = (r - r2s) / s.h_large;
                                bin = (int)floor(interval) + s.n_short;
                            }
                            const double res = interval - floor(interval);
                            const double res2 = res * res;
                            const double inps2 = inps * inps;
                            double tmp[4], l2[4];
                            tmp[0] = -1.0 / 2.0 * (1.0 - 2.0 * res + res2); // HeBulk.cpp:284-287
                            tmp[1] = 1.0 / 6.0 * (-12.0 * res + 9.0 * res2);
                            tmp[2] = 1.0 / 6.0 * (3.0 + 6.0 * res - 9.0 * res2);
                            tmp[3] = 1.0 / 2.0 * res2;
                            const double rinv = he_rcp(r);
                            const double f2 = 2.0 * inps * rinv;
                            l2[0] = inps2 * (1.0 - res) + f2 * tmp[0]; // HeBulk.cpp:299-302
                            l2[1] = inps2 * (1.0 / 6.0 * (-12.0 + 18.0 * res)) + f2 * tmp[1];
                            l2[2] = inps2 * (1.0 / 6.0 * (6.0 - 18.0 * res)) + f2 * tmp[2];
                            l2[3] = inps2 * (res) + f2 * tmp[3];
                            const double ex = vx * rinv, ey = vy * rinv, ez = vz * rinv;
                            double gR = 0.0, gI = 0.0;
#pragma unroll
                            for (int b = 0; b < 4; b++)
                            {
                                const double uRk = utR[bin + b], uIk = utI[bin + b];
                                gR = fma(uRk, tmp[b], gR);
                                gI = fma(uIk, tmp[b], gI);
                                lapR = fma(uRk, l2[b], lapR);
                                lapI = fma(uIk, l2[b], lapI);
                            }
                            gR = gR * inps;
                            gI = gI * inps;
                            gxR = gR * ex; gyR = gR * ey; gzR = gR * ez;
                            gxI = gI * ex; gyI = gI * ey; gzI = gI * ez;
                            const double res3 = res2 * res; // pow(res, 3)
                            // values for splines bin .. bin+3 (HeBulk.cpp:483-486), stored for hist[bin + 3 - p]
                            val[3] = -1.0 / 6.0 * (-1.0 + 3.0 * res - 3.0 * res2 + res3);
                            val[2] = 1.0 / 6.0 * (4.0 - 6.0 * res2 + 3.0 * res3);
                            val[1] = 1.0 / 6.0 * (1.0 + 3.0 * res + 3.0 * res2 - 3.0 * res3);
                            val[0] = 1.0 / 6.0 * res3;
                            spline_val = true;
                        }
                        rRx += gxR; rRy += gyR; rRz += gzR; rIx += gxI; rIy += gyI; rIz += gzI;
                        cRx -= gxR; cRy -= gyR; cRz -= gzR; cIx -= gxI; cIy -= gyI; cIz -= gzI; // the partner sees -v
                    }
                    warp_hist_add4(myhist, bin + 3, spline_val, val, lane);
                }
            }
            // flush: columns of the first role, columns of the second, rows of the first, rows of the second
            const int ic = 32 * J + ((lane + klast) & 31);
            for (int role = 0; role < (shared_shift ? 2 : 1); role++)
            {
                if (tile && (int)second == role)
                {
                    fRx[ic] += cRx; fRy[ic] += cRy; fRz[ic] += cRz;
                    fIx[ic] += cIx; fIy[ic] += cIy; fIz[ic] += cIz;
                }
                __syncthreads();
            }
            for (int role = 0; role < (shared_shift ? 2 : 1); role++)
            {
                if (tile && (int)second == role)
                {
                    fRx[n] += rRx; fRy[n] += rRy; fRz[n] += rRz;
                    fIx[n] += rIx; fIy[n] += rIy; fIz[n] += rIz;
                }
                __syncthreads();
            }
        }
    }
    lapR *= 2.0; // each pair enters the Laplacian of both partners with the same value
    lapI *= 2.0;

    for (int n = tid; n < N; n += blockDim.x)
    {
        if (NR > 0) // density profile around the centre of mass, HeDrop.cpp:482-491
        {
            const double d0 = px[n] - com[0], d1 = py[n] - com[1], d2 = pz[n] - com[2];
            const double rr = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            if (rr < s.gr_max) atomicAdd(&gr[G + min((int)floor(rr / gr_spacing), NR - 1)], 1);
        }
        // the literal 1 of the last parameter, HeBulk.cpp:351
        const double ax = fRx[n] + s.g0R, ay = fRy[n] + s.g0R, az = fRz[n] + s.g0R;
        const double bx = fIx[n] + s.g0I, by = fIy[n] + s.g0I, bz = fIz[n] + s.g0I;
        R1 += ax * ax + ay * ay + az * az;
        I1 += bx * bx + by * by + bz * bz;
        RI += ax * bx + ay * by + az * bz;
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
    __syncthreads(); // the density counts are read by all threads below

    R1 = warp_sum(R1); I1 = warp_sum(I1); RI = warp_sum(RI);
    lapR = warp_sum(lapR); lapI = warp_sum(lapI); pot = warp_sum(pot);
    mcm = warp_sum(mcm); csum = warp_sum(csum); lsum = warp_sum(lsum);
    if (lane == 0)
    {
        double* r = red + warp * 12;
        r[0] = R1; r[1] = I1; r[2] = RI; r[3] = lapR; r[4] = lapI; r[5] = pot; r[6] = mcm; r[7] = csum; r[8] = lsum;
    }
    __syncthreads();
    for (int k = tid; k < K; k += blockDim.x)
    {
        double t = 0.0;
        for (int w = 0; w < nwarps; w++) t += hist[(size_t)w * K + k];
        ext[k] = t;
    }
    if (tid == 0)
    {
        double t6 = 0.0, t7 = 0.0, t8 = 0.0;
        for (int w = 0; w < nwarps; w++)
        {
            t6 += red[w * 12 + 6];
            t7 += red[w * 12 + 7];
            t8 += red[w * 12 + 8];
        }
        ext[MC] = t6;
        ext[CO] = t7;
        ext[LI] = t8;
    }
    __syncthreads();
    if (a.ss_out)
        for (int k = tid; k < NE; k += blockDim.x) a.ss_out[(size_t)cfg * NE + k] = ext[k];

    const long long row = a.row0 + (long long)cfg * a.row_stride;
    double* Arow = a.A + (size_t)row * a.lda;
    double epart = 0.0;
    for (int p = tid; p < P; p += blockDim.x)
    {
        double o = s.map_const[p]; // HeBulk.cpp:376-383, HeDrop.cpp:609-626
        for (int j = s.map_ptr[p]; j < s.map_ptr[p + 1]; j++) o += s.map_val[j] * ext[s.map_col[j]];
        Arow[p] = o;
        epart = fma(s.uR[p], o, epart); // HeBulk.cpp:491-498
    }
    epart = warp_sum(epart);
    if (lane == 0) red[warp * 12 + 9] = epart;
    __syncthreads();
    double* orow = a.other + (size_t)row * s.n_other;
    for (int b = tid; b < G + NR; b += blockDim.x)
    {
        // 1 / grBinVolumes[b], HeBulk.cpp:133-143 (the density profile uses the same shell volumes, HeDrop.cpp:490)
        const int bb = b < G ? b : b - G;
        const double r1 = gr_spacing * (bb + 1), r0 = gr_spacing * bb;
        double vol = 4.0 * M_PI * (r1 * r1 * r1) / 3.0;
        if (bb > 0) vol = vol - 4.0 * M_PI * (r0 * r0 * r0) / 3.0;
        orow[3 + b] = (double)gr[b] * (1.0 / vol);
    }
    if (tid == 0)
    {
        double t[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        for (int w = 0; w < nwarps; w++)
            for (int q = 0; q < 10; q++) t[q] += red[w * 12 + q];
        const double exponent = t[9];
        const double kRI = 2.0 * t[2];
        const double kin_r = -s.hbar * (t[0] - t[1] + t[3]); // HeBulk.cpp:363-364
        const double kin_i = -s.hbar * (kRI + t[4]);
        const double e_r = kin_r + t[5];
        Arow[P] = e_r;
        Arow[P + 1] = kin_i;
        Arow[P + 2] = 1.0;
        orow[0] = kin_r; // HeBulk.cpp:395-397
        orow[1] = t[5];
        orow[2] = s.use_phi ? exp(exponent + s.phiR) : exp(exponent);
        if (a.exponent) a.exponent[row] = exponent;
        if (a.outer_out) a.outer_out[cfg] = 0.0;
    }
}


cudaError_t launch_evaluate_he(const EvalArgs& a, cudaStream_t st)
{
    if (a.n_cfg <= 0) return cudaSuccess;
    const SysDev& s = a.s;
    int threads = ((s.N + 31) / 32) * 32;
    if (threads > 256) threads = 256;
    const int nwarps = threads / 32;
    size_t smem = sizeof(double) * ((size_t)2 * s.n_ext + 3 * (size_t)s.N + (size_t)nwarps * s.K + s.n_ext + (size_t)nwarps * 12 + 4) +
                  sizeof(int) * (size_t)(s.gr_bins + s.rho_bins) + 16;
    if (s.N > 16) // small clusters (HeDrop's six atoms) keep the walk over partners: a 32x32 tile would idle
    {
        const int np32 = ((s.N + 31) / 32) * 32;
        smem = sizeof(double) * ((size_t)2 * s.n_ext + 3 * (size_t)np32 + (size_t)nwarps * s.K + s.n_ext + (size_t)nwarps * 12 + 4 +
                                 (size_t)6 * np32) +
               sizeof(int) * (size_t)(s.gr_bins + s.rho_bins) + 16;
        cudaError_t e = cudaFuncSetAttribute(evaluate_he_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        evaluate_he_tile_kernel<<<a.n_cfg, threads, smem, st>>>(a);
        return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(evaluate_he_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    evaluate_he_kernel<<<a.n_cfg, threads, smem, st>>>(a);
    return cudaGetLastError();
}

} // namespace tdvmc
