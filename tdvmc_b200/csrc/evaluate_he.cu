// Fused evaluation for HeBulk (src/PhysicalSystems/HeBulk.cpp): periodic He-4 with a McMillan r^-5 core
// below rijSplit, uniform cubic B-splines written in the local coordinate res = (r - rs)/h - bin above it,
// the Aziz HFD-B(He) potential inline and g(r) carried in otherExpectationValues[3..102].
//
// Same structure as evaluate.cu (one block per configuration, thread n owns particle n, contraction with
// u~ = M^T u on the fly, per-warp conflict-free histogram for the value sums); the per-pair expressions are
// the reference's (HeBulk.cpp:268-302, 471-486), so each term agrees with the reference to the last bits
// (libm pow/exp differ from glibc by <= 1-2 ulp).
#include "kernels.cuh"

namespace tdvmc
{

// see evaluate.cu; hist[bin - p] += v[p]
__device__ __forceinline__ void warp_hist_add4_he(double* hist, int bin, bool active, const double (&v)[4], int lane)
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

__global__ void __launch_bounds__(256) evaluate_he_kernel(EvalArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SysDev& s = a.s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int cfg = blockIdx.x;
    const int N = s.N, K = s.K, P = s.P, G = s.gr_bins, NE = s.n_ext;

    double* utR = reinterpret_cast<double*>(smem_raw);
    double* utI = utR + NE;
    double* px = utI + NE;
    double* py = px + N;
    double* pz = py + N;
    double* hist = pz + N;                    // [nwarps][K]
    double* ext = hist + (size_t)nwarps * K;  // [NE] extended sums: ss[0..K) then the McMillan sum
    double* red = ext + NE;                   // [nwarps][8]
    int* gr = reinterpret_cast<int*>(red + (size_t)nwarps * 8); // [G]

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
    for (int i = tid; i < G; i += blockDim.x) gr[i] = 0;
    __syncthreads();

    double* myhist = hist + (size_t)warp * K;
    const double rs = s.r0, h = s.h, rmax = s.rmax;
    const double h2 = h * h; // pow(nodePointSpacing, 2), HeBulk.cpp:56
    const double gr_spacing = rmax / (double)G; // HeBulk.cpp:130-131
    // Aziz HFD-B(He), HeBulk.cpp:187-195
    const double e = 10.948, rm = 2.963, aa = 184431.01, alpha = 10.43329537, beta = -2.27965105, dd = 1.4826,
                 c6 = 1.36745214, c8 = 0.42123807, c10 = 0.17473318;
    const double ucR = utR[K], ucI = utI[K]; // McMillan column

    double R1 = 0.0, I1 = 0.0, RI = 0.0, lapR = 0.0, lapI = 0.0, pot = 0.0, mcm = 0.0;

    for (int n0 = warp * 32; n0 < N; n0 += blockDim.x)
    {
        const int n = n0 + lane;
        const bool valid = n < N;
        const double xn = valid ? px[n] : 0.0, yn = valid ? py[n] : 0.0, zn = valid ? pz[n] : 0.0;
        double fRx = 0.0, fRy = 0.0, fRz = 0.0, fIx = 0.0, fIy = 0.0, fIz = 0.0;
        for (int i = 0; i < N; i++)
        {
            double vx, vy, vz;
            const double r = disp_exact(s, xn, yn, zn, px[i], py[i], pz[i], vx, vy, vz);
            const bool act = valid && (i != n) && (r < rmax); // HeBulk.cpp:232
            const bool lower = act && (i < n);
            int bin = 0;
            double val[4] = { 0.0, 0.0, 0.0, 0.0 };
            bool spline_val = false;
            if (lower)
            {
                // Aziz potential, HeBulk.cpp:251-261
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
                pot += e * (aa * exp(-alpha * x + beta * x2) - F * xm6 * (c6 + xm2 * (c8 + xm2 * c10)));
                atomicAdd(&gr[min((int)floor(r / gr_spacing), G - 1)], 1); // g(r) counts, HeBulk.cpp:305-311
            }
            if (act)
            {
                if (r < rs)
                {
                    // McMillan core, HeBulk.cpp:268-276, 471-474
                    const double rm7 = pow(r, -7.0);
                    const double g = -5.0 * rm7;
                    fRx = fma(ucR * g, vx, fRx); fRy = fma(ucR * g, vy, fRy); fRz = fma(ucR * g, vz, fRz);
                    fIx = fma(ucI * g, vx, fIx); fIy = fma(ucI * g, vy, fIy); fIz = fma(ucI * g, vz, fIz);
                    lapR = fma(ucR, 20.0 * rm7, lapR);
                    lapI = fma(ucI, 20.0 * rm7, lapI);
                    if (lower) mcm += pow(r, -5.0);
                }
                else
                {
                    const double interval = (r - rs) / h; // HeBulk.cpp:279-282
                    bin = (int)floor(interval);
                    const double res = interval - bin;
                    const double res2 = res * res;
                    double tmp[4], l2[4];
                    tmp[0] = -1.0 / 2.0 * (1.0 - 2.0 * res + res2); // HeBulk.cpp:284-287
                    tmp[1] = 1.0 / 6.0 * (-12.0 * res + 9.0 * res2);
                    tmp[2] = 1.0 / 6.0 * (3.0 + 6.0 * res - 9.0 * res2);
                    tmp[3] = 1.0 / 2.0 * res2;
                    const double f2 = 2.0 / (h * r);
                    l2[0] = 1.0 / h2 * (1.0 - res) + f2 * tmp[0]; // HeBulk.cpp:299-302
                    l2[1] = 1.0 / h2 * (1.0 / 6.0 * (-12.0 + 18.0 * res)) + f2 * tmp[1];
                    l2[2] = 1.0 / h2 * (1.0 / 6.0 * (6.0 - 18.0 * res)) + f2 * tmp[2];
                    l2[3] = 1.0 / h2 * (res) + f2 * tmp[3];
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
                    gR = gR / h;
                    gI = gI / h;
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
            warp_hist_add4_he(myhist, bin + 3, spline_val, val, lane);
        }
        if (valid)
        {
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
    lapR = warp_sum(lapR); lapI = warp_sum(lapI); pot = warp_sum(pot); mcm = warp_sum(mcm);
    if (lane == 0)
    {
        double* r = red + warp * 8;
        r[0] = R1; r[1] = I1; r[2] = RI; r[3] = lapR; r[4] = lapI; r[5] = pot; r[6] = mcm; r[7] = 0.0;
    }
    __syncthreads();
    for (int k = tid; k < K; k += blockDim.x)
    {
        double t = 0.0;
        for (int w = 0; w < nwarps; w++) t += hist[(size_t)w * K + k];
        ext[k] = t;
        if (a.ss_out) a.ss_out[(size_t)cfg * NE + k] = t;
    }
    if (tid == 0)
    {
        double t = 0.0;
        for (int w = 0; w < nwarps; w++) t += red[w * 8 + 6];
        ext[K] = t;
        if (a.ss_out) a.ss_out[(size_t)cfg * NE + K] = t;
    }
    __syncthreads();

    const long long row = a.row0 + (long long)cfg * a.row_stride;
    double* Arow = a.A + (size_t)row * a.lda;
    double epart = 0.0;
    for (int p = tid; p < P; p += blockDim.x)
    {
        double o = s.map_const[p]; // HeBulk.cpp:376-383
        for (int j = s.map_ptr[p]; j < s.map_ptr[p + 1]; j++) o += s.map_val[j] * ext[s.map_col[j]];
        Arow[p] = o;
        epart = fma(s.uR[p], o, epart); // HeBulk.cpp:491-498
    }
    epart = warp_sum(epart);
    if (lane == 0) red[warp * 8 + 7] = epart;
    __syncthreads();
    double* orow = a.other + (size_t)row * s.n_other;
    for (int b = tid; b < G; b += blockDim.x)
    {
        // 1 / grBinVolumes[b], HeBulk.cpp:133-143
        const double r1 = gr_spacing * (b + 1), r0 = gr_spacing * b;
        double vol = 4.0 * M_PI * (r1 * r1 * r1) / 3.0;
        if (b > 0) vol = vol - 4.0 * M_PI * (r0 * r0 * r0) / 3.0;
        orow[3 + b] = (double)gr[b] * (1.0 / vol);
    }
    if (tid == 0)
    {
        double t[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        for (int w = 0; w < nwarps; w++)
            for (int q = 0; q < 8; q++) t[q] += red[w * 8 + q];
        const double exponent = t[7];
        const double kRI = 2.0 * t[2];
        const double kin_r = -s.hbar * (t[0] - t[1] + t[3]); // HeBulk.cpp:363-364
        const double kin_i = -s.hbar * (kRI + t[4]);
        const double e_r = kin_r + t[5];
        Arow[P] = e_r;
        Arow[P + 1] = kin_i;
        Arow[P + 2] = 1.0;
        orow[0] = kin_r; // HeBulk.cpp:395-397
        orow[1] = t[5];
        orow[2] = exp(exponent);
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
    size_t smem = sizeof(double) * ((size_t)2 * s.n_ext + 3 * (size_t)s.N + (size_t)nwarps * s.K + s.n_ext + (size_t)nwarps * 8) +
                  sizeof(int) * (size_t)s.gr_bins + 16;
    cudaError_t e = cudaFuncSetAttribute(evaluate_he_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    evaluate_he_kernel<<<a.n_cfg, threads, smem, st>>>(a);
    return cudaGetLastError();
}

} // namespace tdvmc
