// K5 — S-matrix / force-vector accumulation as ONE symmetric rank-M update on FP64 tensor cores.
//
// The reference adds, sample by sample, O (x) O / M, O E^R / M, O E^I / M, O / M, E / M into seven
// nested vectors (src/TDVMC.cpp:1103-1109, BosonsBulk.cpp:431-440).  With the sample rows augmented to
//     a_m = [ O_0 .. O_{P-1} | E^R | E^I | 1 ]            (P + 3 columns)
// all of it is the upper triangle of G = sum_m a_m a_m^T:
//     S = G[0:P,0:P],  F^R = G[0:P,P],  F^I = G[0:P,P+1],  sum O = G[0:P,P+2],
//     sum E^R = G[P,P+2],  sum E^I = G[P+1,P+2],  M = G[P+2,P+2].
//
// Kernel: persistent, one CTA of 16 warps per SM.  Each CTA streams its share of the sample rows
// through shared memory with 1-D bulk TMA copies (cp.async.bulk + mbarrier, 3 stages; the rows of a
// chunk are contiguous in HBM, so one copy per stage) and keeps the WHOLE upper triangle of G in
// registers, split into 16x16 super-tiles of 2x2 mma.sync.m8n8k4.f64 (DMMA) tiles.  Every sample row
// is read from HBM exactly once per launch; flops are the symmetric minimum (+ the diagonal blocks).
// tcgen05 has no f64 kind, so mma.sync DMMA is the FP64 tensor path on sm_100a.
// CTAs write their partial G to HBM; a second kernel reduces them in a fixed order (deterministic).
#include "kernels.cuh"

namespace tdvmc
{

constexpr int kAccWarps = 16;
constexpr int kAccStages = 3;
constexpr int kAccMaxSuper = 6; // super-tiles per warp: 16 * 6 = 96 >= 13 * 14 / 2

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kAccWarps * 32, 1) syrk_kernel(AccArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lda = a.lda;
    const uint32_t stage_bytes = (uint32_t)(kAccChunkRows * lda * sizeof(double));
    double* stage0 = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kAccStages * stage_bytes);

    const int T2 = a.ldc / 16;                 // super-tiles per dimension
    const int n_super = T2 * (T2 + 1) / 2;     // upper triangle, row-major enumeration
    int si[kAccMaxSuper], sj[kAccMaxSuper];
    bool ok[kAccMaxSuper];
#pragma unroll
    for (int j = 0; j < kAccMaxSuper; j++)
    {
        int q = warp + kAccWarps * j;
        ok[j] = q < n_super;
        int row = 0, rem = ok[j] ? q : 0;
        while (rem >= T2 - row)
        {
            rem -= T2 - row;
            row++;
        }
        si[j] = row;
        sj[j] = row + rem;
    }
    double acc[kAccMaxSuper][4][2];
#pragma unroll
    for (int j = 0; j < kAccMaxSuper; j++)
#pragma unroll
        for (int t = 0; t < 4; t++) acc[j][t][0] = acc[j][t][1] = 0.0;

    const long long c_begin = (a.n_chunks * blockIdx.x) / gridDim.x;
    const long long c_end = (a.n_chunks * (blockIdx.x + 1)) / gridDim.x;
    const int n_loc = (int)(c_end - c_begin);

    if (tid == 0)
    {
        for (int s = 0; s < kAccStages; s++) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
    {
        for (int s = 0; s < kAccStages && s < n_loc; s++)
        {
            mbar_expect_tx(full + s, stage_bytes);
            tma_load_1d(stage0 + (size_t)s * kAccChunkRows * lda, a.A + (size_t)(c_begin + s) * kAccChunkRows * lda,
                        stage_bytes, full + s);
        }
    }

    const int fr = lane & 3;  // k index inside the fragment
    const int fc = lane >> 2; // row/column index inside the 8x8 tile
    for (int it = 0; it < n_loc; it++)
    {
        const int st = it % kAccStages;
        mbar_wait(full + st, (uint32_t)((it / kAccStages) & 1));
        const double* X = stage0 + (size_t)st * kAccChunkRows * lda;
#pragma unroll 2
        for (int k0 = 0; k0 < kAccChunkRows; k0 += 4)
        {
            const double* xr = X + (size_t)(k0 + fr) * lda + fc;
#pragma unroll
            for (int j = 0; j < kAccMaxSuper; j++)
            {
                if (ok[j])
                {
                    const double a0 = xr[si[j] * 16], a1 = xr[si[j] * 16 + 8];
                    const double b0 = xr[sj[j] * 16], b1 = xr[sj[j] * 16 + 8];
                    dmma_8x8x4(acc[j][0][0], acc[j][0][1], a0, b0);
                    dmma_8x8x4(acc[j][1][0], acc[j][1][1], a0, b1);
                    dmma_8x8x4(acc[j][2][0], acc[j][2][1], a1, b0);
                    dmma_8x8x4(acc[j][3][0], acc[j][3][1], a1, b1);
                }
            }
        }
        __syncthreads(); // everyone is done with stage st
        if (tid == 0 && it + kAccStages < n_loc)
        {
            mbar_expect_tx(full + st, stage_bytes);
            tma_load_1d(stage0 + (size_t)st * kAccChunkRows * lda,
                        a.A + (size_t)(c_begin + it + kAccStages) * kAccChunkRows * lda, stage_bytes, full + st);
        }
    }

    // partial G of this CTA: C fragment layout of m8n8k4: c0 -> (row = lane/4, col = 2*(lane%4)), c1 -> col + 1
    double* G = a.partial + (size_t)blockIdx.x * a.ldc * a.ldc;
#pragma unroll
    for (int j = 0; j < kAccMaxSuper; j++)
    {
        if (!ok[j]) continue;
#pragma unroll
        for (int t = 0; t < 4; t++)
        {
            const int row = si[j] * 16 + (t >> 1) * 8 + fc;
            const int col = sj[j] * 16 + (t & 1) * 8 + 2 * fr;
            *reinterpret_cast<double2*>(G + (size_t)row * a.ldc + col) = make_double2(acc[j][t][0], acc[j][t][1]);
        }
    }
}

size_t accumulate_smem_bytes(int lda)
{
    return (size_t)kAccStages * kAccChunkRows * lda * sizeof(double) + kAccStages * sizeof(uint64_t);
}

cudaError_t launch_accumulate(const AccArgs& a, cudaStream_t st)
{
    if (a.ldc / 16 > 13 || a.ldc % 16 != 0 || a.lda % 16 != 4 || a.lda < a.ldc) return cudaErrorInvalidValue;
    const size_t smem = accumulate_smem_bytes(a.lda);
    cudaError_t e = cudaFuncSetAttribute(syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    syrk_kernel<<<a.n_cta, kAccWarps * 32, smem, st>>>(a);
    return cudaGetLastError();
}

// ---- reduction of the per-CTA partials and packing of the estimator buffer ----
__global__ void acc_reduce_kernel(const double* partial, int n_cta, int ldc, double* G)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ldc * ldc) return;
    const int row = idx / ldc, col = idx % ldc;
    if (row / 16 > col / 16) return; // lower super-tiles were never written
    double t = 0.0;
    for (int c = 0; c < n_cta; c++) t += partial[(size_t)c * ldc * ldc + idx];
    G[idx] = t;
}

// Column sums of the otherExpectationValues rows (HeDrop / the mixture carry ~400 histogram columns per sample: 9 GB per pass
// of 2.8 M samples).  grid (n_blk, strips of 32 columns); block b sums rows [b M / n_blk, (b + 1) M / n_blk): a warp reads 32
// consecutive columns of a row (256 contiguous bytes), the eight warps of the block take the rows in turn with four
// independent accumulators each, and the eight partial sums meet in shared memory in a fixed order (deterministic).
// (r02: the first version gave every thread one column and one load in flight - 0.46 TB/s, 19 ms of config 5's 66 ms pass.)
__global__ void __launch_bounds__(256) other_partial_kernel(const double* __restrict__ other, long long M, int n_other, int n_blk, double* part)
{
    __shared__ double red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + lane;
    const long long r0 = (M * blockIdx.x) / n_blk, r1 = (M * (blockIdx.x + 1)) / n_blk;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    if (c < n_other)
    {
        const double* col = other + c;
        long long r = r0 + warp;
        for (; r + 24 < r1; r += 32)
        {
            t0 += __ldcs(col + (size_t)r * n_other);
            t1 += __ldcs(col + (size_t)(r + 8) * n_other);
            t2 += __ldcs(col + (size_t)(r + 16) * n_other);
            t3 += __ldcs(col + (size_t)(r + 24) * n_other);
        }
        for (; r < r1; r += 8) t0 += __ldcs(col + (size_t)r * n_other);
    }
    red[warp][lane] = (t0 + t1) + (t2 + t3);
    __syncthreads();
    if (warp == 0 && c < n_other)
    {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += red[w][lane];
        part[(size_t)blockIdx.x * n_other + c] = t;
    }
}

// total of the per-walker acceptance counters (integers: the order does not matter)
__global__ void __launch_bounds__(256) accepted_sum_kernel(const unsigned long long* __restrict__ accepted, int W, unsigned long long* total)
{
    unsigned long long t = 0;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x) t += accepted[w];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL_MASK, t, o);
    if ((threadIdx.x & 31) == 0 && t) atomicAdd(total, t);
}

__global__ void pack_est_kernel(AccFinishArgs a, const double* G, const double* other_part, int n_blk, const unsigned long long* accepted_total)
{
    const int P = a.P, ldc = a.ldc;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthr = gridDim.x * blockDim.x;
    double* S = a.est;
    double* FR = S + (size_t)P * P;
    double* FI = FR + P;
    double* O = FI + P;
    double* E = O + P;
    double* oth = E + 2;
    double* cnt = oth + a.n_other;
    for (int idx = tid; idx < P * P; idx += nthr)
    {
        const int i = idx / P, j = idx % P;
        S[idx] = (i / 16 <= j / 16) ? G[(size_t)i * ldc + j] : G[(size_t)j * ldc + i];
    }
    for (int k = tid; k < P; k += nthr)
    {
        FR[k] = G[(size_t)k * ldc + P];
        FI[k] = G[(size_t)k * ldc + P + 1];
        O[k] = G[(size_t)k * ldc + P + 2];
    }
    for (int c = tid; c < a.n_other; c += nthr)
    {
        double t = 0.0;
        for (int b = 0; b < n_blk; b++) t += other_part[(size_t)b * a.n_other + c];
        oth[c] = t;
    }
    if (tid == 0)
    {
        E[0] = G[(size_t)P * ldc + P + 2];
        E[1] = G[(size_t)(P + 1) * ldc + P + 2];
        cnt[0] = (double)*accepted_total; // (r02: a one-thread loop over the walkers here cost 6.5 ms at the 175 k walkers of config 1)
        cnt[1] = a.n_trials;
        cnt[2] = G[(size_t)(P + 2) * ldc + P + 2]; // sum of 1*1 over the samples = M
    }
}

// scratch layout behind a.partial: [n_cta][ldc*ldc] partials | G[ldc*ldc] | other_part[148][n_other] | acceptance total
cudaError_t launch_acc_finish(const AccFinishArgs& a, cudaStream_t st)
{
    const size_t gsz = (size_t)a.ldc * a.ldc;
    double* G = const_cast<double*>(a.partial) + (size_t)a.n_cta * gsz;
    double* other_part = G + gsz;
    int n_blk = (int)((a.M + 255) / 256);
    if (n_blk > 148) n_blk = 148;
    if (n_blk < 1) n_blk = 1;
    acc_reduce_kernel<<<(int)((gsz + 255) / 256), 256, 0, st>>>(a.partial, a.n_cta, a.ldc, G);
    other_partial_kernel<<<dim3(n_blk, a.n_other > 32 ? (a.n_other + 31) / 32 : 1), 256, 0, st>>>(a.other, a.M, a.n_other, n_blk, other_part);
    unsigned long long* accepted_total = reinterpret_cast<unsigned long long*>(other_part + (size_t)148 * a.n_other);
    cudaError_t e = cudaMemsetAsync(accepted_total, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    accepted_sum_kernel<<<64, 256, 0, st>>>(a.accepted, a.W, accepted_total);
    pack_est_kernel<<<64, 256, 0, st>>>(a, G, other_part, n_blk, accepted_total);
    return cudaGetLastError();
}

} // namespace tdvmc
