// Microbenchmark: issue rate of the 64-bit conversions the sweep could use instead of FP64 adds
// (I2F.F64.S64, F2I.F64.TRUNC, I2F.F64.S32) alone and mixed with DFMA, in warp-instructions per clock per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o conv_throughput conv_throughput.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) probe(long long* out, long long seed, int iters, long long* clk)
{
    long long a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7;
    double f0 = 1.0 + threadIdx.x * 1e-9, f1 = f0 * 1.1, f2 = f0 * 1.2, f3 = f0 * 1.3;
    double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int u = 0; u < 8; u++)
        {
            if (MODE == 0 || MODE == 2) // I2F.F64.S64 (x4 independent)
            {
                acc0 += 0; // placeholder to keep structure similar
                double c0 = __ll2double_rn(a0), c1 = __ll2double_rn(a1), c2 = __ll2double_rn(a2), c3 = __ll2double_rn(a3);
                a0 += __double_as_longlong(c0) & 0xff; a1 += __double_as_longlong(c1) & 0xff;
                a2 += __double_as_longlong(c2) & 0xff; a3 += __double_as_longlong(c3) & 0xff;
            }
            if (MODE == 1 || MODE == 2) // DFMA (x4 independent)
            {
                f0 = fma(f0, 1.0000001, 1e-9); f1 = fma(f1, 1.0000001, 1e-9);
                f2 = fma(f2, 1.0000001, 1e-9); f3 = fma(f3, 1.0000001, 1e-9);
            }
            if (MODE == 3) // F2I.F64.TRUNC + I2F.F64 (S32)
            {
                int i0 = (int)f0, i1 = (int)f1, i2 = (int)f2, i3 = (int)f3;
                f0 = (double)(i0 + 1) ; f1 = (double)(i1 + 2); f2 = (double)(i2 + 3); f3 = (double)(i3 + 4);
            }
            if (MODE == 4) // MUFU.RSQ64H
            {
                asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(f0));
                asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(f1));
                asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(f2));
                asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(f3));
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __double_as_longlong(f0 + f1 + f2 + f3 + acc0 + acc1 + acc2 + acc3);
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_unroll)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long *out, *clk;
    cudaMalloc(&out, sizeof(long long) * sms * 1024);
    cudaMalloc(&clk, sizeof(long long) * sms);
    const int iters = 4096;
    probe<MODE><<<sms, 1024>>>(out, 12345, 16, clk);
    probe<MODE><<<sms, 1024>>>(out, 12345, iters, clk);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += (double)h[i];
    avg /= sms;
    const double warp_inst = 32.0 * iters * 8 * ops_per_unroll; // 32 warps per SM
    printf("%-34s %8.3f warp-inst/clk/SM  (%6.1f lanes/clk/SM)\n", name, warp_inst / avg, 32.0 * warp_inst / avg);
    cudaFree(out);
    cudaFree(clk);
}

int main()
{
    run<0>("I2F.F64.S64", 4);
    run<1>("DFMA", 4);
    run<2>("I2F.F64.S64 + DFMA (count each)", 4);
    run<3>("F2I.F64.TRUNC + I2F.F64.S32 (pairs)", 4);
    run<4>("MUFU.RSQ64H", 4);
    return 0;
}
