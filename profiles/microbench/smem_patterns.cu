// Microbenchmark: cost of shared-memory access patterns a bin-sorted evaluation kernel would use, in clocks per
// warp-instruction per SM (the crossbar serves one 128-byte wavefront per clock): broadcast and few-address LDS.128,
// random 16-byte gathers / scatters, 32-bit shared atomics on a 200-entry counter array.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_patterns smem_patterns.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned lcg(unsigned& s)
{
    s = s * 1664525u + 1013904223u;
    return s >> 8;
}

// MODE: 0 LDS.128 distinct contiguous | 1 LDS.128 one address | 2 two rows | 3 four rows | 4 eight rows (same offset)
//       5 LDS.64 one address | 6 LDS.64 distinct | 7 ATOMS.ADD.U32 random of 200 | 8 STS.128 random slot | 9 LDS.128 random slot
//       10 LDS.U16 random | 11 LDS.64 random | 12 STS.64 contiguous | 13 ATOMS.ADD.U32 sorted-ish (lanes in runs of ~8 equal)
template <int MODE>
__global__ void __launch_bounds__(1024) probe(double* out, int iters, long long* clk)
{
    extern __shared__ __align__(128) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<double*>(sm)[i] = i * 1e-3;
    __syncthreads();
    unsigned seed = threadIdx.x * 2654435761u + 12345u;
    double acc = 0.0;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int u = 0; u < 8; u++)
        {
            const unsigned row = (unsigned)((it * 8 + u + warp * 7) & 127); // 128-byte row, warp-uniform, changes every access
            unsigned addr;
            double a = 0, b = 0;
            if (MODE == 0) addr = base + ((row * 4 + (lane >> 3)) & 1023) * 128 + (lane & 7) * 16;
            if (MODE == 1 || MODE == 5) addr = base + row * 128 + 32;
            if (MODE == 2) addr = base + ((row + (lane >> 4) * 37) & 1023) * 128 + 32;
            if (MODE == 3) addr = base + ((row + (lane >> 3) * 37) & 1023) * 128 + 32;
            if (MODE == 4) addr = base + ((row + (lane >> 2) * 37) & 1023) * 128 + 32;
            if (MODE == 6) addr = base + ((row * 2 + (lane >> 4)) & 1023) * 128 + (lane & 15) * 8;
            if (MODE == 7) addr = base + (lcg(seed) % 200u) * 4;
            if (MODE == 13) addr = base + (((lcg(seed) & 1u) + (unsigned)(lane >> 3) * 2u + row) % 200u) * 4;
            if (MODE == 8 || MODE == 9) addr = base + (lcg(seed) & 8191u) * 16;
            if (MODE == 10) addr = base + (lcg(seed) & 32767u) * 2;
            if (MODE == 11) addr = base + (lcg(seed) & 16383u) * 8;
            if (MODE == 12) addr = base + ((row * 2) & 1023) * 128 + lane * 8;
            if (MODE <= 4 || MODE == 9) asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
            if (MODE == 5 || MODE == 6 || MODE == 11) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(addr));
            if (MODE == 7 || MODE == 13)
            {
                unsigned old;
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr));
                a = (double)old;
            }
            if (MODE == 8) asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(acc), "d"(acc + 1.0));
            if (MODE == 12) asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(acc));
            if (MODE == 10)
            {
                unsigned short h;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
                a = (double)h;
            }
            acc += a + b;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    long long* clk;
    cudaMalloc(&out, sizeof(double) * sms * 1024);
    cudaMalloc(&clk, sizeof(long long) * sms);
    const int iters = 2048;
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    probe<MODE><<<sms, 1024, 131072>>>(out, 16, clk);
    probe<MODE><<<sms, 1024, 131072>>>(out, iters, clk);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += (double)h[i];
    avg /= sms;
    const double warp_inst = 32.0 * iters * 8;
    printf("%-52s %7.2f clocks per warp-instruction per SM   (%s)\n", name, avg / warp_inst, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(clk);
}

int main()
{
    run<0>("LDS.128 32 distinct contiguous");
    run<1>("LDS.128 one address (broadcast)");
    run<2>("LDS.128 two rows, same offset");
    run<3>("LDS.128 four rows, same offset");
    run<4>("LDS.128 eight rows, same offset");
    run<5>("LDS.64 one address (broadcast)");
    run<6>("LDS.64 32 distinct contiguous");
    run<11>("LDS.64 random 8-byte slot");
    run<9>("LDS.128 random 16-byte slot");
    run<8>("STS.128 random 16-byte slot");
    run<12>("STS.64 contiguous");
    run<10>("LDS.U16 random");
    run<7>("ATOMS.ADD.U32 random of 200 counters");
    run<13>("ATOMS.ADD.U32 runs of 8 lanes on 2 counters");
    return 0;
}
