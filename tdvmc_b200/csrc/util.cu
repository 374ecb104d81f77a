// Small utility kernels: layout conversion, position wrap, test entry points, FP64 peak probes.
#include "kernels.cuh"

namespace tdvmc
{

// AlignCoordinates (src/TDVMC.cpp:2569-2582): MoveCoordinatesToFirstCell for USE_NIC systems (:787-796,
// R[i][j] = GetCoordinateNIC(R[i][j])), MoveCenterOfMassToZero for USE_MOVE_COM_TO_ZERO systems (:798-809,
// HeDrop::GetCenterOfMass, HeDrop.cpp:255-270: plain mean)
__global__ void wrap_kernel(SysDev s, double* pos, int W)
{
    const size_t total = (size_t)W * 3 * s.Np;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    {
        const int i = (int)(idx % s.Np);
        if (i < s.N) pos[idx] = nic_exact(pos[idx], s.L, s.Linv, s.Lhalf);
    }
}
__global__ void com_kernel(SysDev s, double* pos, int W)
{
    // one warp per (walker, coordinate) row
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= W * 3) return;
    double* p = pos + (size_t)row * s.Np;
    double t = 0.0;
    for (int i = lane; i < s.N; i += 32) t += p[i];
    t = warp_sum(t) / (double)s.N;
    for (int i = lane; i < s.N; i += 32) p[i] -= t;
}
cudaError_t launch_wrap(const SysDev& s, double* pos, int W, cudaStream_t st)
{
    if (s.periodic) wrap_kernel<<<296, 256, 0, st>>>(s, pos, W);
    else com_kernel<<<(W * 3 * 32 + 255) / 256, 256, 0, st>>>(s, pos, W);
    return cudaGetLastError();
}

__global__ void min_image_kernel(SysDev s, const double* a, const double* b, int n, double* norm, double* disp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double vx, vy, vz;
    norm[i] = disp_exact(s, a[3 * i], a[3 * i + 1], a[3 * i + 2], b[3 * i], b[3 * i + 1], b[3 * i + 2], vx, vy, vz);
    disp[3 * i] = vx;
    disp[3 * i + 1] = vy;
    disp[3 * i + 2] = vz;
}
cudaError_t launch_min_image(const SysDev& s0, double L, const double* a, const double* b, int n, double* norm,
                             double* disp, cudaStream_t st)
{
    SysDev s = s0;
    s.L = L;
    s.Linv = 1.0 / L; // src/TDVMC.cpp:535-536
    s.Lhalf = L / 2.0;
    min_image_kernel<<<(n + 127) / 128, 128, 0, st>>>(s, a, b, n, norm, disp);
    return cudaGetLastError();
}

__global__ void proposals_kernel(uint64_t seed, uint32_t walker, uint64_t first_step, int n, int n_particles,
                                 double mc_step, int* particle, double* disp, double* log_u)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Proposal p = make_proposal(seed, walker, first_step + (uint64_t)i, n_particles, mc_step);
    particle[i] = p.particle;
    disp[3 * i] = p.dx;
    disp[3 * i + 1] = p.dy;
    disp[3 * i + 2] = p.dz;
    log_u[i] = p.log_u;
}
cudaError_t launch_proposals(uint64_t seed, uint32_t walker, uint64_t first_step, int n, int n_particles, double mc_step,
                             int* particle, double* disp, double* log_u, cudaStream_t st)
{
    proposals_kernel<<<(n + 127) / 128, 128, 0, st>>>(seed, walker, first_step, n, n_particles, mc_step, particle, disp,
                                                      log_u);
    return cudaGetLastError();
}

// host layout R[c][n][3] <-> device layout pos[c][3][Np]
__global__ void transpose_in_kernel(const double* aos, double* soa, int n_cfg, int N, int Np)
{
    const size_t total = (size_t)n_cfg * N * 3;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    {
        const size_t c = idx / ((size_t)N * 3);
        const int rem = (int)(idx % ((size_t)N * 3));
        const int n = rem / 3, d = rem % 3;
        soa[(c * 3 + d) * Np + n] = aos[idx];
    }
}
__global__ void transpose_out_kernel(const double* soa, double* aos, int n_cfg, int N, int Np)
{
    const size_t total = (size_t)n_cfg * N * 3;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    {
        const size_t c = idx / ((size_t)N * 3);
        const int rem = (int)(idx % ((size_t)N * 3));
        const int n = rem / 3, d = rem % 3;
        aos[idx] = soa[(c * 3 + d) * Np + n];
    }
}
cudaError_t launch_transpose_in(const double* aos, double* soa, int n_cfg, int N, int Np, cudaStream_t st)
{
    transpose_in_kernel<<<592, 256, 0, st>>>(aos, soa, n_cfg, N, Np);
    return cudaGetLastError();
}
cudaError_t launch_transpose_out(const double* soa, double* aos, int n_cfg, int N, int Np, cudaStream_t st)
{
    transpose_out_kernel<<<592, 256, 0, st>>>(soa, aos, n_cfg, N, Np);
    return cudaGetLastError();
}

// A[m] = [O_m | e_r | e_i | 1] from separate arrays (accumulate_fixed)
__global__ void fill_rows_kernel(double* A, int lda, int P, const double* O, const double* e_r, const double* e_i,
                                 long long M)
{
    const size_t total = (size_t)M * (P + 3);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    {
        const size_t m = idx / (P + 3);
        const int c = (int)(idx % (P + 3));
        double v;
        if (c < P) v = O[m * P + c];
        else if (c == P) v = e_r[m];
        else if (c == P + 1) v = e_i[m];
        else v = 1.0;
        A[m * lda + c] = v;
    }
}
cudaError_t launch_fill_rows(double* A, int lda, int P, const double* O, const double* e_r, const double* e_i, long long M,
                             cudaStream_t st)
{
    fill_rows_kernel<<<592, 256, 0, st>>>(A, lda, P, O, e_r, e_i, M);
    return cudaGetLastError();
}

// ---- FP64 peak probes: the driver's MEASURED_PEAKS.json has no FP64 entry ----
__global__ void dfma_probe_kernel(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; i++)
    {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void dmma_probe_kernel(double* out, int iters)
{
    double c[8][2];
    for (int j = 0; j < 8; j++) c[j][0] = c[j][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int i = 0; i < iters; i++)
    {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[j][0]), "+d"(c[j][1])
                         : "d"(a), "d"(b));
    }
    double t = 0;
    for (int j = 0; j < 8; j++) t += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
cudaError_t measure_fp64(double* dfma_tflops, double* dmma_tflops, cudaStream_t st)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 4, threads = 512, iters = 20000;
    double* out = nullptr;
    cudaError_t e = cudaMalloc(&out, sizeof(double) * blocks * threads);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0.f;
    dfma_probe_kernel<<<blocks, threads, 0, st>>>(out, 1000);
    cudaEventRecord(e0, st);
    dfma_probe_kernel<<<blocks, threads, 0, st>>>(out, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    *dfma_tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    dmma_probe_kernel<<<blocks, threads, 0, st>>>(out, 1000);
    cudaEventRecord(e0, st);
    dmma_probe_kernel<<<blocks, threads, 0, st>>>(out, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    // one m8n8k4 = 8*8*4 FMA = 512 flop per warp
    *dmma_tflops = 512.0 * 8.0 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    e = cudaGetLastError();
    cudaFree(out);
    return e;
}

} // namespace tdvmc
