// Additional observables g(r) and S(k), one block per configuration.
//
// Replaces IPhysicalSystem::CalculateAdditionalSystemProperties of the bulk spline systems
// (BosonsBulk.cpp:474-520, NUBosonsBulkPB.cpp:597-639), which the reference calls once per sample of its
// end-of-run pass (src/TDVMC.cpp:1332-1388):
//   g(r): every pair i > j with r < grid.max adds weight / scaling[bin], bin = floor(r / spacing)
//         (Grid.cpp:52-61, ObservableVsOnGridWithScaling.cpp:47-52).  The device counts pairs per bin as
//         integers (exact, order-free); the host multiplies by weight / scaling[bin].
//   S(k): sk[k] = ((sum_{i,kn} cos k_kn.R_i)^2 + (sum_{i,kn} sin k_kn.R_i)^2) / (N n_k): thread q owns wave
//         vector q and walks the particles; a shell's vectors are then summed in index order (deterministic).
// Accumulate mode adds into the configuration's own rows (per-walker sums over samples), so no two blocks
// touch the same address; obs_reduce_kernel sums the rows over walkers in a fixed order.
#include "kernels.cuh"

namespace tdvmc
{

__global__ void __launch_bounds__(256) observables_kernel(ObsArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SysDev& s = a.s;
    const int N = s.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int cfg = blockIdx.x;
    double* px = reinterpret_cast<double*>(smem_raw);
    double* py = px + N;
    double* pz = py + N;
    double* kc = pz + N;               // [n_kvec] cosine sums
    double* ks = kc + a.n_kvec;        // [n_kvec] sine sums
    unsigned int* cnt = reinterpret_cast<unsigned int*>(ks + a.n_kvec); // [gr_count]

    const double* gpos = a.pos + (size_t)cfg * 3 * s.Np;
    for (int i = tid; i < N; i += blockDim.x)
    {
        px[i] = gpos[i];
        py[i] = gpos[s.Np + i];
        pz[i] = gpos[2 * s.Np + i];
    }
    for (int b = tid; b < a.gr_count; b += blockDim.x) cnt[b] = 0u;
    __syncthreads();

    // pair distribution
    if (a.gr_count > 0)
    {
        for (int i = warp; i < N; i += nwarps)
        {
            const double xi = px[i], yi = py[i], zi = pz[i];
            for (int j = lane; j < i; j += 32)
            {
                double r, vx, vy, vz;
                if (s.periodic) r = disp_exact(s, xi, yi, zi, px[j], py[j], pz[j], vx, vy, vz); // BosonsBulk.cpp:486
                else
                {
                    vx = xi - px[j];
                    vy = yi - py[j];
                    vz = zi - pz[j];
                    r = sqrt(vx * vx + vy * vy + vz * vz);
                }
                if (r < a.gr_max)
                {
                    const int bin = (int)floor(r / a.gr_spacing); // Grid.cpp:58-59
                    if (bin < a.gr_count) atomicAdd(&cnt[bin], 1u); // beyond count: the reference writes past its vector
                }
            }
        }
    }

    // structure factor
    for (int q = tid; q < a.n_kvec; q += blockDim.x)
    {
        const double kx = a.kvec[3 * q], ky = a.kvec[3 * q + 1], kz = a.kvec[3 * q + 2];
        double c = 0.0, sn = 0.0;
        for (int i = 0; i < N; i++)
        {
            const double arg = kx * px[i] + ky * py[i] + kz * pz[i]; // NUBosonsBulkPB.cpp:629; VectorDotProduct_DIM
            double sv, cv;
            sincos(arg, &sv, &cv);
            c += cv;
            sn += sv;
        }
        kc[q] = c;
        ks[q] = sn;
    }
    __syncthreads();

    for (int b = tid; b < a.gr_count; b += blockDim.x)
    {
        unsigned long long* row = a.gr_rows + (size_t)cfg * a.gr_count;
        row[b] = (a.accumulate ? row[b] : 0ull) + (unsigned long long)cnt[b];
    }
    for (int k = tid; k < a.n_shells; k += blockDim.x)
    {
        double c = 0.0, sn = 0.0;
        const int q0 = a.shell_ptr[k], q1 = a.shell_ptr[k + 1];
        for (int q = q0; q < q1; q++)
        {
            c += kc[q];
            sn += ks[q];
        }
        const double sk = (c * c + sn * sn) / ((double)(N * (q1 - q0))); // BosonsBulk.cpp:517
        double* row = a.sk_rows + (size_t)cfg * a.n_shells;
        row[k] = (a.accumulate ? row[k] : 0.0) + sk;
    }
}

// out[0..gr_count) = sum over rows of the pair counts, out[gr_count..gr_count+n_shells) = sum of the sk rows
__global__ void obs_reduce_kernel(const unsigned long long* gr_rows, const double* sk_rows, int n_rows, int gr_count,
                                  int n_shells, double* out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < gr_count)
    {
        unsigned long long t = 0ull;
        for (int r = 0; r < n_rows; r++) t += gr_rows[(size_t)r * gr_count + c];
        out[c] = (double)t;
    }
    else if (c < gr_count + n_shells)
    {
        const int k = c - gr_count;
        double t = 0.0;
        for (int r = 0; r < n_rows; r++) t += sk_rows[(size_t)r * n_shells + k];
        out[c] = t;
    }
}

// ---- three-particle cluster observables (BosonMixtureCluster.cpp:680-741) --------------------------
// One thread per walker; the block collects integer histograms in shared memory and adds them to the global
// 64-bit counters once (integer sums: exact, order-free).  r2 goes to the walker's own slot (summed later in
// a fixed order).  Histogram layout: [angle 3 x na | density 3 x nd | distance 3 x np].
__device__ __forceinline__ double cl_dist(const double* a, const double* b)
{
    const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return sqrt(x * x + y * y + z * z); // VectorDisplacement, Utils.cpp:253-263
}

__device__ __forceinline__ double cl_corner_angle(const double* r1, const double* r2, const double* r3)
{
    const double r12 = cl_dist(r1, r2), r13 = cl_dist(r1, r3), r23 = cl_dist(r2, r3);
    double angle = acos((r12 * r12 + r23 * r23 - r13 * r13) / (2 * r12 * r23)); // Utils.cpp:393
    angle = angle / 3.14159265358979323846 * 180.0;
    return angle;
}

__device__ __forceinline__ void cl_count(unsigned int* hist, int count, double spacing, double value)
{
    const int bin = (int)floor(value / spacing); // Grid.cpp:58-59
    if (bin >= 0 && bin < count) atomicAdd(&hist[bin], 1u);
}

__global__ void __launch_bounds__(128) cluster_observables_kernel(ClusterObsArgs a)
{
    extern __shared__ unsigned int cl_hist[];
    const SysDev& s = a.s;
    const int nh = 3 * (a.n_angle + a.n_density + a.n_distance);
    for (int i = threadIdx.x; i < nh; i += blockDim.x) cl_hist[i] = 0u;
    __syncthreads();
    unsigned int* h_angle = cl_hist;
    unsigned int* h_density = h_angle + 3 * a.n_angle;
    unsigned int* h_distance = h_density + 3 * a.n_density;
    const int w = a.per_cfg_hist ? (threadIdx.x == 0 ? (int)blockIdx.x : a.n_cfg) : (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (w < a.n_cfg)
    {
        double R[3][3];
        const double* p = a.pos + (size_t)w * 3 * s.Np;
        for (int i = 0; i < 3; i++)
            for (int c = 0; c < 3; c++) R[i][c] = p[(size_t)c * s.Np + i];
        double msum = 0.0, com[3] = { 0.0, 0.0, 0.0 }; // GetCenterOfMass, BosonMixtureCluster.cpp:348-368
        for (int i = 0; i < 3; i++)
        {
            msum += s.mass_n[i];
            for (int c = 0; c < 3; c++) com[c] += s.mass_n[i] * R[i][c];
        }
        for (int c = 0; c < 3; c++) com[c] /= msum;
        double r2sum = 0.0;
        for (int i = 0; i < 3; i++)
        {
            const double r = cl_dist(R[i], com);
            r2sum += r * r;                                                      // :690-697
            if (r < a.density_max) cl_count(h_density + i * a.n_density, a.n_density, a.density_spacing, r); // :717-724
        }
        a.r2_rows[w] = (a.accumulate ? a.r2_rows[w] : 0.0) + r2sum / 3.0;
        cl_count(h_angle, a.n_angle, a.angle_spacing, cl_corner_angle(R[0], R[1], R[2]));                 // 1-2-3
        cl_count(h_angle + a.n_angle, a.n_angle, a.angle_spacing, cl_corner_angle(R[0], R[2], R[1]));     // 1-3-2
        cl_count(h_angle + 2 * a.n_angle, a.n_angle, a.angle_spacing, cl_corner_angle(R[1], R[0], R[2])); // 2-1-3
        int index = 0;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < i; j++)
            {
                const double r = cl_dist(R[i], R[j]);
                if (r < a.distance_max) cl_count(h_distance + index * a.n_distance, a.n_distance, a.distance_spacing, r); // :728-739
                index++;
            }
    }
    __syncthreads();
    if (a.per_cfg_hist) // parity entry point: one configuration per block
    {
        for (int i = threadIdx.x; i < nh; i += blockDim.x) a.hist[(size_t)blockIdx.x * nh + i] = cl_hist[i];
    }
    else
    {
        for (int i = threadIdx.x; i < nh; i += blockDim.x)
            if (cl_hist[i]) atomicAdd(&a.hist[i], (unsigned long long)cl_hist[i]);
    }
}

// sum of n doubles in a fixed order (deterministic): 256 strided partial sums, then a fixed tree
__global__ void __launch_bounds__(256) sum_rows_kernel(const double* rows, int n, double* out)
{
    __shared__ double red[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) t += rows[i];
    red[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1)
    {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

cudaError_t launch_cluster_observables(const ClusterObsArgs& a, cudaStream_t st)
{
    if (a.n_cfg <= 0) return cudaSuccess;
    const size_t smem = (size_t)3 * (a.n_angle + a.n_density + a.n_distance) * sizeof(unsigned int);
    const int threads = a.per_cfg_hist ? 32 : 128;
    const int blocks = a.per_cfg_hist ? a.n_cfg : (a.n_cfg + threads - 1) / threads;
    cluster_observables_kernel<<<blocks, threads, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_sum_rows(const double* rows, int n, double* out, cudaStream_t st)
{
    sum_rows_kernel<<<1, 256, 0, st>>>(rows, n, out);
    return cudaGetLastError();
}

size_t observables_smem_bytes(const SysDev& s, int n_kvec, int gr_count)
{
    return (size_t)(3 * s.N + 2 * n_kvec) * sizeof(double) + (size_t)gr_count * sizeof(unsigned int);
}

cudaError_t launch_observables(const ObsArgs& a, cudaStream_t st)
{
    if (a.n_cfg <= 0) return cudaSuccess;
    const size_t smem = observables_smem_bytes(a.s, a.n_kvec, a.gr_count);
    cudaError_t e = cudaFuncSetAttribute(observables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    observables_kernel<<<a.n_cfg, 256, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_obs_reduce(const unsigned long long* gr_rows, const double* sk_rows, int n_rows, int gr_count, int n_shells,
                              double* out, cudaStream_t st)
{
    const int n = gr_count + n_shells;
    if (n <= 0) return cudaSuccess;
    obs_reduce_kernel<<<(n + 127) / 128, 128, 0, st>>>(gr_rows, sk_rows, n_rows, gr_count, n_shells, out);
    return cudaGetLastError();
}

} // namespace tdvmc
