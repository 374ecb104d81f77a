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
