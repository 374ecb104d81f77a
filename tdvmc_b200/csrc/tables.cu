// K3 (table form) and K4 (contraction from tables) — the reference's own decomposition.
//
// CalculateOtherLocalOperators (BosonsBulk.cpp:220-336, NUBosonsBulkPB.cpp:279-414) materialises
//     sD[k][n][a] = sum_{i != n} B'_k(r_ni) e_a ,   sD2[k][n] = sum_{i != n} B''_k(r_ni) + (D-1)/r_ni B'_k(r_ni)
// and CalculateExpectationValues (BosonsBulk.cpp:349-458) contracts them with u.  The production path
// (evaluate.cu) never forms these tables; they are kept (a) for the reference's sample-reuse
// semantics, where a stored sample carries its tables (CSDataBulkSplines.h:9-16), (b) for bit-level
// comparison of sD/sD2 with the oracle, and (c) as the HBM-bound exhibit: the table kernel has to
// write 8 K N (D+1) bytes per configuration.
//
// Device layout: T[cfg][n][k][4] = {sD_x, sD_y, sD_z, sD2}: particle-major so that one warp owns one
// particle's 32-byte-record table, builds it in shared memory and streams it out contiguously.
#include "kernels.cuh"

namespace tdvmc
{

// Warps per block: one block per SM with as many row slabs (6.5 KB each at K = 203) as fit beside the tables - 24 at N = 343
// (3.31 against 3.64 ms per 1024 configurations with two blocks of 8), 16 at N = 1728
constexpr int kTabWarpsMax = 24;

struct TabSmem
{
    double* knots;
    double* rec;
    double* px;
    double* py;
    double* pz;
    double* acc; // [warps][K][4]
    unsigned short* lut;
};

__host__ __device__ inline size_t tab_smem_layout(const SysDev& s, int nwarps, TabSmem* out, unsigned char* base)
{
    size_t off = 0;
    const int Npad = (s.N + 1) & ~1;
    TabSmem m;
    m.knots = reinterpret_cast<double*>(base + off); off += (size_t)((s.K + 4 + 1) & ~1) * 8;
    m.rec = reinterpret_cast<double*>(base + off);   off += (size_t)s.nbins * kRecStride * 8;
    m.px = reinterpret_cast<double*>(base + off);    off += (size_t)Npad * 8;
    m.py = reinterpret_cast<double*>(base + off);    off += (size_t)Npad * 8;
    m.pz = reinterpret_cast<double*>(base + off);    off += (size_t)Npad * 8;
    m.acc = reinterpret_cast<double*>(base + off);   off += (size_t)nwarps * s.K * 4 * 8;
    m.lut = reinterpret_cast<unsigned short*>(base + off);
    off += ((size_t)s.ncell * sizeof(unsigned short) + 15) & ~(size_t)15;
    if (out) *out = m;
    return off;
}

// The reference's four IEEE divisions by r_ni (BosonsBulk.cpp:304-307, 319) from ONE reciprocal, correctly rounded all the
// same (Markstein): y = RN(1 / r) - hardware seed, two Newton steps, one residual correction -, then for every numerator
// q0 = x y, q = q0 + (x - r q0) y with the residual exact in an FMA.  (A plain reciprocal-multiply is one ulp off now and then,
// which the perfect-lattice fixture - table entries cancelling to 1e-6 of their terms - shows at the 1e-13 table parity.)
// Exception: a divisor whose significand is all ones (one double in 2^52) - the Newton step cannot reach RN(1 / r) there and
// the quotient may be one ulp off.  tests/test_device_arithmetic_mirrors.py emulates this sequence in exact arithmetic and
// pins both statements; -DTDVMC_TABLES_IEEE_DIV keeps the four divisions.
struct Recip
{
    double r, y;
    __device__ __forceinline__ explicit Recip(double r_) : r(r_)
    {
        double y0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r_));
        double e = fma(-r_, y0, 1.0);
        y0 = fma(y0, e, y0);
        e = fma(-r_, y0, 1.0);
        y0 = fma(y0, e, y0);
        e = fma(-r_, y0, 1.0);
        y = fma(y0, e, y0);
    }
    __device__ __forceinline__ double divide(double x) const
    {
        const double q0 = x * y;
        return fma(fma(-r, q0, x), y, q0);
    }
};

template <bool REFLECT>
__global__ void __launch_bounds__(kTabWarpsMax * 32) tables_kernel(TableArgs a, int particles_per_block)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SysDev& s = a.s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cfg = blockIdx.x;
    const int N = s.N, K = s.K;
    const int nwarps = blockDim.x >> 5;
    TabSmem m;
    tab_smem_layout(s, nwarps, &m, smem_raw);

    for (int i = tid; i < K + 4; i += blockDim.x) m.knots[i] = s.knots[i];
    for (int i = tid; i < s.nbins * kRecStride; i += blockDim.x) m.rec[i] = s.rec[i];
    for (int i = tid; i < s.ncell; i += blockDim.x) m.lut[i] = s.lut[i];
    const double* gpos = a.pos + (size_t)cfg * 3 * s.Np;
    for (int i = tid; i < N; i += blockDim.x)
    {
        m.px[i] = gpos[i];
        m.py[i] = gpos[s.Np + i];
        m.pz[i] = gpos[2 * s.Np + i];
    }
    for (int i = tid; i < nwarps * K * 4; i += blockDim.x) m.acc[i] = 0.0;
    __syncthreads();

    double* acc = m.acc + (size_t)warp * K * 4;
    const double rmax = s.rmax;
    const int n_begin = blockIdx.y * particles_per_block;
    const int n_end = min(N, n_begin + particles_per_block);
    int vcount = 0;

    for (int n = n_begin + warp; n < n_end; n += nwarps)
    {
        const double xn = m.px[n], yn = m.py[n], zn = m.pz[n];
        for (int i0 = 0; i0 < N; i0 += 32)
        {
            const int i = i0 + lane;
            const bool have = i < N;
            double vx, vy, vz;
            double r = disp_exact(s, xn, yn, zn, have ? m.px[i] : 0.0, have ? m.py[i] : 0.0, have ? m.pz[i] : 0.0, vx, vy, vz);
            bool inside;
            if (REFLECT)
            {
                if (!(r < rmax)) r = 2 * rmax - r;
                inside = r < rmax;
            }
            else
            {
                inside = r <= rmax;
            }
            const bool act = have && (i != n) && inside;
            if (act && (i < n) && (r < s.pot_a)) vcount++; // BosonsBulk.cpp:259-271

            int bin = 0;
            double q[4][4];
            if (act)
            {
                // uniform knots: no look-up (find_bin_uniform, same interval); the branch is uniform over the grid
                bin = s.bin_guard > 0.0 ? find_bin_uniform(s, m.knots, m.lut, r) : find_bin_exact(s, m.knots, m.lut, r);
                const double* w = m.rec + (size_t)(bin - s.first_bin) * kRecStride;
                const double r2 = r * r;
                // IEEE divisions as in the reference (BosonsBulk.cpp:304-307, 319): on the perfect-lattice fixture the table
                // entries cancel to ~1e-6 of their terms, and a reciprocal-multiply (one ulp off per term; 3.30 instead of
                // 3.66 ms per 1024 configurations) fails the 1e-13 table parity there
#ifdef TDVMC_TABLES_IEEE_DIV
                const double ex = vx / r, ey = vy / r, ez = vz / r;
                const double f2 = s.dm1 / r;                        // secondDerivativeFactor / rni
#else
                const Recip rc(r);
                const double ex = rc.divide(vx), ey = rc.divide(vy), ez = rc.divide(vz);
                const double f2 = rc.divide(s.dm1);                 // secondDerivativeFactor / rni
#endif
#pragma unroll
                for (int p = 0; p < 4; p++)
                {
                    const double2 w01 = *reinterpret_cast<const double2*>(w + p * 4);
                    const double2 w23 = *reinterpret_cast<const double2*>(w + p * 4 + 2);
                    const double d1 = w01.y + 2.0 * w23.x * r + 3.0 * w23.y * r2;
                    const double d2 = 2.0 * w23.x + 6.0 * w23.y * r;
                    q[p][0] = d1 * ex;
                    q[p][1] = d1 * ey;
                    q[p][2] = d1 * ez;
                    q[p][3] = d2 + f2 * d1;
                }
            }
            // conflict-free accumulation: lanes that share a knot interval take turns (see evaluate.cu)
            const int key = act ? bin : (-1 - lane);
            const unsigned peers = __match_any_sync(FULL_MASK, key);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            for (int round = 0;; round++)
            {
                const bool mine = act && (rank == round);
                if (__ballot_sync(FULL_MASK, mine) == 0u) break;
#pragma unroll
                for (int p = 0; p < 4; p++)
                {
                    if (mine)
                    {
                        double2* t = reinterpret_cast<double2*>(acc + (size_t)(bin - p) * 4);
                        double2 t0 = t[0], t1 = t[1];
                        t0.x += q[p][0];
                        t0.y += q[p][1];
                        t1.x += q[p][2];
                        t1.y += q[p][3];
                        t[0] = t0;
                        t[1] = t1;
                    }
                    __syncwarp();
                }
            }
        }
        // stream this particle's table out (contiguous K*4 doubles) and clear it for the next one
        double2* dst = reinterpret_cast<double2*>(a.T + (((size_t)cfg * N + n) * K) * 4);
        double2* src = reinterpret_cast<double2*>(acc);
        for (int j = lane; j < K * 2; j += 32)
        {
            dst[j] = src[j];
            src[j] = make_double2(0.0, 0.0);
        }
        __syncwarp();
    }
    vcount = warp_sum_int(vcount);
    if (lane == 0 && vcount) atomicAdd(a.v_int + cfg, (double)vcount); // integer-valued: order independent
}

cudaError_t launch_tables(const TableArgs& a, cudaStream_t st)
{
    if (a.n_cfg <= 0) return cudaSuccess;
    int warps = kTabWarpsMax;
    while (warps > 8 && tab_smem_layout(a.s, warps, nullptr, nullptr) + 1024 > (size_t)227 * 1024) warps -= 8;
    const size_t smem = tab_smem_layout(a.s, warps, nullptr, nullptr);
    // split the particles of one configuration over several blocks when there are few configurations
    int chunks = 1;
    if (a.n_cfg < 148) chunks = min((a.s.N + warps - 1) / warps, (148 + a.n_cfg - 1) / a.n_cfg);
    const int ppb = (a.s.N + chunks - 1) / chunks;
    dim3 grid(a.n_cfg, (a.s.N + ppb - 1) / ppb);
    cudaError_t e = cudaMemsetAsync(a.v_int, 0, sizeof(double) * a.n_cfg, st);
    if (e != cudaSuccess) return e;
    if (a.s.pair_rule == 1)
    {
        e = cudaFuncSetAttribute(tables_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tables_kernel<true><<<grid, warps * 32, smem, st>>>(a, ppb);
    }
    else
    {
        e = cudaFuncSetAttribute(tables_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tables_kernel<false><<<grid, warps * 32, smem, st>>>(a, ppb);
    }
    return cudaGetLastError();
}

// K4: F_n = sum_k u~_k T[n][k][0..2], lap = sum_n sum_k u~_k T[n][k][3]; one block per configuration,
// one warp per particle, lanes over k (32-byte records, fully coalesced).
__global__ void __launch_bounds__(256) contract_kernel(ContractArgs a)
{
    __shared__ double red[8][5];
    const SysDev& s = a.s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cfg = blockIdx.x;
    const int N = s.N, K = s.K;
    double R1 = 0.0, I1 = 0.0, RI = 0.0, R2 = 0.0, I2 = 0.0;
    for (int n = warp; n < N; n += 8)
    {
        const double2* t = reinterpret_cast<const double2*>(a.T + (((size_t)cfg * N + n) * K) * 4);
        double fRx = 0, fRy = 0, fRz = 0, fIx = 0, fIy = 0, fIz = 0, lR = 0, lI = 0;
        for (int k = lane; k < K; k += 32)
        {
            const double2 t0 = __ldcs(t + 2 * k), t1 = __ldcs(t + 2 * k + 1);
            const double ur = s.utR[k], ui = s.utI[k];
            fRx = fma(ur, t0.x, fRx);
            fRy = fma(ur, t0.y, fRy);
            fRz = fma(ur, t1.x, fRz);
            lR = fma(ur, t1.y, lR);
            fIx = fma(ui, t0.x, fIx);
            fIy = fma(ui, t0.y, fIy);
            fIz = fma(ui, t1.x, fIz);
            lI = fma(ui, t1.y, lI);
        }
        fRx = warp_sum(fRx); fRy = warp_sum(fRy); fRz = warp_sum(fRz);
        fIx = warp_sum(fIx); fIy = warp_sum(fIy); fIz = warp_sum(fIz);
        lR = warp_sum(lR); lI = warp_sum(lI);
        R1 += fRx * fRx + fRy * fRy + fRz * fRz;
        I1 += fIx * fIx + fIy * fIy + fIz * fIz;
        RI += fRx * fIx + fRy * fIy + fRz * fIz;
        R2 += lR;
        I2 += lI;
    }
    if (lane == 0)
    {
        red[warp][0] = R1; red[warp][1] = I1; red[warp][2] = RI; red[warp][3] = R2; red[warp][4] = I2;
    }
    __syncthreads();
    if (tid == 0)
    {
        double t[5] = { 0, 0, 0, 0, 0 };
        for (int w = 0; w < 8; w++)
            for (int q = 0; q < 5; q++) t[q] += red[w][q];
        const double kRI = 2.0 * t[2];
        a.e_r[cfg] = -(t[0] - t[1] + t[3]) * s.hbar + s.pot_b * a.v_int[cfg];
        a.e_i[cfg] = -(kRI + t[4]) * s.hbar;
        if (a.sums)
        {
            double* o = a.sums + (size_t)cfg * 5;
            o[0] = t[0]; o[1] = t[1]; o[2] = t[3]; o[3] = t[4]; o[4] = kRI;
        }
    }
}

cudaError_t launch_contract(const ContractArgs& a, cudaStream_t st)
{
    if (a.n_cfg <= 0) return cudaSuccess;
    contract_kernel<<<a.n_cfg, 256, 0, st>>>(a);
    return cudaGetLastError();
}

} // namespace tdvmc
