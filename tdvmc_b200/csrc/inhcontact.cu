// InhContactBosons (src/PhysicalSystems/InhContactBosons.cpp) on the device: the ONE-DIMENSIONAL inhomogeneous system of
// config/InhContactBosons*.config - a single-particle spline function ("spf") of the coordinate shifted into [0, L] plus a
// pair-correlation spline function ("pc") of the minimum-image distance (<= L/2); square-well or contact (gamma)
// interaction, lattice potential k^2 V0 sin^2(k x).  A handful of particles per walker (N = 2 ... 20 in the shipped
// configs), so ONE THREAD owns one configuration / walker, as for the three-particle mixture (mixture.cu).
//
// Walkers keep the library's [3][Np] layout; the coordinate is row 0, rows 1 and 2 are ignored.  Knots and spline table
// arrive concatenated, spf first: knots [K1+4 | K2+4], weights [K1+K2][4][4] (s.n_short = K1).  Extended sums
// ext = [ss_spf | ss_pc]; the boundary-condition map of RefreshLocalOperators (:208-247) is the CSR map.  The exponent has
// one parameter-free term, -2 gamma h_pc ss_pc[0] (:764), which also enters the REAL drift and Laplacian (:595-602):
// build_param_tables folds it into u~R of that spline, s.exp_const carries it for the exponent formed from O.
// Interval lookup is the reference's GetBinIndex (upper_bound - 1, src/Utils.cpp:110-114).
#include "kernels.cuh"

namespace tdvmc
{

constexpr int kInhMaxN = 32;   // particles per walker handled by the thread-per-walker kernels
constexpr int kInhMaxExt = 96; // K1 + K2

// upper_bound(nodes, x) - 1: knots[bin] <= x < knots[bin + 1]
__device__ __forceinline__ int inh_find_bin(const double* __restrict__ knots, int nk, double x)
{
    int lo = 0, hi = nk;
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (!(x < knots[mid])) lo = mid + 1;
        else hi = mid;
    }
    return lo - 1;
}

// GetExternalPotential (:448-509) for the four-entry SYSTEM_PARAMS
__device__ __forceinline__ double inh_external(const SysDev& s, double x0)
{
    double value = 0.0;
    if (s.ext_k > 0.0 && s.ext_v0 > 0.0)
    {
        const double x = nic_exact(x0, s.L, s.Linv, s.Lhalf) + s.Lhalf;
        const double k = s.ext_k * 3.14159265358979323846; // kf = pi (:453)
        value = sin(k * x);
        value *= value;
        value *= s.ext_v0;
        value *= k * k;
    }
    return value;
}

__global__ void __launch_bounds__(64) evaluate_inh_kernel(EvalArgs a)
{
    const SysDev& s = a.s;
    const int cfg = blockIdx.x * blockDim.x + threadIdx.x;
    if (cfg >= a.n_cfg) return;
    const int N = s.N, K1 = s.n_short, K2 = s.K - K1, P = s.P, NE = s.n_ext;
    const double* k1 = s.knots;
    const double* k2 = s.knots + K1 + 4;
    const double* w1 = s.rec;
    const double* w2 = s.rec + (size_t)K1 * 16;
    double px[kInhMaxN];
    double ext[kInhMaxExt];
    const double* gpos = a.pos + (size_t)cfg * 3 * s.Np;
    for (int i = 0; i < N; i++) px[i] = gpos[i];
    for (int k = 0; k < NE; k++) ext[k] = 0.0;

    double potExt = 0.0, potInt = 0.0, R1 = 0.0, I1 = 0.0, RI = 0.0, R2 = 0.0, I2 = 0.0;
    for (int n = 0; n < N; n++)
    {
        double fR = 0.0, fI = 0.0;
        potExt += inh_external(s, px[n]);
        {
            // single-particle function (:346-372): unit "vector" 1, second-derivative factor DIM - 1 = 0
            const double r = nic_exact(px[n], s.L, s.Linv, s.Lhalf) + s.Lhalf;
            const int bin = inh_find_bin(k1, K1 + 4, r);
            const double r2 = r * r, r3 = r2 * r;
#pragma unroll
            for (int p = 0; p < 4; p++)
            {
                const double* q = w1 + ((size_t)(bin - p) * 4 + p) * 4;
                const double d1 = q[1] + 2.0 * q[2] * r + 3.0 * q[3] * r2;
                const double d2 = 2.0 * q[2] + 6.0 * q[3] * r;
                fR = fma(s.utR[bin - p], d1, fR);
                fI = fma(s.utI[bin - p], d1, fI);
                R2 = fma(s.utR[bin - p], d2, R2);
                I2 = fma(s.utI[bin - p], d2, I2);
                ext[bin - p] += q[0] + q[1] * r + q[2] * r2 + q[3] * r3; // :264-267
            }
        }
        for (int i = 0; i < N; i++)
        {
            if (i == n) continue;
            const double v = nic_exact(px[n] - px[i], s.L, s.Linv, s.Lhalf); // VectorDisplacementNIC_1D, src/Utils.cpp:352-358
            const double rni = sqrt(v * v);
            if (!(rni <= s.rmax)) continue;
            const bool lower = i < n;
            if (lower && s.gamma == 0.0 && rni < s.pot_a) potInt += s.pot_b; // :383-394
            const int bin = inh_find_bin(k2, K2 + 4, rni);
            const double r2 = rni * rni, r3 = r2 * rni;
            double gR = 0.0, gI = 0.0;
#pragma unroll
            for (int p = 0; p < 4; p++)
            {
                const double* q = w2 + ((size_t)(bin - p) * 4 + p) * 4;
                const double d1 = q[1] + 2.0 * q[2] * rni + 3.0 * q[3] * r2; // :403-406
                const double d2 = 2.0 * q[2] + 6.0 * q[3] * rni;
                gR = fma(s.utR[K1 + bin - p], d1, gR);
                gI = fma(s.utI[K1 + bin - p], d1, gI);
                R2 = fma(s.utR[K1 + bin - p], d2, R2);
                I2 = fma(s.utI[K1 + bin - p], d2, I2);
                if (lower) ext[K1 + bin - p] += q[0] + q[1] * rni + q[2] * r2 + q[3] * r3; // :283-286
            }
            const double e = v / rni; // :410-413
            fR = fma(gR, e, fR);
            fI = fma(gI, e, fI);
        }
        RI += 2.0 * (fR * fI); // :626-628
        R1 += fR * fR;
        I1 += fI * fI;
        if (a.drift_r)
        {
            double* d = a.drift_r + ((size_t)cfg * N + n) * 3;
            d[0] = fR; d[1] = 0.0; d[2] = 0.0;
        }
        if (a.drift_i)
        {
            double* d = a.drift_i + ((size_t)cfg * N + n) * 3;
            d[0] = fI; d[1] = 0.0; d[2] = 0.0;
        }
    }

    const long long row = a.row0 + (long long)cfg * a.row_stride;
    double* Arow = a.A + (size_t)row * a.lda;
    double exponent = 0.0;
    for (int p = 0; p < P; p++)
    {
        double o = 0.0;
        for (int j = s.map_ptr[p]; j < s.map_ptr[p + 1]; j++) o += s.map_val[j] * ext[s.map_col[j]]; // :208-247
        Arow[p] = o;
        exponent = fma(s.uR[p], o, exponent);
    }
    exponent = fma(s.exp_const, ext[K1], exponent); // - 2 gamma h_pc ss_pc[0], :764
    const double kineticR = -(R1 - I1 + R2) * s.hbar; // :631-632
    const double kineticI = -(RI + I2) * s.hbar;
    Arow[P] = kineticR + potInt + potExt;
    Arow[P + 1] = kineticI;
    Arow[P + 2] = 1.0;
    double* o = a.other + (size_t)row * s.n_other; // :660-668
    o[0] = kineticR;
    o[1] = potInt;
    o[2] = exp(exponent + s.phiR);
    o[3] = exponent;
    o[4] = R1;
    o[5] = I1;
    o[6] = R2;
    o[7] = I2;
    o[8] = RI;
    if (a.exponent) a.exponent[row] = exponent;
    if (a.outer_out) a.outer_out[cfg] = 0.0;
    if (a.ss_out)
        for (int k = 0; k < NE; k++) a.ss_out[(size_t)cfg * NE + k] = ext[k];
}

cudaError_t launch_evaluate_inh(const EvalArgs& a, cudaStream_t st)
{
    if (a.n_cfg <= 0) return cudaSuccess;
    if (a.s.N > kInhMaxN || a.s.n_ext > kInhMaxExt) return cudaErrorInvalidValue;
    evaluate_inh_kernel<<<(a.n_cfg + 63) / 64, 64, 0, st>>>(a);
    return cudaGetLastError();
}

// u(x) of one of the two spline functions, sweep form: per knot interval {c0, c1, c2, c3, t_lo, t_hi} (Taylor coefficients
// around the left knot, contracted with the parameters on the host)
__device__ __forceinline__ double inh_poly(const double* __restrict__ knots, int nk, const double* __restrict__ cub, int nb,
                                           double x)
{
    int j = inh_find_bin(knots, nk, x) - 3;
    j = max(0, min(j, nb - 1));
    const double* q = cub + (size_t)j * 6;
    const double t = x - q[4];
    return fma(fma(fma(q[3], t, q[2]), t, q[1]), t, q[0]);
}

// exponent(new x of particle p) - exponent(old x)
__device__ __forceinline__ double inh_delta(const SysDev& s, const double* px, int N, int p, double xo, double xn)
{
    const int K1 = s.n_short, K2 = s.K - K1, nb1 = K1 - 3, nb2 = K2 - 3;
    const double* k1 = s.knots;
    const double* k2 = s.knots + K1 + 4;
    const double* c1 = s.cub;
    const double* c2 = s.cub + (size_t)nb1 * 6;
    double delta = inh_poly(k1, K1 + 4, c1, nb1, nic_exact(xn, s.L, s.Linv, s.Lhalf) + s.Lhalf) -
                   inh_poly(k1, K1 + 4, c1, nb1, nic_exact(xo, s.L, s.Linv, s.Lhalf) + s.Lhalf);
    for (int i = 0; i < N; i++)
    {
        if (i == p) continue;
        double v = nic_exact(px[i] - xn, s.L, s.Linv, s.Lhalf);
        const double rn = sqrt(v * v);
        v = nic_exact(px[i] - xo, s.L, s.Linv, s.Lhalf);
        const double ro = sqrt(v * v);
        const double un = rn <= s.rmax ? inh_poly(k2, K2 + 4, c2, nb2, rn) : 0.0;
        const double uo = ro <= s.rmax ? inh_poly(k2, K2 + 4, c2, nb2, ro) : 0.0;
        delta += un - uo;
    }
    return delta;
}

__global__ void __launch_bounds__(64) sweep_inh_kernel(SweepArgs a)
{
    const SysDev& s = a.s;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= a.W) return;
    const int N = s.N;
    double px[kInhMaxN];
    double* gpos = a.pos + (size_t)w * 3 * s.Np;
    for (int i = 0; i < N; i++) px[i] = gpos[i];
    const uint32_t gw = (uint32_t)(a.first_walker + w);
    unsigned long long n_acc = 0;
    for (long long t = 0; t < a.n_steps; t++)
    {
        // one coordinate per move: the first Gaussian component of the shared proposal stream (src/TDVMC.cpp:870-875)
        const Proposal pr = make_proposal(a.seed, gw, a.first_step + (uint64_t)t, N, a.mc_step);
        const int p = pr.particle;
        double xo = 0.0;
#pragma unroll 4
        for (int i = 0; i < kInhMaxN; i++)
            if (i == p) xo = px[i];
        const double xn = xo + pr.dx;
        const double two_delta = 2.0 * inh_delta(s, px, N, p, xo, xn);
        if ((two_delta >= pr.log_u) && (two_delta <= 709.782712893384))
        {
#pragma unroll 4
            for (int i = 0; i < kInhMaxN; i++)
                if (i == p) px[i] = xn;
            n_acc++;
        }
    }
    for (int i = 0; i < N; i++) gpos[i] = px[i];
    a.accepted[w] += n_acc;
}

cudaError_t launch_sweep_inh(const SweepArgs& a, cudaStream_t st)
{
    if (a.s.N > kInhMaxN) return cudaErrorInvalidValue;
    sweep_inh_kernel<<<(a.W + 63) / 64, 64, 0, st>>>(a);
    return cudaGetLastError();
}

__global__ void quotient_inh_kernel(QuotientArgs a)
{
    const SysDev& s = a.s;
    const int mv = blockIdx.x * blockDim.x + threadIdx.x;
    if (mv >= a.n_moves) return;
    const int p = (int)a.moves[mv * 4];
    a.delta[mv] = inh_delta(s, a.pos, s.N, p, a.pos[p], a.moves[mv * 4 + 1]);
}

cudaError_t launch_quotient_inh(const QuotientArgs& a, cudaStream_t st)
{
    if (a.n_moves <= 0) return cudaSuccess;
    quotient_inh_kernel<<<(a.n_moves + 63) / 64, 64, 0, st>>>(a);
    return cudaGetLastError();
}

} // namespace tdvmc
