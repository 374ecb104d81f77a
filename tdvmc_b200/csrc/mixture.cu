// BosonMixtureCluster (src/PhysicalSystems/BosonMixtureCluster.cpp) on the device: a few particles of several
// species per walker (config/He4He4Na.config: N = 3), so ONE THREAD owns one configuration / walker and keeps
// everything in registers and local memory; walkers are laid out exactly as for the other systems.
//
// Per pair type t: McMillan core r^m below knots[3], monomial-table cubic B-splines on a non-uniform knot vector
// (the reference's SplineFactory table, passed by the caller), constant + linear tails beyond knots[K], a log term
// for every pair; per-species hbar^2/2m; pair potentials HFDB_He_He / KTTY_He_Na / KTTY_He_Cs (src/Potentials).
// Extended sums per type: ext[t*(K+4) + j] = [ss_0..ss_{K-1} | mcMillan | const | linear | log].
//
// ORDER = 3: BosonMixtureCluster (cubic splines, four overlapping pieces, SplineFactory::GetWeights3);
// ORDER = 4: BosonMixtureCluster_4thorder (quartic, five pieces, GetWeights4; rijSplit = nodes[4], rijTail = nodes[size-5],
// BosonMixtureCluster_4thorder.cpp:138-153, :498-517, :890-895).  Knots per type: K + ORDER + 1; for ORDER = 4 the caller
// passes the reference's never-filled 28th spline as a zero spline behind one padding knot (rijTail = knots[K-1]).
#include "kernels.cuh"

namespace tdvmc
{

constexpr int kMixMaxN = 8;    // particles per walker handled by the thread-per-walker kernels
constexpr int kMixMaxExt = 96; // T * (K + 4) <= 96 (3 pair types of 26 splines)

__device__ double mix_hfdb(double r) // HFDB.cpp:23-47 with the HFDB_He_He constants
{
    const double epsil = 10.948, rm = 2.9630, av = 184431.01, alf = 10.43329537, bet = -2.27965105, dv = 1.4826,
                 c6 = 1.36745214, c8 = 0.42123807, c10 = 0.17473318;
    double fpot;
    const double x = r / rm;
    const double x2 = x * x;
    const double xm2 = 1.0 / x2;
    const double xm6 = xm2 * xm2 * xm2;
    const double xm8 = xm6 * xm2;
    const double xm10 = xm8 * xm2;
    const double f3 = c6 * xm6 + c8 * xm8 + c10 * xm10;
    const double f4 = av * exp(-alf * x + bet * x2);
    if (x >= dv) fpot = f4 - f3;
    else
    {
        const double tmp = dv / x - 1.0;
        const double f2 = exp(-(tmp * tmp));
        fpot = f4 - f3 * f2;
    }
    return epsil * fpot;
}

__device__ double mix_ktty(double r, double d, double b1, double b2, double c6, double c8, double c10) // KTTY.cpp:8-87
{
    const double epsil = 3.1577504e8;
    const double q1 = c10 / c8;
    const double c12 = q1 * q1 * q1 * c6;
    const double q2 = c12 / c10;
    const double c14 = q2 * q2 * q2 * c8;
    const double q3 = c14 / c12;
    const double c16 = q3 * q3 * q3 * c10;
    const double x = r / 0.52917721092;
    const double x2 = x * x;
    const double xm2 = 1.0 / x2;
    const double xm6 = xm2 * xm2 * xm2, xm8 = xm6 * xm2, xm10 = xm8 * xm2, xm12 = xm10 * xm2, xm14 = xm12 * xm2,
                 xm16 = xm14 * xm2;
    const double bet = b1 * x + b2 * x * x;
    const double vrep = d * exp(-bet);
    const double br = (b1 + 2.0 * b2 * x) * x;
    const double exbr = exp(-br);
    double p[17], fak[17];
    p[1] = br;
    fak[1] = 1.0;
    fak[2] = 2.0;
#pragma unroll
    for (int k = 2; k <= 16; k++) p[k] = p[k - 1] * br;
#pragma unroll
    for (int k = 3; k <= 16; k++) fak[k] = fak[k - 1] * (double)k;
    const double f6 = 1.0 - exbr * (1.0 + br + p[2] / fak[2] + p[3] / fak[3] + p[4] / fak[4] + p[5] / fak[5] + p[6] / fak[6]);
    const double f8 = f6 - exbr * (p[7] / fak[7] + p[8] / fak[8]);
    const double f10 = f8 - exbr * (p[9] / fak[9] + p[10] / fak[10]);
    const double f12 = f10 - exbr * (p[11] / fak[11] + p[12] / fak[12]);
    const double f14 = f12 - exbr * (p[13] / fak[13] + p[14] / fak[14]);
    const double f16 = f14 - exbr * (p[15] / fak[15] + p[16] / fak[16]);
    const double vatr = f6 * c6 * xm6 + f8 * c8 * xm8 + f10 * c10 * xm10 + f12 * c12 * xm12 + f14 * c14 * xm14 + f16 * c16 * xm16;
    return epsil * (vrep - vatr) / 1000;
}

__device__ double mix_potential(int id, double r)
{
    if (id == 0) return mix_hfdb(r);
    if (id == 1) return mix_ktty(r, 2.218564, 1.00872, 0.00399053, 23.768, 1307.6, 94563.2); // KTTY_He_Na.cpp
    return mix_ktty(r, 1.224951, 0.782095, 0.00513175, 41.417, 3903.4, 453443.0);             // KTTY_He_Cs.cpp
}

// lower_bound(nodes, r) - 1 on a short knot vector
__device__ __forceinline__ int mix_find_bin(const double* __restrict__ knots, int nk, double r)
{
    int lo = 0, hi = nk;
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (knots[mid] < r) lo = mid + 1;
        else hi = mid;
    }
    return lo - 1;
}

template <int ORDER>
__global__ void __launch_bounds__(64) evaluate_mix_kernel(EvalArgs a)
{
    const SysDev& s = a.s;
    const int cfg = blockIdx.x * blockDim.x + threadIdx.x;
    if (cfg >= a.n_cfg) return;
    constexpr int NP = ORDER + 1; // pieces per spline = coefficients per piece
    const int N = s.N, K = s.K, P = s.P, nk = K + ORDER + 1, EXT = K + 4, NE = s.n_ext;
    const int MC = K, CO = K + 1, LI = K + 2, LG = K + 3;

    double px[kMixMaxN], py[kMixMaxN], pz[kMixMaxN];
    double ext[kMixMaxExt];
    const double* gpos = a.pos + (size_t)cfg * 3 * s.Np;
    for (int i = 0; i < N; i++)
    {
        px[i] = gpos[i];
        py[i] = gpos[s.Np + i];
        pz[i] = gpos[2 * s.Np + i];
    }
    for (int k = 0; k < NE; k++) ext[k] = 0.0;

    double potential = 0.0, kin_r = 0.0, kin_i = 0.0, kin1 = 0.0, kin2 = 0.0;
    for (int n = 0; n < N; n++)
    {
        double fRx = 0, fRy = 0, fRz = 0, fIx = 0, fIy = 0, fIz = 0, lR = 0, lI = 0;
        for (int i = 0; i < N; i++)
        {
            if (i == n) continue;
            const int t = s.pair_type[n * N + i];
            const double* knots = s.t_knots + (size_t)t * nk;
            const double* w = s.t_weights + (size_t)t * K * NP * NP;
            const double* uR = s.utR + t * EXT;
            const double* uI = s.utI + t * EXT;
            const double m = s.t_mcm[t], rs = knots[ORDER], rt = knots[K - (ORDER - 3)];
            const double vx = px[n] - px[i], vy = py[n] - py[i], vz = pz[n] - pz[i]; // VectorDisplacement, Utils.cpp:253-263
            const double r = sqrt(vx * vx + vy * vy + vz * vz);
            const double ex = vx / r, ey = vy / r, ez = vz / r;
            double* e = ext + t * EXT;
            const bool lower = i < n;
            if (lower) potential += mix_potential(s.t_pot[t], r); // ppp.potential->GetPotential, :442-445
            double gR = 0.0, gI = 0.0; // radial derivative factors multiplying the unit vector
            if (r < rs)
            {
                // McMillan core, :459-466 (derivatives: strict '<'); values below use '<=' (:869)
                const double rp = pow(r, m - 2.0);
                const double g = m * rp;
                fRx = fma(uR[MC] * g, vx, fRx); fRy = fma(uR[MC] * g, vy, fRy); fRz = fma(uR[MC] * g, vz, fRz);
                fIx = fma(uI[MC] * g, vx, fIx); fIy = fma(uI[MC] * g, vy, fIy); fIz = fma(uI[MC] * g, vz, fIz);
                lR = fma(uR[MC], m * (m + 1.0) * rp, lR);
                lI = fma(uI[MC], m * (m + 1.0) * rp, lI);
            }
            else if (r >= rt)
            {
                gR = uR[LI]; // linear tail: gradient e, Laplacian 2/r (:467-480)
                gI = uI[LI];
                lR = fma(uR[LI], 2.0 / r, lR);
                lI = fma(uI[LI], 2.0 / r, lI);
            }
            else
            {
                const int bin = mix_find_bin(knots, nk, r);
                const double r2 = r * r;
                const double f2 = 2.0 / r;
#pragma unroll
                for (int p = 0; p < NP; p++)
                {
                    const double* q = w + ((size_t)(bin - p) * NP + p) * NP;
                    double d1 = q[1] + 2.0 * q[2] * r + 3.0 * q[3] * r2; // :489-492
                    double d2 = 2.0 * q[2] + 6.0 * q[3] * r;
                    if (ORDER == 4) // BosonMixtureCluster_4thorder.cpp:503-505 (rni3 = rni2 * rni)
                    {
                        d1 = d1 + 4.0 * q[NP - 1] * (r2 * r);
                        d2 = d2 + 12.0 * q[NP - 1] * r2;
                    }
                    gR = fma(uR[bin - p], d1, gR);
                    gI = fma(uI[bin - p], d1, gI);
                    lR = fma(uR[bin - p], d2 + f2 * d1, lR);
                    lI = fma(uI[bin - p], d2 + f2 * d1, lI);
                }
            }
            // log term for every pair: gradient e / r, Laplacian r^-2 (:515-519)
            gR = fma(uR[LG], 1.0 / r, gR);
            gI = fma(uI[LG], 1.0 / r, gI);
            lR = fma(uR[LG], 1.0 / (r * r), lR);
            lI = fma(uI[LG], 1.0 / (r * r), lI);
            fRx = fma(gR, ex, fRx); fRy = fma(gR, ey, fRy); fRz = fma(gR, ez, fRz);
            fIx = fma(gI, ex, fIx); fIy = fma(gI, ey, fIy); fIz = fma(gI, ez, fIz);
            if (lower) // value sums, CalculateWavefunction :861-891
            {
                if (r <= rs) e[MC] += pow(r, m);
                else if (r >= rt)
                {
                    e[CO] += 1.0;
                    e[LI] += r;
                }
                else
                {
                    const int bin = mix_find_bin(knots, nk, r);
                    const double r2 = r * r, r3 = r2 * r;
#pragma unroll
                    for (int p = 0; p < NP; p++)
                    {
                        const double* q = w + ((size_t)(bin - p) * NP + p) * NP;
                        double v = q[0] + q[1] * r + q[2] * r2 + q[3] * r3;
                        if (ORDER == 4) v = v + q[NP - 1] * (r2 * r2); // BosonMixtureCluster_4thorder.cpp:890-894
                        e[bin - p] += v;
                    }
                }
                e[LG] += log(r);
            }
        }
        const double nr = fRx * fRx + fRy * fRy + fRz * fRz;
        const double ni = fIx * fIx + fIy * fIy + fIz * fIz;
        const double dot = fRx * fIx + fRy * fIy + fRz * fIz;
        const double hb = s.hbar_n[n];
        kin_r += -hb * (nr - ni + lR); // :612-613
        kin_i += -hb * (2.0 * dot + lI);
        kin1 += -hb * nr;
        kin2 += -hb * lR;
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

    const long long row = a.row0 + (long long)cfg * a.row_stride;
    double* Arow = a.A + (size_t)row * a.lda;
    double exponent = 0.0;
    for (int p = 0; p < P; p++)
    {
        double o = 0.0;
        for (int j = s.map_ptr[p]; j < s.map_ptr[p + 1]; j++) o += s.map_val[j] * ext[s.map_col[j]]; // :636-645
        Arow[p] = o;
        exponent = fma(s.uR[p], o, exponent);
    }
    Arow[P] = kin_r + potential;
    Arow[P + 1] = kin_i;
    Arow[P + 2] = 1.0;
    double* o = a.other + (size_t)row * s.n_other; // :663-668
    o[0] = kin1;
    o[1] = kin2;
    o[2] = kin_r;
    o[3] = potential;
    o[4] = exp(exponent + s.phiR);
    o[5] = exponent;
    if (a.exponent) a.exponent[row] = exponent;
    if (a.outer_out) a.outer_out[cfg] = 0.0;
    if (a.ss_out)
        for (int k = 0; k < NE; k++) a.ss_out[(size_t)cfg * NE + k] = ext[k];
}

cudaError_t launch_evaluate_mix(const EvalArgs& a, cudaStream_t st)
{
    if (a.n_cfg <= 0) return cudaSuccess;
    if (a.s.N > kMixMaxN || a.s.n_ext > kMixMaxExt) return cudaErrorInvalidValue;
    if (a.s.spline_order == 4) evaluate_mix_kernel<4><<<(a.n_cfg + 63) / 64, 64, 0, st>>>(a);
    else evaluate_mix_kernel<3><<<(a.n_cfg + 63) / 64, 64, 0, st>>>(a);
    return cudaGetLastError();
}

// pair term of the exponent: sum_j u~_j phi_j(r) for pair type t (sweep form): per knot interval the ORDER + 1 Taylor
// coefficients around the left knot, then {t_lo, t_hi}
template <int ORDER>
__device__ __forceinline__ double mix_pair_u(const SysDev& s, int t, double r)
{
    const int K = s.K, nk = K + ORDER + 1, EXT = K + 4, nb = K - 2 * ORDER + 3;
    constexpr int STRIDE = ORDER + 3;
    const double* knots = s.t_knots + (size_t)t * nk;
    const double* u = s.utR + t * EXT;
    const double rs = knots[ORDER], rt = knots[K - (ORDER - 3)];
    double v;
    if (r < rs) v = u[K] * pow(r, s.t_mcm[t]);
    else if (r >= rt) v = fma(u[K + 2], r, u[K + 1]);
    else
    {
        int j = mix_find_bin(knots, nk, r) - ORDER;
        j = max(0, min(j, nb - 1));
        const double* q = s.t_cub + ((size_t)t * nb + j) * STRIDE;
        const double x = r - q[ORDER + 1];
        v = q[ORDER];
#pragma unroll
        for (int c = ORDER - 1; c >= 0; c--) v = fma(v, x, q[c]);
    }
    return fma(u[K + 3], log(r), v);
}

template <int ORDER>
__global__ void __launch_bounds__(64, 16) sweep_mix_kernel(SweepArgs a)
{
    const SysDev& s = a.s;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= a.W) return;
    const int N = s.N;
    double px[kMixMaxN], py[kMixMaxN], pz[kMixMaxN];
    double* gpos = a.pos + (size_t)w * 3 * s.Np;
    for (int i = 0; i < N; i++)
    {
        px[i] = gpos[i];
        py[i] = gpos[s.Np + i];
        pz[i] = gpos[2 * s.Np + i];
    }
    const uint32_t gw = (uint32_t)(a.first_walker + w);
    unsigned long long n_acc = 0;
    for (long long t = 0; t < a.n_steps; t++)
    {
        const Proposal pr = make_proposal(a.seed, gw, a.first_step + (uint64_t)t, N, a.mc_step);
        const int p = pr.particle;
        double ox = 0, oy = 0, oz = 0;
#pragma unroll
        for (int i = 0; i < kMixMaxN; i++)
            if (i == p)
            {
                ox = px[i];
                oy = py[i];
                oz = pz[i];
            }
        const double nx = ox + pr.dx, ny = oy + pr.dy, nz = oz + pr.dz;
        double delta = 0.0;
        for (int i = 0; i < N; i++)
        {
            if (i == p) continue;
            const int ct = s.pair_type[i * N + p];
            double dx = px[i] - ox, dy = py[i] - oy, dz = pz[i] - oz;
            const double r_old = sqrt(dx * dx + dy * dy + dz * dz);
            dx = px[i] - nx; dy = py[i] - ny; dz = pz[i] - nz;
            const double r_new = sqrt(dx * dx + dy * dy + dz * dz);
            delta += mix_pair_u<ORDER>(s, ct, r_new) - mix_pair_u<ORDER>(s, ct, r_old);
        }
        const double two_delta = 2.0 * delta;
        if ((two_delta >= pr.log_u) && (two_delta <= 709.782712893384))
        {
#pragma unroll
            for (int i = 0; i < kMixMaxN; i++)
                if (i == p)
                {
                    px[i] = nx;
                    py[i] = ny;
                    pz[i] = nz;
                }
            n_acc++;
        }
    }
    for (int i = 0; i < N; i++)
    {
        gpos[i] = px[i];
        gpos[s.Np + i] = py[i];
        gpos[2 * s.Np + i] = pz[i];
    }
    a.accepted[w] += n_acc;
}

cudaError_t launch_sweep_mix(const SweepArgs& a, cudaStream_t st)
{
    if (a.s.N > kMixMaxN) return cudaErrorInvalidValue;
    if (a.s.spline_order == 4) sweep_mix_kernel<4><<<(a.W + 63) / 64, 64, 0, st>>>(a);
    else sweep_mix_kernel<3><<<(a.W + 63) / 64, 64, 0, st>>>(a);
    return cudaGetLastError();
}

template <int ORDER>
__global__ void quotient_mix_kernel(QuotientArgs a)
{
    const SysDev& s = a.s;
    const int mv = blockIdx.x * blockDim.x + threadIdx.x;
    if (mv >= a.n_moves) return;
    const int N = s.N;
    const double* px = a.pos;
    const double* py = a.pos + s.Np;
    const double* pz = a.pos + 2 * s.Np;
    const int p = (int)a.moves[mv * 4];
    const double nx = a.moves[mv * 4 + 1], ny = a.moves[mv * 4 + 2], nz = a.moves[mv * 4 + 3];
    double delta = 0.0;
    for (int i = 0; i < N; i++)
    {
        if (i == p) continue;
        const int ct = s.pair_type[i * N + p];
        double dx = px[i] - px[p], dy = py[i] - py[p], dz = pz[i] - pz[p];
        const double r_old = sqrt(dx * dx + dy * dy + dz * dz);
        dx = px[i] - nx; dy = py[i] - ny; dz = pz[i] - nz;
        const double r_new = sqrt(dx * dx + dy * dy + dz * dz);
        delta += mix_pair_u<ORDER>(s, ct, r_new) - mix_pair_u<ORDER>(s, ct, r_old);
    }
    a.delta[mv] = delta;
}

cudaError_t launch_quotient_mix(const QuotientArgs& a, cudaStream_t st)
{
    if (a.n_moves <= 0) return cudaSuccess;
    if (a.s.spline_order == 4) quotient_mix_kernel<4><<<(a.n_moves + 63) / 64, 64, 0, st>>>(a);
    else quotient_mix_kernel<3><<<(a.n_moves + 63) / 64, 64, 0, st>>>(a);
    return cudaGetLastError();
}

// MoveCenterOfMassToZero with the mass-weighted centre of mass (BosonMixtureCluster.cpp:348-368, src/TDVMC.cpp:798-809)
__global__ void com_mix_kernel(SysDev s, double* pos, int W)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    double msum = 0.0;
    for (int i = 0; i < s.N; i++) msum += s.mass_n[i];
    for (int c = 0; c < 3; c++)
    {
        double* p = pos + ((size_t)w * 3 + c) * s.Np;
        double com = 0.0;
        for (int i = 0; i < s.N; i++) com += s.mass_n[i] * p[i];
        com /= msum;
        for (int i = 0; i < s.N; i++) p[i] -= com;
    }
}

cudaError_t launch_com_mix(const SysDev& s, double* pos, int W, cudaStream_t st)
{
    com_mix_kernel<<<(W + 127) / 128, 128, 0, st>>>(s, pos, W);
    return cudaGetLastError();
}

} // namespace tdvmc
