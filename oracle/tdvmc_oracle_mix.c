/*
 * tdvmc_oracle_mix.c — plain-C restatement of the reference's BosonMixtureCluster plugin
 * (src/PhysicalSystems/BosonMixtureCluster.cpp) and of the pair potentials it uses (src/Potentials).
 * TEST INFRASTRUCTURE ONLY; see tdvmc_oracle.h.  Pinned by tests/test_oracle_golden.py.
 */
#include "tdvmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EXT(s) ((s)->n_splines + 4)
#define MC(s) ((s)->n_splines)
#define CO(s) ((s)->n_splines + 1)
#define LI(s) ((s)->n_splines + 2)
#define LG(s) ((s)->n_splines + 3)
/* spline order: 3 = BosonMixtureCluster (GetWeights3), 4 = BosonMixtureCluster_4thorder (GetWeights4, five pieces of degree
 * four; rijSplit = nodes[4], rijTail = nodes[size - 5]: BosonMixtureCluster_4thorder.cpp:138-153).  For order 4 the spec
 * carries the reference's phantom 28th spline as a zero spline behind one padding knot, so rijTail = knots[K - 1]. */
#define ORD(s) ((s)->order == 4 ? 4 : 3)
#define NKN(s) ((s)->n_splines + ORD(s) + 1)
#define WSZ(s) ((ORD(s) + 1) * (ORD(s) + 1))
#define RS(s, knots) ((knots)[ORD(s)])
#define RT(s, knots) ((knots)[(s)->n_splines - (ORD(s) - 3)])

/* piece p of spline k: value, first and second derivative of its monomial form at r, in the reference's expressions */
static double piece_val(const oracle_mix* s, const double* q, double r)
{
    const double r2 = r * r, r3 = r2 * r;
    if (ORD(s) == 4) return q[0] + q[1] * r + q[2] * r2 + q[3] * r3 + q[4] * (r2 * r2); /* _4thorder.cpp:894 */
    return q[0] + q[1] * r + q[2] * r2 + q[3] * r3;
}
static double piece_d1(const oracle_mix* s, const double* q, double r)
{
    const double r2 = r * r;
    if (ORD(s) == 4) return q[1] + 2.0 * q[2] * r + 3.0 * q[3] * r2 + 4.0 * q[4] * (r2 * r); /* _4thorder.cpp:503 */
    return q[1] + 2.0 * q[2] * r + 3.0 * q[3] * r2;
}
static double piece_d2(const oracle_mix* s, const double* q, double r)
{
    if (ORD(s) == 4) return 2.0 * q[2] + 6.0 * q[3] * r + 12.0 * q[4] * (r * r); /* _4thorder.cpp:505 */
    return 2.0 * q[2] + 6.0 * q[3] * r;
}

static double hfdb(double r) /* HFDB.cpp:23-47 with the HFDB_He_He constants */
{
    const double epsil = 10.948, rm = 2.9630, av = 184431.01, alf = 10.43329537, bet = -2.27965105, dv = 1.4826,
                 c6 = 1.36745214, c8 = 0.42123807, c10 = 0.17473318;
    double fpot = 0;
    double x = r / rm;
    double x2 = x * x;
    double xminus2 = 1.0 / x2;
    double xminus6 = xminus2 * xminus2 * xminus2;
    double xminus8 = xminus6 * xminus2;
    double xminus10 = xminus8 * xminus2;
    double f3 = c6 * xminus6 + c8 * xminus8 + c10 * xminus10;
    double f4 = av * exp(-alf * x + bet * x2);
    if (x >= dv) fpot = f4 - f3;
    else
    {
        double tmp = dv / x - 1.0;
        double f2 = exp(-(tmp * tmp));
        fpot = f4 - f3 * f2;
    }
    return epsil * fpot;
}

static double ktty(double r, double d, double b1, double b2, double c6, double c8, double c10) /* KTTY.cpp:8-87 */
{
    const double epsil = 3.1577504e8;
    double c12 = pow(c10 / c8, 3.0) * c6;
    double c14 = pow(c12 / c10, 3.0) * c8;
    double c16 = pow(c14 / c12, 3.0) * c10;
    double fak[17];
    fak[2] = 2.0;
    for (int k = 3; k <= 16; k++) fak[k] = fak[k - 1] * (double)k;
    double x = r / 0.52917721092;
    double x2 = x * x;
    double xm2 = 1.0 / x2;
    double xm6 = xm2 * xm2 * xm2, xm8 = xm6 * xm2, xm10 = xm8 * xm2, xm12 = xm10 * xm2, xm14 = xm12 * xm2, xm16 = xm14 * xm2;
    double bet = b1 * x + b2 * x * x;
    double vrep = d * exp(-bet);
    double br = (b1 + 2.0 * b2 * x) * x;
    double exbr = exp(-br);
    double p[17];
    p[1] = br;
    for (int k = 2; k <= 16; k++) p[k] = p[k - 1] * br;
    double f6 = 1.0 - exbr * (1.0 + br + p[2] / fak[2] + p[3] / fak[3] + p[4] / fak[4] + p[5] / fak[5] + p[6] / fak[6]);
    double f8 = f6 - exbr * (p[7] / fak[7] + p[8] / fak[8]);
    double f10 = f8 - exbr * (p[9] / fak[9] + p[10] / fak[10]);
    double f12 = f10 - exbr * (p[11] / fak[11] + p[12] / fak[12]);
    double f14 = f12 - exbr * (p[13] / fak[13] + p[14] / fak[14]);
    double f16 = f14 - exbr * (p[15] / fak[15] + p[16] / fak[16]);
    double vatr = f6 * c6 * xm6 + f8 * c8 * xm8 + f10 * c10 * xm10 + f12 * c12 * xm12 + f14 * c14 * xm14 + f16 * c16 * xm16;
    double fpot = vrep - vatr;
    return epsil * fpot / 1000;
}

double oracle_pair_potential(int id, double r)
{
    if (id == ORACLE_POT_HFDB_HE_HE) return hfdb(r);
    if (id == ORACLE_POT_KTTY_HE_NA) return ktty(r, 2.218564, 1.00872, 0.00399053, 23.768, 1307.6, 94563.2); /* KTTY_He_Na.cpp */
    return ktty(r, 1.224951, 0.782095, 0.00513175, 41.417, 3903.4, 453443.0);                                   /* KTTY_He_Cs.cpp */
}

static double displacement(const double* a, const double* b, double* vec) /* Utils.cpp:253-263 */
{
    for (int c = 0; c < 3; c++) vec[c] = a[c] - b[c];
    return sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
}

static int find_bin(const double* knots, int nk, double r) /* lower_bound(nodes, r) - 1 */
{
    int lo = 0, hi = nk;
    while (lo < hi)
    {
        int mid = (lo + hi) / 2;
        if (knots[mid] < r) lo = mid + 1;
        else hi = mid;
    }
    return lo - 1;
}

/* value contributions of one pair of type t; core_inclusive: '<=' in CalculateWavefunction (:869), '<' in WFChange (:967) */
static void add_values(const oracle_mix* s, int t, double r, double* ext, int core_inclusive)
{
    const int K = s->n_splines, nk = NKN(s), np = ORD(s) + 1;
    const double* knots = s->knots + (size_t)t * nk;
    const double* w = s->weights + (size_t)t * K * WSZ(s);
    double* e = ext + (size_t)t * EXT(s);
    const double rs = RS(s, knots), rt = RT(s, knots);
    if (core_inclusive ? (r <= rs) : (r < rs))
    {
        e[MC(s)] += pow(r, s->mcm[t]);
    }
    else if (r >= rt)
    {
        e[CO(s)] += 1.0;
        e[LI(s)] += r;
    }
    else
    {
        int bin = find_bin(knots, nk, r);
        for (int p = 0; p < np; p++)
        {
            const double* q = w + ((size_t)(bin - p) * np + p) * np;
            e[bin - p] += piece_val(s, q, r);
        }
    }
    e[LG(s)] += log(r);
}

void oracle_mix_values(const oracle_mix* s, const double* R, double* ext)
{
    const int N = s->n_particles;
    double vec[3];
    memset(ext, 0, sizeof(double) * (size_t)s->n_types * EXT(s));
    for (int n = 0; n < N; n++)
        for (int i = 0; i < n; i++)
        {
            double r = displacement(R + 3 * n, R + 3 * i, vec);
            add_values(s, s->pair_type[n * N + i], r, ext, 1);
        }
}

void oracle_mix_operators(const oracle_mix* s, const double* ext, double* O)
{
    for (int p = 0; p < s->n_params; p++)
    {
        double v = 0.0;
        for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++) v += s->map_val[j] * ext[s->map_col[j]];
        O[p] = v;
    }
}

double oracle_mix_exponent(const oracle_mix* s, const double* ext, const double* uR)
{
    double sum = 0;
    for (int p = 0; p < s->n_params; p++)
    {
        double v = 0.0;
        for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++) v += s->map_val[j] * ext[s->map_col[j]];
        sum += uR[p] * v;
    }
    return sum;
}

void oracle_mix_center_of_mass(const oracle_mix* s, const double* R, double* com)
{
    double msum = 0.0;
    com[0] = com[1] = com[2] = 0.0;
    for (int i = 0; i < s->n_particles; i++)
    {
        msum += s->mass[i];
        for (int a = 0; a < 3; a++) com[a] += s->mass[i] * R[3 * i + a];
    }
    for (int a = 0; a < 3; a++) com[a] /= msum;
}

void oracle_mix_expectation(const oracle_mix* s, const double* R, double wf, double exponent, const double* uR, const double* uI,
                            double* e_r, double* e_i, double* other, double* drift_r, double* drift_i, double* tabD, double* tabD2)
{
    const int N = s->n_particles, P = s->n_params, K = s->n_splines, nk = NKN(s), NE = s->n_types * EXT(s), np = ORD(s) + 1;
    double potential = 0, kin_r = 0, kin_i = 0, kin1 = 0, kin2 = 0;
    double vec[3], evec[3], tmp1[5], tmp2[5];
    memset(tabD, 0, sizeof(double) * (size_t)NE * N * 3);
    memset(tabD2, 0, sizeof(double) * (size_t)NE * N);
#define TD(k, n, a) tabD[((size_t)(k) * N + (n)) * 3 + (a)]
#define TD2(k, n) tabD2[(size_t)(k) * N + (n)]
    for (int n = 0; n < N; n++)
    {
        for (int i = 0; i < N; i++)
        {
            const int t = s->pair_type[n * N + i];
            const double* knots = s->knots + (size_t)t * nk;
            const double* w = s->weights + (size_t)t * K * WSZ(s);
            const int base = t * EXT(s);
            const double m = s->mcm[t], rs = RS(s, knots), rt = RT(s, knots);
            double r = displacement(R + 3 * n, R + 3 * i, vec);
            if (i < n) potential += oracle_pair_potential(s->potential[t], r);
            if (i == n) continue;
            for (int a = 0; a < 3; a++) evec[a] = vec[a] / r;
            if (r < rs)
            {
                double rp = pow(r, m - 2.0);
                for (int a = 0; a < 3; a++) TD(base + MC(s), n, a) += m * rp * vec[a];
                TD2(base + MC(s), n) += m * (m + 1.0) * rp;
            }
            else if (r >= rt)
            {
                for (int a = 0; a < 3; a++) TD(base + LI(s), n, a) += evec[a];
                TD2(base + LI(s), n) += 2.0 / r;
            }
            else
            {
                int bin = find_bin(knots, nk, r);
                for (int p = 0; p < np; p++)
                {
                    const double* q = w + ((size_t)(bin - p) * np + p) * np;
                    tmp1[np - 1 - p] = piece_d1(s, q, r);
                    tmp2[np - 1 - p] = piece_d2(s, q, r);
                }
                for (int a = 0; a < 3; a++)
                    for (int b = 0; b < np; b++) TD(base + bin - b, n, a) += tmp1[np - 1 - b] * evec[a];
                for (int b = 0; b < np; b++) TD2(base + bin - b, n) += tmp2[np - 1 - b] + 2.0 / r * tmp1[np - 1 - b];
            }
            for (int a = 0; a < 3; a++) TD(base + LG(s), n, a) += 1.0 / r * evec[a]; /* :515-519 */
            TD2(base + LG(s), n) += pow(r, -2);
        }
        double vr[3] = { 0, 0, 0 }, vi[3] = { 0, 0, 0 }, R2 = 0, I2 = 0;
        for (int p = 0; p < P; p++)
        {
            for (int a = 0; a < 3; a++)
            {
                double t = 0.0;
                for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++) t += s->map_val[j] * TD(s->map_col[j], n, a);
                vr[a] += uR[p] * t;
                vi[a] += uI[p] * t;
            }
            double t = 0.0;
            for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++) t += s->map_val[j] * TD2(s->map_col[j], n);
            R2 += uR[p] * t;
            I2 += uI[p] * t;
        }
        double dot = 0, nr = 0, ni = 0;
        for (int a = 0; a < 3; a++)
        {
            dot += vr[a] * vi[a];
            nr += vr[a] * vr[a];
            ni += vi[a] * vi[a];
            if (drift_r) drift_r[3 * n + a] = vr[a];
            if (drift_i) drift_i[3 * n + a] = vi[a];
        }
        const double hb = s->hbar[n];
        kin_r += -hb * (nr - ni + R2); /* :612-613 */
        kin_i += -hb * (2.0 * dot + I2);
        kin1 += -hb * nr;
        kin2 += -hb * R2;
    }
#undef TD
#undef TD2
    *e_r = kin_r + potential + 0;
    *e_i = kin_i;
    memset(other, 0, sizeof(double) * (size_t)s->n_other);
    other[0] = kin1; /* :663-668 */
    other[1] = kin2;
    other[2] = kin_r;
    other[3] = potential;
    other[4] = wf;
    other[5] = exponent;
}

double oracle_mix_quotient(const oracle_mix* s, const double* R, int particle, const double* old_pos, const double* ext,
                           double exponent, const double* uR, double* ext_new, double* exponent_new)
{
    const int N = s->n_particles, NE = s->n_types * EXT(s);
    double vec[3];
    double* so = (double*)calloc((size_t)NE, sizeof(double));
    double* sn = (double*)calloc((size_t)NE, sizeof(double));
    for (int i = 0; i < N; i++)
    {
        if (i == particle) continue;
        const int t = s->pair_type[i * N + particle];
        add_values(s, t, displacement(R + 3 * i, old_pos, vec), so, 0);
        add_values(s, t, displacement(R + 3 * i, R + 3 * particle, vec), sn, 0);
    }
    for (int t = 0; t < s->n_types; t++)
    {
        const int b = t * EXT(s);
        for (int k = 0; k < EXT(s); k++) ext_new[b + k] = fmax(0.0, ext[b + k] - so[b + k] + sn[b + k]); /* :1020-1027 */
        ext_new[b + CO(s)] = ext[b + CO(s)] - so[b + CO(s)] + sn[b + CO(s)]; /* const and log are not clamped */
        ext_new[b + LG(s)] = ext[b + LG(s)] - so[b + LG(s)] + sn[b + LG(s)];
    }
    free(so);
    free(sn);
    *exponent_new = oracle_mix_exponent(s, ext_new, uR);
    return exp(2.0 * (*exponent_new - exponent));
}

int64_t oracle_mix_sweep(const oracle_mix* s, double* R, double* ext, double* exponent, const double* uR, uint64_t seed,
                         uint32_t walker, uint64_t first_step, int64_t n_steps, double mc_step)
{
    const int NE = s->n_types * EXT(s);
    int64_t accepted = 0;
    double* ext_new = (double*)malloc(sizeof(double) * (size_t)NE);
    for (int64_t t = 0; t < n_steps; t++)
    {
        int p;
        double disp[3], log_u, old_pos[3], exponent_new;
        oracle_proposal(seed, walker, first_step + (uint64_t)t, s->n_particles, mc_step, &p, disp, &log_u);
        for (int a = 0; a < 3; a++)
        {
            old_pos[a] = R[3 * p + a];
            R[3 * p + a] += disp[a];
        }
        double q = oracle_mix_quotient(s, R, p, old_pos, ext, *exponent, uR, ext_new, &exponent_new);
        int ok = 1, force = 0;
        if (!isfinite(q) || !isfinite(exponent_new) || !isfinite(*exponent))
        {
            ok = 0;
            if (!isfinite(q) && exponent_new > 0 && *exponent == 0)
            {
                ok = 1;
                force = 1;
            }
        }
        if (!ok || (!force && 2.0 * (exponent_new - *exponent) < log_u))
        {
            for (int a = 0; a < 3; a++) R[3 * p + a] = old_pos[a];
        }
        else
        {
            memcpy(ext, ext_new, sizeof(double) * (size_t)NE);
            *exponent = exponent_new;
            accepted++;
        }
    }
    free(ext_new);
    return accepted;
}

int64_t oracle_mix_sample_walker(const oracle_mix* s, double* R, const double* uR, const double* uI, double phiR, uint64_t seed,
                                 uint32_t walker, uint64_t* step_counter, int n_init, int n_samples, int n_therm,
                                 double mc_step, double* est, double* sample_rows)
{
    const int N = s->n_particles, NE = s->n_types * EXT(s), P = s->n_params, NO = s->n_other;
    double* ext = (double*)malloc(sizeof(double) * (size_t)NE);
    double* O = (double*)malloc(sizeof(double) * (size_t)P);
    double* tabD = (double*)malloc(sizeof(double) * (size_t)NE * N * 3);
    double* tabD2 = (double*)malloc(sizeof(double) * (size_t)NE * N);
    double* other = (double*)malloc(sizeof(double) * (size_t)NO);
    double exponent, e_r, e_i;
    int64_t accepted = 0;
    oracle_mix_values(s, R, ext);
    exponent = oracle_mix_exponent(s, ext, uR);
    accepted += oracle_mix_sweep(s, R, ext, &exponent, uR, seed, walker, *step_counter, n_init, mc_step);
    *step_counter += (uint64_t)n_init;
    double* eO = est;
    double* eER = est + P;
    double* eEI = est + P + 1;
    double* eS = est + P + 2;
    double* eOER = eS + (size_t)P * P;
    double* eOEI = eOER + P;
    double* eOther = eOEI + P;
    for (int m = 0; m < n_samples; m++)
    {
        accepted += oracle_mix_sweep(s, R, ext, &exponent, uR, seed, walker, *step_counter, n_therm, mc_step);
        *step_counter += (uint64_t)n_therm;
        oracle_mix_expectation(s, R, exp(exponent + phiR), exponent, uR, uI, &e_r, &e_i, other, NULL, NULL, tabD, tabD2);
        oracle_mix_operators(s, ext, O);
        for (int k = 0; k < P; k++)
        {
            eO[k] += O[k];
            eOER[k] += O[k] * e_r;
            eOEI[k] += O[k] * e_i;
            for (int j = 0; j < P; j++) eS[(size_t)k * P + j] += O[k] * O[j];
        }
        *eER += e_r;
        *eEI += e_i;
        for (int k = 0; k < NO; k++) eOther[k] += other[k];
        if (sample_rows)
        {
            double* row = sample_rows + (size_t)m * (P + 2);
            memcpy(row, O, sizeof(double) * (size_t)P);
            row[P] = e_r;
            row[P + 1] = e_i;
        }
    }
    free(ext); free(O); free(tabD); free(tabD2); free(other);
    return accepted;
}


/* ------------------------------------------------------------------------------------------------
 * Additional observables of the three-particle cluster: BosonMixtureCluster.cpp:680-741
 * ---------------------------------------------------------------------------------------------- */
static double mix_dist(const double* a, const double* b)
{
    const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return sqrt(x * x + y * y + z * z); /* VectorDisplacement + VectorNorm, Utils.cpp:253-263 */
}

static double mix_corner_angle(const double* r1, const double* r2, const double* r3)
{
    const double r12 = mix_dist(r1, r2), r13 = mix_dist(r1, r3), r23 = mix_dist(r2, r3);
    double angle = acos((r12 * r12 + r23 * r23 - r13 * r13) / (2 * r12 * r23)); /* Utils.cpp:393 */
    angle = angle / M_PI * 180.0;
    return angle;
}

static void mix_hist_add(double* hist, const double* grid, double value, double weight)
{
    const int bin = (int)floor(value / grid[1]); /* Grid.cpp:58-59 */
    if (bin >= 0 && bin < (int)grid[0]) hist[bin] += weight;
}

void oracle_mix_observables(const oracle_mix* s, const double* R, const double* angle_grid, const double* density_grid,
                            const double* density_scaling, const double* distance_grid, double* r2, double* angle,
                            double* density, double* distance)
{
    const int N = s->n_particles; /* 3: the reference hard-codes three observables per histogram (:331, 336, 340) */
    const int na = (int)angle_grid[0], nd = (int)density_grid[0], np = (int)distance_grid[0];
    double com[3];
    oracle_mix_center_of_mass(s, R, com);
    double sum = 0.0;
    for (int i = 0; i < N; i++)
    {
        const double r = mix_dist(R + 3 * i, com);
        sum += r * r;
    }
    *r2 = sum / (double)N;
    for (int i = 0; i < 3 * na; i++) angle[i] = 0.0;
    for (int i = 0; i < 3 * nd; i++) density[i] = 0.0;
    for (int i = 0; i < 3 * np; i++) distance[i] = 0.0;
    mix_hist_add(angle, angle_grid, mix_corner_angle(R, R + 3, R + 6), 1.0);          /* 1-2-3, :705 */
    mix_hist_add(angle + na, angle_grid, mix_corner_angle(R, R + 6, R + 3), 1.0);     /* 1-3-2, :709 */
    mix_hist_add(angle + 2 * na, angle_grid, mix_corner_angle(R + 3, R, R + 6), 1.0); /* 2-1-3, :713 */
    for (int i = 0; i < N; i++)
    {
        const double r = mix_dist(R + 3 * i, com);
        if (r < density_grid[2])
        {
            const int bin = (int)floor(r / density_grid[1]);
            if (bin < nd) density[i * nd + bin] += 1.0 / density_scaling[bin];
        }
    }
    int index = 0;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < i; j++)
        {
            const double r = mix_dist(R + 3 * i, R + 3 * j);
            if (r < distance_grid[2]) mix_hist_add(distance + index * np, distance_grid, r, 1.0);
            index++;
        }
}
