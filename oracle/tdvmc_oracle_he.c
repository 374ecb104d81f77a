/*
 * tdvmc_oracle_he.c — plain-C restatement of the reference's He plugins: HeBulk
 * (src/PhysicalSystems/HeBulk.cpp) and HeDrop (src/PhysicalSystems/HeDrop.cpp).
 * TEST INFRASTRUCTURE ONLY; see tdvmc_oracle.h.  Pinned by tests/test_oracle_golden.py against fixtures
 * produced by the unmodified reference.
 */
#include "tdvmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MC(s) ((s)->n_splines)
#define CONST(s) ((s)->n_splines + 1)
#define LIN(s) ((s)->n_splines + 2)
#define NEXT(s) ((s)->n_splines + 3)

/* VectorDisplacementNIC (HeBulk) or VectorDisplacement (HeDrop, Utils.cpp:253-263) */
static double displacement(const oracle_he* s, const double* a, const double* b, double* vec)
{
    if (s->periodic) return oracle_min_image(s->lbox, 3, a, b, vec);
    for (int c = 0; c < 3; c++) vec[c] = a[c] - b[c];
    return sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
}

/* knot interval and local coordinate on the short or the long grid (HeDrop.cpp:739-752) */
static int locate(const oracle_he* s, double r, double* res, double* nps)
{
    double interval;
    int bin;
    if (r < s->r_split2)
    {
        interval = (r - s->rs) / s->h_short;
        bin = (int)floor(interval);
        *res = interval - bin;
        *nps = s->h_short;
    }
    else
    {
        interval = (r - s->r_split2) / s->h_large;
        bin = (int)floor(interval);
        *res = interval - bin;
        bin = bin + s->n_short;
        *nps = s->h_large;
    }
    return bin;
}

static void add_values(const oracle_he* s, double r, double* ext)
{
    if (r < s->rs)
    {
        ext[MC(s)] += pow(r, s->mcm); /* HeBulk.cpp:473, HeDrop.cpp:730 */
    }
    else if (r >= s->r_tail)
    {
        ext[CONST(s)] += 1.0; /* HeDrop.cpp:734-736 */
        ext[LIN(s)] += r;
    }
    else
    {
        double res, nps;
        int bin = locate(s, r, &res, &nps);
        double res2 = res * res; /* pow(res, 2) */
        double res3 = pow(res, 3);
        ext[bin] += -1.0 / 6.0 * (-1.0 + 3.0 * res - 3.0 * res2 + res3); /* HeBulk.cpp:483-486 */
        ext[bin + 1] += 1.0 / 6.0 * (4.0 - 6.0 * res2 + 3.0 * res3);
        ext[bin + 2] += 1.0 / 6.0 * (1.0 + 3.0 * res + 3.0 * res2 - 3.0 * res3);
        ext[bin + 3] += 1.0 / 6.0 * res3;
    }
}

void oracle_he_values(const oracle_he* s, const double* R, double* ext)
{
    int N = s->n_particles;
    double vec[3];
    memset(ext, 0, sizeof(double) * (size_t)NEXT(s));
    for (int n = 0; n < N; n++)
        for (int i = 0; i < n; i++)
        {
            double r = displacement(s, R + 3 * (size_t)n, R + 3 * (size_t)i, vec);
            if (r < s->max_distance) add_values(s, r, ext);
        }
}

void oracle_he_operators(const oracle_he* s, const double* ext, double* O)
{
    for (int p = 0; p < s->n_params; p++)
    {
        double v = s->map_const[p];
        for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++) v += s->map_val[j] * ext[s->map_col[j]];
        O[p] = v;
    }
}

double oracle_he_exponent(const oracle_he* s, const double* ext, const double* uR)
{
    double sum = 0;
    for (int p = 0; p < s->n_params; p++)
    {
        double v = s->map_const[p];
        for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++) v += s->map_val[j] * ext[s->map_col[j]];
        sum += uR[p] * v;
    }
    return sum;
}

static double pair_potential(const oracle_he* s, double r)
{
    if (s->potential == 0)
    {
        /* Aziz HFD-B(He), HeBulk.cpp:187-195, 251-261 */
        const double e = 10.948, rm = 2.963, a = 184431.01, alpha = 10.43329537, beta = -2.27965105, d = 1.4826,
                     c6 = 1.36745214, c8 = 0.42123807, c10 = 0.17473318;
        double x = r / rm;
        double x2 = x * x;
        double xm2 = 1.0 / x2;
        double xm6 = pow(xm2, 3);
        double F = 1;
        if (x < d) F = exp(-pow(d / x - 1, 2));
        return e * (a * exp(-alpha * x + beta * x2) - F * xm6 * (c6 + xm2 * (c8 + xm2 * c10)));
    }
    /* Lennard-Jones, HeDrop.cpp:389-394 */
    double sigma = 4.0, eps = 3.56;
    double s6 = pow(sigma / r, 6);
    return 4.0 * eps * s6 * (s6 - 1.0);
}

void oracle_he_expectation(const oracle_he* s, const double* R, double wf, const double* uR, const double* uI, double* e_r,
                           double* e_i, double* other, double* drift_r, double* drift_i, double* tabD, double* tabD2)
{
    const int N = s->n_particles, P = s->n_params, NE = NEXT(s), G = s->gr_bins, NR = s->rho_bins;
    double potential = 0, R1 = 0, I1 = 0, R1I1 = 0, R2 = 0, I2 = 0;
    double vec[3], evec[3], tmp[4], vr[3], vi[3], com[3] = { 0, 0, 0 };
    double* gr = (double*)calloc((size_t)(G + NR + 1), sizeof(double));
    double* rho = gr + G;
    double* vol = (double*)calloc((size_t)G + 1, sizeof(double));
    const double gr_spacing = s->gr_max / (double)G;
    for (int i = 0; i < G; i++) vol[i] = 4.0 * M_PI * pow(gr_spacing * (i + 1), 3.0) / 3.0;
    for (int i = G - 1; i > 0; i--) vol[i] = vol[i] - vol[i - 1];
    memset(tabD, 0, sizeof(double) * (size_t)NE * N * 3);
    memset(tabD2, 0, sizeof(double) * (size_t)NE * N);
    if (NR > 0) /* GetCenterOfMass, HeDrop.cpp:255-270 */
    {
        for (int i = 0; i < N; i++)
            for (int c = 0; c < 3; c++) com[c] += R[3 * i + c];
        for (int c = 0; c < 3; c++) com[c] /= (double)N;
    }
#define TD(k, n, a) tabD[((size_t)(k) * N + (n)) * 3 + (a)]
#define TD2(k, n) tabD2[(size_t)(k) * N + (n)]
    for (int n = 0; n < N; n++)
    {
        for (int i = 0; i < N; i++)
        {
            double r = displacement(s, R + 3 * (size_t)n, R + 3 * (size_t)i, vec);
            if (r < s->max_distance)
            {
                if (i < n) potential += pair_potential(s, r);
                if (i != n)
                {
                    if (r < s->rs)
                    {
                        double rp = pow(r, s->mcm - 2.0); /* HeBulk.cpp:270 (m = -5), HeDrop.cpp:404 */
                        for (int c = 0; c < 3; c++) TD(MC(s), n, c) += s->mcm * rp * vec[c];
                        TD2(MC(s), n) += s->mcm * (s->mcm + 1.0) * rp;
                    }
                    else if (r >= s->r_tail)
                    {
                        for (int c = 0; c < 3; c++) evec[c] = vec[c] / r; /* HeDrop.cpp:411-426 */
                        for (int c = 0; c < 3; c++) TD(LIN(s), n, c) += evec[c];
                        TD2(LIN(s), n) += 2.0 / r;
                    }
                    else
                    {
                        double res, nps;
                        int bin = locate(s, r, &res, &nps);
                        double res2 = res * res;
                        double nps2 = nps * nps;
                        tmp[0] = -1.0 / 2.0 * (1.0 - 2.0 * res + res2);
                        tmp[1] = 1.0 / 6.0 * (-12.0 * res + 9.0 * res2);
                        tmp[2] = 1.0 / 6.0 * (3.0 + 6.0 * res - 9.0 * res2);
                        tmp[3] = 1.0 / 2.0 * res2;
                        for (int c = 0; c < 3; c++) evec[c] = vec[c] / r;
                        for (int c = 0; c < 3; c++)
                            for (int b = 0; b < 4; b++) TD(bin + b, n, c) += tmp[b] * evec[c] / nps;
                        TD2(bin, n) += 1.0 / nps2 * (1.0 - res) + 2.0 / (nps * r) * tmp[0];
                        TD2(bin + 1, n) += 1.0 / nps2 * (1.0 / 6.0 * (-12.0 + 18.0 * res)) + 2.0 / (nps * r) * tmp[1];
                        TD2(bin + 2, n) += 1.0 / nps2 * (1.0 / 6.0 * (6.0 - 18.0 * res)) + 2.0 / (nps * r) * tmp[2];
                        TD2(bin + 3, n) += 1.0 / nps2 * (res) + 2.0 / (nps * r) * tmp[3];
                    }
                }
            }
            if (i < n && r < s->gr_max)
            {
                int gb = (int)floor(r / gr_spacing);
                gr[gb] += 1.0 / vol[gb];
            }
        }
        if (NR > 0) /* density profile, HeDrop.cpp:482-491 */
        {
            double d0 = R[3 * n] - com[0], d1 = R[3 * n + 1] - com[1], d2 = R[3 * n + 2] - com[2];
            double rr = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            if (rr < s->gr_max)
            {
                int b = (int)floor(rr / gr_spacing);
                rho[b] += 1.0 / vol[b];
            }
        }
        for (int c = 0; c < 3; c++) vr[c] = vi[c] = 0.0;
        for (int p = 0; p < P; p++)
        {
            for (int c = 0; c < 3; c++)
            {
                double t = s->grad_const[p];
                for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++) t += s->map_val[j] * TD(s->map_col[j], n, c);
                vr[c] += uR[p] * t;
                vi[c] += uI[p] * t;
            }
            double t = 0.0;
            for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++) t += s->map_val[j] * TD2(s->map_col[j], n);
            R2 += uR[p] * t;
            I2 += uI[p] * t;
        }
        double dot = 0, nr = 0, ni = 0;
        for (int c = 0; c < 3; c++)
        {
            dot += vr[c] * vi[c];
            nr += vr[c] * vr[c];
            ni += vi[c] * vi[c];
            if (drift_r) drift_r[3 * n + c] = vr[c];
            if (drift_i) drift_i[3 * n + c] = vi[c];
        }
        R1I1 += 2.0 * dot;
        R1 += nr;
        I1 += ni;
    }
#undef TD
#undef TD2
    double kin_r = -s->hbar2_2m * (R1 - I1 + R2);
    double kin_i = -s->hbar2_2m * (R1I1 + I2);
    *e_r = kin_r + potential + 0;
    *e_i = kin_i;
    other[0] = kin_r;
    other[1] = potential;
    other[2] = wf;
    for (int i = 0; i < G; i++) other[3 + i] = gr[i];
    for (int i = 0; i < NR; i++) other[3 + G + i] = rho[i];
    free(gr);
    free(vol);
}

double oracle_he_quotient(const oracle_he* s, const double* R, int particle, const double* old_pos, const double* ext,
                          double exponent, const double* uR, double* ext_new, double* exponent_new)
{
    const int N = s->n_particles, NE = NEXT(s);
    double vec[3];
    double* so = (double*)calloc((size_t)NE, sizeof(double));
    double* sn = (double*)calloc((size_t)NE, sizeof(double));
    for (int i = 0; i < N; i++)
    {
        if (i == particle) continue;
        double r = displacement(s, R + 3 * (size_t)i, old_pos, vec);
        if (r < s->max_distance) add_values(s, r, so);
        r = displacement(s, R + 3 * (size_t)i, R + 3 * (size_t)particle, vec);
        if (r < s->max_distance) add_values(s, r, sn);
    }
    for (int k = 0; k < NE; k++) ext_new[k] = fmax(0.0, ext[k] - so[k] + sn[k]); /* HeBulk.cpp:572-576, HeDrop.cpp:898-905 */
    ext_new[CONST(s)] = ext[CONST(s)] - so[CONST(s)] + sn[CONST(s)];           /* constSum is not clamped (HeDrop.cpp:903) */
    free(so);
    free(sn);
    *exponent_new = oracle_he_exponent(s, ext_new, uR);
    return exp(2.0 * (*exponent_new - exponent));
}

int64_t oracle_he_sweep(const oracle_he* s, double* R, double* ext, double* exponent, const double* uR, uint64_t seed,
                        uint32_t walker, uint64_t first_step, int64_t n_steps, double mc_step)
{
    const int NE = NEXT(s);
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
        double q = oracle_he_quotient(s, R, p, old_pos, ext, *exponent, uR, ext_new, &exponent_new);
        int ok = 1, force = 0;
        if (!isfinite(q) || !isfinite(exponent_new) || !isfinite(*exponent)) /* src/TDVMC.cpp:886-898 */
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

int64_t oracle_he_sample_walker(const oracle_he* s, double* R, const double* uR, const double* uI, double phiR, uint64_t seed,
                                uint32_t walker, uint64_t* step_counter, int n_init, int n_samples, int n_therm,
                                double mc_step, double* est, double* sample_rows)
{
    const int N = s->n_particles, NE = NEXT(s), P = s->n_params, NO = 3 + s->gr_bins + s->rho_bins;
    double* ext = (double*)malloc(sizeof(double) * (size_t)NE);
    double* O = (double*)malloc(sizeof(double) * (size_t)P);
    double* tabD = (double*)malloc(sizeof(double) * (size_t)NE * N * 3);
    double* tabD2 = (double*)malloc(sizeof(double) * (size_t)NE * N);
    double* other = (double*)malloc(sizeof(double) * (size_t)NO);
    double exponent, e_r, e_i;
    int64_t accepted = 0;
    oracle_he_values(s, R, ext);
    exponent = oracle_he_exponent(s, ext, uR);
    accepted += oracle_he_sweep(s, R, ext, &exponent, uR, seed, walker, *step_counter, n_init, mc_step);
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
        accepted += oracle_he_sweep(s, R, ext, &exponent, uR, seed, walker, *step_counter, n_therm, mc_step);
        *step_counter += (uint64_t)n_therm;
        double wf = s->use_phi ? exp(exponent + phiR) : exp(exponent); /* HeDrop.cpp:783 / HeBulk.cpp:500 */
        oracle_he_expectation(s, R, wf, uR, uI, &e_r, &e_i, other, NULL, NULL, tabD, tabD2);
        oracle_he_operators(s, ext, O);
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
