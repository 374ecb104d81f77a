/*
 * tdvmc_oracle_he.c — plain-C restatement of the reference's HeBulk plugin (src/PhysicalSystems/HeBulk.cpp).
 * TEST INFRASTRUCTURE ONLY; see tdvmc_oracle.h.  Pinned by tests/test_oracle_golden.py against fixtures
 * produced by the unmodified reference.
 */
#include "tdvmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

void oracle_hebulk_init(oracle_hebulk* s, int n_particles, double lbox, int n_params)
{
    s->n_particles = n_particles;
    s->n_params = n_params;
    s->lbox = lbox;
    s->gr_bins = 100;                      /* HeBulk.cpp:42 */
    s->rij_split = 1.95;                   /* :48 */
    s->n_splines = n_params - 1 + 3 + 3;   /* :50 */
    double half = lbox / 2.0;
    s->h = (half - s->rij_split) / (double)(s->n_splines - 3.0); /* :53 */
    s->max_distance = half;                /* :54 */
    s->hbar2_2m = 1.0;                     /* Constants.h:12 */
    s->f[0] = 10.0 * s->h / pow(s->rij_split, 6.0);                               /* factorFirstSpline1  :59 */
    s->f[1] = 1.0;                                                                /* factorFirstSpline2  :60 */
    s->f[2] = (-5.0 * s->h + 3.0 * s->rij_split) / (2.0 * pow(s->rij_split, 6.0)); /* factorSecondSpline1 :61 */
    s->f[3] = -1.0 / 2.0;                                                         /* factorSecondSpline2 :62 */
    s->f[4] = -1.0 / 2.0;                                                         /* factorSecondLastSpline :65 */
    s->f[5] = 1.0;                                                                /* factorLastSpline :66 */
    s->f[6] = -3.0 / 2.0;                                                         /* factorSecondLastSplinePhi :67 */
    s->f[7] = 0.0;                                                                /* factorLastSplinePhi :68 */
}

static void add_values(const oracle_hebulk* s, double r, double* sums, double* mcm)
{
    if (r < s->rij_split)
    {
        *mcm += pow(r, -5.0); /* HeBulk.cpp:473 */
    }
    else
    {
        double interval = (r - s->rij_split) / s->h;
        int bin = (int)floor(interval);
        double res = interval - bin;
        double res2 = res * res; /* pow(res, 2) */
        double res3 = pow(res, 3);
        sums[bin] += -1.0 / 6.0 * (-1.0 + 3.0 * res - 3.0 * res2 + res3); /* :483-486 */
        sums[bin + 1] += 1.0 / 6.0 * (4.0 - 6.0 * res2 + 3.0 * res3);
        sums[bin + 2] += 1.0 / 6.0 * (1.0 + 3.0 * res + 3.0 * res2 - 3.0 * res3);
        sums[bin + 3] += 1.0 / 6.0 * res3;
    }
}

void oracle_hebulk_values(const oracle_hebulk* s, const double* R, double* ss, double* mcm)
{
    int N = s->n_particles;
    double vec[3];
    memset(ss, 0, sizeof(double) * (size_t)s->n_splines);
    *mcm = 0.0;
    for (int n = 0; n < N; n++)
        for (int i = 0; i < n; i++)
        {
            double r = oracle_min_image(s->lbox, 3, R + 3 * (size_t)n, R + 3 * (size_t)i, vec);
            if (r < s->max_distance) add_values(s, r, ss, mcm);
        }
}

void oracle_hebulk_operators(const oracle_hebulk* s, const double* ss, double mcm, double* O)
{
    int P = s->n_params, K = s->n_splines;
    const double* f = s->f;
    O[0] = mcm + f[0] * ss[0] + f[2] * ss[1];
    O[1] = ss[2] + f[1] * ss[0] + f[3] * ss[1];
    for (int i = 2; i < P - 2; i++) O[i] = ss[i + 1];
    O[P - 2] = (ss[K - 6] + f[4] * ss[K - 5] + f[5] * ss[K - 4]);
    O[P - 1] = (1.0 + f[6] * ss[K - 5] + f[7] * ss[K - 4]);
}

double oracle_hebulk_exponent(const oracle_hebulk* s, const double* ss, double mcm, const double* uR)
{
    int P = s->n_params, K = s->n_splines;
    const double* f = s->f;
    double sum = 0;
    sum += uR[0] * (mcm + f[0] * ss[0] + f[2] * ss[1]);
    sum += uR[1] * (ss[2] + f[1] * ss[0] + f[3] * ss[1]);
    for (int i = 2; i < P - 2; i++) sum += uR[i] * ss[i + 1];
    sum += uR[P - 2] * (ss[K - 6] + f[4] * ss[K - 5] + f[5] * ss[K - 4]);
    sum += uR[P - 1] * (1.0 + f[6] * ss[K - 5] + f[7] * ss[K - 4]);
    return sum;
}

void oracle_hebulk_expectation(const oracle_hebulk* s, const double* R, double wf, const double* uR, const double* uI,
                               double* e_r, double* e_i, double* other, double* drift_r, double* drift_i, double* sD,
                               double* sD2, double* mcD, double* mcD2)
{
    const int N = s->n_particles, P = s->n_params, K = s->n_splines, G = s->gr_bins;
    const double* f = s->f;
    const double h = s->h, h2 = pow(s->h, 2);
    /* Aziz potential constants, HeBulk.cpp:187-195 */
    const double e = 10.948, rm = 2.963, a = 184431.01, alpha = 10.43329537, beta = -2.27965105, d = 1.4826, c6 = 1.36745214,
                 c8 = 0.42123807, c10 = 0.17473318;
    double potential = 0, R1 = 0, I1 = 0, R1I1 = 0, R2 = 0, I2 = 0;
    double vec[3], evec[3], tmp[4], vr[3], vi[3], temp;
    double* gr = (double*)calloc((size_t)G, sizeof(double));
    double* vol = (double*)calloc((size_t)G, sizeof(double));
    const double gr_spacing = s->max_distance / (double)G; /* :130-131 */
    for (int i = 0; i < G; i++) vol[i] = 4.0 * M_PI * pow(gr_spacing * (i + 1), 3.0) / 3.0;
    for (int i = G - 1; i > 0; i--) vol[i] = vol[i] - vol[i - 1];
    memset(sD, 0, sizeof(double) * (size_t)K * N * 3);
    memset(sD2, 0, sizeof(double) * (size_t)K * N);
    memset(mcD, 0, sizeof(double) * (size_t)N * 3);
    memset(mcD2, 0, sizeof(double) * (size_t)N);
#define SD(k, n, a) sD[((size_t)(k) * N + (n)) * 3 + (a)]
#define SD2(k, n) sD2[(size_t)(k) * N + (n)]
    for (int n = 0; n < N; n++)
    {
        for (int i = 0; i < N; i++)
        {
            double r = oracle_min_image(s->lbox, 3, R + 3 * (size_t)n, R + 3 * (size_t)i, vec);
            if (r < s->max_distance)
            {
                if (i < n)
                {
                    double x = r / rm;
                    double x2 = x * x;
                    double xm2 = 1.0 / x2;
                    double xm6 = pow(xm2, 3);
                    double F = 1;
                    if (x < d) F = exp(-pow(d / x - 1, 2));
                    potential += e * (a * exp(-alpha * x + beta * x2) - F * xm6 * (c6 + xm2 * (c8 + xm2 * c10)));
                }
                if (i != n)
                {
                    if (r < s->rij_split)
                    {
                        double rm7 = pow(r, -7);
                        for (int c = 0; c < 3; c++) mcD[3 * n + c] += -5.0 * rm7 * vec[c];
                        mcD2[n] += 20.0 * rm7;
                    }
                    else
                    {
                        double interval = (r - s->rij_split) / h;
                        int bin = (int)floor(interval);
                        double res = interval - bin;
                        double res2 = res * res;
                        tmp[0] = -1.0 / 2.0 * (1.0 - 2.0 * res + res2);
                        tmp[1] = 1.0 / 6.0 * (-12.0 * res + 9.0 * res2);
                        tmp[2] = 1.0 / 6.0 * (3.0 + 6.0 * res - 9.0 * res2);
                        tmp[3] = 1.0 / 2.0 * res2;
                        for (int c = 0; c < 3; c++) evec[c] = vec[c] / r;
                        for (int c = 0; c < 3; c++)
                            for (int b = 0; b < 4; b++) SD(bin + b, n, c) += tmp[b] * evec[c] / h;
                        SD2(bin, n) += 1.0 / h2 * (1.0 - res) + 2.0 / (h * r) * tmp[0];
                        SD2(bin + 1, n) += 1.0 / h2 * (1.0 / 6.0 * (-12.0 + 18.0 * res)) + 2.0 / (h * r) * tmp[1];
                        SD2(bin + 2, n) += 1.0 / h2 * (1.0 / 6.0 * (6.0 - 18.0 * res)) + 2.0 / (h * r) * tmp[2];
                        SD2(bin + 3, n) += 1.0 / h2 * (res) + 2.0 / (h * r) * tmp[3];
                    }
                }
            }
            if (i < n && r < s->max_distance) /* g(r), grMaxDistance = halfLength */
            {
                int gb = (int)floor(r / gr_spacing);
                gr[gb] += 1.0 / vol[gb];
            }
        }
        for (int c = 0; c < 3; c++) vr[c] = vi[c] = 0.0;
        temp = (mcD2[n] + f[0] * SD2(0, n) + f[2] * SD2(1, n));
        R2 += uR[0] * temp;
        I2 += uI[0] * temp;
        temp = (SD2(2, n) + f[1] * SD2(0, n) + f[3] * SD2(1, n));
        R2 += uR[1] * temp;
        I2 += uI[1] * temp;
        for (int c = 0; c < 3; c++)
        {
            temp = (mcD[3 * n + c] + f[0] * SD(0, n, c) + f[2] * SD(1, n, c));
            vr[c] += uR[0] * temp;
            vi[c] += uI[0] * temp;
            temp = (SD(2, n, c) + f[1] * SD(0, n, c) + f[3] * SD(1, n, c));
            vr[c] += uR[1] * temp;
            vi[c] += uI[1] * temp;
        }
        for (int k = 2; k < P - 2; k++)
        {
            for (int c = 0; c < 3; c++)
            {
                vr[c] += uR[k] * SD(k + 1, n, c);
                vi[c] += uI[k] * SD(k + 1, n, c);
            }
            R2 += uR[k] * SD2(k + 1, n);
            I2 += uI[k] * SD2(k + 1, n);
        }
        for (int c = 0; c < 3; c++)
        {
            temp = (SD(K - 6, n, c) + f[4] * SD(K - 5, n, c) + f[5] * SD(K - 4, n, c));
            vr[c] += uR[P - 2] * temp;
            vi[c] += uI[P - 2] * temp;
            temp = (1 + f[6] * SD(K - 5, n, c) + f[7] * SD(K - 4, n, c)); /* the literal 1 of HeBulk.cpp:351 */
            vr[c] += uR[P - 1] * temp;
            vi[c] += uI[P - 1] * temp;
        }
        temp = (SD2(K - 6, n) + f[4] * SD2(K - 5, n) + f[5] * SD2(K - 4, n));
        R2 += uR[P - 2] * temp;
        I2 += uI[P - 2] * temp;
        temp = (f[6] * SD2(K - 5, n) + f[7] * SD2(K - 4, n));
        R2 += uR[P - 1] * temp;
        I2 += uI[P - 1] * temp;
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
#undef SD
#undef SD2
    double kin_r = -s->hbar2_2m * (R1 - I1 + R2);
    double kin_i = -s->hbar2_2m * (R1I1 + I2);
    *e_r = kin_r + potential + 0;
    *e_i = kin_i;
    other[0] = kin_r;
    other[1] = potential;
    other[2] = wf;
    for (int i = 0; i < G; i++) other[3 + i] = gr[i];
    free(gr);
    free(vol);
}

double oracle_hebulk_quotient(const oracle_hebulk* s, const double* R, int particle, const double* old_pos, const double* ss,
                              double mcm, double exponent, const double* uR, double* ss_new, double* mcm_new,
                              double* exponent_new)
{
    const int N = s->n_particles, K = s->n_splines;
    double vec[3], mc_old = 0, mc_new = 0;
    double* so = (double*)calloc((size_t)K, sizeof(double));
    double* sn = (double*)calloc((size_t)K, sizeof(double));
    for (int i = 0; i < N; i++)
    {
        if (i == particle) continue;
        double r = oracle_min_image(s->lbox, 3, R + 3 * (size_t)i, old_pos, vec);
        if (r < s->max_distance) add_values(s, r, so, &mc_old);
        r = oracle_min_image(s->lbox, 3, R + 3 * (size_t)i, R + 3 * (size_t)particle, vec);
        if (r < s->max_distance) add_values(s, r, sn, &mc_new);
    }
    *mcm_new = fmax(0.0, mcm - mc_old + mc_new);
    for (int k = 0; k < K; k++) ss_new[k] = fmax(0.0, ss[k] - so[k] + sn[k]);
    free(so);
    free(sn);
    *exponent_new = oracle_hebulk_exponent(s, ss_new, *mcm_new, uR);
    return exp(2.0 * (*exponent_new - exponent));
}

int64_t oracle_hebulk_sweep(const oracle_hebulk* s, double* R, double* ss, double* mcm, double* exponent, const double* uR,
                            uint64_t seed, uint32_t walker, uint64_t first_step, int64_t n_steps, double mc_step)
{
    const int K = s->n_splines;
    int64_t accepted = 0;
    double* ss_new = (double*)malloc(sizeof(double) * (size_t)K);
    for (int64_t t = 0; t < n_steps; t++)
    {
        int p;
        double disp[3], log_u, old_pos[3], mcm_new, exponent_new;
        oracle_proposal(seed, walker, first_step + (uint64_t)t, s->n_particles, mc_step, &p, disp, &log_u);
        for (int a = 0; a < 3; a++)
        {
            old_pos[a] = R[3 * p + a];
            R[3 * p + a] += disp[a];
        }
        double q = oracle_hebulk_quotient(s, R, p, old_pos, ss, *mcm, *exponent, uR, ss_new, &mcm_new, &exponent_new);
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
            memcpy(ss, ss_new, sizeof(double) * (size_t)K);
            *mcm = mcm_new;
            *exponent = exponent_new;
            accepted++;
        }
    }
    free(ss_new);
    return accepted;
}

int64_t oracle_hebulk_sample_walker(const oracle_hebulk* s, double* R, const double* uR, const double* uI, uint64_t seed,
                                    uint32_t walker, uint64_t* step_counter, int n_init, int n_samples, int n_therm,
                                    double mc_step, double* est, double* sample_rows)
{
    const int N = s->n_particles, K = s->n_splines, P = s->n_params, NO = 3 + s->gr_bins;
    double* ss = (double*)malloc(sizeof(double) * (size_t)K);
    double* O = (double*)malloc(sizeof(double) * (size_t)P);
    double* sD = (double*)malloc(sizeof(double) * (size_t)K * N * 3);
    double* sD2 = (double*)malloc(sizeof(double) * (size_t)K * N);
    double* mcD = (double*)malloc(sizeof(double) * (size_t)N * 3);
    double* mcD2 = (double*)malloc(sizeof(double) * (size_t)N);
    double* other = (double*)malloc(sizeof(double) * (size_t)NO);
    double mcm, exponent, e_r, e_i;
    int64_t accepted = 0;
    oracle_hebulk_values(s, R, ss, &mcm);
    exponent = oracle_hebulk_exponent(s, ss, mcm, uR);
    accepted += oracle_hebulk_sweep(s, R, ss, &mcm, &exponent, uR, seed, walker, *step_counter, n_init, mc_step);
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
        accepted += oracle_hebulk_sweep(s, R, ss, &mcm, &exponent, uR, seed, walker, *step_counter, n_therm, mc_step);
        *step_counter += (uint64_t)n_therm;
        oracle_hebulk_expectation(s, R, exp(exponent), uR, uI, &e_r, &e_i, other, NULL, NULL, sD, sD2, mcD, mcD2);
        oracle_hebulk_operators(s, ss, mcm, O);
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
    free(ss); free(O); free(sD); free(sD2); free(mcD); free(mcD2); free(other);
    return accepted;
}
