/*
 * TEST INFRASTRUCTURE ONLY - CPU oracle, never linked into or called from the product path.
 *
 * Plain-C restatement of NUBosonsBulkPBBoxAndRadial (reference: src/PhysicalSystems/NUBosonsBulkPBBoxAndRadial.cpp),
 * the "radial + box splines" periodic system: a radial spline basis in r_ij (inside maxDistanceRad) and a "box" spline
 * basis evaluated at |x_ij|, |y_ij|, |z_ij| of the minimum-image displacement, both on the same knots.  Pinned against
 * fixtures dumped from the unmodified reference (oracle/gen_golden.py gen_boxradial -> tests/golden/boxradial_*.npz) by
 * tests/test_oracle_golden.py.  Each function cites the lines it follows; expressions keep the reference's order.
 */
#include "tdvmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* std::lower_bound(nodes, x) - nodes.begin() - 1 (:232-233, :247-248): knots[bin] < x <= knots[bin + 1] */
static int br_bin(const oracle_br* s, double x)
{
    int lo = 0, hi = s->n_splines + 4; /* first index with knots[idx] >= x */
    while (lo < hi)
    {
        int mid = (lo + hi) / 2;
        if (s->knots[mid] < x) lo = mid + 1;
        else hi = mid;
    }
    return lo - 1;
}

static double br_w(const oracle_br* s, int k, int p, int c) { return s->weights[((size_t)k * 4 + p) * 4 + c]; }

/* sum of the four overlapping spline pieces at x into sums[bin - p] (:238-241, :253-256) */
static void br_add_values(const oracle_br* s, double x, double* sums)
{
    const int bin = br_bin(s, x);
    const double x2 = x * x, x3 = x2 * x;
    for (int p = 0; p < 4; p++)
        sums[bin - p] += br_w(s, bin - p, p, 0) + br_w(s, bin - p, p, 1) * x + br_w(s, bin - p, p, 2) * x2 + br_w(s, bin - p, p, 3) * x3;
}

/* CalculateLocalOperators (:213-262): ext = [splineSumsRad (K) | splineSums (K)] over unordered pairs */
void oracle_br_values(const oracle_br* s, const double* R, double* ext)
{
    const int N = s->n_particles, K = s->n_splines;
    memset(ext, 0, sizeof(double) * 2 * (size_t)K);
    for (int n = 0; n < N; n++)
        for (int i = 0; i < n; i++)
        {
            double vec[3];
            const double rni = oracle_min_image(s->lbox, 3, R + 3 * n, R + 3 * i, vec);
            if (rni < s->r_max) br_add_values(s, rni, ext);
            for (int a = 0; a < s->dim; a++) br_add_values(s, fabs(vec[a]), ext + K);
        }
}

/* RefreshLocalOperators (:193-211) through the CSR form of the same map */
void oracle_br_operators(const oracle_br* s, const double* ext, double* O)
{
    for (int p = 0; p < s->n_params; p++)
    {
        double o = 0.0;
        for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++)
        {
            /* the reference adds ss[i+1] first, then the boundary terms, one `+=` each; "/ (-2.0)" == "* -0.5" exactly */
            o += ext[s->map_col[j]] * s->map_val[j];
        }
        O[p] = o;
    }
}

/* CalculateWavefunction (:610-621): exponent = sum_i uR[i] O[i], in index order */
double oracle_br_exponent(const oracle_br* s, const double* ext, const double* uR)
{
    double* O = (double*)malloc(sizeof(double) * (size_t)s->n_params);
    oracle_br_operators(s, ext, O);
    double sum = 0.0;
    for (int i = 0; i < s->n_params; i++) sum += uR[i] * O[i];
    free(O);
    return sum;
}

/* CalculateOtherLocalOperators (:264-425) + CalculateExpectationValues (:438-580).
 * tabD [2K][N][3], tabD2 [2K][N] (may be NULL): the four tables, radial first.  other: kinetic, potential, wf, gr[gr_bins]. */
void oracle_br_expectation(const oracle_br* s, const double* R, double wf, const double* uR, const double* uI, double* e_r,
                           double* e_i, double* other, double* drift_r, double* drift_i, double* tabD, double* tabD2)
{
    const int N = s->n_particles, K = s->n_splines, P = s->n_params, PR = P / 2;
    double* sD = (double*)calloc((size_t)2 * K * N * 3, sizeof(double));
    double* sD2 = (double*)calloc((size_t)2 * K * N, sizeof(double));
    double* gr = (double*)calloc((size_t)s->gr_bins, sizeof(double));
    double* sDr = sD;                       /* splineSumsDRad */
    double* sDb = sD + (size_t)K * N * 3;   /* splineSumsD */
    double* sD2r = sD2;
    double* sD2b = sD2 + (size_t)K * N;
    double potentialIntern = 0.0;
    const double a = s->pot_a, b = s->pot_b;

    for (int n = 0; n < N; n++)
        for (int i = 0; i < N; i++)
        {
            double vec[3], evec[3], tmp1[4], tmp2[4];
            const double rni = oracle_min_image(s->lbox, 3, R + 3 * n, R + 3 * i, vec);
            if (i < n && rni < s->gr_max) /* :312-318 */
            {
                const double grBinInterval = rni / s->gr_spacing;
                const int grBin = (int)grBinInterval;
                gr[grBin] += 1.0 / s->gr_volumes[grBin];
            }
            if (rni < s->r_max && i < n) /* :321-343 Gauss potential */
            {
                const double rnia = rni / a;
                potentialIntern += b * exp(-(rnia * rnia) / 2.0);
            }
            if (rni < s->r_max && i != n) /* :346-376 radial basis */
            {
                const int bin = br_bin(s, rni);
                const double rni2 = rni * rni;
                for (int p = 0; p < 4; p++)
                {
                    tmp1[3 - p] = br_w(s, bin - p, p, 1) + 2.0 * br_w(s, bin - p, p, 2) * rni + 3.0 * br_w(s, bin - p, p, 3) * rni2;
                    tmp2[3 - p] = 2.0 * br_w(s, bin - p, p, 2) + 6.0 * br_w(s, bin - p, p, 3) * rni;
                }
                for (int c = 0; c < 3; c++) evec[c] = vec[c] / rni;
                for (int c = 0; c < 3; c++)
                    for (int q = 0; q < 4; q++) sDr[((size_t)(bin - q) * N + n) * 3 + c] += tmp1[3 - q] * evec[c];
                const double secondDerivativeFactor = s->dim - 1.0;
                for (int q = 0; q < 4; q++) sD2r[(size_t)(bin - q) * N + n] += tmp2[3 - q] + secondDerivativeFactor / rni * tmp1[3 - q];
            }
            if (i != n) /* :378-416 box basis, per coordinate */
                for (int c = 0; c < s->dim; c++)
                {
                    const double rnia = fabs(vec[c]);
                    const int bin = br_bin(s, rnia);
                    const double rnia2 = rnia * rnia;
                    for (int p = 0; p < 4; p++)
                    {
                        tmp1[3 - p] = br_w(s, bin - p, p, 1) + 2.0 * br_w(s, bin - p, p, 2) * rnia + 3.0 * br_w(s, bin - p, p, 3) * rnia2;
                        tmp2[3 - p] = 2.0 * br_w(s, bin - p, p, 2) + 6.0 * br_w(s, bin - p, p, 3) * rnia;
                    }
                    const int sign = vec[c] < 0 ? -1 : 1;
                    for (int q = 0; q < 4; q++) sDb[((size_t)(bin - q) * N + n) * 3 + c] += tmp1[3 - q] * sign;
                    for (int q = 0; q < 4; q++) sD2b[(size_t)(bin - q) * N + n] += tmp2[3 - q];
                }
        }

    /* CalculateExpectationValues (:438-580) */
    double kineticSumR1 = 0, kineticSumI1 = 0, kineticSumR1I1 = 0, kineticSumR2 = 0, kineticSumI2 = 0;
    for (int n = 0; n < N; n++)
    {
        double vR[3] = { 0, 0, 0 }, vI[3] = { 0, 0, 0 };
#define SDR(k, c) sDr[((size_t)(k) * N + n) * 3 + (c)]
#define SDB(k, c) sDb[((size_t)(k) * N + n) * 3 + (c)]
        for (int k = 0; k < PR; k++)
        {
            for (int c = 0; c < 3; c++)
            {
                vR[c] += uR[k] * SDR(k + 1, c);
                vI[c] += uI[k] * SDR(k + 1, c);
            }
            kineticSumR2 += uR[k] * sD2r[(size_t)(k + 1) * N + n];
            kineticSumI2 += uI[k] * sD2r[(size_t)(k + 1) * N + n];
        }
        for (int c = 0; c < 3; c++)
        {
            vR[c] += uR[1] * SDR(0, c);
            vI[c] += uI[1] * SDR(0, c);
        }
        kineticSumR2 += uR[1] * sD2r[n];
        kineticSumI2 += uI[1] * sD2r[n];
        for (int c = 0; c < 3; c++)
        {
            vR[c] += uR[PR - 1] * SDR(K - 2, c) / (-2.0);
            vI[c] += uI[PR - 1] * SDR(K - 2, c) / (-2.0);
        }
        kineticSumR2 += uR[PR - 1] * sD2r[(size_t)(K - 2) * N + n] / (-2.0);
        kineticSumI2 += uI[PR - 1] * sD2r[(size_t)(K - 2) * N + n] / (-2.0);
        for (int c = 0; c < 3; c++)
        {
            vR[c] += uR[PR - 1] * SDB(K - 1, c); /* :493-497: the BOX table, as the reference has it */
            vI[c] += uI[PR - 1] * SDB(K - 1, c);
        }
        kineticSumR2 += uR[PR - 1] * sD2r[(size_t)(K - 1) * N + n];
        kineticSumI2 += uI[PR - 1] * sD2r[(size_t)(K - 1) * N + n];
        for (int k = 0; k < PR; k++)
        {
            for (int c = 0; c < 3; c++)
            {
                vR[c] += uR[k + PR] * SDB(k + 1, c);
                vI[c] += uI[k + PR] * SDB(k + 1, c);
            }
            kineticSumR2 += uR[k + PR] * sD2b[(size_t)(k + 1) * N + n];
            kineticSumI2 += uI[k + PR] * sD2b[(size_t)(k + 1) * N + n];
        }
        for (int c = 0; c < 3; c++)
        {
            vR[c] += uR[1 + PR] * SDB(0, c);
            vI[c] += uI[1 + PR] * SDB(0, c);
        }
        kineticSumR2 += uR[1 + PR] * sD2b[n];
        kineticSumI2 += uI[1 + PR] * sD2b[n];
        for (int c = 0; c < 3; c++)
        {
            vR[c] += uR[P - 1] * SDB(K - 2, c);
            vI[c] += uI[P - 1] * SDB(K - 2, c);
        }
        kineticSumR2 += uR[P - 1] * sD2b[(size_t)(K - 2) * N + n];
        kineticSumI2 += uI[P - 1] * sD2b[(size_t)(K - 2) * N + n];
        for (int c = 0; c < 3; c++)
        {
            vR[c] += uR[P - 1] * SDB(K - 1, c);
            vI[c] += uI[P - 1] * SDB(K - 1, c);
        }
        kineticSumR2 += uR[P - 1] * sD2b[(size_t)(K - 1) * N + n];
        kineticSumI2 += uI[P - 1] * sD2b[(size_t)(K - 1) * N + n];
#undef SDR
#undef SDB
        kineticSumR1I1 += 2.0 * (vR[0] * vI[0] + vR[1] * vI[1] + vR[2] * vI[2]);
        kineticSumR1 += vR[0] * vR[0] + vR[1] * vR[1] + vR[2] * vR[2];
        kineticSumI1 += vI[0] * vI[0] + vI[1] * vI[1] + vI[2] * vI[2];
        if (drift_r)
            for (int c = 0; c < 3; c++)
            {
                drift_r[3 * n + c] = vR[c];
                drift_i[3 * n + c] = vI[c];
            }
    }
    const double kineticR = -(kineticSumR1 - kineticSumI1 + kineticSumR2);
    const double kineticI = -(kineticSumR1I1 + kineticSumI2);
    *e_r = kineticR + potentialIntern + 0.0; /* otherO[0] + otherO[2] (:535) */
    *e_i = kineticI + 0.0;                   /* otherO[1] (:536) */
    other[0] = kineticR;
    other[1] = potentialIntern;
    other[2] = wf;
    for (int g = 0; g < s->gr_bins; g++) other[3 + g] = gr[g];
    if (tabD) memcpy(tabD, sD, sizeof(double) * (size_t)2 * K * N * 3);
    if (tabD2) memcpy(tabD2, sD2, sizeof(double) * (size_t)2 * K * N);
    free(sD);
    free(sD2);
    free(gr);
}

/* CalculateWFChange + CalculateWFQuotient (:632-747).  R holds the NEW position of `particle`. */
double oracle_br_quotient(const oracle_br* s, const double* R, int particle, const double* old_pos, const double* ext,
                          double exponent, const double* uR, double* ext_new, double* exponent_new)
{
    const int N = s->n_particles, K = s->n_splines;
    double* oldb = (double*)calloc((size_t)4 * K, sizeof(double)); /* sumOldPerBinRad | sumOldPerBin | sumNew... */
    double* newb = oldb + 2 * K;
    for (int i = 0; i < N; i++)
    {
        if (i == particle) continue;
        double vec[3];
        double rni = oracle_min_image(s->lbox, 3, R + 3 * i, old_pos, vec);
        if (rni < s->r_max) br_add_values(s, rni, oldb);
        for (int a = 0; a < s->dim; a++) br_add_values(s, fabs(vec[a]), oldb + K);
        rni = oracle_min_image(s->lbox, 3, R + 3 * i, R + 3 * particle, vec);
        if (rni < s->r_max) br_add_values(s, rni, newb);
        for (int a = 0; a < s->dim; a++) br_add_values(s, fabs(vec[a]), newb + K);
    }
    for (int k = 0; k < 2 * K; k++) ext_new[k] = fmax(0.0, ext[k] - oldb[k] + newb[k]); /* :713-720 */
    free(oldb);
    /* :722-735: the boundary map contracted term by term, in the reference's order */
    {
        const int PR = s->n_params / 2, P = s->n_params;
        const double* nr = ext_new;
        const double* nb = ext_new + K;
        double sum = 0.0;
        for (int i = 0; i < PR; i++) sum += uR[i] * nr[i + 1];
        sum += uR[1] * nr[0];
        sum += uR[PR - 1] * nr[K - 2] / (-2.0);
        sum += uR[PR - 1] * nr[K - 1];
        for (int i = 0; i < PR; i++) sum += uR[i + PR] * nb[i + 1];
        sum += uR[1 + PR] * nb[0];
        sum += uR[P - 1] * nb[K - 2];
        sum += uR[P - 1] * nb[K - 1];
        *exponent_new = sum;
    }
    return exp(2.0 * (*exponent_new - exponent));
}

/* DoMetropolisStep (src/TDVMC.cpp:858-916) with the shared Philox proposal stream; see oracle_sweep */
int64_t oracle_br_sweep(const oracle_br* s, double* R, double* ext, double* exponent, const double* uR, uint64_t seed,
                        uint32_t walker, uint64_t first_step, int64_t n_steps, double mc_step)
{
    const int K = s->n_splines;
    int64_t accepted = 0;
    double* ext_new = (double*)malloc(sizeof(double) * 2 * (size_t)K);
    for (int64_t t = 0; t < n_steps; t++)
    {
        int p;
        double disp[3], log_u, old_pos[3], exponent_new;
        oracle_proposal(seed, walker, first_step + (uint64_t)t, s->n_particles, mc_step, &p, disp, &log_u);
        for (int a = 0; a < 3; a++)
        {
            old_pos[a] = R[(size_t)p * 3 + a];
            if (a < s->dim) R[(size_t)p * 3 + a] += disp[a];
        }
        const double q = oracle_br_quotient(s, R, p, old_pos, ext, *exponent, uR, ext_new, &exponent_new);
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
            for (int a = 0; a < 3; a++) R[(size_t)p * 3 + a] = old_pos[a];
        }
        else
        {
            memcpy(ext, ext_new, sizeof(double) * 2 * (size_t)K); /* AcceptMove (:749-755) */
            *exponent = exponent_new;
            accepted++;
        }
    }
    free(ext_new);
    return accepted;
}

/* UpdateExpectationValues (src/TDVMC.cpp:1038-1150) for one walker; est layout as oracle_sample_walker with
 * 3 + gr_bins other values */
int64_t oracle_br_sample_walker(const oracle_br* s, double* R, const double* uR, const double* uI, double phiR, uint64_t seed,
                                uint32_t walker, uint64_t* step_counter, int n_init, int n_samples, int n_therm, double mc_step,
                                double* est, double* sample_rows)
{
    const int K = s->n_splines, P = s->n_params, NO = 3 + s->gr_bins;
    double* ext = (double*)malloc(sizeof(double) * 2 * (size_t)K);
    double* O = (double*)malloc(sizeof(double) * (size_t)P);
    double* other = (double*)malloc(sizeof(double) * (size_t)NO);
    double exponent, e_r, e_i;
    int64_t accepted = 0;
    oracle_br_values(s, R, ext);
    exponent = oracle_br_exponent(s, ext, uR);
    accepted += oracle_br_sweep(s, R, ext, &exponent, uR, seed, walker, *step_counter, n_init, mc_step);
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
        accepted += oracle_br_sweep(s, R, ext, &exponent, uR, seed, walker, *step_counter, n_therm, mc_step);
        *step_counter += (uint64_t)n_therm;
        oracle_br_operators(s, ext, O);
        oracle_br_expectation(s, R, exp(exponent + phiR), uR, uI, &e_r, &e_i, other, NULL, NULL, NULL, NULL);
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
    free(ext);
    free(O);
    free(other);
    return accepted;
}
