/*
 * TEST INFRASTRUCTURE ONLY - CPU oracle, never linked into or called from the product path.
 *
 * Plain-C restatement of InhContactBosons (reference: src/PhysicalSystems/InhContactBosons.cpp), the ONE-DIMENSIONAL
 * inhomogeneous system of the only shipped v0.2x configs: a single-particle spline function ("spf") of the coordinate in
 * [0, L] plus a pair-correlation spline function ("pc") of the minimum-image distance in [0, L/2]; square-well or contact
 * (gamma) interaction, k^2 V_0 sin^2(k x) lattice potential.  R is [N][3] like everywhere else, coordinate in component 0.
 * Extended sums ext = [ss_spf (K1) | ss_pc (K2)].  Pinned against fixtures dumped from the unmodified reference
 * (oracle/gen_golden.py gen_inhcontact -> tests/golden/inhcontact_*.npz) by tests/test_oracle_golden.py.
 */
#include "tdvmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* GetBinIndex (src/Utils.cpp:110-114): upper_bound(list, value) - begin - 1, i.e. knots[bin] <= value < knots[bin + 1] */
static int inh_bin(const double* knots, int nk, double x)
{
    int lo = 0, hi = nk; /* first index with knots[idx] > x */
    while (lo < hi)
    {
        int mid = (lo + hi) / 2;
        if (!(x < knots[mid])) lo = mid + 1;
        else hi = mid;
    }
    return lo - 1;
}

/* GetCoordinateNIC (src/Utils.cpp:266-281) */
static double inh_nic(double r, double L)
{
    const double Linv = 1.0 / L, Lhalf = L / 2.0;
    int k = (int)(r * Linv + ((r >= 0.0) ? 0.5 : -0.5));
    double result = r - k * L;
    if (result == Lhalf) result -= 1e-10;
    else if (result == -Lhalf) result += 1e-10;
    return result;
}

static void inh_add_values(const double* knots, int nk, const double* w, double x, double* sums)
{
    const int bin = inh_bin(knots, nk, x);
    const double x2 = x * x, x3 = x2 * x;
    for (int p = 0; p < 4; p++)
    {
        const double* q = w + ((size_t)(bin - p) * 4 + p) * 4;
        sums[bin - p] += q[0] + q[1] * x + q[2] * x2 + q[3] * x3;
    }
}

/* CalculateLocalOperators (InhContactBosons.cpp:249-304) */
void oracle_inh_values(const oracle_inh* s, const double* R, double* ext)
{
    const int N = s->n_particles, K1 = s->n_splines_spf, K2 = s->n_splines_pc;
    memset(ext, 0, sizeof(double) * (size_t)(K1 + K2));
    for (int n = 0; n < N; n++)
    {
        const double r = inh_nic(R[3 * n], s->lbox) + s->lbox / 2.0;
        inh_add_values(s->knots_spf, K1 + 4, s->weights_spf, r, ext);
        for (int i = 0; i < n; i++)
        {
            const double v = inh_nic(R[3 * n] - R[3 * i], s->lbox);
            const double rni = sqrt(v * v);
            if (rni <= s->r_max) inh_add_values(s->knots_pc, K2 + 4, s->weights_pc, rni, ext + K1);
        }
    }
}

/* RefreshLocalOperators (:208-247) through the CSR form of the same map */
void oracle_inh_operators(const oracle_inh* s, const double* ext, double* O)
{
    for (int p = 0; p < s->n_params; p++)
    {
        double o = 0.0;
        for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++) o += s->map_val[j] * ext[s->map_col[j]];
        O[p] = o;
    }
}

/* CalculateWavefunction (:756-769): sum_i uR[i] O[i] - 2 gamma h_pc ss_pc[0] */
double oracle_inh_exponent(const oracle_inh* s, const double* ext, const double* uR)
{
    double* O = (double*)malloc(sizeof(double) * (size_t)s->n_params);
    oracle_inh_operators(s, ext, O);
    double sum = 0.0;
    for (int i = 0; i < s->n_params; i++) sum += uR[i] * O[i];
    sum += -2.0 * s->gamma * s->h_pc * ext[s->n_splines_spf];
    free(O);
    return sum;
}

/* GetExternalPotential (:448-509) for the four-entry SYSTEM_PARAMS of the shipped config */
static double inh_external(const oracle_inh* s, double x0)
{
    double value = 0.0;
    const double kf = M_PI;
    const double x = inh_nic(x0, s->lbox) + s->lbox / 2.0;
    if (s->ext_k > 0.0 && s->ext_v0 > 0.0)
    {
        const double k = s->ext_k * kf;
        value = sin(k * x);
        value *= value;
        value *= s->ext_v0;
        value *= k * k;
    }
    return value;
}

/* CalculateOtherLocalOperators (:306-440) + CalculateExpectationValues (:511-671).
 * tabD [K1+K2][N], tabD2 [K1+K2][N] (DIM = 1; may be NULL).  other[9] as :660-668. */
void oracle_inh_expectation(const oracle_inh* s, const double* R, double wf, double exponent, const double* uR, const double* uI,
                            double* e_r, double* e_i, double* other, double* drift_r, double* drift_i, double* tabD, double* tabD2)
{
    const int N = s->n_particles, K1 = s->n_splines_spf, K2 = s->n_splines_pc, P = s->n_params, NE = K1 + K2;
    double* sD = (double*)calloc((size_t)NE * N, sizeof(double));
    double* sD2 = (double*)calloc((size_t)NE * N, sizeof(double));
    double potentialExtern = 0, potentialIntern = 0;
    double tmp1[4], tmp2[4];
    for (int n = 0; n < N; n++)
    {
        potentialExtern += inh_external(s, R[3 * n]);
        {
            const double r = inh_nic(R[3 * n], s->lbox) + s->lbox / 2.0;
            const int bin = inh_bin(s->knots_spf, K1 + 4, r);
            const double r2 = r * r;
            for (int p = 0; p < 4; p++)
            {
                const double* q = s->weights_spf + ((size_t)(bin - p) * 4 + p) * 4;
                tmp1[3 - p] = q[1] + 2.0 * q[2] * r + 3.0 * q[3] * r2;
                tmp2[3 - p] = 2.0 * q[2] + 6.0 * q[3] * r;
            }
            for (int b = 0; b < 4; b++) sD[(size_t)(bin - b) * N + n] += tmp1[3 - b] * 1.0 * 1.0;
            const double secondDerivativeFactor = 1 - 1.0;
            for (int b = 0; b < 4; b++) sD2[(size_t)(bin - b) * N + n] += tmp2[3 - b] + secondDerivativeFactor / r * tmp1[3 - b];
        }
        for (int i = 0; i < N; i++)
        {
            const double v = inh_nic(R[3 * n] - R[3 * i], s->lbox);
            const double rni = sqrt(v * v);
            if (rni <= s->r_max)
            {
                if (i < n && s->gamma == 0.0 && rni < s->pot_range) potentialIntern += s->pot_strength; /* :383-394 */
                if (i != n)
                {
                    const int bin = inh_bin(s->knots_pc, K2 + 4, rni);
                    const double rni2 = rni * rni;
                    for (int p = 0; p < 4; p++)
                    {
                        const double* q = s->weights_pc + ((size_t)(bin - p) * 4 + p) * 4;
                        tmp1[3 - p] = q[1] + 2.0 * q[2] * rni + 3.0 * q[3] * rni2;
                        tmp2[3 - p] = 2.0 * q[2] + 6.0 * q[3] * rni;
                    }
                    const double evec = v / rni;
                    for (int b = 0; b < 4; b++) sD[(size_t)(K1 + bin - b) * N + n] += tmp1[3 - b] * evec * 1.0;
                    const double secondDerivativeFactor = 1 - 1.0;
                    for (int b = 0; b < 4; b++) sD2[(size_t)(K1 + bin - b) * N + n] += tmp2[3 - b] + secondDerivativeFactor / rni * tmp1[3 - b];
                }
            }
        }
    }
    /* contraction: the boundary-condition map of :536-624 is the CSR map; the contact term enters the REAL sums only */
    const double cg = -2.0 * s->gamma * s->h_pc;
    double kR1 = 0, kI1 = 0, kRI = 0, kR2 = 0, kI2 = 0;
    for (int n = 0; n < N; n++)
    {
        double vR = 0, vI = 0;
        for (int p = 0; p < P; p++)
        {
            double t = 0.0, t2 = 0.0;
            for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++)
            {
                t += s->map_val[j] * sD[(size_t)s->map_col[j] * N + n];
                t2 += s->map_val[j] * sD2[(size_t)s->map_col[j] * N + n];
            }
            vR += uR[p] * t;
            vI += uI[p] * t;
            kR2 += uR[p] * t2;
            kI2 += uI[p] * t2;
        }
        vR += cg * sD[(size_t)K1 * N + n];   /* :595-602 */
        kR2 += cg * sD2[(size_t)K1 * N + n];
        kRI += 2.0 * (vR * vI);
        kR1 += vR * vR;
        kI1 += vI * vI;
        if (drift_r)
        {
            drift_r[3 * n] = vR;
            drift_r[3 * n + 1] = drift_r[3 * n + 2] = 0.0;
            drift_i[3 * n] = vI;
            drift_i[3 * n + 1] = drift_i[3 * n + 2] = 0.0;
        }
    }
    const double kineticR = -(kR1 - kI1 + kR2) * s->hbar2_2m;
    const double kineticI = -(kRI + kI2) * s->hbar2_2m;
    *e_r = kineticR + potentialIntern + potentialExtern;
    *e_i = kineticI + 0.0;
    other[0] = kineticR;
    other[1] = potentialIntern;
    other[2] = wf;
    other[3] = exponent;
    other[4] = kR1;
    other[5] = kI1;
    other[6] = kR2;
    other[7] = kI2;
    other[8] = kRI;
    if (tabD) memcpy(tabD, sD, sizeof(double) * (size_t)NE * N);
    if (tabD2) memcpy(tabD2, sD2, sizeof(double) * (size_t)NE * N);
    free(sD);
    free(sD2);
}

/* CalculateWFChange + CalculateWFQuotient (:783-919).  R holds the NEW coordinate of `particle`. */
double oracle_inh_quotient(const oracle_inh* s, const double* R, int particle, const double* old_pos, const double* ext,
                           double exponent, const double* uR, double* ext_new, double* exponent_new)
{
    const int N = s->n_particles, K1 = s->n_splines_spf, K2 = s->n_splines_pc, NE = K1 + K2;
    double* oldb = (double*)calloc((size_t)2 * NE, sizeof(double));
    double* newb = oldb + NE;
    inh_add_values(s->knots_spf, K1 + 4, s->weights_spf, inh_nic(old_pos[0], s->lbox) + s->lbox / 2.0, oldb);
    inh_add_values(s->knots_spf, K1 + 4, s->weights_spf, inh_nic(R[3 * particle], s->lbox) + s->lbox / 2.0, newb);
    for (int i = 0; i < N; i++)
    {
        if (i == particle) continue;
        double v = inh_nic(R[3 * i] - old_pos[0], s->lbox);
        double rni = sqrt(v * v);
        if (rni <= s->r_max) inh_add_values(s->knots_pc, K2 + 4, s->weights_pc, rni, oldb + K1);
        v = inh_nic(R[3 * i] - R[3 * particle], s->lbox);
        rni = sqrt(v * v);
        if (rni <= s->r_max) inh_add_values(s->knots_pc, K2 + 4, s->weights_pc, rni, newb + K1);
    }
    for (int k = 0; k < NE; k++) ext_new[k] = fmax(0.0, ext[k] - oldb[k] + newb[k]); /* :860-867 */
    free(oldb);
    *exponent_new = oracle_inh_exponent(s, ext_new, uR); /* :869-895: the same map, term by term */
    return exp(2.0 * (*exponent_new - exponent));
}

/* DoMetropolisStep (src/TDVMC.cpp:858-916), one coordinate per move; proposal stream shared with the device (only the
 * first of the three Gaussian components is used) */
int64_t oracle_inh_sweep(const oracle_inh* s, double* R, double* ext, double* exponent, const double* uR, uint64_t seed,
                         uint32_t walker, uint64_t first_step, int64_t n_steps, double mc_step)
{
    const int NE = s->n_splines_spf + s->n_splines_pc;
    int64_t accepted = 0;
    double* ext_new = (double*)malloc(sizeof(double) * (size_t)NE);
    for (int64_t t = 0; t < n_steps; t++)
    {
        int p;
        double disp[3], log_u, old_pos[3], exponent_new;
        oracle_proposal(seed, walker, first_step + (uint64_t)t, s->n_particles, mc_step, &p, disp, &log_u);
        old_pos[0] = R[(size_t)p * 3];
        R[(size_t)p * 3] += disp[0];
        const double q = oracle_inh_quotient(s, R, p, old_pos, ext, *exponent, uR, ext_new, &exponent_new);
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
        if (!ok || (!force && 2.0 * (exponent_new - *exponent) < log_u)) R[(size_t)p * 3] = old_pos[0];
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

int64_t oracle_inh_sample_walker(const oracle_inh* s, double* R, const double* uR, const double* uI, double phiR, uint64_t seed,
                                 uint32_t walker, uint64_t* step_counter, int n_init, int n_samples, int n_therm, double mc_step,
                                 double* est, double* sample_rows)
{
    const int P = s->n_params, NE = s->n_splines_spf + s->n_splines_pc, NO = 9;
    double* ext = (double*)malloc(sizeof(double) * (size_t)NE);
    double* O = (double*)malloc(sizeof(double) * (size_t)P);
    double other[9], exponent, e_r, e_i;
    int64_t accepted = 0;
    oracle_inh_values(s, R, ext);
    exponent = oracle_inh_exponent(s, ext, uR);
    accepted += oracle_inh_sweep(s, R, ext, &exponent, uR, seed, walker, *step_counter, n_init, mc_step);
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
        accepted += oracle_inh_sweep(s, R, ext, &exponent, uR, seed, walker, *step_counter, n_therm, mc_step);
        *step_counter += (uint64_t)n_therm;
        oracle_inh_operators(s, ext, O);
        oracle_inh_expectation(s, R, exp(exponent + phiR), exponent, uR, uI, &e_r, &e_i, other, NULL, NULL, NULL, NULL);
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
    return accepted;
}
