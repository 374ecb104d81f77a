/*
 * tdvmc_oracle.c — plain-C restatement of the reference's walker hot path.
 * TEST INFRASTRUCTURE ONLY; see tdvmc_oracle.h for scope, parity status and citations.
 */
#include "tdvmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- geometry: Utils.cpp:266-281 (GetCoordinateNIC), :329-338, :368-374 ---- */

static double nic_coordinate(double r, double lbox, double lbox_r, double lbox_2)
{
    int k = (int)(r * lbox_r + ((r >= 0.0) ? 0.5 : -0.5));
    double result = r - k * lbox;
    if (result == lbox_2) result -= 1e-10;
    else if (result == -lbox_2) result += 1e-10;
    return result;
}

double oracle_min_image(double lbox, int dim, const double* a, const double* b, double* disp)
{
    double lbox_r = 1.0 / lbox; /* src/TDVMC.cpp:535-536 */
    double lbox_2 = lbox / 2.0;
    double sum = 0.0;
    for (int c = 0; c < dim; c++)
    {
        double delta = a[c] - b[c];
        disp[c] = nic_coordinate(delta, lbox, lbox_r, lbox_2);
    }
    if (dim == 3) sum = disp[0] * disp[0] + disp[1] * disp[1] + disp[2] * disp[2]; /* Utils.cpp:174-179 */
    else
        for (int c = 0; c < dim; c++) sum += disp[c] * disp[c];
    return sqrt(sum);
}

/* std::lower_bound(nodes, r) - 1: knots[bin] < r <= knots[bin+1] (BosonsBulk.cpp:197-198) */
static int find_bin(const oracle_system* s, double r)
{
    int lo = 0, hi = s->n_splines + 4; /* first index with knots[idx] >= r */
    while (lo < hi)
    {
        int mid = (lo + hi) / 2;
        if (s->knots[mid] < r) lo = mid + 1;
        else hi = mid;
    }
    return lo - 1;
}

static inline const double* piece(const oracle_system* s, int spline, int part)
{
    return s->weights + ((size_t)spline * 4 + part) * 4;
}

/* pair rule: returns 1 when the (possibly reflected) distance takes the spline branch */
static int in_spline_range(const oracle_system* s, double* r, int strict)
{
    if (s->pair_rule == ORACLE_PAIR_RULE_REFLECT)
    {
        if (!(*r < s->r_max)) *r = 2 * s->r_max - *r; /* NUBosonsBulkPB.cpp:250-253 */
        return *r < s->r_max;
    }
    return strict ? (*r < s->r_max) : (*r <= s->r_max); /* BosonsBulk.cpp:195, :571 vs :593 */
}

static void add_basis_values(const oracle_system* s, double r, double* sums)
{
    int bin = find_bin(s, r);
    double r2 = r * r;
    double r3 = r2 * r;
    for (int p = 0; p < 4; p++)
    {
        const double* w = piece(s, bin - p, p);
        sums[bin - p] += w[0] + w[1] * r + w[2] * r2 + w[3] * r3; /* BosonsBulk.cpp:204 */
    }
}

void oracle_basis_sums(const oracle_system* s, const double* R, double* ss, double* outer)
{
    int N = s->n_particles, D = s->dim;
    double vec[3];
    *outer = 0.0;
    memset(ss, 0, sizeof(double) * (size_t)s->n_splines);
    for (int n = 0; n < N; n++)
    {
        for (int i = 0; i < n; i++)
        {
            double r = oracle_min_image(s->lbox, D, R + (size_t)n * D, R + (size_t)i * D, vec);
            if (in_spline_range(s, &r, 0)) add_basis_values(s, r, ss);
            else *outer += 1.0;
        }
    }
}

void oracle_local_operators(const oracle_system* s, const double* ss, double* O)
{
    for (int p = 0; p < s->n_params; p++)
    {
        double v = 0.0;
        for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++) v += s->map_val[j] * ss[s->map_col[j]];
        O[p] = v;
    }
}

double oracle_exponent(const oracle_system* s, const double* O, double outer, const double* uR)
{
    double sum = 0.0;
    for (int i = 0; i < s->n_params; i++) sum += uR[i] * O[i];
    sum += uR[s->tail_param] * outer; /* BosonsBulk.cpp:532-534 */
    return sum;
}

void oracle_tables(const oracle_system* s, const double* R, double* sD, double* sD2, double* v_int)
{
    int N = s->n_particles, D = s->dim, K = s->n_splines;
    double vec[3], evec[3], tmp1[4], tmp2[4];
    double potential = 0.0;
    memset(sD, 0, sizeof(double) * (size_t)K * N * D);
    memset(sD2, 0, sizeof(double) * (size_t)K * N);
    for (int n = 0; n < N; n++)
    {
        for (int i = 0; i < N; i++)
        {
            double r = oracle_min_image(s->lbox, D, R + (size_t)n * D, R + (size_t)i * D, vec);
            if (!in_spline_range(s, &r, 0)) continue;
            if (i < n && r < s->pot_a) potential += s->pot_b; /* BosonsBulk.cpp:268-271 */
            if (i == n) continue;
            int bin = find_bin(s, r);
            double r2 = r * r;
            for (int p = 0; p < 4; p++)
            {
                const double* w = piece(s, bin - p, p);
                tmp1[3 - p] = w[1] + 2.0 * w[2] * r + 3.0 * w[3] * r2; /* BosonsBulk.cpp:299 */
                tmp2[3 - p] = 2.0 * w[2] + 6.0 * w[3] * r;             /* BosonsBulk.cpp:301 */
            }
            for (int a = 0; a < D; a++) evec[a] = vec[a] / r; /* NU: unreflected vec / reflected r */
            for (int a = 0; a < D; a++)
                for (int b = 0; b < 4; b++) sD[((size_t)(bin - b) * N + n) * D + a] += tmp1[3 - b] * evec[a] * 1.0;
            double f2 = D - 1.0;
            for (int b = 0; b < 4; b++) sD2[(size_t)(bin - b) * N + n] += tmp2[3 - b] + f2 / r * tmp1[3 - b];
        }
    }
    *v_int = potential;
}

void oracle_expectation(const oracle_system* s, const double* O, const double* sD, const double* sD2,
                        double v_int, double exponent, double phiR, const double* uR, const double* uI,
                        double* e_r, double* e_i, double* other9, double* drift_r, double* drift_i)
{
    int N = s->n_particles, D = s->dim, P = s->n_params;
    double R1 = 0, I1 = 0, R1I1 = 0, R2 = 0, I2 = 0;
    double vr[3], vi[3];
    (void)O;
    for (int n = 0; n < N; n++)
    {
        for (int a = 0; a < D; a++) vr[a] = vi[a] = 0.0;
        for (int p = 0; p < P; p++)
        {
            for (int a = 0; a < D; a++)
            {
                double tmp = 0.0;
                for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++)
                    tmp += s->map_val[j] * sD[((size_t)s->map_col[j] * N + n) * D + a];
                vr[a] += uR[p] * tmp;
                vi[a] += uI[p] * tmp;
            }
            double tmp = 0.0;
            for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++)
                tmp += s->map_val[j] * sD2[(size_t)s->map_col[j] * N + n];
            R2 += uR[p] * tmp;
            I2 += uI[p] * tmp;
        }
        double dot = 0, nr = 0, ni = 0;
        for (int a = 0; a < D; a++)
        {
            dot += vr[a] * vi[a];
            nr += vr[a] * vr[a];
            ni += vi[a] * vi[a];
            if (drift_r) drift_r[(size_t)n * D + a] = vr[a];
            if (drift_i) drift_i[(size_t)n * D + a] = vi[a];
        }
        R1I1 += 2.0 * dot;
        R1 += nr;
        I1 += ni;
    }
    double kin_r = -(R1 - I1 + R2) * s->hbar2_2m; /* BosonsBulk.cpp:422-423 */
    double kin_i = -(R1I1 + I2) * s->hbar2_2m;
    *e_r = kin_r + v_int + 0.0; /* external potential is 0 (BosonsBulk.cpp:344-347) */
    *e_i = kin_i + 0.0;
    other9[0] = kin_r;
    other9[1] = v_int;
    other9[2] = exp(exponent + phiR);
    other9[3] = exponent;
    other9[4] = R1;
    other9[5] = I1;
    other9[6] = R2;
    other9[7] = I2;
    other9[8] = R1I1;
}

double oracle_wf_quotient(const oracle_system* s, const double* R, int particle, const double* old_pos,
                          const double* ss, double outer, double exponent, const double* uR,
                          double* ss_new, double* outer_new, double* exponent_new)
{
    int N = s->n_particles, D = s->dim, K = s->n_splines;
    double vec[3];
    double old_outer = 0.0, new_outer = 0.0;
    double* sum_old = (double*)calloc((size_t)K, sizeof(double));
    double* sum_new = (double*)calloc((size_t)K, sizeof(double));
    for (int i = 0; i < N; i++)
    {
        if (i == particle) continue;
        double r = oracle_min_image(s->lbox, D, R + (size_t)i * D, old_pos, vec);
        if (in_spline_range(s, &r, 0)) add_basis_values(s, r, sum_old);
        else old_outer += 1.0;
        r = oracle_min_image(s->lbox, D, R + (size_t)i * D, R + (size_t)particle * D, vec);
        if (in_spline_range(s, &r, 1)) add_basis_values(s, r, sum_new);
        else new_outer += 1.0;
    }
    for (int k = 0; k < K; k++) ss_new[k] = fmax(0.0, ss[k] - sum_old[k] + sum_new[k]); /* :619-621 */
    *outer_new = fmax(0.0, outer - old_outer + new_outer);
    free(sum_old);
    free(sum_new);
    double sum = 0.0;
    for (int p = 0; p < s->n_params; p++)
    {
        double v = 0.0;
        for (int j = s->map_ptr[p]; j < s->map_ptr[p + 1]; j++) v += s->map_val[j] * ss_new[s->map_col[j]];
        sum += uR[p] * v;
    }
    sum += uR[s->tail_param] * *outer_new;
    *exponent_new = sum;
    return exp(2.0 * (sum - exponent));
}

/* ---- proposal stream (ours, shared with the CUDA path; see header) ---- */

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; round++)
    {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double u53(uint32_t hi, uint32_t lo) /* (0, 1] */
{
    uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)x + 1.0) * (1.0 / 9007199254740992.0);
}

void oracle_proposal(uint64_t seed, uint32_t walker, uint64_t step, int n_particles, double mc_step,
                     int* particle, double disp[3], double* log_u)
{
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t ctr[4] = { (uint32_t)step, (uint32_t)(step >> 32), walker, 0u };
    uint32_t a[4], b[4], c[4];
    oracle_philox4x32_10(ctr, key, a);
    ctr[3] = 1u;
    oracle_philox4x32_10(ctr, key, b);
    ctr[3] = 2u;
    oracle_philox4x32_10(ctr, key, c);
    *particle = (int)(((uint64_t)a[0] * (uint64_t)n_particles) >> 32);
    *log_u = log(u53(a[1], a[2]));
    const double two_pi = 6.283185307179586476925286766559;
    double rad0 = sqrt(-2.0 * log(u53(a[3], b[0])));
    double ang0 = two_pi * (u53(b[1], b[2]) - 1.0 / 9007199254740992.0);
    double rad1 = sqrt(-2.0 * log(u53(b[3], c[0])));
    double ang1 = two_pi * (u53(c[1], c[2]) - 1.0 / 9007199254740992.0);
    disp[0] = rad0 * cos(ang0) * mc_step;
    disp[1] = rad0 * sin(ang0) * mc_step;
    disp[2] = rad1 * cos(ang1) * mc_step;
}

int64_t oracle_sweep(const oracle_system* s, double* R, double* ss, double* outer, double* exponent,
                     const double* uR, uint64_t seed, uint32_t walker, uint64_t first_step, int64_t n_steps,
                     double mc_step)
{
    int D = s->dim, K = s->n_splines;
    int64_t accepted = 0;
    double* ss_new = (double*)malloc(sizeof(double) * (size_t)K);
    for (int64_t t = 0; t < n_steps; t++)
    {
        int p;
        double disp[3], log_u, old_pos[3], outer_new, exponent_new;
        oracle_proposal(seed, walker, first_step + (uint64_t)t, s->n_particles, mc_step, &p, disp, &log_u);
        for (int a = 0; a < D; a++)
        {
            old_pos[a] = R[(size_t)p * D + a];
            R[(size_t)p * D + a] += disp[a]; /* src/TDVMC.cpp:872-875 */
        }
        double q = oracle_wf_quotient(s, R, p, old_pos, ss, *outer, *exponent, uR, ss_new, &outer_new, &exponent_new);
        int ok = 1, force = 0;
        if (!isfinite(q) || !isfinite(exponent_new) || !isfinite(*exponent)) /* src/TDVMC.cpp:886-898 */
        {
            ok = 0;
            if (!isfinite(q) && exponent_new > 0 && *exponent == 0)
            {
                ok = 1;
                force = 1; /* "accept by 100%" */
            }
        }
        /* quotient < p  <=>  2 (e_new - e) < log p; compared in the log domain like the CUDA path */
        if (!ok || (!force && 2.0 * (exponent_new - *exponent) < log_u))
        {
            for (int a = 0; a < D; a++) R[(size_t)p * D + a] = old_pos[a];
        }
        else
        {
            memcpy(ss, ss_new, sizeof(double) * (size_t)K); /* AcceptMove, BosonsBulk.cpp:659-665 */
            *outer = outer_new;
            *exponent = exponent_new;
            accepted++;
        }
    }
    free(ss_new);
    return accepted;
}

int64_t oracle_sample_walker(const oracle_system* s, double* R, const double* uR, const double* uI, double phiR,
                             uint64_t seed, uint32_t walker, uint64_t* step_counter, int n_init, int n_samples,
                             int n_therm, double mc_step, double* est, double* sample_rows)
{
    int N = s->n_particles, D = s->dim, K = s->n_splines, P = s->n_params;
    double* ss = (double*)malloc(sizeof(double) * (size_t)K);
    double* O = (double*)malloc(sizeof(double) * (size_t)P);
    double* sD = (double*)malloc(sizeof(double) * (size_t)K * N * D);
    double* sD2 = (double*)malloc(sizeof(double) * (size_t)K * N);
    double outer, exponent, v_int, e_r, e_i, other[9];
    int64_t accepted = 0;

    oracle_basis_sums(s, R, ss, &outer); /* sys->CalculateWavefunction, src/TDVMC.cpp:1060 */
    oracle_local_operators(s, ss, O);
    exponent = oracle_exponent(s, O, outer, uR);
    accepted += oracle_sweep(s, R, ss, &outer, &exponent, uR, seed, walker, *step_counter, n_init, mc_step);
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
        accepted += oracle_sweep(s, R, ss, &outer, &exponent, uR, seed, walker, *step_counter, n_therm, mc_step);
        *step_counter += (uint64_t)n_therm;
        oracle_local_operators(s, ss, O); /* RefreshLocalOperators from the maintained sums */
        oracle_tables(s, R, sD, sD2, &v_int);
        oracle_expectation(s, O, sD, sD2, v_int, exponent, phiR, uR, uI, &e_r, &e_i, other, NULL, NULL);
        for (int k = 0; k < P; k++)
        {
            eO[k] += O[k];
            eOER[k] += O[k] * e_r;
            eOEI[k] += O[k] * e_i;
            for (int j = 0; j < P; j++) eS[(size_t)k * P + j] += O[k] * O[j];
        }
        *eER += e_r;
        *eEI += e_i;
        for (int k = 0; k < 9; k++) eOther[k] += other[k];
        if (sample_rows)
        {
            double* row = sample_rows + (size_t)m * (P + 2);
            memcpy(row, O, sizeof(double) * (size_t)P);
            row[P] = e_r;
            row[P + 1] = e_i;
        }
    }
    free(ss);
    free(O);
    free(sD);
    free(sD2);
    return accepted;
}


/* ------------------------------------------------------------------------------------------------
 * Additional observables: BosonsBulk.cpp:474-520, NUBosonsBulkPB.cpp:597-639
 * ---------------------------------------------------------------------------------------------- */
void oracle_observables(double lbox, int n_particles, const double* R, int gr_count, double gr_spacing, double gr_max,
                        double gr_weight, const double* gr_scaling, int n_shells, const int32_t* shell_ptr,
                        const double* kvec, double* gr, double* sk)
{
    const int N = n_particles;
    for (int b = 0; b < gr_count; b++) gr[b] = 0.0;
    for (int i = 0; i < N; i++)
    {
        for (int j = 0; j < i; j++)
        {
            double vec[3];
            const double r = oracle_min_image(lbox, 3, R + 3 * i, R + 3 * j, vec); /* BosonsBulk.cpp:486 */
            if (r < gr_max)
            {
                const int bin = (int)floor(r / gr_spacing); /* Grid.cpp:58-59 */
                if (bin < gr_count) gr[bin] += gr_weight / gr_scaling[bin];
            }
        }
    }
    for (int k = 0; k < n_shells; k++)
    {
        double c = 0.0, s = 0.0;
        for (int i = 0; i < N; i++) /* the reference's loop order is i, k, kn; each shell's sum sees i outermost */
        {
            for (int q = shell_ptr[k]; q < shell_ptr[k + 1]; q++)
            {
                const double* kv = kvec + 3 * q;
                const double arg = kv[0] * R[3 * i] + kv[1] * R[3 * i + 1] + kv[2] * R[3 * i + 2];
                c += cos(arg);
                s += sin(arg);
            }
        }
        sk[k] = (c * c + s * s) / ((double)(N * (shell_ptr[k + 1] - shell_ptr[k])));
    }
}
