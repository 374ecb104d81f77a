// Parameter linear solve on the device: S u' = F for the P variational parameters, from the all-reduced estimator
// sums that already sit in HBM.
//
// Replaces, for LINEAR_EQUATION_SOLVER_TYPE = 0 (src/TDVMC.cpp:1713-1763):
//   BuildSystemOfEquationsForParametersIncludePhi  src/TDVMC.cpp:1506-1537
//   PreconditionEquationSystemByScaling            :1684-1701
//   RegularizeEquationSystem                       :1703-1711
//   PerformCholeskyDecomposition                   :1560-1592
//   SolveCholeskyDecomposedEquationSystem          :1594-1622
//   CalculatePhiDot                                :1658-1682
//
// One CTA.  The lower triangle lives packed (row i at i(i+1)/2) in shared memory when it fits (P <= 230), else in a
// global scratch buffer that stays in L2.  The factorisation goes column by column with one thread per row, each element
// formed by the reference's own `sum -= matrix[i][k] * matrix[j][k]` loop (k ascending), and the two substitutions are
// column-oriented so that every row's sum grows in the reference's order; with -fmad=false every rounding step is the
// reference's and the result is bit-identical to the host code it replaces (IEEE sqrt and division on both sides).
// P <= 1024 (one thread per unknown).
#include "kernels.cuh"

namespace tdvmc
{

__device__ __forceinline__ size_t tri(int i, int j) { return (size_t)i * (i + 1) / 2 + j; }

__global__ void __launch_bounds__(1024, 1) solve_kernel(SolveArgs a)
{
    extern __shared__ __align__(16) double sm[];
    const int P = a.P, tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    double* vec = sm;               // 8 vectors of P
    double* O = vec;
    double* bR = vec + P;
    double* bI = vec + 2 * P;
    double* scal = vec + 3 * P;
    double* diag0 = vec + 4 * P;
    double* col = vec + 5 * P;      // current column of L, later the forward solution of the first right-hand side
    double* tI = vec + 6 * P;       // forward solution of the second right-hand side
    double* xR = vec + 7 * P;
    double* xI = vec + 8 * P;
    double* Lp = a.L_in_smem ? vec + 9 * P : a.L_global;
    __shared__ int s_flag;

    const double* S = a.est;
    const double* FR = S + (size_t)P * P;
    const double* FI = FR + P;
    const double* Os = FI + P;
    const double* E = Os + P;
    const double n = a.est[a.cnt_offset + 2];
    const double inv = 1.0 / n; // ReduceToAverage (MPIMethods.h:199-203), as tdvmc_gpu_allreduce_and_fetch divides
    const double ER = E[0] * inv, EI = E[1] * inv;

    if (tid == 0) s_flag = 0;
    for (int i = tid; i < P; i += T) O[i] = Os[i] * inv;
    __syncthreads();
    for (int i = tid; i < P; i += T)
    {
        const double oer = FR[i] * inv, oei = FI[i] * inv;
        if (a.imaginary_time == -1) // BuildSystemOfEquationsForParametersIncludePhiWithTimeRotation, src/TDVMC.cpp:1475-1504
        {
            const double rotation = 1.499 * 3.14159265358979323846; // 3/2 Pi -> real time; Pi -> imaginary time (:1477)
            const double c = cos(rotation), sn = sin(rotation);
            bR[i] = c * (oer - ER * O[i]) - sn * (oei);
            bI[i] = sn * (oer - ER * O[i]) + c * (oei);
        }
        else if (a.imaginary_time == 0) // src/TDVMC.cpp:1518-1523
        {
            bR[i] = oei - EI * O[i];
            bI[i] = -oer + ER * O[i];
        }
        else                       // :1524-1528
        {
            bR[i] = -oer + ER * O[i];
            bI[i] = -oei;
        }
    }
    // matrix[i][j] = <O_i O_j> - <O_i><O_j>, lower triangle (:1529-1534)
    for (int i = warp; i < P; i += nwarp)
        for (int j = lane; j <= i; j += 32) Lp[tri(i, j)] = S[(size_t)i * P + j] * inv - O[i] * O[j];
    __syncthreads();

    if (a.use_preconditioning) // :1684-1701
    {
        for (int i = tid; i < P; i += T)
        {
            double s = sqrt(Lp[tri(i, i)]);
            if (a.min_scaling > 0.0 && !(s >= a.min_scaling)) s = a.min_scaling; // not in the reference (it would divide by 0)
            scal[i] = s;
        }
        __syncthreads();
        for (int i = warp; i < P; i += nwarp)
            for (int j = lane; j <= i; j += 32) Lp[tri(i, j)] = Lp[tri(i, j)] / (scal[i] * scal[j]);
        for (int i = tid; i < P; i += T)
        {
            bR[i] = bR[i] / scal[i];
            bI[i] = bI[i] / scal[i];
        }
        __syncthreads();
    }
    for (int i = tid; i < P; i += T) // RegularizeEquationSystem :1703-1711
    {
        const double d = Lp[tri(i, i)] + a.regularization;
        Lp[tri(i, i)] = d;
        diag0[i] = d;
    }
    __syncthreads();

    // Cholesky, thread i owns row i: s = A[i][j] - sum_{k<j} L[i][k] L[j][k] with k ascending in registers - literally
    // the reference's inner loop (:1568-1573), so the rounding sequence is its own; the rows j are read by every thread at
    // the same address (broadcast), row i is private.  Each sum is one dependent chain of FP64 subtractions (about 40
    // cycles per link), so kColBlock columns advance together: their chains are independent up to k = j0 and share the
    // load of L[i][k]; then the columns of the block are finished in order, each adding its new term to the chains of
    // the columns after it.  Two barriers per column (pivot known, column complete).  A pivot that is not positive
    // leaves the ORIGINAL diagonal entry in place, as the reference does (:1577-1589, it only logs and sets
    // doNotAcceptStep); the flag goes back to the caller.
    // (History, P = 201: right-looking rank-1 updates by 32-wide rows 0.44 ms - 1.8 M warp-instructions, half the
    // lanes idle, profiles/r01g_solve_ncu.txt; one column at a time with this row ownership 0.46 ms - the chain.)
    constexpr int kColBlock = 4;
    double* pivot = col; // col[0]: the pivot of the current column
    for (int j0 = 0; j0 < P; j0 += kColBlock)
    {
        const int i = tid;
        const int nb = min(kColBlock, P - j0);
        const bool row = i >= j0 && i < P;
        double sdot[kColBlock];
        const double* ri = Lp + tri(row ? i : 0, 0);
        if (row)
        {
            const double* rj[kColBlock];
#pragma unroll
            for (int c = 0; c < kColBlock; c++)
            {
                const int j = min(j0 + c, P - 1);
                rj[c] = Lp + tri(j, 0);
                sdot[c] = (c < nb && i >= j0 + c) ? ri[j0 + c] : 0.0;
            }
#pragma unroll 2
            for (int k = 0; k < j0; k++)
            {
                const double l = ri[k];
#pragma unroll
                for (int c = 0; c < kColBlock; c++) sdot[c] -= l * rj[c][k];
            }
        }
#pragma unroll
        for (int c = 0; c < kColBlock; c++)
        {
            if (c < nb) // (uniform over the block)
            {
                const int j = j0 + c;
                if (row && i == j)
                {
                    double d;
                    if (sdot[c] > 0.0) d = sqrt(sdot[c]);
                    else
                    {
                        d = diag0[j];
                        s_flag = 1;
                    }
                    pivot[0] = d;
                    Lp[tri(j, j)] = d;
                }
                __syncthreads();
                double lij = 0.0;
                if (row && i > j)
                {
                    lij = sdot[c] / pivot[0];
                    Lp[tri(i, j)] = lij;
                }
                __syncthreads();
                // the new column enters the chains of the block's later columns as their term k = j
                if (row)
                {
#pragma unroll
                    for (int c2 = c + 1; c2 < kColBlock; c2++)
                        if (c2 < nb && i >= j0 + c2) sdot[c2] -= lij * Lp[tri(j0 + c2, j)];
                }
            }
        }
    }

    // forward substitution L y = b, both right-hand sides; thread i owns the running sum of row i
    // (sum += matrix[i][j] * tmp[j] for j = 0 .. i-1, :1605-1611)
    double* tR = col;
    double accR = 0.0, accI = 0.0;
    for (int j = 0; j < P; j++)
    {
        if (tid == j)
        {
            const double dj = Lp[tri(j, j)];
            tR[j] = 1.0 / dj * (bR[j] - accR);
            tI[j] = 1.0 / dj * (bI[j] - accI);
        }
        __syncthreads();
        if (tid > j && tid < P)
        {
            const double l = Lp[tri(tid, j)];
            accR += l * tR[j];
            accI += l * tI[j];
        }
    }
    // backward substitution L^T x = y (sum += matrix[j][i] * solution[j] for j = P-1 .. i+1, :1613-1621)
    accR = 0.0;
    accI = 0.0;
    for (int j = P - 1; j >= 0; j--)
    {
        if (tid == j)
        {
            const double dj = Lp[tri(j, j)];
            xR[j] = 1.0 / dj * (tR[j] - accR);
            xI[j] = 1.0 / dj * (tI[j] - accI);
        }
        __syncthreads();
        if (tid < j)
        {
            const double l = Lp[tri(j, tid)];
            accR += l * xR[j];
            accI += l * xI[j];
        }
    }
    __syncthreads();

    // CalculatePhiDot (:1658-1682) runs on the solution BEFORE the scalings are divided out (:1743-1752)
    if (tid == 0)
    {
        double pr = 0.0, pi = 0.0;
        for (int i = 0; i < P; i++)
        {
            pr -= O[i] * xR[i];
            pi -= O[i] * xI[i];
        }
        if (a.imaginary_time == -1) // CalculatePhiDot, :1666-1673
        {
            const double rotation = 1.499 * 3.14159265358979323846;
            pi -= cos(rotation) * ER;
            pr -= sin(rotation) * ER;
        }
        else if (a.imaginary_time == 0) pi -= ER;
        else pr -= ER;
        double* tail = a.out + 2 * (size_t)P;
        tail[0] = pr;
        tail[1] = pi;
        tail[2] = (double)s_flag;
        tail[3] = ER;
        tail[4] = EI;
    }
    for (int i = tid; i < P; i += T)
    {
        a.out[i] = a.use_preconditioning ? xR[i] / scal[i] : xR[i];
        a.out[P + i] = a.use_preconditioning ? xI[i] / scal[i] : xI[i];
    }
}

size_t solve_smem_bytes(int P, bool l_in_smem)
{
    size_t n = 9 * (size_t)P;
    if (l_in_smem) n += (size_t)P * (P + 1) / 2;
    return n * sizeof(double);
}

cudaError_t launch_solve(SolveArgs a, int smem_optin, cudaStream_t st)
{
    if (a.P < 1 || a.P > 1024) return cudaErrorInvalidValue;
    a.L_in_smem = (solve_smem_bytes(a.P, true) + 1024 <= (size_t)smem_optin && !a.force_global) ? 1 : 0;
    if (!a.L_in_smem && !a.L_global) return cudaErrorInvalidValue;
    const size_t smem = solve_smem_bytes(a.P, a.L_in_smem != 0);
    cudaError_t e = cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    solve_kernel<<<1, 1024, smem, st>>>(a);
    return cudaGetLastError();
}

} // namespace tdvmc
