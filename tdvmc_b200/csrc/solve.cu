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
// global scratch buffer that stays in L2.  The factorisation is right-looking (column j is scaled, then subtracted from
// the trailing triangle), which applies the products l_ik l_jk to element (i, j) in the order k = 0, 1, ... j-1 - the
// order of the reference's `sum -= matrix[i][k] * matrix[j][k]` loop - and the two substitutions are column-oriented
// for the same reason, so with -fmad=false every rounding step is the reference's and the result is bit-identical to
// the host code it replaces (IEEE sqrt and division on both sides).  P <= 1024 (one thread per unknown in the solves).
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
        if (a.imaginary_time == 0) // src/TDVMC.cpp:1518-1523
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

    // Cholesky, right-looking.  A pivot that is not positive leaves the ORIGINAL diagonal entry in place, as the
    // reference does (:1577-1589, it only logs and sets doNotAcceptStep); the flag goes back to the caller.
    for (int j = 0; j < P; j++)
    {
        const double piv = Lp[tri(j, j)];
        double d;
        if (piv > 0.0) d = sqrt(piv);
        else
        {
            d = diag0[j];
            if (tid == 0) s_flag = 1;
        }
        for (int i = j + 1 + tid; i < P; i += T)
        {
            const double l = Lp[tri(i, j)] / d;
            Lp[tri(i, j)] = l;
            col[i] = l;
        }
        __syncthreads(); // every thread has read the pivot, the column is complete
        if (tid == 0) Lp[tri(j, j)] = d;
        for (int i = j + 1 + warp; i < P; i += nwarp)
        {
            const double li = col[i];
            double* row = Lp + tri(i, 0);
            for (int k = j + 1 + lane; k <= i; k += 32) row[k] = row[k] - li * col[k];
        }
        __syncthreads();
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
        if (a.imaginary_time == 0) pi -= ER;
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
