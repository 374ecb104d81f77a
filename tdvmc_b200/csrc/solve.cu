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

#include <cooperative_groups.h>

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

// ---------------------------------------------------------------------------------------------------------------
// LINEAR_EQUATION_SOLVER_TYPE = 1: Eigen::FullPivHouseholderQR (src/TDVMC.cpp:1763-1827), r02.
//
// The reference builds the same system, scales it and adds 0.002 to the diagonal only with USE_PRECONDITIONING = 1
// (:1767-1771), factorises with Eigen 3.3.7's full-pivoting Householder QR and solves both right-hand sides, subtracts
// the mean of each solution (:1802-1809), then CalculatePhiDot and the scalings.  This kernel follows Eigen's
// algorithm step by step (resources/Eigen/src/QR/FullPivHouseholderQR.h computeInPlace / _solve_impl,
// Householder/Householder.h makeHouseholder / applyHouseholderOnTheLeft): pivot = the first maximum of |a_ij| over the
// trailing block in column-major order, early exit when it is negligible against the first pivot, row and column
// transpositions, beta / tau / essential part as there, tmp = essential^T bottom + row 0, row 0 -= tau tmp,
// bottom -= (tau essential) tmp; rank from |R_ii| > eps P max|R_ii|; Q^T applied to the right-hand sides, back
// substitution on the leading rank x rank triangle, column permutation undone, zeros beyond the rank.  Dot products and
// norms are summed in a different order than Eigen's packet reductions, so the results agree to rounding times the
// condition number, not bit for bit.
//
// Two instances of one kernel.  CLUSTER = true (the default where it fits, P <= 228): a thread-block CLUSTER of two CTAs
// holds the columns alternately in shared memory (column j in CTA j & 1, P doubles each: 162 KB per CTA at P = 201 - the
// 323 KB matrix fits no single SM), so that the 201 pivot searches, swaps and reflections never wait for L2.  What crosses
// between the two SMs goes through distributed shared memory: each CTA publishes the biggest entry of its half of the
// trailing block, the owner of column k publishes the Householder vector, a column transposition between the halves is an
// exchange, the back substitution reads the peer's columns in place.  CLUSTER = false: one CTA, the matrix column-major in
// global memory (it stays in L2).  The 201 steps are a chain of short phases, so the cost is barriers and reduction
// latency: every reduction is a warp's (the pivot's 32 candidates, the norm of the pivot column), and the right-hand
// sides travel as two more columns of the reflection - replicated in both CTAs, Q^T applied as the reflections are formed
// (reflection k only touches rows >= k, so going past the rank changes nothing that is used).  Both instances form every
// sum with the same lanes in the same order and agree bit for bit.
namespace cg = cooperative_groups;

// reflection k on one column (or right-hand side), by one warp (Householder.h applyHouseholderOnTheLeft)
__device__ __forceinline__ void qr_reflect_column(double* cj, const double* v, int k, int P, double tau, int lane, double& tmp)
{
    double d = 0.0;
    for (int i = k + 1 + lane; i < P; i += 32) d += v[i] * cj[i];
    d = warp_sum(d);
    tmp = d + cj[k];
    __syncwarp();
    if (lane == 0) cj[k] = cj[k] - tau * tmp;
}

template <bool CLUSTER>
__global__ void __launch_bounds__(1024, 1) solve_qr_kernel(SolveArgs a)
{
    extern __shared__ __align__(16) double sm[];
    const int me = CLUSTER ? (int)cg::this_cluster().block_rank() : 0;
    const int P = a.P, tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    constexpr int STEP = CLUSTER ? 2 : 1;   // distance between two columns of this CTA
    const int ncl = CLUSTER ? (P + 1) >> 1 : 0;
    double* Aloc = CLUSTER ? sm : a.L_global; // column j at Aloc + (j / STEP) P
    double* O = sm + (size_t)ncl * P;
    double* cR = O + P;                     // right-hand sides, then Q^T b, then the triangular solution
    double* cI = cR + P;
    double* scal = cI + P;
    double* hco = scal + P;                 // Householder coefficients tau_k
    double* xR = hco + P;
    double* xI = xR + P;
    double* diag = xI + P;                  // R_kk
    double* vbuf = diag + P;                // essential part of the step's reflection (rows k + 1 ..)
    double* xbuf = vbuf + P;                // column on its way to the peer
    double* red = xbuf + P;                 // 64 doubles of reduction scratch
    double* pub = red + 64;                 // [0..1] biggest entry found by CTA 0 / 1, [2..3] its rank (as double), [4] tau, [5] beta, [6] den
    int* rowT = reinterpret_cast<int*>(pub + 8);
    int* colT = rowT + P;
    int* perm = colT + P;
    double* pub_peer = pub;
    double* vbuf_peer = vbuf;
    const double* xbuf_peer = xbuf;
    const double* Aloc_peer = Aloc;
    if (CLUSTER)
    {
        cg::cluster_group cluster = cg::this_cluster();
        pub_peer = cluster.map_shared_rank(pub, me ^ 1);
        vbuf_peer = cluster.map_shared_rank(vbuf, me ^ 1);
        xbuf_peer = cluster.map_shared_rank(xbuf, me ^ 1);
        Aloc_peer = cluster.map_shared_rank(Aloc, me ^ 1);
    }
    __shared__ int s_nonzero;
    __shared__ double s_maxpivot;

    const double* S = a.est;
    const double* FR = S + (size_t)P * P;
    const double* FI = FR + P;
    const double* Os = FI + P;
    const double* E = Os + P;
    const double n = a.est[a.cnt_offset + 2];
    const double inv = 1.0 / n;
    const double ER = E[0] * inv, EI = E[1] * inv;

    for (int i = tid; i < P; i += T) O[i] = Os[i] * inv;
    __syncthreads();
    for (int i = tid; i < P; i += T)
    {
        const double oer = FR[i] * inv, oei = FI[i] * inv;
        if (a.imaginary_time == -1)
        {
            const double rotation = 1.499 * 3.14159265358979323846;
            const double c = cos(rotation), sn = sin(rotation);
            cR[i] = c * (oer - ER * O[i]) - sn * (oei);
            cI[i] = sn * (oer - ER * O[i]) + c * (oei);
        }
        else if (a.imaginary_time == 0)
        {
            cR[i] = oei - EI * O[i];
            cI[i] = -oer + ER * O[i];
        }
        else
        {
            cR[i] = -oer + ER * O[i];
            cI[i] = -oei;
        }
        double sc = 1.0;
        if (a.use_preconditioning) // :1767-1771; the diagonal entry is formed exactly as the matrix entry below
        {
            sc = sqrt(S[(size_t)i * P + i] * inv - O[i] * O[i]);
            if (a.min_scaling > 0.0 && !(sc >= a.min_scaling)) sc = a.min_scaling;
        }
        scal[i] = sc;
    }
    __syncthreads();
    // matrix[i][j] = <O_i O_j> - <O_i><O_j> for j <= i, mirrored (:1529-1534), scaled and regularised (:1684-1711)
    const int nloc = CLUSTER ? ncl : P;
    for (int idx = tid; idx < nloc * P; idx += T)
    {
        const int i = idx % P, j = STEP * (idx / P) + me;
        if (j >= P) continue;
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        double v = S[(size_t)hi * P + lo] * inv - O[hi] * O[lo];
        if (a.use_preconditioning)
        {
            v = v / (scal[i] * scal[j]);
            if (i == j) v += a.regularization; // RegularizeEquationSystem(matrix, 0.002)
        }
        Aloc[idx] = v;
    }
    if (a.use_preconditioning)
        for (int i = tid; i < P; i += T)
        {
            cR[i] = cR[i] / scal[i];
            cI[i] = cI[i] / scal[i];
        }
    if (tid == 0)
    {
        s_nonzero = P;
        s_maxpivot = 0.0;
    }
    if (CLUSTER) cg::this_cluster().sync(); // (also: the peer has started, its shared memory may be written from here on)
    else __syncthreads();
    const double precision = 2.220446049250313e-16 * (double)P; // NumTraits<double>::epsilon() * size
    double biggest = 0.0;                                       // the first pivot (every thread keeps its copy)
    double carry_best = -1.0;
    int carry_idx = 0x7fffffff;

    for (int k = 0; k < P; k++)
    {
        // ---- biggest |a_ij| of the trailing block, first in column-major order (rank (j - k) m + (i - k)) ----
        const int m = P - k;
        double best = carry_best;
        int bidx = carry_idx;
        if (k == 0) // later steps: found while the previous reflection was applied (one pass over the block less)
        {
            for (int idx = tid; idx < nloc * P; idx += T)
            {
                const int i = idx % P, j = STEP * (idx / P) + me;
                if (j >= P) continue;
                const double v = fabs(Aloc[idx]);
                const int rk = j * P + i;
                if (v > best || (v == best && rk < bidx))
                {
                    best = v;
                    bidx = rk;
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1)
        {
            const double ob = __shfl_xor_sync(FULL_MASK, best, o);
            const int oi = __shfl_xor_sync(FULL_MASK, bidx, o);
            if (ob > best || (ob == best && oi < bidx))
            {
                best = ob;
                bidx = oi;
            }
        }
        if (lane == 0)
        {
            red[warp] = best;
            reinterpret_cast<int*>(red + 32)[warp] = bidx;
        }
        __syncthreads();
        if (warp == 0)
        {
            double b = lane < nwarp ? red[lane] : -1.0;
            int bi = lane < nwarp ? reinterpret_cast<int*>(red + 32)[lane] : 0x7fffffff;
            for (int o = 16; o > 0; o >>= 1)
            {
                const double ob = __shfl_xor_sync(FULL_MASK, b, o);
                const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
                if (ob > b || (ob == b && oi < bi))
                {
                    b = ob;
                    bi = oi;
                }
            }
            if (lane == 0)
            {
                pub[me] = b;
                pub[2 + me] = (double)bi;
                if (CLUSTER)
                {
                    pub_peer[me] = b;
                    pub_peer[2 + me] = (double)bi;
                }
            }
        }
        if (CLUSTER) cg::this_cluster().sync();
        else __syncthreads();
        double gbest = pub[0];
        int gidx = (int)pub[2];
        if (CLUSTER)
        {
            const double ob = pub[1];
            const int oi = (int)pub[3];
            if (ob > gbest || (ob == gbest && oi < gidx))
            {
                gbest = ob;
                gidx = oi;
            }
        }
        if (k == 0) biggest = gbest;
        if (gbest <= biggest * precision) // isMuchSmallerThan(biggest_in_corner, biggest, precision): |x| <= |y| * prec
        {
            if (tid == 0) s_nonzero = k;
            for (int i = k + tid; i < P; i += T)
            {
                rowT[i] = i;
                colT[i] = i;
                hco[i] = 0.0;
            }
            __syncthreads();
            break;
        }
        const int pr = k + gidx % m, pc = k + gidx / m;
        if (tid == 0)
        {
            rowT[k] = pr;
            colT[k] = pc;
        }
        if (pr != k) // m_qr.row(k).tail(cols-k).swap(m_qr.row(pr).tail(cols-k)), and the same rows of the right-hand sides
        {
            for (int j = k + (CLUSTER && (k & 1) != me ? 1 : 0) + STEP * tid; j < P; j += STEP * T)
            {
                double* cj = Aloc + (size_t)(j / STEP) * P;
                const double t = cj[k];
                cj[k] = cj[pr];
                cj[pr] = t;
            }
            if (tid == T - 1)
            {
                double t = cR[k]; cR[k] = cR[pr]; cR[pr] = t;
                t = cI[k]; cI[k] = cI[pr]; cI[pr] = t;
            }
        }
        __syncthreads();
        if (pc != k) // m_qr.col(k).swap(m_qr.col(pc)): whole columns
        {
            const bool own_k = !CLUSTER || (k & 1) == me, own_pc = !CLUSTER || (pc & 1) == me;
            if (own_k && own_pc)
            {
                double* ck = Aloc + (size_t)(k / STEP) * P;
                double* cp = Aloc + (size_t)(pc / STEP) * P;
                for (int i = tid; i < P; i += T)
                {
                    const double t = ck[i];
                    ck[i] = cp[i];
                    cp[i] = t;
                }
            }
            else if (own_k != own_pc) // one column each: exchange through distributed shared memory
            {
                double* mine = Aloc + (size_t)((own_k ? k : pc) / STEP) * P;
                for (int i = tid; i < P; i += T) xbuf[i] = mine[i];
                cg::this_cluster().sync();
                for (int i = tid; i < P; i += T) mine[i] = xbuf_peer[i];
                cg::this_cluster().sync();
            }
            __syncthreads();
        }
        // ---- makeHouseholderInPlace on column k (by its owner): norm of the tail by one warp ----
        if (!CLUSTER || (k & 1) == me)
        {
            double* colk = Aloc + (size_t)(k / STEP) * P;
            if (warp == 0)
            {
                double part = 0.0;
                for (int i = k + 1 + lane; i < P; i += 32) part += colk[i] * colk[i];
                const double tailSq = warp_sum(part);
                if (lane == 0)
                {
                    const double c0 = colk[k];
                    double beta, tau, den;
                    if (tailSq <= 2.2250738585072014e-308)
                    {
                        tau = 0.0;
                        beta = c0;
                        den = 0.0; // essential part := 0
                    }
                    else
                    {
                        beta = sqrt(c0 * c0 + tailSq);
                        if (c0 >= 0.0) beta = -beta;
                        den = c0 - beta;
                        tau = (beta - c0) / beta;
                    }
                    pub[4] = tau;
                    pub[5] = beta;
                    pub[6] = den;
                    if (CLUSTER)
                    {
                        pub_peer[4] = tau;
                        pub_peer[5] = beta;
                    }
                }
            }
            __syncthreads();
            const double den = pub[6];
            for (int i = k + 1 + tid; i < P; i += T)
            {
                const double v = den == 0.0 ? 0.0 : colk[i] / den;
                colk[i] = v;
                vbuf[i] = v;
                if (CLUSTER) vbuf_peer[i] = v;
            }
            if (tid == 0) colk[k] = pub[5];
        }
        if (CLUSTER) cg::this_cluster().sync();
        else __syncthreads();
        const double tau = pub[4], beta = pub[5];
        if (tid == 0)
        {
            hco[k] = tau;
            diag[k] = beta;
            if (fabs(beta) > s_maxpivot) s_maxpivot = fabs(beta);
        }
        // ---- applyHouseholderOnTheLeft to the columns j > k of this CTA (one warp per column; the new trailing block's
        //      biggest entry on the way) and to the two right-hand sides (the last two warps) ----
        carry_best = -1.0;
        carry_idx = 0x7fffffff;
        const int m1 = P - k - 1;
        for (int j = k + 1 + (CLUSTER && (k & 1) == me ? 1 : 0) + STEP * warp; j < P; j += STEP * nwarp)
        {
            double* cj = Aloc + (size_t)(j / STEP) * P;
            double tmp = 0.0;
            if (tau != 0.0) qr_reflect_column(cj, vbuf, k, P, tau, lane, tmp);
            for (int i = k + 1 + lane; i < P; i += 32)
            {
                double v = cj[i];
                if (tau != 0.0)
                {
                    v = v - (tau * vbuf[i]) * tmp;
                    cj[i] = v;
                }
                const double av = fabs(v);
                const int idx = (j - k - 1) * m1 + (i - k - 1);
                if (av > carry_best || (av == carry_best && idx < carry_idx))
                {
                    carry_best = av;
                    carry_idx = idx;
                }
            }
        }
        if (tau != 0.0 && warp >= nwarp - 2) // Q^T on the right-hand sides, step k (FullPivHouseholderQR::_solve_impl)
        {
            double* c = warp == nwarp - 1 ? cR : cI;
            double tmp;
            qr_reflect_column(c, vbuf, k, P, tau, lane, tmp);
            for (int i = k + 1 + lane; i < P; i += 32) c[i] = c[i] - (tau * vbuf[i]) * tmp;
        }
        __syncthreads();
    }

    // ---- rank, back substitution, permutation ----
    __syncthreads();
    const int nonzero = s_nonzero;
    const double premult = fabs(s_maxpivot) * (2.220446049250313e-16 * (double)P);
    int rank = 0;
    for (int i = 0; i < nonzero; i++) rank += (fabs(diag[i]) > premult) ? 1 : 0; // (every thread counts alike)
    // m_cols_permutation = product of the column transpositions (applyTranspositionOnTheRight)
    if (tid == 0)
    {
        for (int i = 0; i < P; i++) perm[i] = i;
        for (int k = 0; k < P; k++)
        {
            const int t = perm[k];
            perm[k] = perm[colT[k]];
            perm[colT[k]] = t;
        }
    }
    if (CLUSTER) cg::this_cluster().sync(); // both halves complete before either CTA reads the other's columns
    else __syncthreads();
    // upper-triangular solve R[0:rank, 0:rank] y = c[0:rank] (column-oriented back substitution)
    for (int i = rank - 1; i >= 0; i--)
    {
        if (tid == 0)
        {
            cR[i] = cR[i] / diag[i];
            cI[i] = cI[i] / diag[i];
        }
        __syncthreads();
        const double yR = cR[i], yI = cI[i];
        const double* ci = ((!CLUSTER || (i & 1) == me) ? Aloc : Aloc_peer) + (size_t)(i / STEP) * P;
        for (int r = tid; r < i; r += T)
        {
            const double rij = ci[r];
            cR[r] = cR[r] - rij * yR;
            cI[r] = cI[r] - rij * yI;
        }
        __syncthreads();
    }
    if (CLUSTER)
    {
        cg::this_cluster().sync(); // no CTA leaves while its shared memory may still be read
        if (me != 0) return;
    }
    for (int i = tid; i < P; i += T)
    {
        xR[perm[i]] = i < rank ? cR[i] : 0.0;
        xI[perm[i]] = i < rank ? cI[i] : 0.0;
    }
    __syncthreads();
    // mean subtraction (:1800-1809), CalculatePhiDot (:1658-1682), scalings (:1818-1825)
    if (tid == 0)
    {
        double mR = 0.0, mI = 0.0;
        for (int i = 0; i < P; i++)
        {
            mR += xR[i];
            mI += xI[i];
        }
        mR = mR / (double)P;
        mI = mI / (double)P;
        double pr = 0.0, pi = 0.0;
        for (int i = 0; i < P; i++)
        {
            xR[i] = xR[i] - mR;
            xI[i] = xI[i] - mI;
            pr -= O[i] * xR[i];
            pi -= O[i] * xI[i];
        }
        if (a.imaginary_time == -1)
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
        tail[2] = 0.0;
        tail[3] = ER;
        tail[4] = EI;
    }
    __syncthreads();
    for (int i = tid; i < P; i += T)
    {
        a.out[i] = a.use_preconditioning ? xR[i] / scal[i] : xR[i];
        a.out[P + i] = a.use_preconditioning ? xI[i] / scal[i] : xI[i];
    }
}

static size_t solve_qr_smem(int P, bool cluster)
{
    const size_t ncl = cluster ? ((size_t)P + 1) / 2 : 0;
    return (ncl * P + 11 * (size_t)P + 64 + 8) * sizeof(double) + 3 * (size_t)P * sizeof(int);
}

cudaError_t launch_solve_qr(SolveArgs a, cudaStream_t st)
{
    if (a.P < 1 || a.P > 1024 || !a.L_global) return cudaErrorInvalidValue;
    const bool cluster = !a.force_global && a.P >= 2 && solve_qr_smem(a.P, true) + 1024 <= (size_t)227 * 1024;
    const size_t smem = solve_qr_smem(a.P, cluster);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster ? 2 : 1);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cluster ? 1 : 0;
    void (*fn)(SolveArgs) = cluster ? solve_qr_kernel<true> : solve_qr_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, fn, a);
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
