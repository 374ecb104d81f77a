"""Host side of one TDVMC time step, mirroring the reference driver (numpy; nothing here runs on the GPU and nothing
here is on the accelerated path - BASELINE north_star keeps the small parameter solve on the host):

    BuildSystemOfEquationsForParametersIncludePhi   src/TDVMC.cpp:1506-1537
    PreconditionEquationSystemByScaling             :1684-1701
    RegularizeEquationSystem                        :1703-1711
    PerformCholeskyDecomposition / Solve...         :1560-1622
    CalculatePhiDot                                 :1658-1682
    SolveForParametersDot (Cholesky branch)         :1713-1763
    CalculateNextParametersEuler                    :1834-1853

Input is the estimator dictionary GpuEnsembleSystem.ParallelUpdateExpectationValues returns (the reference's seven
global arrays by name).  Used by the time-evolution parity test and by examples; the reference driver keeps using its
own C++ versions of these functions."""
import numpy as np


def build_system_of_equations(est, imaginary_time=1):
    O = np.asarray(est["localOperators"], np.float64)
    S = np.asarray(est["localOperatorsMatrix"], np.float64)
    OER = np.asarray(est["localOperatorlocalEnergyR"], np.float64)
    OEI = np.asarray(est["localOperatorlocalEnergyI"], np.float64)
    ER, EI = float(est["localEnergyR"]), float(est["localEnergyI"])
    if imaginary_time == 0:
        b_r = OEI - EI * O
        b_i = -OER + ER * O
    elif imaginary_time == 1:
        b_r = -OER + ER * O
        b_i = -OEI                      # :1526 - the imaginary right-hand side is not centred in imaginary time
    elif imaginary_time == -1:          # BuildSystemOfEquationsForParametersIncludePhiWithTimeRotation, :1475-1504
        rotation = 1.499 * np.pi
        c, sn = np.cos(rotation), np.sin(rotation)
        b_r = c * (OER - ER * O) - sn * OEI
        b_i = sn * (OER - ER * O) + c * OEI
    else:
        raise ValueError("IMAGINARY_TIME must be -1, 0 or 1")
    A = S - np.outer(O, O)
    A = np.tril(A) + np.tril(A, -1).T    # the reference fills the lower triangle and mirrors it (:1528-1533)
    return A, b_r, b_i


def cholesky_solve(A, rhs_list):
    """Cholesky-Banachiewicz as :1560-1592 (no pivoting; raises if not positive definite, where the reference sets
    doNotAcceptStep) and the two triangular solves of :1594-1622."""
    n = A.shape[0]
    L = np.zeros_like(A)
    for i in range(n):
        for j in range(i + 1):
            s = A[i, j] - np.dot(L[i, :j], L[j, :j])
            if i > j:
                L[i, j] = s / L[j, j]
            elif s > 0:
                L[i, i] = np.sqrt(s)
            else:
                raise np.linalg.LinAlgError(f"not positive definite at i={i}")
    out = []
    for rhs in rhs_list:
        tmp = np.zeros(n)
        for i in range(n):
            tmp[i] = 1.0 / L[i, i] * (rhs[i] - np.dot(L[i, :i], tmp[:i]))
        x = np.zeros(n)
        for i in range(n - 1, -1, -1):
            x[i] = 1.0 / L[i, i] * (tmp[i] - np.dot(L[i + 1:, i], x[i + 1:]))
        out.append(x)
    return out


def solve_for_parameters_dot(est, imaginary_time=1, use_preconditioning=True, regularization=None, lapack=False,
                             min_scaling=0.0, solver_type=0):
    """SolveForParametersDot with LINEAR_EQUATION_SOLVER_TYPE = 0: returns (uDotR, uDotI, phiDotR, phiDotI).
    lapack=True factorises with numpy's LAPACK (dpotrf/dpotrs through numpy.linalg) instead of the reference's
    hand-written loops - same matrix, same right-hand sides, results equal to rounding."""
    if regularization is None:
        regularization = 0.002 if solver_type == 1 else 0.001          # :1770, :1737
    A, b_r, b_i = build_system_of_equations(est, imaginary_time)
    if solver_type == 1:
        return _solve_qr_branch(est, A, b_r, b_i, imaginary_time, use_preconditioning, regularization, min_scaling)
    scal = np.ones(len(b_r))
    if use_preconditioning:
        scal = np.sqrt(np.maximum(np.diag(A), 0.0))
        if min_scaling > 0.0:            # a parameter whose operator never varied (empty knot interval): the reference would
            scal = np.maximum(scal, min_scaling)   # divide by zero here; callers that cannot rule it out pass a floor
        A = A / np.outer(scal, scal)
        b_r = b_r / scal
        b_i = b_i / scal
    A = A + regularization * np.eye(len(b_r))
    if lapack:
        L = np.linalg.cholesky(A)
        y = np.linalg.solve(L, np.stack([b_r, b_i], axis=1))
        x = np.linalg.solve(L.T, y)
        u_r, u_i = x[:, 0], x[:, 1]
    else:
        u_r, u_i = cholesky_solve(A, [b_r, b_i])
    O = np.asarray(est["localOperators"], np.float64)
    phi_r = -float(np.dot(O, u_r))      # CalculatePhiDot runs BEFORE the scalings are divided out (:1743-1752)
    phi_i = -float(np.dot(O, u_i))
    if imaginary_time == -1:            # CalculatePhiDot, :1666-1673
        rotation = 1.499 * np.pi
        phi_i -= np.cos(rotation) * float(est["localEnergyR"])
        phi_r -= np.sin(rotation) * float(est["localEnergyR"])
    elif imaginary_time == 0:
        phi_i -= float(est["localEnergyR"])
    else:
        phi_r -= float(est["localEnergyR"])
    return u_r / scal, u_i / scal, phi_r, phi_i


def _phi_dot(est, u_r, u_i, imaginary_time):
    O = np.asarray(est["localOperators"], np.float64)
    phi_r = -float(np.dot(O, u_r))
    phi_i = -float(np.dot(O, u_i))
    if imaginary_time == -1:
        rotation = 1.499 * np.pi
        phi_i -= np.cos(rotation) * float(est["localEnergyR"])
        phi_r -= np.sin(rotation) * float(est["localEnergyR"])
    elif imaginary_time == 0:
        phi_i -= float(est["localEnergyR"])
    else:
        phi_r -= float(est["localEnergyR"])
    return phi_r, phi_i


def _solve_qr_branch(est, A, b_r, b_i, imaginary_time, use_preconditioning, regularization, min_scaling):
    """LINEAR_EQUATION_SOLVER_TYPE = 1 (src/TDVMC.cpp:1763-1827): scaling and +0.002 only with USE_PRECONDITIONING, a
    rank-revealing solve (Eigen FullPivHouseholderQR there, LAPACK's least-squares here - the same solution wherever the
    matrix has full numerical rank), the mean of each solution subtracted, CalculatePhiDot, scalings divided out."""
    scal = np.ones(len(b_r))
    if use_preconditioning:
        scal = np.sqrt(np.maximum(np.diag(A), 0.0))
        if min_scaling > 0.0:
            scal = np.maximum(scal, min_scaling)
        A = A / np.outer(scal, scal) + regularization * np.eye(len(b_r))
        b_r, b_i = b_r / scal, b_i / scal
    x = np.linalg.lstsq(A, np.stack([b_r, b_i], axis=1), rcond=len(b_r) * np.finfo(float).eps)[0]
    u_r = x[:, 0] - x[:, 0].mean()
    u_i = x[:, 1] - x[:, 1].mean()
    phi_r, phi_i = _phi_dot(est, u_r, u_i, imaginary_time)
    return u_r / scal, u_i / scal, phi_r, phi_i


def euler_step(dt, uR, uI, phiR, phiI, est, **kw):
    """CalculateNextParametersEuler: returns the new (uR, uI, phiR, phiI)."""
    du_r, du_i, dphi_r, dphi_i = solve_for_parameters_dot(est, **kw)
    return uR + du_r * dt, uI + du_i * dt, phiR + dphi_r * dt, phiI + dphi_i * dt
