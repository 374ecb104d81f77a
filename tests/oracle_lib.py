"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE (the checker, never the product)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


class OracleSystem(C.Structure):
    _fields_ = [("n_particles", C.c_int32), ("dim", C.c_int32), ("n_params", C.c_int32), ("n_splines", C.c_int32),
                ("pair_rule", C.c_int32), ("tail_param", C.c_int32), ("lbox", C.c_double), ("r_max", C.c_double),
                ("hbar2_2m", C.c_double), ("pot_a", C.c_double), ("pot_b", C.c_double), ("knots", dp), ("weights", dp),
                ("map_ptr", ip), ("map_col", ip), ("map_val", dp)]


class OracleHeBulkStruct(C.Structure):
    _fields_ = [("n_particles", C.c_int32), ("n_params", C.c_int32), ("n_splines", C.c_int32), ("gr_bins", C.c_int32),
                ("lbox", C.c_double), ("rij_split", C.c_double), ("h", C.c_double), ("max_distance", C.c_double),
                ("hbar2_2m", C.c_double), ("f", C.c_double * 8)]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
        L = C.CDLL(path)
        L.oracle_min_image.restype = C.c_double
        L.oracle_exponent.restype = C.c_double
        L.oracle_wf_quotient.restype = C.c_double
        L.oracle_sweep.restype = C.c_int64
        L.oracle_sample_walker.restype = C.c_int64
        L.oracle_hebulk_exponent.restype = C.c_double
        L.oracle_hebulk_quotient.restype = C.c_double
        L.oracle_hebulk_sweep.restype = C.c_int64
        L.oracle_hebulk_sample_walker.restype = C.c_int64
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(dp)


class Oracle:
    """Holds the arrays alive and exposes the oracle functions with numpy in/out."""

    def __init__(self, spec, time=0.0):
        self.spec = spec
        a, b = spec.potential(time)
        self._keep = [np.ascontiguousarray(spec.knots, np.float64), np.ascontiguousarray(spec.weights, np.float64),
                      np.ascontiguousarray(spec.map_ptr, np.int32), np.ascontiguousarray(spec.map_col, np.int32),
                      np.ascontiguousarray(spec.map_val, np.float64)]
        k = self._keep
        self.sys = OracleSystem(spec.n_particles, spec.dim, spec.n_params, spec.n_splines, spec.pair_rule,
                                spec.tail_param, spec.lbox, spec.r_max, spec.hbar2_2m, a, b, _d(k[0]), _d(k[1]),
                                k[2].ctypes.data_as(ip), k[3].ctypes.data_as(ip), _d(k[4]))
        self.N, self.D, self.P, self.K = spec.n_particles, spec.dim, spec.n_params, spec.n_splines

    def min_image(self, L, a, b):
        a = np.ascontiguousarray(a, np.float64)
        b = np.ascontiguousarray(b, np.float64)
        disp = np.zeros(3)
        n = lib().oracle_min_image(C.c_double(L), 3, _d(a), _d(b), _d(disp))
        return n, disp

    def basis_sums(self, R):
        R = np.ascontiguousarray(R, np.float64)
        ss = np.zeros(self.K)
        outer = C.c_double(0)
        lib().oracle_basis_sums(C.byref(self.sys), _d(R), _d(ss), C.byref(outer))
        return ss, outer.value

    def local_operators(self, ss):
        O = np.zeros(self.P)
        lib().oracle_local_operators(C.byref(self.sys), _d(np.ascontiguousarray(ss)), _d(O))
        return O

    def exponent(self, O, outer, uR):
        return lib().oracle_exponent(C.byref(self.sys), _d(np.ascontiguousarray(O)), C.c_double(outer),
                                     _d(np.ascontiguousarray(uR, np.float64)))

    def tables(self, R):
        R = np.ascontiguousarray(R, np.float64)
        sD = np.zeros((self.K, self.N, self.D))
        sD2 = np.zeros((self.K, self.N))
        v = C.c_double(0)
        lib().oracle_tables(C.byref(self.sys), _d(R), _d(sD), _d(sD2), C.byref(v))
        return sD, sD2, v.value

    def expectation(self, O, sD, sD2, v_int, exponent, phiR, uR, uI):
        er, ei = C.c_double(0), C.c_double(0)
        other = np.zeros(9)
        dr = np.zeros((self.N, self.D))
        di = np.zeros((self.N, self.D))
        lib().oracle_expectation(C.byref(self.sys), _d(np.ascontiguousarray(O)), _d(sD), _d(sD2), C.c_double(v_int),
                                 C.c_double(exponent), C.c_double(phiR), _d(np.ascontiguousarray(uR, np.float64)),
                                 _d(np.ascontiguousarray(uI, np.float64)), C.byref(er), C.byref(ei), _d(other), _d(dr), _d(di))
        return er.value, ei.value, other, dr, di

    def evaluate(self, R, uR, uI, phiR=0.0):
        """CalculateWavefunction + CalculateExpectationValues on one configuration."""
        ss, outer = self.basis_sums(R)
        O = self.local_operators(ss)
        ex = self.exponent(O, outer, uR)
        sD, sD2, v = self.tables(R)
        er, ei, other, dr, di = self.expectation(O, sD, sD2, v, ex, phiR, uR, uI)
        return dict(ss=ss, outer=outer, O=O, exponent=ex, sD=sD, sD2=sD2, v_int=v, e_r=er, e_i=ei, other=other,
                    drift_r=dr, drift_i=di)

    def quotient(self, R, particle, new_pos, uR):
        R = np.array(R, np.float64)
        ss, outer = self.basis_sums(R)
        ex = self.exponent(self.local_operators(ss), outer, uR)
        old = R[particle].copy()
        R[particle] = new_pos
        ss_new = np.zeros(self.K)
        on, en = C.c_double(0), C.c_double(0)
        q = lib().oracle_wf_quotient(C.byref(self.sys), _d(R), int(particle), _d(old), _d(ss), C.c_double(outer),
                                     C.c_double(ex), _d(np.ascontiguousarray(uR, np.float64)), _d(ss_new), C.byref(on), C.byref(en))
        return q, en.value, ex

    def proposal(self, seed, walker, step, mc_step):
        p = C.c_int(0)
        disp = np.zeros(3)
        lu = C.c_double(0)
        lib().oracle_proposal(C.c_uint64(seed), C.c_uint32(walker), C.c_uint64(step), self.N, C.c_double(mc_step),
                              C.byref(p), _d(disp), C.byref(lu))
        return p.value, disp, lu.value

    def sweep(self, R, uR, seed, walker, first_step, n_steps, mc_step):
        """Runs n_steps Metropolis steps in place on a copy of R; returns (R, accepted)."""
        R = np.array(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        ss, outer = self.basis_sums(R)
        ex = C.c_double(self.exponent(self.local_operators(ss), outer, uR))
        o = C.c_double(outer)
        acc = lib().oracle_sweep(C.byref(self.sys), _d(R), _d(ss), C.byref(o), C.byref(ex), _d(uR), C.c_uint64(seed),
                                 C.c_uint32(walker), C.c_uint64(first_step), C.c_int64(n_steps), C.c_double(mc_step))
        return R, int(acc)

    def est_size(self):
        return self.P * self.P + 3 * self.P + 2 + 9

    def sample_walker(self, R, uR, uI, phiR, seed, walker, step0, n_init, n_samples, n_therm, mc_step, est=None):
        R = np.array(R, np.float64)
        if est is None:
            est = np.zeros(self.est_size())
        rows = np.zeros((n_samples, self.P + 2))
        sc = C.c_uint64(step0)
        acc = lib().oracle_sample_walker(C.byref(self.sys), _d(R), _d(np.ascontiguousarray(uR, np.float64)),
                                         _d(np.ascontiguousarray(uI, np.float64)), C.c_double(phiR), C.c_uint64(seed),
                                         C.c_uint32(walker), C.byref(sc), n_init, n_samples, n_therm, C.c_double(mc_step),
                                         _d(est), _d(rows))
        return dict(R=R, est=est, rows=rows, accepted=int(acc), steps=sc.value)

    def unpack_est(self, est, n):
        P = self.P
        return dict(O=est[:P] / n, e_r=est[P] / n, e_i=est[P + 1] / n, S=est[P + 2:P + 2 + P * P].reshape(P, P) / n,
                    OER=est[P + 2 + P * P:P + 2 + P * P + P] / n, OEI=est[P + 2 + P * P + P:P + 2 + P * P + 2 * P] / n,
                    other=est[P + 2 + P * P + 2 * P:] / n)


class OracleHeBulk:
    """HeBulk restatement (oracle/tdvmc_oracle_he.c) with numpy in/out."""

    def __init__(self, spec, time=0.0):
        self.spec = spec
        self.sys = OracleHeBulkStruct()
        lib().oracle_hebulk_init(C.byref(self.sys), spec.n_particles, C.c_double(spec.lbox), spec.n_params)
        self.N, self.P, self.K, self.NO = spec.n_particles, spec.n_params, self.sys.n_splines, 3 + self.sys.gr_bins

    def values(self, R):
        R = np.ascontiguousarray(R, np.float64)
        ss = np.zeros(self.K)
        mcm = C.c_double(0)
        lib().oracle_hebulk_values(C.byref(self.sys), _d(R), _d(ss), C.byref(mcm))
        return ss, mcm.value

    def operators(self, ss, mcm):
        O = np.zeros(self.P)
        lib().oracle_hebulk_operators(C.byref(self.sys), _d(np.ascontiguousarray(ss)), C.c_double(mcm), _d(O))
        return O

    def exponent(self, ss, mcm, uR):
        return lib().oracle_hebulk_exponent(C.byref(self.sys), _d(np.ascontiguousarray(ss)), C.c_double(mcm),
                                            _d(np.ascontiguousarray(uR, np.float64)))

    def evaluate(self, R, uR, uI, phiR=0.0):
        R = np.ascontiguousarray(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        uI = np.ascontiguousarray(uI, np.float64)
        ss, mcm = self.values(R)
        ex = self.exponent(ss, mcm, uR)
        er, ei = C.c_double(0), C.c_double(0)
        other = np.zeros(self.NO)
        dr, di = np.zeros((self.N, 3)), np.zeros((self.N, 3))
        sD, sD2 = np.zeros((self.K, self.N, 3)), np.zeros((self.K, self.N))
        mcD, mcD2 = np.zeros((self.N, 3)), np.zeros(self.N)
        lib().oracle_hebulk_expectation(C.byref(self.sys), _d(R), C.c_double(np.exp(ex)), _d(uR), _d(uI), C.byref(er), C.byref(ei),
                                        _d(other), _d(dr), _d(di), _d(sD), _d(sD2), _d(mcD), _d(mcD2))
        return dict(ss=ss, mcm=mcm, O=self.operators(ss, mcm), exponent=ex, e_r=er.value, e_i=ei.value, other=other,
                    drift_r=dr, drift_i=di, sD=sD, sD2=sD2, mcD=mcD, mcD2=mcD2)

    def quotient(self, R, particle, new_pos, uR):
        R = np.array(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        ss, mcm = self.values(R)
        ex = self.exponent(ss, mcm, uR)
        old = R[particle].copy()
        R[particle] = new_pos
        ss_new = np.zeros(self.K)
        mn, en = C.c_double(0), C.c_double(0)
        q = lib().oracle_hebulk_quotient(C.byref(self.sys), _d(R), int(particle), _d(old), _d(ss), C.c_double(mcm),
                                         C.c_double(ex), _d(uR), _d(ss_new), C.byref(mn), C.byref(en))
        return q, en.value, ex

    def sweep(self, R, uR, seed, walker, first_step, n_steps, mc_step):
        R = np.array(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        ss, mcm = self.values(R)
        ex = C.c_double(self.exponent(ss, mcm, uR))
        m = C.c_double(mcm)
        acc = lib().oracle_hebulk_sweep(C.byref(self.sys), _d(R), _d(ss), C.byref(m), C.byref(ex), _d(uR), C.c_uint64(seed),
                                        C.c_uint32(walker), C.c_uint64(first_step), C.c_int64(n_steps), C.c_double(mc_step))
        return R, int(acc)

    def est_size(self):
        return self.P * self.P + 3 * self.P + 2 + self.NO

    def sample_walker(self, R, uR, uI, phiR, seed, walker, step0, n_init, n_samples, n_therm, mc_step, est=None):
        R = np.array(R, np.float64)
        if est is None:
            est = np.zeros(self.est_size())
        rows = np.zeros((n_samples, self.P + 2))
        sc = C.c_uint64(step0)
        acc = lib().oracle_hebulk_sample_walker(C.byref(self.sys), _d(R), _d(np.ascontiguousarray(uR, np.float64)),
                                                _d(np.ascontiguousarray(uI, np.float64)), C.c_uint64(seed), C.c_uint32(walker),
                                                C.byref(sc), n_init, n_samples, n_therm, C.c_double(mc_step), _d(est), _d(rows))
        return dict(R=R, est=est, rows=rows, accepted=int(acc), steps=sc.value)

    def unpack_est(self, est, n):
        return Oracle.unpack_est(self, est, n)


def make_oracle(spec, time=0.0):
    from tdvmc_b200 import systems

    return OracleHeBulk(spec, time) if spec.kind == systems.KIND_HE_BULK else Oracle(spec, time)
