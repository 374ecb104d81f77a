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


class OracleHeStruct(C.Structure):
    _fields_ = [("n_particles", C.c_int32), ("n_params", C.c_int32), ("n_splines", C.c_int32), ("n_short", C.c_int32),
                ("periodic", C.c_int32), ("potential", C.c_int32), ("gr_bins", C.c_int32), ("rho_bins", C.c_int32),
                ("use_phi", C.c_int32), ("pad", C.c_int32), ("lbox", C.c_double), ("rs", C.c_double), ("r_split2", C.c_double),
                ("r_tail", C.c_double), ("h_short", C.c_double), ("h_large", C.c_double), ("max_distance", C.c_double),
                ("mcm", C.c_double), ("gr_max", C.c_double), ("hbar2_2m", C.c_double), ("map_ptr", ip), ("map_col", ip),
                ("map_val", dp), ("map_const", dp), ("grad_const", dp)]


class OracleMixStruct(C.Structure):
    _fields_ = [("n_particles", C.c_int32), ("n_params", C.c_int32), ("n_types", C.c_int32), ("n_splines", C.c_int32),
                ("n_other", C.c_int32), ("order", C.c_int32), ("pair_type", ip), ("hbar", dp), ("mass", dp), ("knots", dp),
                ("weights", dp), ("mcm", dp), ("potential", ip), ("map_ptr", ip), ("map_col", ip), ("map_val", dp)]


class OracleBRStruct(C.Structure):
    _fields_ = [("n_particles", C.c_int32), ("n_params", C.c_int32), ("n_splines", C.c_int32), ("gr_bins", C.c_int32),
                ("dim", C.c_int32), ("pad", C.c_int32),
                ("lbox", C.c_double), ("r_max", C.c_double), ("pot_a", C.c_double), ("pot_b", C.c_double), ("gr_max", C.c_double),
                ("gr_spacing", C.c_double), ("gr_volumes", dp), ("knots", dp), ("weights", dp), ("map_ptr", ip), ("map_col", ip),
                ("map_val", dp)]


class OracleInhStruct(C.Structure):
    _fields_ = [("n_particles", C.c_int32), ("n_params", C.c_int32), ("n_splines_spf", C.c_int32), ("n_splines_pc", C.c_int32),
                ("lbox", C.c_double), ("r_max", C.c_double), ("h_pc", C.c_double), ("gamma", C.c_double), ("pot_range", C.c_double),
                ("pot_strength", C.c_double), ("ext_k", C.c_double), ("ext_v0", C.c_double), ("hbar2_2m", C.c_double),
                ("knots_spf", dp), ("weights_spf", dp), ("knots_pc", dp), ("weights_pc", dp), ("map_ptr", ip), ("map_col", ip),
                ("map_val", dp)]


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
        L.oracle_mix_exponent.restype = C.c_double
        L.oracle_mix_quotient.restype = C.c_double
        L.oracle_mix_sweep.restype = C.c_int64
        L.oracle_mix_sample_walker.restype = C.c_int64
        L.oracle_pair_potential.restype = C.c_double
        L.oracle_he_exponent.restype = C.c_double
        L.oracle_he_quotient.restype = C.c_double
        L.oracle_he_sweep.restype = C.c_int64
        L.oracle_he_sample_walker.restype = C.c_int64
        L.oracle_inh_exponent.restype = C.c_double
        L.oracle_inh_quotient.restype = C.c_double
        L.oracle_inh_sweep.restype = C.c_int64
        L.oracle_inh_sample_walker.restype = C.c_int64
        L.oracle_br_exponent.restype = C.c_double
        L.oracle_br_quotient.restype = C.c_double
        L.oracle_br_sweep.restype = C.c_int64
        L.oracle_br_sample_walker.restype = C.c_int64
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


class OracleHe:
    """He family restatement (oracle/tdvmc_oracle_he.c: HeBulk, HeDrop) with numpy in/out."""

    def __init__(self, spec, time=0.0):
        self.spec = spec
        e = spec.extra
        self._keep = [np.ascontiguousarray(spec.map_ptr, np.int32), np.ascontiguousarray(spec.map_col, np.int32),
                      np.ascontiguousarray(spec.map_val, np.float64), np.ascontiguousarray(spec.map_const, np.float64),
                      np.ascontiguousarray(spec.grad_const, np.float64)]
        k = self._keep
        big = 1e300
        f = lambda x: float(min(x, big))
        self.sys = OracleHeStruct(spec.n_particles, spec.n_params, e["n_splines"], e["n_short"], e["periodic"], e["potential"],
                                  e["gr_bins"], e["rho_bins"], e["use_phi"], 0, spec.lbox, e["rij_split"], f(e["r_split2"]),
                                  f(e["r_tail"]), e["h"], e["h_large"], f(e["r_max"]), e["mcm"], e["gr_max"], spec.hbar2_2m,
                                  k[0].ctypes.data_as(ip), k[1].ctypes.data_as(ip), _d(k[2]), _d(k[3]), _d(k[4]))
        self.N, self.P, self.K, self.NE = spec.n_particles, spec.n_params, e["n_splines"], e["n_splines"] + 3
        self.NO = 3 + e["gr_bins"] + e["rho_bins"]
        self.use_phi = e["use_phi"]

    def values(self, R):
        R = np.ascontiguousarray(R, np.float64)
        ext = np.zeros(self.NE)
        lib().oracle_he_values(C.byref(self.sys), _d(R), _d(ext))
        return ext

    def operators(self, ext):
        O = np.zeros(self.P)
        lib().oracle_he_operators(C.byref(self.sys), _d(np.ascontiguousarray(ext)), _d(O))
        return O

    def exponent(self, ext, uR):
        return lib().oracle_he_exponent(C.byref(self.sys), _d(np.ascontiguousarray(ext)), _d(np.ascontiguousarray(uR, np.float64)))

    def evaluate(self, R, uR, uI, phiR=0.0):
        R = np.ascontiguousarray(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        uI = np.ascontiguousarray(uI, np.float64)
        ext = self.values(R)
        ex = self.exponent(ext, uR)
        wf = np.exp(ex + phiR) if self.use_phi else np.exp(ex)
        er, ei = C.c_double(0), C.c_double(0)
        other = np.zeros(self.NO)
        dr, di = np.zeros((self.N, 3)), np.zeros((self.N, 3))
        tD, tD2 = np.zeros((self.NE, self.N, 3)), np.zeros((self.NE, self.N))
        lib().oracle_he_expectation(C.byref(self.sys), _d(R), C.c_double(wf), _d(uR), _d(uI), C.byref(er), C.byref(ei), _d(other),
                                    _d(dr), _d(di), _d(tD), _d(tD2))
        return dict(ext=ext, ss=ext[:self.K], mcm=ext[self.K], O=self.operators(ext), exponent=ex, e_r=er.value, e_i=ei.value,
                    other=other, drift_r=dr, drift_i=di, tabD=tD, tabD2=tD2)

    def quotient(self, R, particle, new_pos, uR):
        R = np.array(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        ext = self.values(R)
        ex = self.exponent(ext, uR)
        old = R[particle].copy()
        R[particle] = new_pos
        ext_new = np.zeros(self.NE)
        en = C.c_double(0)
        q = lib().oracle_he_quotient(C.byref(self.sys), _d(R), int(particle), _d(old), _d(ext), C.c_double(ex), _d(uR),
                                     _d(ext_new), C.byref(en))
        return q, en.value, ex

    def sweep(self, R, uR, seed, walker, first_step, n_steps, mc_step):
        R = np.array(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        ext = self.values(R)
        ex = C.c_double(self.exponent(ext, uR))
        acc = lib().oracle_he_sweep(C.byref(self.sys), _d(R), _d(ext), C.byref(ex), _d(uR), C.c_uint64(seed), C.c_uint32(walker),
                                    C.c_uint64(first_step), C.c_int64(n_steps), C.c_double(mc_step))
        return R, int(acc)

    def est_size(self):
        return self.P * self.P + 3 * self.P + 2 + self.NO

    def sample_walker(self, R, uR, uI, phiR, seed, walker, step0, n_init, n_samples, n_therm, mc_step, est=None):
        R = np.array(R, np.float64)
        if est is None:
            est = np.zeros(self.est_size())
        rows = np.zeros((n_samples, self.P + 2))
        sc = C.c_uint64(step0)
        acc = lib().oracle_he_sample_walker(C.byref(self.sys), _d(R), _d(np.ascontiguousarray(uR, np.float64)),
                                            _d(np.ascontiguousarray(uI, np.float64)), C.c_double(phiR), C.c_uint64(seed),
                                            C.c_uint32(walker), C.byref(sc), n_init, n_samples, n_therm, C.c_double(mc_step),
                                            _d(est), _d(rows))
        return dict(R=R, est=est, rows=rows, accepted=int(acc), steps=sc.value)

    def unpack_est(self, est, n):
        return Oracle.unpack_est(self, est, n)


def br_shell_volumes(half, n_bins, dim=3):
    """grBinVolumes of NUBosonsBulkPBBoxAndRadial::InitSystem (:146-169), same operations in the same order."""
    import math
    spacing = half / float(n_bins)
    if dim == 3:
        v = [4.0 * math.pi * math.pow(spacing * (i + 1), 3.0) / 3.0 for i in range(n_bins)]
    elif dim == 2:
        v = [math.pi * math.pow(spacing * (i + 1), 2.0) for i in range(n_bins)]
    else:
        v = [2.0 * (spacing * (i + 1)) for i in range(n_bins)]
    for i in range(n_bins - 1, 0, -1):
        v[i] = v[i] - v[i - 1]
    return np.array(v), spacing


class OracleBR:
    """NUBosonsBulkPBBoxAndRadial restatement (oracle/tdvmc_oracle_br.c) with numpy in/out; ext = [ssRad | ss]."""

    def __init__(self, spec, time=0.0):
        self.spec = spec
        e = spec.extra
        a, b = spec.potential(time)
        vol, spacing = br_shell_volumes(e["half"], e["gr_bins"], spec.dim)
        self._keep = [vol, np.ascontiguousarray(spec.knots, np.float64), np.ascontiguousarray(spec.weights, np.float64),
                      np.ascontiguousarray(spec.map_ptr, np.int32), np.ascontiguousarray(spec.map_col, np.int32),
                      np.ascontiguousarray(spec.map_val, np.float64)]
        k = self._keep
        self.sys = OracleBRStruct(spec.n_particles, spec.n_params, e["n_splines"], e["gr_bins"], spec.dim, 0, spec.lbox, spec.r_max, a, b,
                                  e["half"], spacing, _d(k[0]), _d(k[1]), _d(k[2]), k[3].ctypes.data_as(ip),
                                  k[4].ctypes.data_as(ip), _d(k[5]))
        self.N, self.P, self.K = spec.n_particles, spec.n_params, e["n_splines"]
        self.NE, self.NO = 2 * self.K, 3 + e["gr_bins"]

    def values(self, R):
        ext = np.zeros(self.NE)
        lib().oracle_br_values(C.byref(self.sys), _d(np.ascontiguousarray(R, np.float64)), _d(ext))
        return ext

    def operators(self, ext):
        O = np.zeros(self.P)
        lib().oracle_br_operators(C.byref(self.sys), _d(np.ascontiguousarray(ext)), _d(O))
        return O

    def exponent(self, ext, uR):
        return lib().oracle_br_exponent(C.byref(self.sys), _d(np.ascontiguousarray(ext)), _d(np.ascontiguousarray(uR, np.float64)))

    def evaluate(self, R, uR, uI, phiR=0.0):
        R = np.ascontiguousarray(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        uI = np.ascontiguousarray(uI, np.float64)
        ext = self.values(R)
        ex = self.exponent(ext, uR)
        er, ei = C.c_double(0), C.c_double(0)
        other = np.zeros(self.NO)
        dr, di = np.zeros((self.N, 3)), np.zeros((self.N, 3))
        tD, tD2 = np.zeros((self.NE, self.N, 3)), np.zeros((self.NE, self.N))
        lib().oracle_br_expectation(C.byref(self.sys), _d(R), C.c_double(np.exp(ex + phiR)), _d(uR), _d(uI), C.byref(er),
                                    C.byref(ei), _d(other), _d(dr), _d(di), _d(tD), _d(tD2))
        return dict(ext=ext, O=self.operators(ext), exponent=ex, e_r=er.value, e_i=ei.value, other=other, drift_r=dr,
                    drift_i=di, tabD=tD, tabD2=tD2)

    def quotient(self, R, particle, new_pos, uR):
        R = np.array(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        ext = self.values(R)
        ex = self.exponent(ext, uR)
        old = R[particle].copy()
        R[particle] = new_pos
        ext_new = np.zeros(self.NE)
        en = C.c_double(0)
        q = lib().oracle_br_quotient(C.byref(self.sys), _d(R), int(particle), _d(old), _d(ext), C.c_double(ex), _d(uR),
                                     _d(ext_new), C.byref(en))
        return q, en.value, ex

    def sweep(self, R, uR, seed, walker, first_step, n_steps, mc_step):
        R = np.array(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        ext = self.values(R)
        ex = C.c_double(self.exponent(ext, uR))
        acc = lib().oracle_br_sweep(C.byref(self.sys), _d(R), _d(ext), C.byref(ex), _d(uR), C.c_uint64(seed), C.c_uint32(walker),
                                    C.c_uint64(first_step), C.c_int64(n_steps), C.c_double(mc_step))
        return R, int(acc)

    def est_size(self):
        return self.P * self.P + 3 * self.P + 2 + self.NO

    def sample_walker(self, R, uR, uI, phiR, seed, walker, step0, n_init, n_samples, n_therm, mc_step, est=None):
        R = np.array(R, np.float64)
        if est is None:
            est = np.zeros(self.est_size())
        rows = np.zeros((n_samples, self.P + 2))
        sc = C.c_uint64(step0)
        acc = lib().oracle_br_sample_walker(C.byref(self.sys), _d(R), _d(np.ascontiguousarray(uR, np.float64)),
                                            _d(np.ascontiguousarray(uI, np.float64)), C.c_double(phiR), C.c_uint64(seed),
                                            C.c_uint32(walker), C.byref(sc), n_init, n_samples, n_therm, C.c_double(mc_step),
                                            _d(est), _d(rows))
        return dict(R=R, est=est, rows=rows, accepted=int(acc), steps=sc.value)

    def unpack_est(self, est, n):
        return Oracle.unpack_est(self, est, n)


class OracleInh:
    """InhContactBosons restatement (oracle/tdvmc_oracle_inh.c), one-dimensional; R is [N][3] with the coordinate in
    component 0; ext = [ss_spf | ss_pc]."""

    def __init__(self, spec, time=0.0):
        self.spec = spec
        e = spec.extra
        k1, k2 = e["n_splines_spf"], e["n_splines_pc"]
        kn = np.ascontiguousarray(spec.knots, np.float64)
        w = np.ascontiguousarray(spec.weights, np.float64).reshape(-1)
        self._keep = [np.ascontiguousarray(kn[:k1 + 4]), np.ascontiguousarray(w[:k1 * 16]), np.ascontiguousarray(kn[k1 + 4:]),
                      np.ascontiguousarray(w[k1 * 16:]), np.ascontiguousarray(spec.map_ptr, np.int32),
                      np.ascontiguousarray(spec.map_col, np.int32), np.ascontiguousarray(spec.map_val, np.float64)]
        k = self._keep
        sp = spec.system_params
        self.sys = OracleInhStruct(spec.n_particles, spec.n_params, k1, k2, spec.lbox, spec.r_max, e["h_pc"], e["gamma"],
                                   float(sp[0]), float(sp[1]), float(sp[2]), float(sp[3]), spec.hbar2_2m, _d(k[0]), _d(k[1]),
                                   _d(k[2]), _d(k[3]), k[4].ctypes.data_as(ip), k[5].ctypes.data_as(ip), _d(k[6]))
        self.N, self.P, self.NE, self.NO = spec.n_particles, spec.n_params, k1 + k2, 9

    def values(self, R):
        ext = np.zeros(self.NE)
        lib().oracle_inh_values(C.byref(self.sys), _d(np.ascontiguousarray(R, np.float64)), _d(ext))
        return ext

    def operators(self, ext):
        O = np.zeros(self.P)
        lib().oracle_inh_operators(C.byref(self.sys), _d(np.ascontiguousarray(ext)), _d(O))
        return O

    def exponent(self, ext, uR):
        return lib().oracle_inh_exponent(C.byref(self.sys), _d(np.ascontiguousarray(ext)), _d(np.ascontiguousarray(uR, np.float64)))

    def evaluate(self, R, uR, uI, phiR=0.0):
        R = np.ascontiguousarray(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        uI = np.ascontiguousarray(uI, np.float64)
        ext = self.values(R)
        ex = self.exponent(ext, uR)
        er, ei = C.c_double(0), C.c_double(0)
        other = np.zeros(self.NO)
        dr, di = np.zeros((self.N, 3)), np.zeros((self.N, 3))
        tD, tD2 = np.zeros((self.NE, self.N)), np.zeros((self.NE, self.N))
        lib().oracle_inh_expectation(C.byref(self.sys), _d(R), C.c_double(np.exp(ex + phiR)), C.c_double(ex), _d(uR), _d(uI),
                                     C.byref(er), C.byref(ei), _d(other), _d(dr), _d(di), _d(tD), _d(tD2))
        return dict(ext=ext, O=self.operators(ext), exponent=ex, e_r=er.value, e_i=ei.value, other=other, drift_r=dr,
                    drift_i=di, tabD=tD, tabD2=tD2)

    def quotient(self, R, particle, new_pos, uR):
        R = np.array(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        ext = self.values(R)
        ex = self.exponent(ext, uR)
        old = R[particle].copy()
        R[particle] = new_pos
        ext_new = np.zeros(self.NE)
        en = C.c_double(0)
        q = lib().oracle_inh_quotient(C.byref(self.sys), _d(R), int(particle), _d(old), _d(ext), C.c_double(ex), _d(uR),
                                      _d(ext_new), C.byref(en))
        return q, en.value, ex

    def sweep(self, R, uR, seed, walker, first_step, n_steps, mc_step):
        R = np.array(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        ext = self.values(R)
        ex = C.c_double(self.exponent(ext, uR))
        acc = lib().oracle_inh_sweep(C.byref(self.sys), _d(R), _d(ext), C.byref(ex), _d(uR), C.c_uint64(seed), C.c_uint32(walker),
                                     C.c_uint64(first_step), C.c_int64(n_steps), C.c_double(mc_step))
        return R, int(acc)

    def est_size(self):
        return self.P * self.P + 3 * self.P + 2 + self.NO

    def sample_walker(self, R, uR, uI, phiR, seed, walker, step0, n_init, n_samples, n_therm, mc_step, est=None):
        R = np.array(R, np.float64)
        if est is None:
            est = np.zeros(self.est_size())
        rows = np.zeros((n_samples, self.P + 2))
        sc = C.c_uint64(step0)
        acc = lib().oracle_inh_sample_walker(C.byref(self.sys), _d(R), _d(np.ascontiguousarray(uR, np.float64)),
                                             _d(np.ascontiguousarray(uI, np.float64)), C.c_double(phiR), C.c_uint64(seed),
                                             C.c_uint32(walker), C.byref(sc), n_init, n_samples, n_therm, C.c_double(mc_step),
                                             _d(est), _d(rows))
        return dict(R=R, est=est, rows=rows, accepted=int(acc), steps=sc.value)

    def unpack_est(self, est, n):
        return Oracle.unpack_est(self, est, n)


class OracleMix:
    """BosonMixtureCluster restatement (oracle/tdvmc_oracle_mix.c) with numpy in/out."""

    def __init__(self, spec, time=0.0):
        self.spec = spec
        e = spec.extra
        self._keep = [np.ascontiguousarray(e["pair_type"], np.int32), np.ascontiguousarray(e["hbar"], np.float64),
                      np.ascontiguousarray(e["mass"], np.float64), np.ascontiguousarray(e["type_knots"], np.float64),
                      np.ascontiguousarray(e["type_weights"], np.float64), np.ascontiguousarray(e["type_mcm"], np.float64),
                      np.ascontiguousarray(e["type_potential"], np.int32), np.ascontiguousarray(spec.map_ptr, np.int32),
                      np.ascontiguousarray(spec.map_col, np.int32), np.ascontiguousarray(spec.map_val, np.float64)]
        k = self._keep
        I = lambda a: a.ctypes.data_as(ip)
        self.sys = OracleMixStruct(spec.n_particles, spec.n_params, e["n_types"], e["n_splines"], spec.n_other,
                                   int(e.get("order", 3)), I(k[0]),
                                   _d(k[1]), _d(k[2]), _d(k[3]), _d(k[4]), _d(k[5]), I(k[6]), I(k[7]), I(k[8]), _d(k[9]))
        self.N, self.P, self.K, self.T = spec.n_particles, spec.n_params, e["n_splines"], e["n_types"]
        self.NE, self.NO = self.T * (self.K + 4), spec.n_other

    def values(self, R):
        R = np.ascontiguousarray(R, np.float64)
        ext = np.zeros(self.NE)
        lib().oracle_mix_values(C.byref(self.sys), _d(R), _d(ext))
        return ext

    def operators(self, ext):
        O = np.zeros(self.P)
        lib().oracle_mix_operators(C.byref(self.sys), _d(np.ascontiguousarray(ext)), _d(O))
        return O

    def exponent(self, ext, uR):
        return lib().oracle_mix_exponent(C.byref(self.sys), _d(np.ascontiguousarray(ext)), _d(np.ascontiguousarray(uR, np.float64)))

    def evaluate(self, R, uR, uI, phiR=0.0):
        R = np.ascontiguousarray(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        uI = np.ascontiguousarray(uI, np.float64)
        ext = self.values(R)
        ex = self.exponent(ext, uR)
        er, ei = C.c_double(0), C.c_double(0)
        other = np.zeros(self.NO)
        dr, di = np.zeros((self.N, 3)), np.zeros((self.N, 3))
        tD, tD2 = np.zeros((self.NE, self.N, 3)), np.zeros((self.NE, self.N))
        lib().oracle_mix_expectation(C.byref(self.sys), _d(R), C.c_double(np.exp(ex + phiR)), C.c_double(ex), _d(uR), _d(uI),
                                     C.byref(er), C.byref(ei), _d(other), _d(dr), _d(di), _d(tD), _d(tD2))
        return dict(ext=ext, O=self.operators(ext), exponent=ex, e_r=er.value, e_i=ei.value, other=other, drift_r=dr,
                    drift_i=di, tabD=tD, tabD2=tD2)

    def quotient(self, R, particle, new_pos, uR):
        R = np.array(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        ext = self.values(R)
        ex = self.exponent(ext, uR)
        old = R[particle].copy()
        R[particle] = new_pos
        ext_new = np.zeros(self.NE)
        en = C.c_double(0)
        q = lib().oracle_mix_quotient(C.byref(self.sys), _d(R), int(particle), _d(old), _d(ext), C.c_double(ex), _d(uR),
                                      _d(ext_new), C.byref(en))
        return q, en.value, ex

    def est_size(self):
        return self.P * self.P + 3 * self.P + 2 + self.NO

    def sample_walker(self, R, uR, uI, phiR, seed, walker, step0, n_init, n_samples, n_therm, mc_step, est=None):
        R = np.array(R, np.float64)
        if est is None:
            est = np.zeros(self.est_size())
        rows = np.zeros((n_samples, self.P + 2))
        sc = C.c_uint64(step0)
        acc = lib().oracle_mix_sample_walker(C.byref(self.sys), _d(R), _d(np.ascontiguousarray(uR, np.float64)),
                                             _d(np.ascontiguousarray(uI, np.float64)), C.c_double(phiR), C.c_uint64(seed),
                                             C.c_uint32(walker), C.byref(sc), n_init, n_samples, n_therm, C.c_double(mc_step),
                                             _d(est), _d(rows))
        return dict(R=R, est=est, rows=rows, accepted=int(acc), steps=sc.value)

    def sweep(self, R, uR, seed, walker, first_step, n_steps, mc_step):
        R = np.array(R, np.float64)
        uR = np.ascontiguousarray(uR, np.float64)
        ext = self.values(R)
        ex = C.c_double(self.exponent(ext, uR))
        acc = lib().oracle_mix_sweep(C.byref(self.sys), _d(R), _d(ext), C.byref(ex), _d(uR), C.c_uint64(seed), C.c_uint32(walker),
                                     C.c_uint64(first_step), C.c_int64(n_steps), C.c_double(mc_step))
        return R, int(acc)

    def unpack_est(self, est, n):
        return Oracle.unpack_est(self, est, n)

    def center_of_mass(self, R):
        com = np.zeros(3)
        lib().oracle_mix_center_of_mass(C.byref(self.sys), _d(np.ascontiguousarray(R, np.float64)), _d(com))
        return com


def oracle_mix_observables(omix, R, grids):
    """oracle_mix_observables for one configuration; grids: dict(angle_grid, density_grid, density_scaling, distance_grid)."""
    R = np.ascontiguousarray(R, np.float64)
    ag, dg, ds, pg = (np.ascontiguousarray(grids[k], np.float64) for k in ("angle_grid", "density_grid", "density_scaling", "distance_grid"))
    r2 = C.c_double(0)
    angle = np.zeros((3, int(ag[0])))
    density = np.zeros((3, int(dg[0])))
    distance = np.zeros((3, int(pg[0])))
    lib().oracle_mix_observables(C.byref(omix.sys), _d(R), _d(ag), _d(dg), _d(ds), _d(pg), C.byref(r2), _d(angle), _d(density),
                                 _d(distance))
    return r2.value, angle, density, distance


OracleHeBulk = OracleHe


def oracle_observables(lbox, R, obs):
    """oracle_observables (oracle/tdvmc_oracle.c) for one configuration; obs: tdvmc_b200.observables.ObservableSpec."""
    R = np.ascontiguousarray(R, np.float64)
    sc = np.ascontiguousarray(obs.gr_scaling, np.float64)
    ptr = np.ascontiguousarray(obs.shell_ptr, np.int32)
    kv = np.ascontiguousarray(obs.kvec, np.float64)
    gr = np.zeros(obs.gr_count)
    sk = np.zeros(obs.n_shells)
    lib().oracle_observables(C.c_double(lbox), C.c_int(R.shape[0]), _d(R), C.c_int(obs.gr_count), C.c_double(obs.gr_spacing),
                             C.c_double(obs.gr_max), C.c_double(obs.gr_weight), _d(sc), C.c_int(obs.n_shells),
                             ptr.ctypes.data_as(ip), _d(kv), _d(gr), _d(sk))
    return gr, sk


def make_oracle(spec, time=0.0):
    from tdvmc_b200 import systems

    if spec.kind == systems.KIND_MIXTURE:
        return OracleMix(spec, time)
    if spec.kind == systems.KIND_BOX_RADIAL:
        return OracleBR(spec, time)
    if spec.kind == systems.KIND_INH_CONTACT:
        return OracleInh(spec, time)
    return Oracle(spec, time) if spec.kind == systems.KIND_SPLINE_TABLE else OracleHe(spec, time)
