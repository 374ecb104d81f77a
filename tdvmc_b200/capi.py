"""ctypes binding of libtdvmc_b200.so (include/tdvmc_gpu.h).

This is the product path: every call lands in the hand-written CUDA library.  There is no CPU
fallback -- if the shared library is missing or no CUDA device is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtdvmc_b200.so")

UNIQUE_ID_BYTES = 128
KERNELS = {"sweep": 0, "evaluate": 1, "accumulate": 2, "tables": 3, "contract": 4, "other": 5, "solve": 6}

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


class MixtureDesc(C.Structure):
    _fields_ = [("n_pair_types", C.c_int32), ("spline_order", C.c_int32), ("pair_type", ip), ("hbar_over_2m", dp), ("mass", dp),
                ("knots", dp), ("spline_weights", dp), ("mcmillan_factor", dp), ("potential", ip)]


class SystemDesc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_particles", C.c_int32), ("dim", C.c_int32), ("n_params", C.c_int32),
                ("n_splines", C.c_int32), ("pair_rule", C.c_int32), ("tail_param", C.c_int32), ("n_other", C.c_int32),
                ("lbox", C.c_double), ("hbar2_2m", C.c_double), ("knots", dp), ("spline_weights", dp), ("map_ptr", ip),
                ("map_col", ip), ("map_val", dp), ("system_params", dp), ("n_system_params", C.c_int32),
                ("system_kind", C.c_int32), ("n_ext", C.c_int32), ("n_splines_first", C.c_int32), ("map_const", dp),
                ("grad_const", dp), ("mixture", C.POINTER(MixtureDesc))]


class EnsembleDesc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("n_walkers", C.c_int32), ("first_walker", C.c_int32),
                ("max_samples_per_walker", C.c_int32), ("keep_sample_positions", C.c_int32), ("seed", C.c_uint64),
                ("mc_step", C.c_double)]


class Estimators(C.Structure):
    _fields_ = [("local_operators", dp), ("local_energy_r", dp), ("local_energy_i", dp), ("local_operators_matrix", dp),
                ("local_operator_energy_r", dp), ("local_operator_energy_i", dp), ("other_expectation_values", dp),
                ("n_acceptances", C.c_int64), ("n_trials", C.c_int64), ("n_samples", C.c_int64)]


class SolverDesc(C.Structure):
    """tdvmc_solver_desc"""
    _fields_ = [("struct_size", C.c_uint32), ("imaginary_time", C.c_int32), ("use_preconditioning", C.c_int32),
                ("force_global_scratch", C.c_int32), ("regularization", C.c_double), ("min_scaling", C.c_double),
                ("solver_type", C.c_int32), ("reserved", C.c_int32)]


class ParametersDot(C.Structure):
    """tdvmc_parameters_dot"""
    _fields_ = [("u_dot_r", dp), ("u_dot_i", dp), ("phi_dot_r", C.c_double), ("phi_dot_i", C.c_double),
                ("local_energy_r", C.c_double), ("local_energy_i", C.c_double), ("not_positive_definite", C.c_int32)]


class ObservableDesc(C.Structure):
    """tdvmc_observable_desc"""
    _fields_ = [("gr_count", C.c_int32), ("n_shells", C.c_int32), ("gr_spacing", C.c_double), ("gr_max", C.c_double),
                ("gr_weight", C.c_double), ("gr_scaling", dp), ("shell_ptr", ip), ("kvec", dp)]


class ClusterObservableDesc(C.Structure):
    """tdvmc_cluster_observable_desc"""
    _fields_ = [("n_angle", C.c_int32), ("n_density", C.c_int32), ("n_distance", C.c_int32), ("reserved", C.c_int32),
                ("angle_spacing", C.c_double), ("density_spacing", C.c_double), ("density_max", C.c_double),
                ("distance_spacing", C.c_double), ("distance_max", C.c_double), ("density_scaling", dp)]


# every symbol include/tdvmc_gpu.h declares: (name, restype, argtypes)
_VP = C.c_void_p
SYMBOLS = [
    ("tdvmc_gpu_abi_version", C.c_int, []),
    ("tdvmc_gpu_device_count", C.c_int, []),
    ("tdvmc_gpu_create", C.c_int, [C.POINTER(SystemDesc), C.POINTER(EnsembleDesc), C.POINTER(_VP)]),
    ("tdvmc_gpu_destroy", None, [_VP]),
    ("tdvmc_gpu_last_error", C.c_char_p, [_VP]),
    ("tdvmc_gpu_set_positions", C.c_int, [_VP, dp, C.c_int32, C.c_int32]),
    ("tdvmc_gpu_get_positions", C.c_int, [_VP, dp, C.c_int32, C.c_int32]),
    ("tdvmc_gpu_set_params", C.c_int, [_VP, dp, dp, C.c_double, C.c_double, C.c_double]),
    ("tdvmc_gpu_wrap_positions", C.c_int, [_VP]),
    ("tdvmc_gpu_reset_counters", C.c_int, [_VP]),
    ("tdvmc_gpu_set_mc_step", C.c_int, [_VP, C.c_double]),
    ("tdvmc_gpu_sweep", C.c_int, [_VP, C.c_int64]),
    ("tdvmc_gpu_sample_and_accumulate", C.c_int, [_VP, C.c_int32, C.c_int32, C.c_int32]),
    ("tdvmc_gpu_reevaluate_stored", C.c_int, [_VP]),
    ("tdvmc_gpu_update_stored", C.c_int, [_VP, C.c_int32, C.c_int32]),
    ("tdvmc_gpu_allreduce_and_fetch", C.c_int, [_VP, C.POINTER(Estimators)]),
    ("tdvmc_gpu_solve_parameters_dot", C.c_int, [_VP, C.POINTER(SolverDesc), C.POINTER(ParametersDot)]),
    ("tdvmc_gpu_euler_step", C.c_int, [_VP, C.POINTER(SolverDesc), C.c_double, C.c_double, dp, dp, dp, dp, C.POINTER(ParametersDot)]),
    ("tdvmc_gpu_solve_fixed", C.c_int, [_VP, C.POINTER(SolverDesc), C.POINTER(Estimators), C.POINTER(ParametersDot)]),
    ("tdvmc_gpu_last_exponent", C.c_int, [_VP, dp]),
    ("tdvmc_gpu_comm_unique_id", C.c_int, [C.POINTER(C.c_uint8)]),
    ("tdvmc_gpu_comm_init", C.c_int, [_VP, C.POINTER(C.c_uint8), C.c_int32, C.c_int32]),
    ("tdvmc_gpu_evaluate_fixed", C.c_int, [_VP, dp, C.c_int32] + [dp] * 9),
    ("tdvmc_gpu_quotient_fixed", C.c_int, [_VP, dp, dp, C.c_int32, dp, dp]),
    ("tdvmc_gpu_tables_fixed", C.c_int, [_VP, dp, dp, dp]),
    ("tdvmc_gpu_min_image", C.c_int, [_VP, C.c_double, dp, dp, C.c_int32, dp, dp]),
    ("tdvmc_gpu_accumulate_fixed", C.c_int, [_VP, dp, dp, dp, C.c_int64, dp, dp, dp, dp]),
    ("tdvmc_gpu_proposals", C.c_int, [_VP, C.c_int32, C.c_int64, C.c_int32, ip, dp, dp]),
    ("tdvmc_gpu_observables_fixed", C.c_int, [_VP, C.POINTER(ObservableDesc), dp, C.c_int32, dp, dp]),
    ("tdvmc_gpu_sample_observables", C.c_int, [_VP, C.POINTER(ObservableDesc), C.c_int32, C.c_int32, C.c_int32, dp, dp]),
    ("tdvmc_gpu_cluster_observables_fixed", C.c_int, [_VP, C.POINTER(ClusterObservableDesc), dp, C.c_int32, dp, dp, dp, dp]),
    ("tdvmc_gpu_sample_cluster_observables", C.c_int,
     [_VP, C.POINTER(ClusterObservableDesc), C.c_int32, C.c_int32, C.c_int32, dp, dp, dp, dp]),
    ("tdvmc_gpu_profile", C.c_int, [_VP, C.c_int32, C.c_int32]),
    ("tdvmc_gpu_kernel_stats", C.c_int, [_VP, C.c_int32, C.POINTER(C.c_int64), dp]),
    ("tdvmc_gpu_synchronize", C.c_int, [_VP]),
    ("tdvmc_gpu_timer_start", C.c_int, [_VP]),
    ("tdvmc_gpu_timer_stop", C.c_int, [_VP, dp]),
    ("tdvmc_gpu_launch_count", C.c_int, [_VP, C.POINTER(C.c_int64)]),
    ("tdvmc_gpu_flush_l2", C.c_int, [_VP, C.c_int64]),
    ("tdvmc_gpu_tables_resident", C.c_int, [_VP, C.c_int32]),
    ("tdvmc_gpu_contract_resident", C.c_int, [_VP, C.c_int32, dp, dp]),
    ("tdvmc_gpu_resident_walkers", C.c_int, [_VP, ip, ip]),
    ("tdvmc_gpu_measure_fp64_peak", C.c_int, [_VP, dp, dp]),
]

_lib = None


def load():
    """Load the CUDA library; raises if it was not built (run ``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `make -C tdvmc_b200/csrc` (no CPU fallback exists)")
        lib = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _d(a):
    return a.ctypes.data_as(dp) if a is not None else None


class TdvmcError(RuntimeError):
    pass


class Handle:
    """Thin object wrapper over the opaque ``tdvmc_gpu_handle``."""

    def __init__(self, spec, n_walkers, seed=1, mc_step=0.5, first_walker=0, max_samples=1, keep_sample_positions=False,
                 device=0):
        lib = load()
        self.lib = lib
        self.spec = spec
        self.N, self.P, self.K, self.n_other = spec.n_particles, spec.n_params, spec.n_splines, spec.n_other
        self.n_ext = spec.n_ext
        self.W = int(n_walkers)
        self._keep = dict(knots=np.ascontiguousarray(spec.knots, np.float64),
                          w=np.ascontiguousarray(spec.weights, np.float64).reshape(-1),
                          mp=np.ascontiguousarray(spec.map_ptr, np.int32), mc=np.ascontiguousarray(spec.map_col, np.int32),
                          mv=np.ascontiguousarray(spec.map_val, np.float64),
                          sp=np.ascontiguousarray(spec.system_params, np.float64),
                          mk=np.ascontiguousarray(spec.map_const, np.float64), gk=np.ascontiguousarray(spec.grad_const, np.float64))
        k = self._keep
        mix = None
        if spec.kind == 3:
            e = spec.extra
            k.update(pt=np.ascontiguousarray(e["pair_type"], np.int32), hb=np.ascontiguousarray(e["hbar"], np.float64),
                     ms=np.ascontiguousarray(e["mass"], np.float64), tk=np.ascontiguousarray(e["type_knots"], np.float64),
                     tw=np.ascontiguousarray(e["type_weights"], np.float64), tm=np.ascontiguousarray(e["type_mcm"], np.float64),
                     tp=np.ascontiguousarray(e["type_potential"], np.int32))
            self._mix = MixtureDesc(e["n_types"], int(e.get("order", 3)), k["pt"].ctypes.data_as(ip), _d(k["hb"]), _d(k["ms"]), _d(k["tk"]),
                                    _d(k["tw"]), _d(k["tm"]), k["tp"].ctypes.data_as(ip))
            mix = C.pointer(self._mix)
        sd = SystemDesc(C.sizeof(SystemDesc), spec.n_particles, spec.dim, spec.n_params, spec.n_splines, spec.pair_rule,
                        spec.tail_param, spec.n_other, spec.lbox, spec.hbar2_2m, _d(k["knots"]), _d(k["w"]),
                        k["mp"].ctypes.data_as(ip), k["mc"].ctypes.data_as(ip), _d(k["mv"]), _d(k["sp"]), len(k["sp"]),
                        spec.kind, spec.n_ext, int(spec.extra.get("n_splines_spf", 0)), _d(k["mk"]), _d(k["gk"]), mix)
        ed = EnsembleDesc(C.sizeof(EnsembleDesc), device, self.W, first_walker, max_samples, int(keep_sample_positions),
                          seed, mc_step)
        h = _VP()
        rc = lib.tdvmc_gpu_create(C.byref(sd), C.byref(ed), C.byref(h))
        if rc != 0:
            raise TdvmcError(f"tdvmc_gpu_create failed ({rc}): {lib.tdvmc_gpu_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.tdvmc_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise TdvmcError(f"{what} failed ({rc}): {self.lib.tdvmc_gpu_last_error(self.h).decode()}")

    # ---- state ----
    def set_positions(self, R, first=0):
        R = np.ascontiguousarray(R, np.float64).reshape(-1, self.N, 3)
        self._ck(self.lib.tdvmc_gpu_set_positions(self.h, _d(R), first, R.shape[0]), "set_positions")

    def set_positions_raw(self, ptr, first, n):
        """ptr: address of a (pinned) host buffer [n][N][3]."""
        self._ck(self.lib.tdvmc_gpu_set_positions(self.h, C.cast(ptr, dp), first, n), "set_positions")

    def get_positions(self, first=0, n=None, out=None):
        n = self.W - first if n is None else n
        R = np.empty((n, self.N, 3)) if out is None else out
        self._ck(self.lib.tdvmc_gpu_get_positions(self.h, _d(R), first, n), "get_positions")
        return R

    def set_params(self, uR, uI, phiR=0.0, phiI=0.0, time=0.0):
        uR = np.ascontiguousarray(uR, np.float64)
        uI = np.ascontiguousarray(uI, np.float64)
        assert uR.size == self.P and uI.size == self.P
        self._ck(self.lib.tdvmc_gpu_set_params(self.h, _d(uR), _d(uI), phiR, phiI, time), "set_params")

    def wrap_positions(self):
        self._ck(self.lib.tdvmc_gpu_wrap_positions(self.h), "wrap_positions")

    def reset_counters(self):
        """nAcceptances = nTrials = 0 at the start of a time step (src/TDVMC.cpp:3428-3429)."""
        self._ck(self.lib.tdvmc_gpu_reset_counters(self.h), "reset_counters")

    def set_mc_step(self, mc_step):
        self._ck(self.lib.tdvmc_gpu_set_mc_step(self.h, float(mc_step)), "set_mc_step")

    # ---- sampling ----
    def sweep(self, n_steps):
        self._ck(self.lib.tdvmc_gpu_sweep(self.h, int(n_steps)), "sweep")

    def sample_and_accumulate(self, n_samples, n_therm, n_init=0):
        self._ck(self.lib.tdvmc_gpu_sample_and_accumulate(self.h, n_samples, n_therm, n_init), "sample_and_accumulate")

    def reevaluate_stored(self):
        self._ck(self.lib.tdvmc_gpu_reevaluate_stored(self.h), "reevaluate_stored")

    def update_stored(self, n_update, n_therm):
        self._ck(self.lib.tdvmc_gpu_update_stored(self.h, int(n_update), int(n_therm)), "update_stored")

    def allreduce_and_fetch(self, out=None):
        P = self.P
        if out is None:
            out = dict(O=np.empty(P), e_r=np.empty(1), e_i=np.empty(1), S=np.empty((P, P)), OER=np.empty(P), OEI=np.empty(P),
                       other=np.empty(self.n_other))
        est = Estimators(_d(out["O"]), _d(out["e_r"]), _d(out["e_i"]), _d(out["S"]), _d(out["OER"]), _d(out["OEI"]),
                         _d(out["other"]), 0, 0, 0)
        self._ck(self.lib.tdvmc_gpu_allreduce_and_fetch(self.h, C.byref(est)), "allreduce_and_fetch")
        out["n_acceptances"], out["n_trials"], out["n_samples"] = est.n_acceptances, est.n_trials, est.n_samples
        return out

    # ---- parameter derivatives / Euler step on the device ----
    @staticmethod
    def _solver_desc(imaginary_time=1, use_preconditioning=True, regularization=None, min_scaling=0.0, force_global=False,
                     solver_type=0):
        if regularization is None:       # the reference's hard-coded values (src/TDVMC.cpp:1737, :1770)
            regularization = 0.002 if solver_type == 1 else 0.001
        return SolverDesc(C.sizeof(SolverDesc), int(imaginary_time), int(bool(use_preconditioning)), int(bool(force_global)),
                          float(regularization), float(min_scaling), int(solver_type), 0)

    def _dot_out(self):
        o = dict(u_dot_r=np.empty(self.P), u_dot_i=np.empty(self.P))
        return o, ParametersDot(_d(o["u_dot_r"]), _d(o["u_dot_i"]), 0.0, 0.0, 0.0, 0.0, 0)

    @staticmethod
    def _dot_fill(o, pd):
        o.update(phi_dot_r=pd.phi_dot_r, phi_dot_i=pd.phi_dot_i, e_r=pd.local_energy_r, e_i=pd.local_energy_i,
                 not_positive_definite=bool(pd.not_positive_definite))
        return o

    def solve_parameters_dot(self, **kw):
        """SolveForParametersDot (Cholesky branch) on the device-resident estimators of the last accumulation."""
        sd = self._solver_desc(**kw)
        o, pd = self._dot_out()
        self._ck(self.lib.tdvmc_gpu_solve_parameters_dot(self.h, C.byref(sd), C.byref(pd)), "solve_parameters_dot")
        return self._dot_fill(o, pd)

    def solve_fixed(self, est, **kw):
        """The same solve on caller-given averages: est has O, S, OER, OEI, e_r, e_i (as allreduce_and_fetch returns)."""
        sd = self._solver_desc(**kw)
        o, pd = self._dot_out()
        k = {n: np.ascontiguousarray(np.atleast_1d(est[n]), np.float64) for n in ("O", "S", "OER", "OEI", "e_r", "e_i")}
        e = Estimators(_d(k["O"]), _d(k["e_r"]), _d(k["e_i"]), _d(k["S"]), _d(k["OER"]), _d(k["OEI"]), None, 0, 0, 0)
        self._ck(self.lib.tdvmc_gpu_solve_fixed(self.h, C.byref(sd), C.byref(e), C.byref(pd)), "solve_fixed")
        return self._dot_fill(o, pd)

    def euler_step(self, dt, uR, uI, phiR, phiI, time=0.0, **kw):
        """CalculateNextParametersEuler + parameter feedback; returns (uR, uI, phiR, phiI, dot)."""
        sd = self._solver_desc(**kw)
        o, pd = self._dot_out()
        uR = np.array(uR, np.float64)
        uI = np.array(uI, np.float64)
        pr, pi = C.c_double(phiR), C.c_double(phiI)
        self._ck(self.lib.tdvmc_gpu_euler_step(self.h, C.byref(sd), dt, time, _d(uR), _d(uI), C.byref(pr), C.byref(pi),
                                               C.byref(pd)), "euler_step")
        return uR, uI, pr.value, pi.value, self._dot_fill(o, pd)

    def last_exponent(self):
        x = C.c_double(0)
        self._ck(self.lib.tdvmc_gpu_last_exponent(self.h, C.byref(x)), "last_exponent")
        return x.value

    def comm_init(self, unique_id, rank, n_ranks):
        buf = (C.c_uint8 * UNIQUE_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.tdvmc_gpu_comm_init(self.h, buf, rank, n_ranks), "comm_init")

    # ---- fixed-configuration entry points ----
    def evaluate_fixed(self, R):
        R = np.ascontiguousarray(R, np.float64).reshape(-1, self.N, 3)
        n = R.shape[0]
        o = dict(e_r=np.empty(n), e_i=np.empty(n), O=np.empty((n, self.P)), other=np.empty((n, self.n_other)),
                 exponent=np.empty(n), drift_r=np.empty((n, self.N, 3)), drift_i=np.empty((n, self.N, 3)),
                 ss=np.empty((n, self.n_ext)), outer=np.empty(n))
        self._ck(self.lib.tdvmc_gpu_evaluate_fixed(self.h, _d(R), n, _d(o["e_r"]), _d(o["e_i"]), _d(o["O"]), _d(o["other"]),
                                                   _d(o["exponent"]), _d(o["drift_r"]), _d(o["drift_i"]), _d(o["ss"]),
                                                   _d(o["outer"])), "evaluate_fixed")
        return o

    def quotient_fixed(self, R, moves):
        R = np.ascontiguousarray(R, np.float64).reshape(self.N, 3)
        moves = np.ascontiguousarray(moves, np.float64).reshape(-1, 4)
        q = np.empty(len(moves))
        d = np.empty(len(moves))
        self._ck(self.lib.tdvmc_gpu_quotient_fixed(self.h, _d(R), _d(moves), len(moves), _d(q), _d(d)), "quotient_fixed")
        return q, d

    @staticmethod
    def _observable_desc(obs):
        keep = [np.ascontiguousarray(obs.gr_scaling, np.float64), np.ascontiguousarray(obs.shell_ptr, np.int32),
                np.ascontiguousarray(obs.kvec, np.float64)]
        od = ObservableDesc(int(obs.gr_count), int(obs.n_shells), float(obs.gr_spacing), float(obs.gr_max), float(obs.gr_weight),
                            _d(keep[0]), keep[1].ctypes.data_as(ip), _d(keep[2]))
        return od, keep

    def observables_fixed(self, obs, R):
        """g(r) and S(k) of given configurations; obs: tdvmc_b200.observables.ObservableSpec."""
        R = np.ascontiguousarray(R, np.float64).reshape(-1, self.N, 3)
        od, keep = self._observable_desc(obs)
        gr = np.empty((len(R), obs.gr_count))
        sk = np.empty((len(R), obs.n_shells))
        self._ck(self.lib.tdvmc_gpu_observables_fixed(self.h, C.byref(od), _d(R), len(R), _d(gr), _d(sk)), "observables_fixed")
        return gr, sk

    def sample_observables(self, obs, n_samples, n_therm, n_init):
        """The reference's end-of-run observable pass over the resident walkers: mean g(r), S(k)."""
        od, keep = self._observable_desc(obs)
        gr = np.empty(obs.gr_count)
        sk = np.empty(obs.n_shells)
        self._ck(self.lib.tdvmc_gpu_sample_observables(self.h, C.byref(od), int(n_samples), int(n_therm), int(n_init), _d(gr), _d(sk)),
                 "sample_observables")
        return gr, sk

    @staticmethod
    def _cluster_desc(grids):
        """grids: dict(angle_grid, density_grid, distance_grid = (count, spacing, max), density_scaling)."""
        ag, dg, pg = (np.asarray(grids[k], np.float64) for k in ("angle_grid", "density_grid", "distance_grid"))
        sc = np.ascontiguousarray(grids["density_scaling"], np.float64)
        od = ClusterObservableDesc(int(ag[0]), int(dg[0]), int(pg[0]), 0, float(ag[1]), float(dg[1]), float(dg[2]), float(pg[1]),
                                   float(pg[2]), _d(sc))
        return od, sc

    def cluster_observables_fixed(self, grids, R):
        R = np.ascontiguousarray(R, np.float64).reshape(-1, 3, 3)
        od, keep = self._cluster_desc(grids)
        n = len(R)
        r2 = np.empty(n)
        angle = np.empty((n, 3, od.n_angle))
        density = np.empty((n, 3, od.n_density))
        distance = np.empty((n, 3, od.n_distance))
        self._ck(self.lib.tdvmc_gpu_cluster_observables_fixed(self.h, C.byref(od), _d(R), n, _d(r2), _d(angle), _d(density),
                                                              _d(distance)), "cluster_observables_fixed")
        return r2, angle, density, distance

    def sample_cluster_observables(self, grids, n_samples, n_therm, n_init):
        od, keep = self._cluster_desc(grids)
        r2 = np.empty(1)
        angle = np.empty((3, od.n_angle))
        density = np.empty((3, od.n_density))
        distance = np.empty((3, od.n_distance))
        self._ck(self.lib.tdvmc_gpu_sample_cluster_observables(self.h, C.byref(od), int(n_samples), int(n_therm), int(n_init),
                                                               _d(r2), _d(angle), _d(density), _d(distance)),
                 "sample_cluster_observables")
        return float(r2[0]), angle, density, distance

    def tables_fixed(self, R):
        R = np.ascontiguousarray(R, np.float64).reshape(self.N, 3)
        sD = np.empty((self.K, self.N, 3))
        sD2 = np.empty((self.K, self.N))
        self._ck(self.lib.tdvmc_gpu_tables_fixed(self.h, _d(R), _d(sD), _d(sD2)), "tables_fixed")
        return sD, sD2

    def min_image(self, L, a, b):
        a = np.ascontiguousarray(a, np.float64).reshape(-1, 3)
        b = np.ascontiguousarray(b, np.float64).reshape(-1, 3)
        norm = np.empty(len(a))
        disp = np.empty((len(a), 3))
        self._ck(self.lib.tdvmc_gpu_min_image(self.h, float(L), _d(a), _d(b), len(a), _d(norm), _d(disp)), "min_image")
        return norm, disp

    def accumulate_fixed(self, O, e_r, e_i):
        O = np.ascontiguousarray(O, np.float64)
        M, P = O.shape
        assert P == self.P
        e_r = np.ascontiguousarray(e_r, np.float64)
        e_i = np.ascontiguousarray(e_i, np.float64)
        S, fr, fi, o = np.empty((P, P)), np.empty(P), np.empty(P), np.empty(P)
        self._ck(self.lib.tdvmc_gpu_accumulate_fixed(self.h, _d(O), _d(e_r), _d(e_i), M, _d(S), _d(fr), _d(fi), _d(o)),
                 "accumulate_fixed")
        return S, fr, fi, o

    def proposals(self, global_walker, first_step, n):
        p = np.empty(n, np.int32)
        d = np.empty((n, 3))
        lu = np.empty(n)
        self._ck(self.lib.tdvmc_gpu_proposals(self.h, global_walker, first_step, n, p.ctypes.data_as(ip), _d(d), _d(lu)),
                 "proposals")
        return p, d, lu

    # ---- measurement ----
    def profile(self, enable=True, reset=True):
        self._ck(self.lib.tdvmc_gpu_profile(self.h, int(enable), int(reset)), "profile")

    def kernel_stats(self):
        out = {}
        for name, kid in KERNELS.items():
            n = C.c_int64(0)
            ms = C.c_double(0)
            self._ck(self.lib.tdvmc_gpu_kernel_stats(self.h, kid, C.byref(n), C.byref(ms)), "kernel_stats")
            out[name] = (n.value, ms.value)
        return out

    def synchronize(self):
        self._ck(self.lib.tdvmc_gpu_synchronize(self.h), "synchronize")

    def timer_start(self):
        self._ck(self.lib.tdvmc_gpu_timer_start(self.h), "timer_start")

    def timer_stop(self):
        ms = C.c_double(0)
        self._ck(self.lib.tdvmc_gpu_timer_stop(self.h, C.byref(ms)), "timer_stop")
        return ms.value

    def launch_count(self):
        n = C.c_int64(0)
        self._ck(self.lib.tdvmc_gpu_launch_count(self.h, C.byref(n)), "launch_count")
        return n.value

    def flush_l2(self, n_bytes=256 << 20):
        self._ck(self.lib.tdvmc_gpu_flush_l2(self.h, n_bytes), "flush_l2")

    def tables_resident(self, n_walkers):
        self._ck(self.lib.tdvmc_gpu_tables_resident(self.h, n_walkers), "tables_resident")

    def contract_resident(self, n_walkers, fetch=True):
        e_r = np.empty(n_walkers) if fetch else None
        e_i = np.empty(n_walkers) if fetch else None
        self._ck(self.lib.tdvmc_gpu_contract_resident(self.h, n_walkers, _d(e_r), _d(e_i)), "contract_resident")
        return e_r, e_i

    def resident_walkers(self):
        a, b = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.tdvmc_gpu_resident_walkers(self.h, C.byref(a), C.byref(b)), "resident_walkers")
        return a.value, b.value

    def measure_fp64_peak(self):
        a, b = C.c_double(0), C.c_double(0)
        self._ck(self.lib.tdvmc_gpu_measure_fp64_peak(self.h, C.byref(a), C.byref(b)), "measure_fp64_peak")
        return a.value, b.value


def comm_unique_id():
    lib = load()
    buf = (C.c_uint8 * UNIQUE_ID_BYTES)()
    rc = lib.tdvmc_gpu_comm_unique_id(buf)
    if rc != 0:
        raise TdvmcError(f"comm_unique_id failed: {lib.tdvmc_gpu_last_error(None).decode()}")
    return bytes(buf)
