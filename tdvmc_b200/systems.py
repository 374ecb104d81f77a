"""Host-side description of a physical system for the walker hot path.

Mirrors what the reference's ``IPhysicalSystem`` plugins set up in ``InitSystem()`` -- knots,
spline table, boundary-condition map, cut-off rule, pair potential -- as plain data that is
handed to the CUDA library through ``tdvmc_system_desc`` (include/tdvmc_gpu.h).  Systems are data,
not kernels: the device code is driven by the flags in this description.

The five systems of the BASELINE configs:
  * ``BosonsBulk``           src/PhysicalSystems/BosonsBulk.cpp:49-156          (config 3, headline)
  * ``NUBosonsBulkPB``       src/PhysicalSystems/NUBosonsBulkPB.cpp:53-216      (config 4)
  * ``HeBulk``               src/PhysicalSystems/HeBulk.cpp:40-70, 376-383      (config 2)
  * ``HeDrop``               src/PhysicalSystems/HeDrop.cpp:71-135, 609-626     (config 1)
  * ``BosonMixtureCluster``  src/PhysicalSystems/BosonMixtureCluster.cpp:58-346 (config 5)
and, widening per SURVEY.md section 8(f) rank 4:
  * ``NUBosonsBulkPBBoxAndRadial``  src/PhysicalSystems/NUBosonsBulkPBBoxAndRadial.cpp:36-191 (the "radial+box splines" system)
  * ``BosonMixtureCluster_4thorder`` src/PhysicalSystems/BosonMixtureCluster_4thorder.cpp (quartic splines)
  * ``InhContactBosons``            src/PhysicalSystems/InhContactBosons.cpp:64-247 (1-D, one-body + pair splines)
"""
from dataclasses import dataclass, field

import numpy as np

from . import splines

PAIR_RULE_CUT = 0      # r <= r_max -> spline, else tail count        (BosonsBulk.cpp:195-210)
PAIR_RULE_REFLECT = 1  # r -> 2 r_max - r beyond r_max, then r < r_max (NUBosonsBulkPB.cpp:249-269)

KIND_SPLINE_TABLE = 0   # BosonsBulk, NUBosonsBulkPB: monomial spline table + boundary map
KIND_HE_BULK = 1        # HeBulk: McMillan core + uniform B-splines in the local coordinate + Aziz potential
KIND_HE_DROP = 2        # HeDrop: open boundary, McMillan core, two uniform grids, const + linear tails, LJ potential
KIND_MIXTURE = 3        # BosonMixtureCluster: species, per-pair-type spline tables + McMillan/const/linear/log, pair potentials
KIND_INH_CONTACT = 5    # InhContactBosons: ONE-dimensional, single-particle spline function + pair-correlation splines
KIND_BOX_RADIAL = 4     # NUBosonsBulkPBBoxAndRadial: radial splines in r_ij + "box" splines in |x_ij|, |y_ij|, |z_ij|, Gauss potential

POT_HFDB_HE_HE, POT_KTTY_HE_NA, POT_KTTY_HE_CS = 0, 1, 2
# BosonMixtureCluster.h:28-36
SPECIES = {"He3": 0, "He4": 1, "Na": 2, "Li": 3, "Cs": 4}
SPECIES_MASS = {0: 3.0160293191, 1: 4.00260325415, 2: 22.9897692809, 4: 132.905451932}   # BosonMixtureCluster.cpp:108-135

HBAR2_2M = 1.0  # src/Constants.h:12


@dataclass
class SystemSpec:
    name: str
    n_particles: int
    n_params: int
    lbox: float
    knots: np.ndarray              # [K+4]
    weights: np.ndarray            # [K][4][4], SplineFactory::GetWeights3 layout
    map_ptr: np.ndarray            # [P+1] CSR rows: O_p = sum_j map_val[j] * ss[map_col[j]]
    map_col: np.ndarray
    map_val: np.ndarray
    pair_rule: int
    system_params: np.ndarray      # SYSTEM_PARAMS as in the config
    n_other: int = 9               # length of otherExpectationValues
    dim: int = 3
    hbar2_2m: float = HBAR2_2M
    tail_param: int = -1
    kind: int = KIND_SPLINE_TABLE
    n_ext: int = 0                  # columns of the map: K spline sums (+ analytic extras, e.g. the McMillan sum)
    map_const: np.ndarray = None    # [P] constant added to O_p (HeBulk.cpp:383: the literal 1.0 of the last operator)
    grad_const: np.ndarray = None   # [P] constant added to every gradient component of parameter p (HeBulk.cpp:351)
    extra: dict = field(default_factory=dict)

    def __post_init__(self):
        if not self.n_ext:
            self.n_ext = len(self.knots) - 4
        if self.map_const is None:
            self.map_const = np.zeros(self.n_params)
        if self.grad_const is None:
            self.grad_const = np.zeros(self.n_params)

    @property
    def n_splines(self):
        return int(self.extra["n_splines"]) if "n_splines" in self.extra else len(self.knots) - 4

    @property
    def r_max(self):
        return float(self.extra["r_max"]) if "r_max" in self.extra else float(self.knots[len(self.knots) - 4])

    def potential(self, time):
        """Square-well (a, b) with the time switch of BosonsBulk.cpp:237-243 / NUBosonsBulkPB.cpp:300-307."""
        p = self.system_params
        a, b = float(p[0]), float(p[1])
        if len(p) > 2 and time >= p[2]:
            a, b = float(p[3]), float(p[4])
        return a, b

    def map_rows(self):
        return [[(int(self.map_col[j]), float(self.map_val[j])) for j in range(self.map_ptr[p], self.map_ptr[p + 1])]
                for p in range(self.n_params)]

    def spline_space(self, u):
        """u~_k = sum_p u_p M[p][k]: parameters pushed through the transposed boundary map."""
        ut = np.zeros(self.n_ext, dtype=np.float64)
        for p, row in enumerate(self.map_rows()):
            for k, f in row:
                ut[k] += u[p] * f
        return ut


def _csr(rows):
    ptr = np.zeros(len(rows) + 1, dtype=np.int32)
    col, val = [], []
    for p, row in enumerate(rows):
        for k, f in row:
            col.append(k)
            val.append(f)
        ptr[p + 1] = len(col)
    return ptr, np.array(col, dtype=np.int32), np.array(val, dtype=np.float64)


def bosons_bulk(n_particles, lbox, n_params, system_params=(1.0, 1.0), nurbs_grid=None, weights=None, dim=3):
    """``BosonsBulk`` (BosonsBulk.cpp:49-156).

    Knots: uniform ``(i L/2)/(P-1)``, i=-3..P+2 (:61-67) or a mirrored NURBS grid (:35-44).
    Boundary map: ``SetBoundaryConditions3_1D_OR_2`` at the origin and ``..._CO_2`` at the cut
    (SplineFactory.cpp:423-456, 563-596; both are called with ``uniform=true``, BosonsBulk.cpp:75-76),
    applied as in ``RefreshLocalOperators`` (:158-177).  ``P`` must equal ``K - 2`` (:85-91).
    """
    knots = splines.uniform_knots(n_params, lbox / 2.0) if nurbs_grid is None else splines.extend_knots_mirrored(nurbs_grid)
    K = len(knots) - 4
    if n_params != K - 2:
        raise ValueError(f"BosonsBulk needs N_PARAM = K - 2 = {K - 2}, got {n_params}")
    if weights is None:
        weights = splines.bspline_monomial_weights(knots)
    rows = [[(1, 1.0)], [(0, 1.0), (2, 1.0)]]
    rows += [[(3 + i, 1.0)] for i in range(n_params - 4)]
    rows += [[(K - 3, 1.0), (K - 1, 1.0)], [(K - 2, 1.0)]]
    ptr, col, val = _csr(rows)
    return SystemSpec("BosonsBulk", n_particles, n_params, float(lbox), knots, np.ascontiguousarray(weights), ptr, col, val,
                      PAIR_RULE_CUT, np.asarray(system_params, dtype=np.float64), n_other=9, tail_param=n_params - 1, dim=int(dim))


def nu_bosons_bulk_pb(n_particles, lbox, n_params, nurbs_grid, system_params=(0.0, 0.0, 0.0, 0.1, 50.0),
                      gr_bin_count=400, weights=None, dim=3):
    """``NUBosonsBulkPB`` (NUBosonsBulkPB.cpp:53-216): non-uniform knots, periodic-box reflection.

    ``K = P + 3`` (:71); map ``O_i = ss[i+1]``, ``O_1 += ss[0]``, ``O_{P-1} += ss[K-2] + ss[K-1]`` (:219-232);
    ``otherExpectationValues`` has ``9 + GR_BIN_COUNT`` entries of which only the first nine are filled (:61, :578-581).
    """
    knots = splines.extend_knots_mirrored(nurbs_grid)
    K = len(knots) - 4
    if K != n_params + 3:
        raise ValueError(f"NUBosonsBulkPB needs K = N_PARAM + 3, got K={K}, N_PARAM={n_params}")
    if weights is None:
        weights = splines.bspline_monomial_weights(knots)
    rows = [[(i + 1, 1.0)] for i in range(n_params)]
    rows[1].append((0, 1.0))
    rows[n_params - 1] += [(K - 2, 1.0), (K - 1, 1.0)]
    ptr, col, val = _csr(rows)
    return SystemSpec("NUBosonsBulkPB", n_particles, n_params, float(lbox), knots, np.ascontiguousarray(weights), ptr, col, val,
                      PAIR_RULE_REFLECT, np.asarray(system_params, dtype=np.float64), n_other=9 + int(gr_bin_count),
                      tail_param=n_params - 1, dim=int(dim))


def nu_bosons_bulk_pb_box_and_radial(n_particles, lbox, n_params, nurbs_grid, system_params=(0.1, 50.0), gr_bin_count=400,
                                     weights=None, dim=3):
    """``NUBosonsBulkPBBoxAndRadial`` (NUBosonsBulkPBBoxAndRadial.cpp:36-191).

    ``SetNodes`` (:36-62) mirrors the same NURBS grid into ``nodes`` (box splines, argument ``|x_ij|`` per coordinate) and
    ``nodesRad`` (radial splines, argument ``r_ij < maxDistanceRad``), so one knot vector and one spline table serve both
    bases (without ``USE_NURBS`` the reference reads the empty ``nodesRad``, :96 - the grid is mandatory).  ``N_PARAM/2``
    parameters each (:84-85), ``K = N_PARAM/2 + 3`` splines each (:88-89).  Extended sums ``[ssRad_0..K-1 | ss_0..K-1]``;
    map of ``RefreshLocalOperators`` (:193-211): ``O_i = ssRad[i+1]``, ``O_1 += ssRad[0]``,
    ``O_{PR-1} += ssRad[K-2]/(-2) + ssRad[K-1]``; ``O_{PR+i} = ss[i+1]``, ``O_{PR+1} += ss[0]``, ``O_{P-1} += ss[K-2] + ss[K-1]``.
    One deviation the reference has and this reproduces: the drift contracts the LAST radial spline's parameter with the
    box table ``sD[K-1]`` instead of ``sDRad[K-1]`` (:493-497), the Laplacian does not (:498-499) -- ``extra["grad_swap"]``.
    ``otherExpectationValues`` = kinetic, potential, wf, then ``GR_BIN_COUNT`` g(r) bins weighted by 1/shell volume (:574-580).
    Gauss pair potential ``b exp(-(r/a)^2/2)`` inside ``maxDistanceRad`` with the time switch of :290-297."""
    if n_params % 2:
        raise ValueError("NUBosonsBulkPBBoxAndRadial splits N_PARAM evenly between the radial and the box basis")
    PR = n_params // 2
    knots = splines.extend_knots_mirrored(nurbs_grid)
    K = len(knots) - 4
    if K != PR + 3:
        raise ValueError(f"NUBosonsBulkPBBoxAndRadial needs K = N_PARAM/2 + 3, got K={K}, N_PARAM={n_params}")
    if weights is None:
        weights = splines.bspline_monomial_weights(knots)
    rows = [[(i + 1, 1.0)] for i in range(PR)]
    rows[1].append((0, 1.0))
    rows[PR - 1] += [(K - 2, 1.0 / (-2.0)), (K - 1, 1.0)]
    rows += [[(K + i + 1, 1.0)] for i in range(PR)]
    rows[PR + 1].append((K, 1.0))
    rows[n_params - 1] += [(K + K - 2, 1.0), (K + K - 1, 1.0)]
    ptr, col, val = _csr(rows)
    half = lbox / 2.0
    return SystemSpec("NUBosonsBulkPBBoxAndRadial", n_particles, n_params, float(lbox), knots, np.ascontiguousarray(weights),
                      ptr, col, val, PAIR_RULE_CUT, np.asarray(system_params, dtype=np.float64), n_other=3 + int(gr_bin_count),
                      tail_param=-1, kind=KIND_BOX_RADIAL, n_ext=2 * K, dim=int(dim),
                      extra=dict(n_splines=K, gr_bins=int(gr_bin_count), half=half, gr_spacing=half / float(gr_bin_count),
                                 grad_swap=(K - 1, 2 * K - 1, PR - 1)))


def inh_contact_bosons(n_particles, lbox, n_params, system_params, spf, pc):
    """``InhContactBosons`` (InhContactBosons.cpp:64-247), one-dimensional.  ``spf`` / ``pc``: dicts with the reference's
    ``knots``, ``weights`` (SplineFactory::GetWeights3), ``bc_start``, ``bc_end`` (SetBoundaryConditions3_1D_*), ``np``
    = (np1, np2, np3, numberOfSplines) and ``node_spacing`` of the single-particle function (argument: the coordinate
    shifted into [0, L]) and of the pair correlation (argument: minimum-image distance <= L/2).  Extended sums
    ``[ss_spf | ss_pc]``; map of ``RefreshLocalOperators`` (:208-247): the first three single-particle operators take the
    first AND the last three splines (periodicity), the pair part has start and end boundary rows.  The exponent carries
    the parameter-free contact term ``-2 gamma h_pc ss_pc[0]`` (:764), ``gamma = SYSTEM_PARAMS[1] * SYSTEM_PARAMS[2] * pi``
    if ``SYSTEM_PARAMS[0] == 0`` (:25-29).  Coordinates travel as R[N][3] with the coordinate in component 0."""
    sp = np.asarray(system_params, dtype=np.float64)
    if len(sp) != 4:
        raise ValueError("InhContactBosons: the four-entry SYSTEM_PARAMS {range, strength, k, V0} is supported")
    k1, k2 = int(spf["np"][3]), int(pc["np"][3])
    s1, s2, s3 = (int(x) for x in spf["np"][:3])
    p1, p2, p3 = (int(x) for x in pc["np"][:3])
    if n_params != s3 + p3:
        raise ValueError(f"InhContactBosons needs N_PARAM = {s3 + p3}, got {n_params}")
    bs, be = np.asarray(spf["bc_start"], np.float64), np.asarray(spf["bc_end"], np.float64)
    rows = [[(j, bs[i][j]) for j in range(3)] + [(k1 - 3 + j, be[i][j]) for j in range(3)] for i in range(s1)]
    rows += [[(3 + (i - s1), 1.0)] for i in range(s1, s2)]
    bs, be = np.asarray(pc["bc_start"], np.float64), np.asarray(pc["bc_end"], np.float64)
    rows += [[(k1 + j, bs[i][j]) for j in range(3)] for i in range(p1)]
    rows += [[(k1 + 3 + (i - p1), 1.0)] for i in range(p1, p2)]
    rows += [[(k1 + k2 - 3 + j, be[i][j]) for j in range(3)] for i in range(p3 - p2)]
    ptr, col, val = _csr(rows)
    gamma = (float(sp[1]) if sp[0] == 0.0 else 0.0) * (float(sp[2]) * np.pi)   # gamma *= potentialK * kf (:29)
    knots = np.concatenate([np.asarray(spf["knots"], np.float64), np.asarray(pc["knots"], np.float64)])
    weights = np.concatenate([np.asarray(spf["weights"], np.float64).reshape(k1, 4, 4), np.asarray(pc["weights"], np.float64).reshape(k2, 4, 4)])
    return SystemSpec("InhContactBosons", n_particles, n_params, float(lbox), knots, np.ascontiguousarray(weights), ptr, col, val,
                      PAIR_RULE_CUT, sp, n_other=9, dim=1, tail_param=-1, kind=KIND_INH_CONTACT, n_ext=k1 + k2,
                      extra=dict(n_splines=k1 + k2, n_splines_spf=k1, n_splines_pc=k2, h_pc=float(pc["node_spacing"]), gamma=gamma,
                                 r_max=float(np.asarray(pc["knots"])[-4])))


def he_bulk(n_particles, lbox, n_params):
    """``HeBulk`` (HeBulk.cpp:40-70): ``K = P + 5`` uniform splines of spacing ``h = (L/2 - rs)/(K - 3)`` starting at the
    McMillan split ``rs = 1.95``; extended sums are ``[ss_0 .. ss_{K-1}, mcMillanSum, constSum, linearSum]`` (the last two
    unused); parameter map of :376-383 with the constant 1 of the last operator and the literal 1 added to its gradient (:351)."""
    P = n_params
    K = P - 1 + 3 + 3
    rs = 1.95
    half = lbox / 2.0
    h = (half - rs) / float(K - 3.0)
    f11 = 10.0 * h / rs ** 6.0
    f21 = (-5.0 * h + 3.0 * rs) / (2.0 * rs ** 6.0)
    MC = K  # column of the McMillan sum
    rows = [[(MC, 1.0), (0, f11), (1, f21)], [(2, 1.0), (0, 1.0), (1, -0.5)]]
    rows += [[(i + 1, 1.0)] for i in range(2, P - 2)]
    rows += [[(K - 6, 1.0), (K - 5, -0.5), (K - 4, 1.0)], [(K - 5, -1.5), (K - 4, 0.0)]]
    ptr, col, val = _csr(rows)
    mconst = np.zeros(P)
    mconst[P - 1] = 1.0
    gconst = np.zeros(P)
    gconst[P - 1] = 1.0
    knots = rs + h * np.arange(-3, K + 1, dtype=np.float64)   # informational only: the kernels use (rs, h)
    inf = float("inf")
    return SystemSpec("HeBulk", n_particles, P, float(lbox), knots, np.zeros((K, 4, 4)), ptr, col, val, PAIR_RULE_CUT,
                      np.zeros(0), n_other=3 + 100, tail_param=-1, kind=KIND_HE_BULK, n_ext=K + 3, map_const=mconst,
                      grad_const=gconst,
                      extra=dict(n_splines=K, n_short=K, r_max=half, rij_split=rs, h=h, h_large=h, r_split2=inf, r_tail=inf,
                                 mcm=-5.0, gr_bins=100, rho_bins=0, gr_max=half, periodic=1, potential=0, use_phi=0,
                                 factors=(f11, 1.0, f21, -0.5, -0.5, 1.0, -1.5, 0.0)))


def he_drop(n_particles, n_params):
    """``HeDrop`` (HeDrop.cpp:71-135): open boundary, McMillan ``r^-4.7`` below ``rs = 3``, 70 splines of spacing 0.1, then
    spacing 0.5 up to ``r_tail``, constant + linear tails beyond; ``K = P + 3``; parameter map of :609-626."""
    P = n_params
    m, rs, hS, hL, nS = -4.7, 3.0, 0.1, 0.5, 70
    K = P + 1 + 2
    nL = K - nS
    r2 = hS * (nS - 3.0) + rs
    rt = hL * (nL - 3.0) + r2
    fFS1 = -2.0 * m * hS * rs ** (m - 1.0)
    fSS1 = (m * hS + 3.0 * rs) * rs ** (m - 1.0) / 2.0
    d = 1.0 / (hS + hL)
    fSLS, fSLL = (-hS + hL) * d, (2.0 * hL) * d
    fLS, fLL = (-4.0 * hS) * d, (4.0 * hL) * d
    fFS, fFL = (4.0 * hS) * d, (-4.0 * hL) * d
    fSS, fSL = (2.0 * hS) * d, (hS - hL) * d
    MC, CO, LI = K, K + 1, K + 2
    rows = [[(MC, 1.0), (0, fFS1), (1, fSS1)], [(2, 1.0), (0, 1.0), (1, -0.5)]]
    rows += [[(i + 1, 1.0)] for i in range(2, nS - 4)]
    rows += [[(nS - 3, 1.0), (nS - 1, fSLS), (nS, fSLL)], [(nS - 2, 1.0), (nS - 1, fLS), (nS, fLL)],
             [(nS + 1, 1.0), (nS - 1, fFS), (nS, fFL)], [(nS + 2, 1.0), (nS - 1, fSS), (nS, fSL)]]
    rows += [[(i + 3, 1.0)] for i in range(nS, P - 3)]
    rows += [[(K - 3, 1.0), (K - 2, -0.5), (K - 1, 1.0)], [(CO, 1.0), (K - 2, 1.5), (K - 1, 0.0)],
             [(LI, 1.0), (K - 2, 1.5 * rt - hL / 2.0), (K - 1, 2.0 * hL)]]
    assert len(rows) == P
    ptr, col, val = _csr(rows)
    knots = np.concatenate([rs + hS * np.arange(0, nS - 3), r2 + hL * np.arange(0, nL - 2)])
    inf = float("inf")
    return SystemSpec("HeDrop", n_particles, P, 0.0, knots, np.zeros((K, 4, 4)), ptr, col, val, PAIR_RULE_CUT, np.zeros(0),
                      n_other=3 + 200 + 200, tail_param=-1, kind=KIND_HE_DROP, n_ext=K + 3,
                      extra=dict(n_splines=K, n_short=nS, r_max=inf, rij_split=rs, h=hS, h_large=hL, r_split2=r2, r_tail=rt,
                                 mcm=m, gr_bins=200, rho_bins=200, gr_max=2.0 * rt, periodic=0, potential=1, use_phi=1))


def species_hbar_over_2m(mass):
    """hbar^2 / (2 m u) / (A^2 k_B) with the constants of src/Constants.h:26-33 (BosonMixtureCluster.cpp:113)."""
    hbar, u, A2m, kb = 1.054571628e-34, 1.660538782e-27, 1e-10, 1.3806504e-23
    return hbar ** 2.0 / (2.0 * mass * u) / (A2m ** 2.0 * kb)


def boson_mixture_cluster(particle_types, type_knots, type_weights, type_bc, n_other=403, type_mcm=None, order=3):
    """``BosonMixtureCluster`` (BosonMixtureCluster.cpp:58-346) and, with ``order=4``, ``BosonMixtureCluster_4thorder``
    (BosonMixtureCluster_4thorder.cpp:104-346, config/He4He4Na_4thOrder.config).  ``particle_types``: the config's
    PARTICLE_TYPES (enum values).  Species and pair types are numbered in order of first appearance (:58-102).  Per pair
    type the caller passes the reference's knots (K + order + 1), spline table (K x (order+1) x (order+1),
    SplineFactory::GetWeights3 / GetWeights4) and boundary factors bcFactors (5 x (order-1); SetBoundaryConditions1_MM_1 /
    1_EXP_2 and their _4thorder twins).  K = 26 cubic / 28 quartic splines; 26 parameters per pair type either way (:543);
    extended sums per type: ``[ss_0..ss_{K-1} | mcMillan | const | linear | log]``; map of :636-645 (4th order: :641-650, the
    plain operators are ``ss[i + 2]``)."""
    if order not in (3, 4):
        raise ValueError("spline order must be 3 or 4")
    pt = [int(x) for x in particle_types]
    N = len(pt)
    pair_index, pair_type = {}, np.zeros((N, N), dtype=np.int32)
    for i in range(N):
        for j in range(i):
            key = frozenset((pt[i], pt[j])) if pt[i] != pt[j] else (pt[i],)
            if key not in pair_index:
                pair_index[key] = len(pair_index)
            pair_type[i, j] = pair_type[j, i] = pair_index[key]
    T = len(pair_index)
    nb = order - 1                 # boundary factors per row: the first / last nb splines are tied to their neighbours
    K = 26 + (order - 3) * 2
    EXT = K + 4
    MC, CO, LI, LG = K, K + 1, K + 2, K + 3
    rows = []
    for t in range(T):
        bc = np.asarray(type_bc[t], dtype=np.float64)
        b = t * EXT
        first = [(b + j, bc[0][j]) for j in range(nb)]
        rows.append([(b + MC, 1.0)] + first)
        rows.append([(b + nb, 1.0)] + [(b + j, bc[1][j]) for j in range(nb)])
        rows += [[(b + i + nb - 1, 1.0)] for i in range(2, 22)]
        last = lambda r: [(b + K - nb + j, bc[r][j]) for j in range(nb)]
        rows.append([(b + K - nb - 1, 1.0)] + last(2))
        rows.append([(b + CO, 1.0)] + last(3))
        rows.append([(b + LI, 1.0)] + last(4))
        rows.append([(b + LG, 1.0)])
    ptr, col, val = _csr(rows)
    pots = []
    for key, t in sorted(pair_index.items(), key=lambda kv: kv[1]):
        sp = set(key)
        pots.append(POT_KTTY_HE_CS if 4 in sp else (POT_KTTY_HE_NA if 2 in sp else POT_HFDB_HE_HE))   # :233-282
    mass = np.array([SPECIES_MASS[x] for x in pt])
    hbar = np.array([species_hbar_over_2m(m) for m in mass])
    knots = np.ascontiguousarray(type_knots, np.float64).reshape(T, -1)
    weights = np.ascontiguousarray(type_weights, np.float64).reshape(T, -1, order + 1, order + 1)
    if order == 4 and knots.shape[1] == K + order and weights.shape[1] == K - 1:
        # The reference sizes its sums for numberOfSplines = 28 (BosonMixtureCluster_4thorder.cpp:138-146) while the
        # config's 32 knots carry 27 quartic splines (SplineFactory.cpp:112-114): spline 27 never receives a term but
        # is a column of the operator map (:647-649).  Kept as a zero spline behind one padding knot, so that the
        # layout stays [K splines | extras]; rijTail = nodes[size - 5] is knots[K - 1] of the padded vector.
        knots = np.concatenate([knots, knots[:, -1:] + 1.0], axis=1)
        weights = np.concatenate([weights, np.zeros((T, 1, order + 1, order + 1))], axis=1)
    if knots.shape[1] != K + order + 1 or weights.shape[1] != K:
        raise ValueError(f"expected {K} splines on {K + order + 1} knots per pair type")
    mcm = np.full(T, -4.7) if type_mcm is None else np.asarray(type_mcm, np.float64)
    name = "BosonMixtureCluster" if order == 3 else "BosonMixtureCluster_4thorder"
    return SystemSpec(name, N, 26 * T, 0.0, knots[0], weights[0], ptr, col, val, PAIR_RULE_CUT, np.zeros(0),
                      n_other=n_other, tail_param=-1, kind=KIND_MIXTURE, n_ext=T * EXT,
                      extra=dict(n_splines=K, n_types=T, pair_type=pair_type, type_knots=knots, type_weights=weights,
                                 type_mcm=mcm, type_potential=np.array(pots, dtype=np.int32), hbar=hbar, mass=mass,
                                 r_max=float("inf"), order=order))


def from_golden(g):
    """Build the spec of a tests/golden fixture, taking knots and spline table from the reference dump."""
    name = str(g["system"])
    N, L, P = int(g["N"]), float(g["LBOX"]), int(g["N_PARAM"])
    dim = int(g["DIM"]) if "DIM" in getattr(g, "files", g) else 3
    if name == "HeBulk":
        spec = he_bulk(N, L, P)
        if spec.extra["h"] != float(g["node_point_spacing"]) or not np.array_equal(spec.extra["factors"], g["bc_factors"]):
            raise AssertionError("HeBulk set-up differs from the reference dump")
        return spec
    if name in ("BosonMixtureCluster", "BosonMixtureCluster_4thorder"):
        T = int(g["n_pair_types"])
        spec = boson_mixture_cluster(g["PARTICLE_TYPES"], [g[f"knots_{t}"] for t in range(T)],
                                     [g[f"spline_weights_{t}"] for t in range(T)], [g[f"bc_factors_{t}"] for t in range(T)],
                                     n_other=len(g["other_expectation_values"]), type_mcm=[g[f"extras_{t}"][6] for t in range(T)],
                                     order=3 if name == "BosonMixtureCluster" else 4)
        if not np.array_equal(spec.extra["pair_type"].ravel(), g["correlation_types"].astype(np.int32)):
            raise AssertionError("pair-type numbering differs from the reference dump")
        ref_hb = g["type_hbar_over_2m"][g["particle_types"].astype(int)]
        if not np.allclose(spec.extra["hbar"], ref_hb, rtol=1e-15, atol=0):
            raise AssertionError("hbar^2/2m differs from the reference dump")
        spec.extra["hbar"] = ref_hb
        return spec
    if name == "HeDrop":
        spec = he_drop(N, P)
        ref = g["bc_factors"]   # every factor* member of the reference object, in declaration order (HeDrop.h:70-89)
        rows = spec.map_rows()
        nS, K = 70, spec.n_splines
        mine = [rows[0][1][1], rows[1][1][1], rows[0][2][1], rows[1][2][1], rows[P - 3][1][1], rows[P - 2][1][1],
                rows[P - 1][1][1], rows[P - 3][2][1], rows[P - 2][2][1], rows[P - 1][2][1], rows[nS - 4][1][1],
                rows[nS - 4][2][1], rows[nS - 3][1][1], rows[nS - 3][2][1], rows[nS - 2][1][1], rows[nS - 2][2][1],
                rows[nS - 1][1][1], rows[nS - 1][2][1]]
        if not np.array_equal(np.array(mine), ref) or spec.extra["r_tail"] != float(g["rij_tail"]):
            raise AssertionError("HeDrop set-up differs from the reference dump")
        return spec
    if name == "BosonsBulk":
        spec = bosons_bulk(N, L, P, g["SYSTEM_PARAMS"], weights=g["spline_weights"], dim=dim)
    elif name == "NUBosonsBulkPB":
        spec = nu_bosons_bulk_pb(N, L, P, g["NURBS_GRID"], g["SYSTEM_PARAMS"], weights=g["spline_weights"],
                                 gr_bin_count=len(g["other_expectation_values"]) - 9, dim=dim)
    elif name == "InhContactBosons":
        part = lambda q: dict(knots=g["knots_" + q], weights=g["spline_weights_" + q], bc_start=g["bc_start_" + q],
                              bc_end=g["bc_end_" + q], np=g["np_" + q], node_spacing=float(g["node_spacing_" + q]))
        spec = inh_contact_bosons(N, L, P, g["SYSTEM_PARAMS"], part("spf"), part("pc"))
        if spec.extra["gamma"] != float(g["gamma"]) or spec.r_max != float(g["max_distance"]):
            raise AssertionError("InhContactBosons set-up differs from the reference dump")
        return spec
    elif name == "NUBosonsBulkPBBoxAndRadial":
        spec = nu_bosons_bulk_pb_box_and_radial(N, L, P, g["NURBS_GRID"], g["SYSTEM_PARAMS"], weights=g["spline_weights"],
                                                gr_bin_count=len(g["other_expectation_values"]) - 3, dim=dim)
        if not (np.array_equal(g["knots"], g["knots_rad"]) and np.array_equal(g["spline_weights"], g["spline_weights_rad"])):
            raise AssertionError("the reference's radial and box bases are expected to share knots and table")
    else:
        raise ValueError(name)
    if not np.array_equal(spec.knots, g["knots"]):
        raise AssertionError("knot construction differs from the reference dump")
    return spec


def smooth_params(n_params, r_max, a_r=-0.5, w_r=0.8, a_i=0.05, c_i=1.5, w_i=0.5):
    """The fixed smooth parameter profile used for synthetic workloads."""
    h = r_max / (n_params - 1)
    k = np.arange(n_params)
    uR = a_r * np.exp(-((k * h / w_r) ** 2))
    uI = a_i * np.exp(-(((k * h - c_i) / w_i) ** 2))
    return uR, uI


def jittered_lattice(n_particles, lbox, rng):
    """Start-up lattice + U(-0.05, 0.05) l jitter, the shape of src/TDVMC.cpp:727-739."""
    m = int(round(n_particles ** (1.0 / 3.0)))
    if m ** 3 != n_particles:
        raise ValueError("cubic lattice needs N = m^3")
    l = lbox / m
    g = (np.arange(m) + 0.5) * l - lbox / 2
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    R = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    return R + rng.uniform(-0.05, 0.05, R.shape) * l
