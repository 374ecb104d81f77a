"""Host-side mirror of the reference's per-rank estimator loops on top of the CUDA library.

Same names, argument meaning and results as the reference's L3 entry points (src/TDVMC.cpp):

    ParallelUpdateExpectationValues(R, uR, uI, phiR, phiI)                  :1152-1220
    ParallelUpdateExpectationValuesForGivenSamples(samples, uR, uI, ...)    :1305-1330
    DoMetropolisSteps(R, uR, uI, phiR, phiI, n)                             :918-924
    MoveCoordinatesToFirstCell(R)                                           :787-796

with one difference in kind: the reference owns ONE walker per rank and loops MC_NSTEPS samples
over it, this class owns ``n_walkers`` device-resident walkers per GPU and draws MC_NSTEPS samples
from each.  The seven estimator arrays come back under the reference's global names
(src/TDVMC.cpp:147-153).  Everything here calls libtdvmc_b200.so; there is no CPU path.
"""
import numpy as np

from . import capi
from .estimators import shard_walkers


class GpuEnsembleSystem:
    def __init__(self, spec, n_walkers_total, mc_step, seed=1, rank=0, world=1, device=None, mc_nsteps=1,
                 update_samples_every_nth_step=0, unique_id=None):
        self.spec = spec
        self.rank, self.world = rank, world
        self.first_walker, self.n_local = shard_walkers(n_walkers_total, rank, world)
        self.n_walkers_total = n_walkers_total
        self.handle = capi.Handle(spec, self.n_local, seed=seed, mc_step=mc_step, first_walker=self.first_walker,
                                  max_samples=mc_nsteps, keep_sample_positions=update_samples_every_nth_step > 0,
                                  device=rank if device is None else device)
        if world > 1:
            if unique_id is None:
                raise ValueError("world > 1 needs the NCCL unique id of rank 0 (capi.comm_unique_id(), broadcast by the host)")
            self.handle.comm_init(unique_id, rank, world)

    # -- state ------------------------------------------------------------------------------
    def SetPositions(self, R):
        """R: [n_local][N][3] (this rank's walkers)."""
        self.handle.set_positions(R)

    def GetPositions(self):
        return self.handle.get_positions()

    def BroadcastNewParameters(self, uR, uI, phiR, phiI, time=0.0):
        """Every rank passes the same values (the reference broadcasts from root, :506-512)."""
        self.handle.set_params(uR, uI, phiR, phiI, time)

    def MoveCoordinatesToFirstCell(self):
        self.handle.wrap_positions()

    def DoMetropolisSteps(self, n):
        self.handle.sweep(n)

    # -- estimator loops ----------------------------------------------------------------------
    def ParallelUpdateExpectationValues(self, uR, uI, phiR, phiI, MC_NSTEPS, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS=0,
                                        time=0.0):
        self.handle.set_params(uR, uI, phiR, phiI, time)
        self.handle.sample_and_accumulate(MC_NSTEPS, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS)
        return self._fetch()

    def ParallelUpdateExpectationValuesForGivenSamples(self, uR, uI, phiR, phiI, time=0.0):
        self.handle.set_params(uR, uI, phiR, phiI, time)
        self.handle.reevaluate_stored()
        return self._fetch()

    def ParallelCalculateAdditionalSystemProperties(self, uR, uI, phiR, phiI, observables, MC_NADDITIONALSTEPS,
                                                    MC_NADDITIONALTHERMSTEPS, MC_NADDITIONALINITIALIZATIONSTEPS, time=0.0):
        """src/TDVMC.cpp:1438-1444: the end-of-run observable pass; returns the reference's additionalObservablesMean
        as ``dict(pairDistribution=g(r), structureFactor=S(k))``; observables: tdvmc_b200.observables.ObservableSpec."""
        self.handle.set_params(uR, uI, phiR, phiI, time)
        gr, sk = self.handle.sample_observables(observables, MC_NADDITIONALSTEPS, MC_NADDITIONALTHERMSTEPS,
                                                MC_NADDITIONALINITIALIZATIONSTEPS)
        return dict(pairDistribution=gr, structureFactor=sk)

    def UpdateSamplesConsecutive(self, nrOfSamplesToUpdate, uR, uI, phiR, phiI, MC_NTHERMSTEPS, time=0.0):
        """src/TDVMC.cpp:975-983: the next stored samples of every walker advance by MC_NTHERMSTEPS steps."""
        self.handle.set_params(uR, uI, phiR, phiI, time)
        self.handle.update_stored(nrOfSamplesToUpdate, MC_NTHERMSTEPS)

    def ParallelCalculateAdditionalSystemPropertiesCluster(self, uR, uI, phiR, phiI, grids, MC_NADDITIONALSTEPS,
                                                           MC_NADDITIONALTHERMSTEPS, MC_NADDITIONALINITIALIZATIONSTEPS, time=0.0):
        """The same pass for the three-particle mixture cluster (BosonMixtureCluster.cpp:680-741): r2, angularDistribution,
        densityFromCOM, particleDistances means; grids: dict(angle_grid, density_grid, distance_grid, density_scaling)."""
        self.handle.set_params(uR, uI, phiR, phiI, time)
        r2, angle, density, distance = self.handle.sample_cluster_observables(grids, MC_NADDITIONALSTEPS, MC_NADDITIONALTHERMSTEPS,
                                                                              MC_NADDITIONALINITIALIZATIONSTEPS)
        return dict(r2=r2, angularDistribution=angle, densityFromCOM=density, particleDistances=distance)

    # -- parameter derivatives and the Euler step, solved on the device ------------------------------
    def SampleExpectationValues(self, uR, uI, phiR, phiI, MC_NSTEPS, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS=0, time=0.0):
        """ParallelUpdateExpectationValues without the fetch: the estimator sums stay on the device for
        SolveForParametersDot / CalculateNextParametersEuler below."""
        self.handle.set_params(uR, uI, phiR, phiI, time)
        self.handle.sample_and_accumulate(MC_NSTEPS, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS)

    def SolveForParametersDot(self, IMAGINARY_TIME=1, USE_PRECONDITIONING=1, regularization=0.001):
        """src/TDVMC.cpp:1713-1763 (Cholesky branch) on the all-reduced device estimators:
        returns (uDotR, uDotI, phiDotR, phiDotI, info)."""
        d = self.handle.solve_parameters_dot(imaginary_time=IMAGINARY_TIME, use_preconditioning=USE_PRECONDITIONING,
                                             regularization=regularization)
        return d["u_dot_r"], d["u_dot_i"], d["phi_dot_r"], d["phi_dot_i"], d

    def CalculateNextParametersEuler(self, dt, uR, uI, phiR, phiI, IMAGINARY_TIME=1, USE_PRECONDITIONING=1, time=0.0,
                                     regularization=0.001):
        """src/TDVMC.cpp:1834-1853 + BroadcastNewParameters (:506-512): returns the new (uR, uI, phiR, phiI, info) and
        leaves them current on the device; info carries <E^R>, <E^I> of the step and doNotAcceptStep."""
        return self.handle.euler_step(dt, uR, uI, phiR, phiI, time=time, imaginary_time=IMAGINARY_TIME,
                                      use_preconditioning=USE_PRECONDITIONING, regularization=regularization)

    # -- the explicit multi-stage integrators: the same solve after every stage, estimators never fetched --------------
    def _stage(self, uR, uI, phiR, phiI, sampling, time, solver):
        """One intermediate stage: estimators at the given parameters (fresh sampling with the MC counts in `sampling`,
        or re-evaluation of the stored samples when `sampling` is None), then SolveForParametersDot on the device."""
        if sampling is None:
            self.handle.set_params(uR, uI, phiR, phiI, time)
            self.handle.reevaluate_stored()                       # ParallelUpdateExpectationValuesForGivenSamples(..., true)
        else:
            self.SampleExpectationValues(uR, uI, phiR, phiI, *sampling, time=time)   # ParallelUpdateExpectationValues(..., true)
        return self.handle.solve_parameters_dot(**solver)

    def _predictor_corrector(self, dt, uR, uI, phiR, phiI, pc_steps, sampling, time, solver):
        uR, uI = np.asarray(uR, np.float64), np.asarray(uI, np.float64)
        dt_2 = dt / 2.0
        d0 = self.handle.solve_parameters_dot(**solver)
        tR, tI = uR + d0["u_dot_r"] * dt, uI + d0["u_dot_i"] * dt
        tpR, tpI = phiR + d0["phi_dot_r"] * dt, phiI + d0["phi_dot_i"] * dt
        for _ in range(pc_steps):
            d1 = self._stage(tR, tI, tpR, tpI, sampling, time, solver)
            tR, tI = uR + (d0["u_dot_r"] + d1["u_dot_r"]) * dt_2, uI + (d0["u_dot_i"] + d1["u_dot_i"]) * dt_2
            tpR, tpI = phiR + (d0["phi_dot_r"] + d1["phi_dot_r"]) * dt_2, phiI + (d0["phi_dot_i"] + d1["phi_dot_i"]) * dt_2
        self.handle.set_params(tR, tI, tpR, tpI, time)
        return tR, tI, tpR, tpI, d0

    def _rk4(self, dt, uR, uI, phiR, phiI, sampling, time, solver):
        uR, uI = np.asarray(uR, np.float64), np.asarray(uI, np.float64)
        dt_2 = dt / 2.0
        d = [self.handle.solve_parameters_dot(**solver)]
        for step in (dt_2, dt_2, dt):
            k = d[-1]
            d.append(self._stage(uR + k["u_dot_r"] * step, uI + k["u_dot_i"] * step, phiR + k["phi_dot_r"] * step,
                                 phiI + k["phi_dot_i"] * step, sampling, time, solver))
        comb = lambda key: (d[0][key] + d[1][key] * 2.0 + d[2][key] * 2.0 + d[3][key]) / 6.0
        nR, nI = uR + comb("u_dot_r") * dt, uI + comb("u_dot_i") * dt
        npR, npI = phiR + comb("phi_dot_r") * dt, phiI + comb("phi_dot_i") * dt
        self.handle.set_params(nR, nI, npR, npI, time)
        return nR, nI, npR, npI, d[0]

    @staticmethod
    def _solver(IMAGINARY_TIME, USE_PRECONDITIONING, regularization):
        return dict(imaginary_time=IMAGINARY_TIME, use_preconditioning=USE_PRECONDITIONING, regularization=regularization)

    def CalculateNextParametersPC(self, dt, uR, uI, phiR, phiI, MC_NSTEPS, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS=0,
                                  IMAGINARY_TIME=1, USE_PRECONDITIONING=1, time=0.0, regularization=0.001):
        """src/TDVMC.cpp:1855-1910 (ODE_SOLVER_TYPE 1): predictor-corrector, one corrector with a fresh sampling pass.  Like
        the Euler step it starts from the estimators of the current parameters already accumulated on the device."""
        return self._predictor_corrector(dt, uR, uI, phiR, phiI, 1, (MC_NSTEPS, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS), time,
                                         self._solver(IMAGINARY_TIME, USE_PRECONDITIONING, regularization))

    def CalculateNextParametersPCReuseSamples(self, dt, uR, uI, phiR, phiI, IMAGINARY_TIME=1, USE_PRECONDITIONING=1, time=0.0,
                                              regularization=0.001):
        """src/TDVMC.cpp:1912-1967: six correctors on the stored samples (needs update_samples_every_nth_step > 0)."""
        return self._predictor_corrector(dt, uR, uI, phiR, phiI, 6, None, time,
                                         self._solver(IMAGINARY_TIME, USE_PRECONDITIONING, regularization))

    def CalculateNextParametersRK4(self, dt, uR, uI, phiR, phiI, MC_NSTEPS, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS=0,
                                   IMAGINARY_TIME=1, USE_PRECONDITIONING=1, time=0.0, regularization=0.001):
        """src/TDVMC.cpp:1969-2035: classical Runge-Kutta, a fresh sampling pass per stage."""
        return self._rk4(dt, uR, uI, phiR, phiI, (MC_NSTEPS, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS), time,
                         self._solver(IMAGINARY_TIME, USE_PRECONDITIONING, regularization))

    def CalculateNextParametersRK4ReuseSamples(self, dt, uR, uI, phiR, phiI, IMAGINARY_TIME=1, USE_PRECONDITIONING=1, time=0.0,
                                               regularization=0.001):
        """src/TDVMC.cpp:2037-2103: Runge-Kutta on the stored samples."""
        return self._rk4(dt, uR, uI, phiR, phiI, None, time, self._solver(IMAGINARY_TIME, USE_PRECONDITIONING, regularization))

    def GetExponent(self):
        return self.handle.last_exponent()

    def _fetch(self):
        o = self.handle.allreduce_and_fetch()
        return dict(localOperators=o["O"], localEnergyR=float(o["e_r"][0]), localEnergyI=float(o["e_i"][0]),
                    localOperatorsMatrix=o["S"], localOperatorlocalEnergyR=o["OER"], localOperatorlocalEnergyI=o["OEI"],
                    otherExpectationValues=o["other"], nAcceptances=o["n_acceptances"], nTrials=o["n_trials"],
                    nSamples=o["n_samples"])

    def close(self):
        self.handle.close()
