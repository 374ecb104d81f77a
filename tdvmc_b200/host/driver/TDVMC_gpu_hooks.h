// TDVMC_gpu_hooks.h — the GPU binding of the reference driver (mathiasgartner/TDVMC, src/TDVMC.cpp).
//
// This file is OUR code.  It is textually included twice into a patched copy of the reference's src/TDVMC.cpp
// (patch_driver.py inserts the two #include lines and a dozen one-line call sites; nothing else of the driver
// changes): once after the driver's globals with TDVMC_GPU_HOOKS_DECL defined (declarations, new config globals),
// once before mainMPI with TDVMC_GPU_HOOKS_IMPL defined (definitions, which use the driver's own helpers).
//
// With GPU_WALKERS = 0 (or absent from the config) every hook returns false / does nothing and the driver runs the
// reference's CPU path unchanged.  With GPU_WALKERS > 0 the per-rank sampling loops are served by
// tdvmc_host::GpuEnsembleSystem (the C ABI of include/tdvmc_gpu.h): GPU_WALKERS walkers in total, split over the
// MPI ranks, each drawing MC_NSTEPS samples per estimator pass; everything downstream - SolveForParametersDot
// (Cholesky or Eigen QR), the ODE integrators, AcceptNewParams, rollback, the .dat output - is the driver's own code
// working on the seven estimator globals as before (src/TDVMC.cpp:147-153).
//
// The reference's plugin classes keep their knots / spline tables private and have no getters; a maintainer would
// add two getters.  To leave every reference header untouched, this file reads those members through
// explicit template instantiation (member-pointer arguments of an explicit instantiation are exempt from access
// checking, [temp.explicit]/12) - no macro redefinition, no modified header, the reference objects link unchanged.

#ifdef TDVMC_GPU_HOOKS_DECL
#undef TDVMC_GPU_HOOKS_DECL

#include "GpuEnsembleSystem.h"

namespace tdvmc_gpu_access
{
template <typename Tag> struct Slot
{
	static typename Tag::type ptr;
};
template <typename Tag> typename Tag::type Slot<Tag>::ptr;
template <typename Tag, typename Tag::type P> struct Fill
{
	Fill() { Slot<Tag>::ptr = P; }
	static Fill instance;
};
template <typename Tag, typename Tag::type P> Fill<Tag, P> Fill<Tag, P>::instance;
}
#define TDVMC_GPU_EXPOSE(Tag, Class, Type, member)         \
	struct Tag                                             \
	{                                                      \
		typedef Type Class::*type;                         \
	};                                                     \
	template struct tdvmc_gpu_access::Fill<Tag, &Class::member>;
#define TDVMC_GPU_MEMBER(Tag, object) ((object).*tdvmc_gpu_access::Slot<Tag>::ptr)

typedef vector<vector<vector<double> > > TdvmcGpuTensor3;
TDVMC_GPU_EXPOSE(BB_nodes, PhysicalSystems::BosonsBulk, vector<double>, nodes)
TDVMC_GPU_EXPOSE(BB_weights, PhysicalSystems::BosonsBulk, TdvmcGpuTensor3, splineWeights)
TDVMC_GPU_EXPOSE(BB_kValues, PhysicalSystems::BosonsBulk, TdvmcGpuTensor3, kValues)
TDVMC_GPU_EXPOSE(BB_gr, PhysicalSystems::BosonsBulk, Observables::ObservableVsOnGridWithScaling, pairDistribution)
TDVMC_GPU_EXPOSE(NU_nodes, PhysicalSystems::NUBosonsBulkPB, vector<double>, nodes)
TDVMC_GPU_EXPOSE(NU_weights, PhysicalSystems::NUBosonsBulkPB, TdvmcGpuTensor3, splineWeights)
TDVMC_GPU_EXPOSE(NU_kValues, PhysicalSystems::NUBosonsBulkPB, TdvmcGpuTensor3, kValues)
TDVMC_GPU_EXPOSE(NU_gr, PhysicalSystems::NUBosonsBulkPB, Observables::ObservableVsOnGridWithScaling, pairDistribution)
TDVMC_GPU_EXPOSE(NU_grBinCount, PhysicalSystems::NUBosonsBulkPB, int, grBinCount)
// BosonMixtureCluster keeps its per-species / per-pair-type data protected (BosonMixtureCluster.h:42-58)
TDVMC_GPU_EXPOSE(MX_corrTypes, PhysicalSystems::BosonMixtureCluster, vector<vector<int> >, correlationTypes)
TDVMC_GPU_EXPOSE(MX_particleTypes, PhysicalSystems::BosonMixtureCluster, vector<int>, particleTypes)
TDVMC_GPU_EXPOSE(MX_cfd, PhysicalSystems::BosonMixtureCluster, vector<CorrelationFunctionData>, corrFuncData)
TDVMC_GPU_EXPOSE(MX_pp, PhysicalSystems::BosonMixtureCluster, vector<ParticleProperties>, particleProperties)
TDVMC_GPU_EXPOSE(MX_ppp, PhysicalSystems::BosonMixtureCluster, vector<ParticlePairProperties>, particlePairProperties)
TDVMC_GPU_EXPOSE(MX_angular, PhysicalSystems::BosonMixtureCluster, Observables::ObservableVsOnGrid, angularDistribution)
TDVMC_GPU_EXPOSE(MX_density, PhysicalSystems::BosonMixtureCluster, Observables::ObservableVsOnGridWithScaling, densityFromCOM)
TDVMC_GPU_EXPOSE(MX_distances, PhysicalSystems::BosonMixtureCluster, Observables::ObservableVsOnGrid, particleDistances)
// BosonMixtureCluster_4thorder: the same members under the same names (BosonMixtureCluster_4thorder.h:26-62)
TDVMC_GPU_EXPOSE(M4_corrTypes, PhysicalSystems::BosonMixtureCluster_4thorder, vector<vector<int> >, correlationTypes)
TDVMC_GPU_EXPOSE(M4_particleTypes, PhysicalSystems::BosonMixtureCluster_4thorder, vector<int>, particleTypes)
TDVMC_GPU_EXPOSE(M4_cfd, PhysicalSystems::BosonMixtureCluster_4thorder, vector<CorrelationFunctionData>, corrFuncData)
TDVMC_GPU_EXPOSE(M4_pp, PhysicalSystems::BosonMixtureCluster_4thorder, vector<ParticleProperties>, particleProperties)
TDVMC_GPU_EXPOSE(M4_ppp, PhysicalSystems::BosonMixtureCluster_4thorder, vector<ParticlePairProperties>, particlePairProperties)
TDVMC_GPU_EXPOSE(M4_angular, PhysicalSystems::BosonMixtureCluster_4thorder, Observables::ObservableVsOnGrid, angularDistribution)
TDVMC_GPU_EXPOSE(M4_density, PhysicalSystems::BosonMixtureCluster_4thorder, Observables::ObservableVsOnGridWithScaling, densityFromCOM)
TDVMC_GPU_EXPOSE(M4_distances, PhysicalSystems::BosonMixtureCluster_4thorder, Observables::ObservableVsOnGrid, particleDistances)
// NUBosonsBulkPBBoxAndRadial (NUBosonsBulkPBBoxAndRadial.h:22, 41-42) and InhContactBosons (InhContactBosons.h:26-27)
TDVMC_GPU_EXPOSE(BR_nodes, PhysicalSystems::NUBosonsBulkPBBoxAndRadial, vector<double>, nodes)
TDVMC_GPU_EXPOSE(BR_weights, PhysicalSystems::NUBosonsBulkPBBoxAndRadial, TdvmcGpuTensor3, splineWeights)
TDVMC_GPU_EXPOSE(BR_grBinCount, PhysicalSystems::NUBosonsBulkPBBoxAndRadial, int, grBinCount)
TDVMC_GPU_EXPOSE(IC_spf, PhysicalSystems::InhContactBosons, WFParts::SingleParticleFunction, spf)
TDVMC_GPU_EXPOSE(IC_pc, PhysicalSystems::InhContactBosons, WFParts::PairCorrelation, pc)

// new config items (registered next to the reference's, src/TDVMC.cpp:297-349; absent keys stay 0)
int GPU_WALKERS = 0;      // total number of device-resident walkers over all ranks; 0: reference CPU path
int GPU_SEED = 0;         // Philox seed of the ensemble; 0: 1
int GPU_DEVICE_SOLVE = 0; // 1: SolveForParametersDot of the Euler step on the device (solve_kernel / solve_qr_kernel)

tdvmc_host::GpuEnsembleSystem* gpu = nullptr;
tdvmc_host::ObservableTables gpuObservableTables;
tdvmc_host::ClusterObservableTables gpuClusterObservableTables;
bool gpuHasObservables = false;
bool gpuHasClusterObservables = false;
bool gpuSamplesStored = false;

void GpuRegisterConfigItems();
void GpuInit(vector<vector<double> >& R);
void GpuBeginTimeStep();
void GpuAlignCoordinates();
void GpuVeryFirstInitialization(vector<double>& uR, vector<double>& uI, double phiR, double phiI);
bool GpuParallelUpdateExpectationValues(vector<double>& uR, vector<double>& uI, double phiR, double phiI, bool intermediateStep);
bool GpuParallelUpdateExpectationValuesForGivenSamples(vector<double>& uR, vector<double>& uI, double phiR, double phiI);
bool GpuUpdateSamplesConsecutive(int nrOfSamplesToUpdate, vector<double>& uR, vector<double>& uI, double phiR, double phiI);
bool GpuParallelCalculateAdditionalSystemProperties(vector<double>& uR, vector<double>& uI, double phiR, double phiI);
bool GpuCalculateNextParametersEuler(double dt, vector<double>& uR, vector<double>& uI, double* phiR, double* phiI);
double GpuExponentOr(double cpuExponent);
void GpuCopyWalkerToDriver(vector<vector<double> >& R);
void GpuShutdown();

#endif // TDVMC_GPU_HOOKS_DECL

#ifdef TDVMC_GPU_HOOKS_IMPL
#undef TDVMC_GPU_HOOKS_IMPL

void GpuRegisterConfigItems()
{
	configItems.push_back(ConfigItem("GPU_WALKERS", &GPU_WALKERS, ConfigItemType::INT));
	configItems.push_back(ConfigItem("GPU_SEED", &GPU_SEED, ConfigItemType::INT));
	configItems.push_back(ConfigItem("GPU_DEVICE_SOLVE", &GPU_DEVICE_SOLVE, ConfigItemType::INT));
}

static void GpuCopyEstimators(const tdvmc_host::Estimators& e)
{
	localOperators = e.localOperators;
	localEnergyR = e.localEnergyR;
	localEnergyI = e.localEnergyI;
	localOperatorsMatrix = e.localOperatorsMatrix;
	localOperatorlocalEnergyR = e.localOperatorlocalEnergyR;
	localOperatorlocalEnergyI = e.localOperatorlocalEnergyI;
	otherExpectationValues = e.otherExpectationValues;
	// already summed over ranks; the driver's ReduceToAverage(&nAcceptances) (:3727) then divides by numOfProcesses
	// for its log line only
	nAcceptances = e.nAcceptances;
	nTrials = e.nTrials;
}

// BosonMixtureCluster and BosonMixtureCluster_4thorder: per-species data per particle, per-pair-type spline sets and
// potentials as InitSystem() left them (BosonMixtureCluster.cpp:104-346; the 4th-order class differs in the spline
// order and the 5 x 3 boundary-condition factors only)
static bool GpuBindMixture(tdvmc_host::SystemTables& t, int splineOrder, vector<int>& pt, vector<ParticleProperties>& pp,
		vector<CorrelationFunctionData>& cfd, vector<ParticlePairProperties>& ppp, vector<vector<int> >& corrTypes,
		Observables::ObservableVsOnGrid& ang, Observables::ObservableVsOnGridWithScaling& den, Observables::ObservableVsOnGrid& dis)
{
	vector<double> hbarOver2m(N), mass(N);
	for (int n = 0; n < N; n++)
	{
		hbarOver2m[n] = pp[pt[n]].hbarOver2m;
		mass[n] = pp[pt[n]].mass;
	}
	vector<tdvmc_host::MixturePairType> types(cfd.size());
	for (size_t c = 0; c < cfd.size(); c++)
	{
		types[c].nodes = cfd[c].nodes;
		types[c].splineWeights = cfd[c].splineWeights;
		types[c].bcFactors = cfd[c].bcFactors;
		types[c].mcMillanFactor = cfd[c].mcMillanFactor;
		types[c].potential = dynamic_cast<Potentials::KTTY_He_Cs*>(ppp[c].potential) ? 2 : (dynamic_cast<Potentials::KTTY_He_Na*>(ppp[c].potential) ? 1 : 0);
		if (types[c].potential == 0 && !dynamic_cast<Potentials::HFDB_He_He*>(ppp[c].potential))
		{
			return false; // a pair potential the device does not carry (e.g. LJ_He_He)
		}
	}
	t = tdvmc_host::MakeBosonMixtureClusterTables(N, corrTypes, hbarOver2m, mass, types, sys->GetNumOfOtherExpectationValues(), splineOrder);
	gpuClusterObservableTables.angleCount = ang.grid.count;
	gpuClusterObservableTables.angleSpacing = ang.grid.spacing;
	gpuClusterObservableTables.densityCount = den.grid.count;
	gpuClusterObservableTables.densitySpacing = den.grid.spacing;
	gpuClusterObservableTables.densityMax = den.grid.max;
	gpuClusterObservableTables.densityScaling = den.scalingGrid;
	gpuClusterObservableTables.distanceCount = dis.grid.count;
	gpuClusterObservableTables.distanceSpacing = dis.grid.spacing;
	gpuClusterObservableTables.distanceMax = dis.grid.max;
	gpuHasClusterObservables = N == 3; // the reference's own pass hard-codes three particles (:696-706)
	return true;
}

// Called after sys->InitSystem(); PostSystemInit(); (src/TDVMC.cpp:3132-3133): the system's own InitSystem() results
// (knots, SplineFactory table, observable grids) become the device-side description.
void GpuInit(vector<vector<double> >& R)
{
	if (GPU_WALKERS <= 0)
	{
		return;
	}
	tdvmc_host::SystemTables t;
	bool known = true;
	if (auto s = dynamic_cast<PhysicalSystems::BosonsBulk*>(sys))
	{
		t = tdvmc_host::MakeBosonsBulkTables(N, LBOX, N_PARAM, TDVMC_GPU_MEMBER(BB_nodes, *s), TDVMC_GPU_MEMBER(BB_weights, *s), SYSTEM_PARAMS);
		t.dim = DIM;
		auto& gr = TDVMC_GPU_MEMBER(BB_gr, *s);
		gpuObservableTables.grCount = gr.grid.count;
		gpuObservableTables.grSpacing = gr.grid.spacing;
		gpuObservableTables.grMax = gr.grid.max;
		gpuObservableTables.grWeight = 1.0 / ((double) (N - 1)) * DIM; // BosonsBulk.cpp:481
		gpuObservableTables.grScaling = gr.scalingGrid;
		gpuObservableTables.kValues = TDVMC_GPU_MEMBER(BB_kValues, *s);
		gpuHasObservables = DIM == 3;
	}
	else if (auto s = dynamic_cast<PhysicalSystems::NUBosonsBulkPB*>(sys))
	{
		t = tdvmc_host::MakeNUBosonsBulkPBTables(N, LBOX, N_PARAM, TDVMC_GPU_MEMBER(NU_nodes, *s), TDVMC_GPU_MEMBER(NU_weights, *s), SYSTEM_PARAMS,
				TDVMC_GPU_MEMBER(NU_grBinCount, *s));
		t.dim = DIM;
		auto& gr = TDVMC_GPU_MEMBER(NU_gr, *s);
		gpuObservableTables.grCount = gr.grid.count;
		gpuObservableTables.grSpacing = gr.grid.spacing;
		gpuObservableTables.grMax = gr.grid.max;
		gpuObservableTables.grWeight = 1.0; // NUBosonsBulkPB.cpp:611
		gpuObservableTables.grScaling = gr.scalingGrid;
		gpuObservableTables.kValues = TDVMC_GPU_MEMBER(NU_kValues, *s);
		gpuHasObservables = DIM == 3;
	}
	else if (auto s = dynamic_cast<PhysicalSystems::BosonMixtureCluster*>(sys))
	{
		known = GpuBindMixture(t, 3, TDVMC_GPU_MEMBER(MX_particleTypes, *s), TDVMC_GPU_MEMBER(MX_pp, *s), TDVMC_GPU_MEMBER(MX_cfd, *s),
				TDVMC_GPU_MEMBER(MX_ppp, *s), TDVMC_GPU_MEMBER(MX_corrTypes, *s), TDVMC_GPU_MEMBER(MX_angular, *s),
				TDVMC_GPU_MEMBER(MX_density, *s), TDVMC_GPU_MEMBER(MX_distances, *s));
	}
	else if (auto s = dynamic_cast<PhysicalSystems::BosonMixtureCluster_4thorder*>(sys))
	{
		known = GpuBindMixture(t, 4, TDVMC_GPU_MEMBER(M4_particleTypes, *s), TDVMC_GPU_MEMBER(M4_pp, *s), TDVMC_GPU_MEMBER(M4_cfd, *s),
				TDVMC_GPU_MEMBER(M4_ppp, *s), TDVMC_GPU_MEMBER(M4_corrTypes, *s), TDVMC_GPU_MEMBER(M4_angular, *s),
				TDVMC_GPU_MEMBER(M4_density, *s), TDVMC_GPU_MEMBER(M4_distances, *s));
	}
	else if (auto s = dynamic_cast<PhysicalSystems::NUBosonsBulkPBBoxAndRadial*>(sys))
	{
		// g(r) rides in otherExpectationValues for this class, so no separate observable pass is bound
		t = tdvmc_host::MakeNUBosonsBulkPBBoxAndRadialTables(N, LBOX, N_PARAM, TDVMC_GPU_MEMBER(BR_nodes, *s), TDVMC_GPU_MEMBER(BR_weights, *s),
				SYSTEM_PARAMS, TDVMC_GPU_MEMBER(BR_grBinCount, *s));
		t.dim = DIM;
	}
	else if (auto s = dynamic_cast<PhysicalSystems::InhContactBosons*>(sys))
	{
		if (DIM != 1)
		{
			known = false;
		}
		else
		{
			tdvmc_host::SplinedFunctionTables f[2];
			WFParts::SplinedFunction* src[2] = { &TDVMC_GPU_MEMBER(IC_spf, *s), &TDVMC_GPU_MEMBER(IC_pc, *s) };
			for (int i = 0; i < 2; i++)
			{
				f[i].nodes = src[i]->nodes;
				f[i].splineWeights = src[i]->splineWeights;
				f[i].bcFactorsStart = src[i]->bcFactorsStart;
				f[i].bcFactorsEnd = src[i]->bcFactorsEnd;
				f[i].np1 = src[i]->np1;
				f[i].np2 = src[i]->np2;
				f[i].np3 = src[i]->np3;
			}
			t = tdvmc_host::MakeInhContactBosonsTables(N, LBOX, N_PARAM, SYSTEM_PARAMS, f[0], f[1]);
		}
	}
	else if (dynamic_cast<PhysicalSystems::HeBulk*>(sys))
	{
		t = tdvmc_host::MakeHeBulkTables(N, LBOX, N_PARAM);
	}
	else if (dynamic_cast<PhysicalSystems::HeDrop*>(sys))
	{
		t = tdvmc_host::MakeHeDropTables(N, N_PARAM);
	}
	else
	{
		known = false;
	}
	if (!known)
	{
		Log("GPU_WALKERS > 0 but SYSTEM_TYPE " + SYSTEM_TYPE + " has no device binding in this driver: running the CPU path", WARNING);
		return;
	}

	// sample capacity per walker: MC_NSTEPS x the write-data factor (UpdateExpectationValues, :1047) x the retry
	// factor of the acceptance loop (:3650-3651, at most maxNrOfAcceptParameterTrials = 2)
	int capacity = mc_nsteps_original * (MC_NSTEP_MULTIPLICATION_FACTOR_FOR_WRITE_DATA > 1 ? MC_NSTEP_MULTIPLICATION_FACTOR_FOR_WRITE_DATA : 1) * 2;
	int deviceCount = tdvmc_gpu_device_count();
	if (deviceCount < 1)
	{
		Log("GPU_WALKERS > 0 but no CUDA device is visible (libtdvmc_b200.so has no CPU fallback)", ERROR);
		MPI_Abort(MPI_COMM_WORLD, 3);
		exit(3);
	}
	try
	{
		gpu = new tdvmc_host::GpuEnsembleSystem(t, GPU_WALKERS, MC_STEP, capacity, UPDATE_SAMPLES_EVERY_NTH_STEP,
				(unsigned long long) (GPU_SEED > 0 ? GPU_SEED : 1), processRank, numOfProcesses, processRank % deviceCount);
		if (numOfProcesses > 1)
		{
			// NCCL communicator over the same ranks as MPI_COMM_WORLD: rank 0 creates the id, MPI carries it
			vector<unsigned char> id(TDVMC_GPU_UNIQUE_ID_BYTES);
			if (isRootRank)
			{
				id = tdvmc_host::GpuEnsembleSystem::CreateCommunicatorId();
			}
			MPI_Bcast(id.data(), (int) id.size(), MPI_CHAR, rootRank, MPI_COMM_WORLD);
			gpu->JoinCommunicator(id);
		}
		// start configurations: the driver's own R (InitCoordinateConfiguration, :672-785: restart file or lattice / drop
		// with jitter from the rank-seeded generator) for the first local walker, re-jittered copies for the others
		int nLocal = gpu->LocalWalkers();
		double l = sys->USE_NIC ? LBOX / round(pow(N, 1.0 / ((double) DIM))) : pow(LBOX, 1.0 / 3.0);
		vector<vector<vector<double> > > Rw(nLocal, vector<vector<double> >(N, vector<double>(3, 0.0)));
		for (int w = 0; w < nLocal; w++)
		{
			for (int i = 0; i < N; i++)
			{
				for (int a = 0; a < DIM; a++)
				{
					Rw[w][i][a] = R[i][a] + (w == 0 ? 0.0 : (random01() - 0.5) * l / 10.0);
				}
			}
		}
		gpu->SetPositions(Rw);
	}
	catch (const std::exception& ex)
	{
		Log(string("GPU initialisation failed: ") + ex.what(), ERROR);
		MPI_Abort(MPI_COMM_WORLD, 3);
		exit(3);
	}
	if (isRootRank)
	{
		Log("GPU ensemble: " + to_string(GPU_WALKERS) + " walkers over " + to_string(numOfProcesses) + " rank(s), " + to_string(gpu->LocalWalkers()) + " on this rank, sample capacity " + to_string(capacity) + " per walker");
	}
}

// nAcceptances = 0; nTrials = 0 at the start of a time step (src/TDVMC.cpp:3428-3429)
void GpuBeginTimeStep()
{
	if (gpu)
	{
		gpu->SetMCStep(MC_STEP); // runtime-changeable through ./param (:2718-2744)
		gpu->ResetCounters();
	}
}

// AlignCoordinates (src/TDVMC.cpp:2569-2582) for the device-resident walkers
void GpuAlignCoordinates()
{
	if (gpu && (sys->USE_NIC || sys->USE_MOVE_COM_TO_ZERO))
	{
		gpu->MoveCoordinatesToFirstCell(); // first cell for the periodic systems, (mass-weighted) centre of mass to zero for the open ones
	}
}

// MC_VERY_FIRST_NINITIALIZATIONSTEPS Metropolis steps of every walker (src/TDVMC.cpp:3411-3418)
void GpuVeryFirstInitialization(vector<double>& uR, vector<double>& uI, double phiR, double phiI)
{
	if (gpu)
	{
		gpu->DoMetropolisSteps(MC_VERY_FIRST_NINITIALIZATIONSTEPS, uR, uI, phiR, phiI);
		gpu->ResetCounters();
	}
}

// ParallelUpdateExpectationValues (src/TDVMC.cpp:1152-1188): the estimator pass and the seven reductions
bool GpuParallelUpdateExpectationValues(vector<double>& uR, vector<double>& uI, double phiR, double phiI, bool intermediateStep)
{
	if (!gpu)
	{
		return false;
	}
	// UpdateExpectationValues multiplies MC_NSTEPS on the steps that are written to file and divides again (:1047, :1149)
	int nSamples = MC_NSTEPS * (sys->GetStep() % WRITE_EVERY_NTH_STEP_TO_FILE == 0 ? MC_NSTEP_MULTIPLICATION_FACTOR_FOR_WRITE_DATA : 1);
	if (nSamples < 1)
	{
		nSamples = MC_NSTEPS;
	}
	mc_nsteps = (double) nSamples;
	try
	{
		GpuCopyEstimators(gpu->ParallelUpdateExpectationValues(uR, uI, phiR, phiI, nSamples, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS, sys->GetTime()));
	}
	catch (const std::exception& ex)
	{
		Log(string("GPU estimator pass failed: ") + ex.what(), ERROR);
		MPI_Abort(MPI_COMM_WORLD, 3);
		exit(3);
	}
	gpuSamplesStored = UPDATE_SAMPLES_EVERY_NTH_STEP > 0;
	if (isRootRank && !intermediateStep)
	{
		cout << "Acceptance: " << (nAcceptances / (nTrials / 100.0)) << "% (" << nAcceptances << "/" << nTrials << ")" << endl;
	}
	return true;
}

// ParallelUpdateExpectationValuesForGivenSamples (src/TDVMC.cpp:1305-1330)
bool GpuParallelUpdateExpectationValuesForGivenSamples(vector<double>& uR, vector<double>& uI, double phiR, double phiI)
{
	if (!gpu)
	{
		return false;
	}
	try
	{
		GpuCopyEstimators(gpu->ParallelUpdateExpectationValuesForGivenSamples(uR, uI, phiR, phiI, sys->GetTime()));
	}
	catch (const std::exception& ex)
	{
		Log(string("GPU re-evaluation of the stored samples failed: ") + ex.what(), ERROR);
		MPI_Abort(MPI_COMM_WORLD, 3);
		exit(3);
	}
	return true;
}

// UpdateSamplesConsecutive (src/TDVMC.cpp:975-983)
bool GpuUpdateSamplesConsecutive(int nrOfSamplesToUpdate, vector<double>& uR, vector<double>& uI, double phiR, double phiI)
{
	if (!gpu)
	{
		return false;
	}
	if (nrOfSamplesToUpdate > 0)
	{
		gpu->UpdateSamplesConsecutive(nrOfSamplesToUpdate, uR, uI, phiR, phiI, MC_NTHERMSTEPS, sys->GetTime());
	}
	return true;
}

// ParallelCalculateAdditionalSystemProperties (src/TDVMC.cpp:1438-1444) for the bulk spline systems: g(r) and S(k)
bool GpuParallelCalculateAdditionalSystemProperties(vector<double>& uR, vector<double>& uI, double phiR, double phiI)
{
	if (gpu && gpuHasClusterObservables && MC_NADDITIONALSTEPS > 0)
	{
		// BosonMixtureCluster::CalculateAdditionalSystemProperties (BosonMixtureCluster.cpp:680-741): r2, the three corner
		// angles, density from the centre of mass, pair distances - additionalObservables[0..3] in that order (:342-345)
		auto o = gpu->ParallelCalculateAdditionalSystemPropertiesCluster(uR, uI, phiR, phiI, gpuClusterObservableTables, MC_NADDITIONALSTEPS,
				MC_NADDITIONALTHERMSTEPS, MC_NADDITIONALINITIALIZATIONSTEPS, sys->GetTime());
		additionalObservablesMean.ClearValues();
		if (auto r = dynamic_cast<Observables::Observable*>(additionalObservablesMean.observables[0]))
		{
			r->value = o.r2;
		}
		const vector<double>* src[3] = { &o.angularDistribution, &o.densityFromCOM, &o.particleDistances };
		for (int q = 0; q < 3; q++)
		{
			if (auto g = dynamic_cast<Observables::ObservableVsOnGrid*>(additionalObservablesMean.observables[q + 1]))
			{
				const size_t count = src[q]->size() / 3;
				for (size_t v = 0; v < g->observablesV.size() && v < 3; v++)
				{
					for (size_t i = 0; i < g->observablesV[v].values.size() && i < count; i++)
					{
						g->observablesV[v].values[i] = (*src[q])[v * count + i];
					}
				}
			}
		}
		return true;
	}
	if (!gpu || !gpuHasObservables || MC_NADDITIONALSTEPS <= 0)
	{
		return false;
	}
	auto o = gpu->ParallelCalculateAdditionalSystemProperties(uR, uI, phiR, phiI, gpuObservableTables, MC_NADDITIONALSTEPS, MC_NADDITIONALTHERMSTEPS,
			MC_NADDITIONALINITIALIZATIONSTEPS, sys->GetTime());
	additionalObservablesMean.ClearValues();
	if (auto g = dynamic_cast<Observables::ObservableVsOnGrid*>(additionalObservablesMean.observables[0]))
	{
		for (size_t i = 0; i < g->observablesV[0].values.size() && i < o.pairDistribution.size(); i++)
		{
			g->observablesV[0].values[i] = o.pairDistribution[i];
		}
	}
	if (auto s = dynamic_cast<Observables::ObservableVsOnGrid*>(additionalObservablesMean.observables[1]))
	{
		for (size_t i = 0; i < s->observablesV[0].values.size() && i < o.structureFactor.size(); i++)
		{
			s->observablesV[0].values[i] = o.structureFactor[i];
		}
	}
	return true;
}

// CalculateNextParametersEuler (src/TDVMC.cpp:1834-1853) with SolveForParametersDot (either branch) on the device
bool GpuCalculateNextParametersEuler(double dt, vector<double>& uR, vector<double>& uI, double* phiR, double* phiI)
{
	if (!gpu || GPU_DEVICE_SOLVE != 1 || (LINEAR_EQUATION_SOLVER_TYPE != 0 && LINEAR_EQUATION_SOLVER_TYPE != 1) || USE_PARAM_START != 0
			|| (USE_PARAM_END != 0 && USE_PARAM_END != N_PARAM - 1))
	{
		return false;
	}
	gpu->SetLinearEquationSolverType(LINEAR_EQUATION_SOLVER_TYPE);
	doNotAcceptStep = gpu->CalculateNextParametersEuler(dt, uR, uI, phiR, phiI, IMAGINARY_TIME, USE_PRECONDITIONING, sys->GetTime() + dt, nullptr, nullptr) || doNotAcceptStep;
	return true;
}

// sys->GetExponent() for NormalizeWavefunction (src/TDVMC.cpp:3763)
double GpuExponentOr(double cpuExponent)
{
	return gpu ? gpu->GetExponent() : cpuExponent;
}

// end of run: the coordinates the driver writes for the next run (src/TDVMC.cpp:4018-4022) are the first local walker's
void GpuCopyWalkerToDriver(vector<vector<double> >& R)
{
	if (!gpu)
	{
		return;
	}
	vector<vector<vector<double> > > Rw;
	gpu->GetPositions(Rw);
	for (int i = 0; i < N; i++)
	{
		for (int a = 0; a < DIM; a++)
		{
			R[i][a] = Rw[0][i][a];
		}
	}
}

void GpuShutdown()
{
	delete gpu;
	gpu = nullptr;
}

#endif // TDVMC_GPU_HOOKS_IMPL
