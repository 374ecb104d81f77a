// GpuEnsembleSystem — C++ host adapter between the reference's driver and libtdvmc_b200.so.
//
// The reference driver (src/TDVMC.cpp) keeps its seven estimator globals (:147-153) and calls
//     ParallelUpdateExpectationValues(R, uR, uI, phiR, phiI)                  :1152-1220
//     ParallelUpdateExpectationValuesForGivenSamples(samples, uR, uI, ...)    :1305-1330
//     MoveCoordinatesToFirstCell(R)                                           :787-796
//     sys->GetExponent()  (NormalizeWavefunction)                             :3763
// once per estimator evaluation.  This class offers the same calls on std::vector arguments (the
// reference's own types) and forwards them to the C ABI of include/tdvmc_gpu.h; INTEGRATION.md shows
// the few lines that re-point the driver.  It owns no physics: knots, spline table and boundary map
// come from the reference's own InitSystem() results (or from SystemTables.h for the two covered
// systems when the caller has only the config values).
//
// One instance per process / GPU.  Errors throw std::runtime_error carrying tdvmc_gpu_last_error().
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/tdvmc_gpu.h"

namespace tdvmc_host
{

// Plain-data copy of what IPhysicalSystem::InitSystem() computed (BosonsBulk.cpp:49-156).
struct SystemTables
{
    int n_particles = 0;
    int n_params = 0;
    int pair_rule = TDVMC_PAIR_RULE_CUT;
    int tail_param = 0;
    int n_other = 9;
    double lbox = 0.0;
    double hbar2_2m = 1.0;                 // HBAR2_2M, src/Constants.h:12
    std::vector<double> knots;             // nodes
    std::vector<double> spline_weights;    // splineWeights flattened [K][4][4]
    std::vector<int32_t> map_ptr, map_col; // boundary-condition map, CSR over parameters
    std::vector<double> map_val;
    std::vector<double> system_params;     // SYSTEM_PARAMS
    int system_kind = TDVMC_SYSTEM_SPLINE_TABLE;
    int n_ext = 0;                         // 0: number of splines
    std::vector<double> map_const, grad_const; // empty: zeros
    // BosonMixtureCluster only (system_kind == TDVMC_SYSTEM_MIXTURE)
    int n_pair_types = 0;
    int spline_order = 3;                  // 4: BosonMixtureCluster_4thorder
    int dim = 3;                           // 1: InhContactBosons (positions still [walker][particle][3], coordinate first)
    int n_splines_first = 0;               // InhContactBosons: splines of the single-particle function
    std::vector<int32_t> pair_type;        // correlationTypes flattened [N][N]
    std::vector<int32_t> pair_potential;   // [T] 0 HFDB_He_He, 1 KTTY_He_Na, 2 KTTY_He_Cs
    std::vector<double> hbar_over_2m, mass; // [N]
    std::vector<double> type_knots, type_weights, type_mcmillan;
};

// Flattens the reference's vector<vector<vector<double>>> splineWeights (SplineFactory::GetWeights3).
std::vector<double> FlattenWeights(const std::vector<std::vector<std::vector<double> > >& w);

// BosonsBulk (BosonsBulk.cpp:61-106, 158-177): knots/weights as computed by the reference.
SystemTables MakeBosonsBulkTables(int N, double LBOX, int N_PARAM, const std::vector<double>& nodes,
                                  const std::vector<std::vector<std::vector<double> > >& splineWeights,
                                  const std::vector<double>& SYSTEM_PARAMS);
// NUBosonsBulkPB (NUBosonsBulkPB.cpp:53-135, 219-232).
SystemTables MakeNUBosonsBulkPBTables(int N, double LBOX, int N_PARAM, const std::vector<double>& nodes,
                                      const std::vector<std::vector<std::vector<double> > >& splineWeights,
                                      const std::vector<double>& SYSTEM_PARAMS, int grBinCount);
// NUBosonsBulkPBBoxAndRadial (NUBosonsBulkPBBoxAndRadial.cpp:64-191, 193-211): nodes == nodesRad and
// splineWeights == splineWeightsRad after SetNodes (:36-62), so one knot vector and one table are passed.
SystemTables MakeNUBosonsBulkPBBoxAndRadialTables(int N, double LBOX, int N_PARAM, const std::vector<double>& nodes,
                                                  const std::vector<std::vector<std::vector<double> > >& splineWeights,
                                                  const std::vector<double>& SYSTEM_PARAMS, int grBinCount);
// One WFParts::SplinedFunction of InhContactBosons after InitSystem() (WFParts/SplinedFunction.h:11-33).
struct SplinedFunctionTables
{
    std::vector<double> nodes;
    std::vector<std::vector<std::vector<double> > > splineWeights;
    std::vector<std::vector<double> > bcFactorsStart, bcFactorsEnd;
    int np1 = 0, np2 = 0, np3 = 0;
};
// InhContactBosons (InhContactBosons.cpp:64-247), one-dimensional: spf = single-particle function, pc = pair correlation.
SystemTables MakeInhContactBosonsTables(int N, double LBOX, int N_PARAM, const std::vector<double>& SYSTEM_PARAMS,
                                        const SplinedFunctionTables& spf, const SplinedFunctionTables& pc);

// HeBulk (HeBulk.cpp:40-70, 376-383): everything follows from N, LBOX and N_PARAM.
SystemTables MakeHeBulkTables(int N, double LBOX, int N_PARAM);
// HeDrop (HeDrop.cpp:71-135, 609-626): open boundary; everything follows from N and N_PARAM.
SystemTables MakeHeDropTables(int N, int N_PARAM);
// One pair type of BosonMixtureCluster: the reference's CorrelationFunctionData after InitSystem().
struct MixturePairType
{
    std::vector<double> nodes;                                     // cfd.nodes
    std::vector<std::vector<std::vector<double> > > splineWeights; // cfd.splineWeights
    std::vector<std::vector<double> > bcFactors;                   // cfd.bcFactors (5 x 2; 5 x 3 for the 4th-order class)
    double mcMillanFactor;
    int potential;                                                 // 0 HFDB_He_He, 1 KTTY_He_Na, 2 KTTY_He_Cs
};
// BosonMixtureCluster (BosonMixtureCluster.cpp:58-346, 636-645); correlationTypes [N][N], hbarOver2m/mass per particle.
SystemTables MakeBosonMixtureClusterTables(int N, const std::vector<std::vector<int> >& correlationTypes,
                                           const std::vector<double>& hbarOver2mPerParticle,
                                           const std::vector<double>& massPerParticle,
                                           const std::vector<MixturePairType>& pairTypes, int numOfOtherExpectationValues,
                                           int splineOrder = 3); // 4: BosonMixtureCluster_4thorder (bcFactors 5 x 3)

// g(r) / S(k) description: what BosonsBulk::InitSystem builds (BosonsBulk.cpp:124-153).
struct ObservableTables
{
    int grCount = 0;                 // pairDistribution.grid.count
    double grSpacing = 0, grMax = 0; // pairDistribution.grid.spacing / .max
    double grWeight = 1;             // DIM/(N-1) (BosonsBulk.cpp:481) or 1 (NUBosonsBulkPB.cpp:611)
    std::vector<double> grScaling;   // pairDistribution.scalingGrid
    std::vector<std::vector<std::vector<double> > > kValues; // [numOfkValues][kn][3], already times 2 pi / LBOX
};
// Grid::Init + InitScaling for the pair distribution (Grid.cpp:16-31, ObservableVsOnGridWithScaling.cpp:19-44)
ObservableTables MakePairDistributionGrid(double rMax, int numOfPairDistributionValues, double weight);

struct AdditionalObservables
{
    std::vector<double> pairDistribution, structureFactor;
};

// Histogram grids of the three-particle cluster observables (BosonMixtureCluster.cpp:327-340): {count, spacing, max}.
struct ClusterObservableTables
{
    int angleCount = 180, densityCount = 800, distanceCount = 160;
    double angleSpacing = 1.0, densitySpacing = 0.1, densityMax = 80.0, distanceSpacing = 0.5, distanceMax = 80.0;
    std::vector<double> densityScaling; // densityFromCOM.scalingGrid
};
struct ClusterObservables
{
    double r2 = 0;
    std::vector<double> angularDistribution, densityFromCOM, particleDistances; // [3][count] each
};

// The seven estimator arrays under the reference's global names (src/TDVMC.cpp:147-153).
struct Estimators
{
    std::vector<double> localOperators;
    double localEnergyR = 0.0;
    double localEnergyI = 0.0;
    std::vector<std::vector<double> > localOperatorsMatrix;
    std::vector<double> localOperatorlocalEnergyR;
    std::vector<double> localOperatorlocalEnergyI;
    std::vector<double> otherExpectationValues;
    long long nAcceptances = 0;
    long long nTrials = 0;
    long long nSamples = 0;
};

class GpuEnsembleSystem
{
public:
    // walkersTotal walkers are split evenly over numOfProcesses ranks (processRank owns a contiguous block);
    // MC_NSTEPS is the per-walker sample capacity, UPDATE_SAMPLES_EVERY_NTH_STEP > 0 keeps the sample positions.
    GpuEnsembleSystem(const SystemTables& tables, int walkersTotal, double MC_STEP, int MC_NSTEPS,
                      int UPDATE_SAMPLES_EVERY_NTH_STEP, unsigned long long seed, int processRank, int numOfProcesses,
                      int device);
    ~GpuEnsembleSystem();
    GpuEnsembleSystem(const GpuEnsembleSystem&) = delete;
    GpuEnsembleSystem& operator=(const GpuEnsembleSystem&) = delete;

    int LocalWalkers() const { return nLocal; }
    int FirstWalker() const { return firstWalker; }

    // rank 0 creates the id, the driver broadcasts it (MPI_Bcast of 128 bytes), every rank joins
    static std::vector<unsigned char> CreateCommunicatorId();
    void JoinCommunicator(const std::vector<unsigned char>& id);

    // R[w][n][a] for the local walkers
    void SetPositions(const std::vector<std::vector<std::vector<double> > >& R);
    void GetPositions(std::vector<std::vector<std::vector<double> > >& R);
    void MoveCoordinatesToFirstCell();
    // nAcceptances = 0; nTrials = 0 at the start of a time step (src/TDVMC.cpp:3428-3429)
    void ResetCounters();
    // MC_STEP changed at run time (./param file, src/TDVMC.cpp:2718-2744)
    void SetMCStep(double MC_STEP);
    void DoMetropolisSteps(long long n, const std::vector<double>& uR, const std::vector<double>& uI, double phiR,
                           double phiI);

    Estimators ParallelUpdateExpectationValues(const std::vector<double>& uR, const std::vector<double>& uI, double phiR,
                                               double phiI, int MC_NSTEPS, int MC_NTHERMSTEPS,
                                               int MC_NINITIALIZATIONSTEPS, double time);
    Estimators ParallelUpdateExpectationValuesForGivenSamples(const std::vector<double>& uR, const std::vector<double>& uI,
                                                              double phiR, double phiI, double time);
    // The same pass for BosonMixtureCluster (BosonMixtureCluster.cpp:680-741).
    ClusterObservables ParallelCalculateAdditionalSystemPropertiesCluster(const std::vector<double>& uR, const std::vector<double>& uI,
                                                                          double phiR, double phiI, const ClusterObservableTables& obs,
                                                                          int MC_NADDITIONALSTEPS, int MC_NADDITIONALTHERMSTEPS,
                                                                          int MC_NADDITIONALINITIALIZATIONSTEPS, double time);
    // UpdateSamplesConsecutive (src/TDVMC.cpp:975-983): the next nrOfSamplesToUpdate stored samples of every walker
    // advance by MC_NTHERMSTEPS steps at the given parameters; call ParallelUpdateExpectationValuesForGivenSamples next.
    void UpdateSamplesConsecutive(int nrOfSamplesToUpdate, const std::vector<double>& uR, const std::vector<double>& uI,
                                  double phiR, double phiI, int MC_NTHERMSTEPS, double time);
    double GetExponent();

    // ---- parameter derivatives on the device ----
    // LINEAR_EQUATION_SOLVER_TYPE of the calls below: 0 the hand-written Cholesky (default), 1 Eigen's FullPivHouseholderQR
    // with the mean subtraction of src/TDVMC.cpp:1800-1809
    void SetLinearEquationSolverType(int LINEAR_EQUATION_SOLVER_TYPE) { solverType = LINEAR_EQUATION_SOLVER_TYPE; }
    // ParallelUpdateExpectationValues without the fetch: the estimator sums stay in HBM for the two calls below.
    void SampleExpectationValues(const std::vector<double>& uR, const std::vector<double>& uI, double phiR, double phiI,
                                 int MC_NSTEPS, int MC_NTHERMSTEPS, int MC_NINITIALIZATIONSTEPS, double time);
    // SolveForParametersDot, LINEAR_EQUATION_SOLVER_TYPE = 0 (src/TDVMC.cpp:1713-1763), on the all-reduced estimators;
    // returns true if the matrix was not positive definite (the reference's doNotAcceptStep, :1577-1589).
    bool SolveForParametersDot(std::vector<double>& uDotR, std::vector<double>& uDotI, double* phiDotR, double* phiDotI,
                               int IMAGINARY_TIME, int USE_PRECONDITIONING);
    // CalculateNextParametersEuler (src/TDVMC.cpp:1834-1853) followed by BroadcastNewParameters (:506-512): every rank
    // ends with the same new parameters, current on its device.  localEnergyR/I (may be null) receive <E> of the step.
    bool CalculateNextParametersEuler(double dt, std::vector<double>& uR, std::vector<double>& uI, double* phiR, double* phiI,
                                      int IMAGINARY_TIME, int USE_PRECONDITIONING, double time, double* localEnergyR,
                                      double* localEnergyI);
    // The explicit multi-stage integrators composed from device stages - fresh sampling (MC counts given) or re-evaluation of
    // the stored samples (ReuseSamples), each followed by the device solve; no estimator is fetched.  All start from the
    // estimators of (uR, uI) already accumulated on the device and leave the new parameters current there.  Return value:
    // some stage's matrix was not positive definite.
    //   CalculateNextParametersPC              src/TDVMC.cpp:1855-1910     CalculateNextParametersRK4              :1969-2035
    //   CalculateNextParametersPCReuseSamples  :1912-1967                  CalculateNextParametersRK4ReuseSamples  :2037-2103
    bool CalculateNextParametersPC(double dt, std::vector<double>& uR, std::vector<double>& uI, double* phiR, double* phiI,
                                   int MC_NSTEPS, int MC_NTHERMSTEPS, int MC_NINITIALIZATIONSTEPS, int IMAGINARY_TIME,
                                   int USE_PRECONDITIONING, double time);
    bool CalculateNextParametersPCReuseSamples(double dt, std::vector<double>& uR, std::vector<double>& uI, double* phiR,
                                               double* phiI, int IMAGINARY_TIME, int USE_PRECONDITIONING, double time);
    bool CalculateNextParametersRK4(double dt, std::vector<double>& uR, std::vector<double>& uI, double* phiR, double* phiI,
                                    int MC_NSTEPS, int MC_NTHERMSTEPS, int MC_NINITIALIZATIONSTEPS, int IMAGINARY_TIME,
                                    int USE_PRECONDITIONING, double time);
    bool CalculateNextParametersRK4ReuseSamples(double dt, std::vector<double>& uR, std::vector<double>& uI, double* phiR,
                                                double* phiI, int IMAGINARY_TIME, int USE_PRECONDITIONING, double time);
    // ParallelCalculateAdditionalSystemProperties (src/TDVMC.cpp:1438-1444) for the bulk spline systems: the mean
    // pairDistribution / structureFactor values (additionalObservablesMean.observables[0], [1]) over samples, walkers, ranks.
    AdditionalObservables ParallelCalculateAdditionalSystemProperties(const std::vector<double>& uR, const std::vector<double>& uI,
                                                                      double phiR, double phiI, const ObservableTables& obs,
                                                                      int MC_NADDITIONALSTEPS, int MC_NADDITIONALTHERMSTEPS,
                                                                      int MC_NADDITIONALINITIALIZATIONSTEPS, double time);

private:
    void Check(int rc, const char* what);
    Estimators Fetch();
    struct Dot
    {
        std::vector<double> uR, uI;
        double phiR = 0, phiI = 0;
        bool notPD = false;
    };
    Dot Stage(const std::vector<double>& uR, const std::vector<double>& uI, double phiR, double phiI, const int* mcCounts,
              int IMAGINARY_TIME, int USE_PRECONDITIONING, double time);
    Dot SolveNow(int IMAGINARY_TIME, int USE_PRECONDITIONING);
    bool PredictorCorrector(double dt, std::vector<double>& uR, std::vector<double>& uI, double* phiR, double* phiI, int pcSteps,
                            const int* mcCounts, int IMAGINARY_TIME, int USE_PRECONDITIONING, double time);
    bool RungeKutta4(double dt, std::vector<double>& uR, std::vector<double>& uI, double* phiR, double* phiI, const int* mcCounts,
                     int IMAGINARY_TIME, int USE_PRECONDITIONING, double time);

    tdvmc_gpu_handle* handle = nullptr;
    int N = 0, P = 0, nOther = 9;
    int nLocal = 0, firstWalker = 0, rank = 0, world = 1;
    int solverType = 0;
    std::vector<double> flat;
};

} // namespace tdvmc_host
