#include "GpuEnsembleSystem.h"

#include <stdexcept>

namespace tdvmc_host
{

std::vector<double> FlattenWeights(const std::vector<std::vector<std::vector<double> > >& w)
{
    std::vector<double> out;
    out.reserve(w.size() * 16);
    for (const auto& spline : w)
        for (const auto& part : spline)
            for (double c : part) out.push_back(c);
    return out;
}

namespace
{
void PushRow(SystemTables& t, std::initializer_list<std::pair<int, double> > row)
{
    for (const auto& e : row)
    {
        t.map_col.push_back(e.first);
        t.map_val.push_back(e.second);
    }
    t.map_ptr.push_back((int32_t)t.map_col.size());
}
} // namespace

SystemTables MakeBosonsBulkTables(int N, double LBOX, int N_PARAM, const std::vector<double>& nodes,
                                  const std::vector<std::vector<std::vector<double> > >& splineWeights,
                                  const std::vector<double>& SYSTEM_PARAMS)
{
    SystemTables t;
    t.n_particles = N;
    t.n_params = N_PARAM;
    t.lbox = LBOX;
    t.pair_rule = TDVMC_PAIR_RULE_CUT;
    t.tail_param = N_PARAM - 1; // BosonsBulk.cpp:532-534
    t.n_other = 9;              // BosonsBulk.cpp:55
    t.knots = nodes;
    t.spline_weights = FlattenWeights(splineWeights);
    t.system_params = SYSTEM_PARAMS;
    const int K = (int)nodes.size() - 4;
    if (N_PARAM != K - 2) throw std::runtime_error("BosonsBulk: N_PARAM must equal numberOfSplines - 2 (BosonsBulk.cpp:85-91)");
    // RefreshLocalOperators with the uniform factor tables of SetBoundaryConditions3_1D_OR_2 / _CO_2
    // (BosonsBulk.cpp:158-177, SplineFactory.cpp:428-438, 568-578)
    t.map_ptr.push_back(0);
    PushRow(t, { { 1, 1.0 } });
    PushRow(t, { { 0, 1.0 }, { 2, 1.0 } });
    for (int i = 2; i < N_PARAM - 2; i++) PushRow(t, { { i + 1, 1.0 } });
    PushRow(t, { { K - 3, 1.0 }, { K - 1, 1.0 } });
    PushRow(t, { { K - 2, 1.0 } });
    return t;
}

SystemTables MakeNUBosonsBulkPBTables(int N, double LBOX, int N_PARAM, const std::vector<double>& nodes,
                                      const std::vector<std::vector<std::vector<double> > >& splineWeights,
                                      const std::vector<double>& SYSTEM_PARAMS, int grBinCount)
{
    SystemTables t;
    t.n_particles = N;
    t.n_params = N_PARAM;
    t.lbox = LBOX;
    t.pair_rule = TDVMC_PAIR_RULE_REFLECT; // NUBosonsBulkPB.cpp:249-269
    t.tail_param = N_PARAM - 1;            // :651-653
    t.n_other = 9 + (grBinCount == 0 ? 400 : grBinCount); // :56-61
    t.knots = nodes;
    t.spline_weights = FlattenWeights(splineWeights);
    t.system_params = SYSTEM_PARAMS;
    const int K = (int)nodes.size() - 4;
    if (K != N_PARAM + 3) throw std::runtime_error("NUBosonsBulkPB: numberOfSplines must equal N_PARAM + 3 (NUBosonsBulkPB.cpp:71)");
    // RefreshLocalOperators (NUBosonsBulkPB.cpp:219-232)
    t.map_ptr.push_back(0);
    for (int i = 0; i < N_PARAM; i++)
    {
        if (i == 1) PushRow(t, { { 2, 1.0 }, { 0, 1.0 } });
        else if (i == N_PARAM - 1) PushRow(t, { { N_PARAM, 1.0 }, { K - 2, 1.0 }, { K - 1, 1.0 } });
        else PushRow(t, { { i + 1, 1.0 } });
    }
    return t;
}

GpuEnsembleSystem::GpuEnsembleSystem(const SystemTables& tb, int walkersTotal, double MC_STEP, int MC_NSTEPS,
                                     int UPDATE_SAMPLES_EVERY_NTH_STEP, unsigned long long seed, int processRank,
                                     int numOfProcesses, int device)
{
    N = tb.n_particles;
    P = tb.n_params;
    nOther = tb.n_other;
    rank = processRank;
    world = numOfProcesses;
    const int base = walkersTotal / world, rem = walkersTotal % world;
    firstWalker = rank * base + (rank < rem ? rank : rem);
    nLocal = base + (rank < rem ? 1 : 0);

    tdvmc_system_desc sd;
    sd.struct_size = sizeof(sd);
    sd.n_particles = tb.n_particles;
    sd.dim = 3;
    sd.n_params = tb.n_params;
    sd.n_splines = (int32_t)tb.knots.size() - 4;
    sd.pair_rule = tb.pair_rule;
    sd.tail_param = tb.tail_param;
    sd.n_other = tb.n_other;
    sd.lbox = tb.lbox;
    sd.hbar2_2m = tb.hbar2_2m;
    sd.knots = tb.knots.data();
    sd.spline_weights = tb.spline_weights.data();
    sd.map_ptr = tb.map_ptr.data();
    sd.map_col = tb.map_col.data();
    sd.map_val = tb.map_val.data();
    sd.system_params = tb.system_params.data();
    sd.n_system_params = (int32_t)tb.system_params.size();
    sd.system_kind = tb.system_kind;
    sd.n_ext = tb.n_ext > 0 ? tb.n_ext : sd.n_splines;
    sd.reserved = 0;
    sd.map_const = tb.map_const.empty() ? nullptr : tb.map_const.data();
    sd.grad_const = tb.grad_const.empty() ? nullptr : tb.grad_const.data();
    sd.mixture = nullptr;
    tdvmc_ensemble_desc ed;
    ed.struct_size = sizeof(ed);
    ed.device = device;
    ed.n_walkers = nLocal;
    ed.first_walker = firstWalker;
    ed.max_samples_per_walker = MC_NSTEPS;
    ed.keep_sample_positions = UPDATE_SAMPLES_EVERY_NTH_STEP > 0 ? 1 : 0;
    ed.seed = seed;
    ed.mc_step = MC_STEP;
    int rc = tdvmc_gpu_create(&sd, &ed, &handle);
    if (rc != 0) throw std::runtime_error(std::string("tdvmc_gpu_create: ") + tdvmc_gpu_last_error(nullptr));
}

GpuEnsembleSystem::~GpuEnsembleSystem()
{
    if (handle) tdvmc_gpu_destroy(handle);
}

void GpuEnsembleSystem::Check(int rc, const char* what)
{
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + tdvmc_gpu_last_error(handle));
}

std::vector<unsigned char> GpuEnsembleSystem::CreateCommunicatorId()
{
    std::vector<unsigned char> id(TDVMC_GPU_UNIQUE_ID_BYTES);
    if (tdvmc_gpu_comm_unique_id(id.data()) != 0) throw std::runtime_error(std::string("comm_unique_id: ") + tdvmc_gpu_last_error(nullptr));
    return id;
}

void GpuEnsembleSystem::JoinCommunicator(const std::vector<unsigned char>& id)
{
    Check(tdvmc_gpu_comm_init(handle, id.data(), rank, world), "comm_init");
}

void GpuEnsembleSystem::SetPositions(const std::vector<std::vector<std::vector<double> > >& R)
{
    flat.resize((size_t)nLocal * N * 3);
    for (int w = 0; w < nLocal; w++)
        for (int n = 0; n < N; n++)
            for (int a = 0; a < 3; a++) flat[((size_t)w * N + n) * 3 + a] = R[w][n][a];
    Check(tdvmc_gpu_set_positions(handle, flat.data(), 0, nLocal), "set_positions");
}

void GpuEnsembleSystem::GetPositions(std::vector<std::vector<std::vector<double> > >& R)
{
    flat.resize((size_t)nLocal * N * 3);
    Check(tdvmc_gpu_get_positions(handle, flat.data(), 0, nLocal), "get_positions");
    R.assign(nLocal, std::vector<std::vector<double> >(N, std::vector<double>(3)));
    for (int w = 0; w < nLocal; w++)
        for (int n = 0; n < N; n++)
            for (int a = 0; a < 3; a++) R[w][n][a] = flat[((size_t)w * N + n) * 3 + a];
}

void GpuEnsembleSystem::MoveCoordinatesToFirstCell() { Check(tdvmc_gpu_wrap_positions(handle), "wrap_positions"); }

void GpuEnsembleSystem::DoMetropolisSteps(long long n, const std::vector<double>& uR, const std::vector<double>& uI,
                                          double phiR, double phiI)
{
    Check(tdvmc_gpu_set_params(handle, uR.data(), uI.data(), phiR, phiI, 0.0), "set_params");
    Check(tdvmc_gpu_sweep(handle, n), "sweep");
}

Estimators GpuEnsembleSystem::Fetch()
{
    Estimators e;
    e.localOperators.resize(P);
    e.localOperatorlocalEnergyR.resize(P);
    e.localOperatorlocalEnergyI.resize(P);
    e.otherExpectationValues.resize(nOther);
    std::vector<double> S((size_t)P * P);
    tdvmc_estimators out;
    out.local_operators = e.localOperators.data();
    out.local_energy_r = &e.localEnergyR;
    out.local_energy_i = &e.localEnergyI;
    out.local_operators_matrix = S.data();
    out.local_operator_energy_r = e.localOperatorlocalEnergyR.data();
    out.local_operator_energy_i = e.localOperatorlocalEnergyI.data();
    out.other_expectation_values = e.otherExpectationValues.data();
    out.n_acceptances = out.n_trials = out.n_samples = 0;
    Check(tdvmc_gpu_allreduce_and_fetch(handle, &out), "allreduce_and_fetch");
    e.localOperatorsMatrix.assign(P, std::vector<double>(P));
    for (int k = 0; k < P; k++)
        for (int j = 0; j < P; j++) e.localOperatorsMatrix[k][j] = S[(size_t)k * P + j];
    e.nAcceptances = out.n_acceptances;
    e.nTrials = out.n_trials;
    e.nSamples = out.n_samples;
    return e;
}

Estimators GpuEnsembleSystem::ParallelUpdateExpectationValues(const std::vector<double>& uR, const std::vector<double>& uI,
                                                              double phiR, double phiI, int MC_NSTEPS, int MC_NTHERMSTEPS,
                                                              int MC_NINITIALIZATIONSTEPS, double time)
{
    Check(tdvmc_gpu_set_params(handle, uR.data(), uI.data(), phiR, phiI, time), "set_params");
    Check(tdvmc_gpu_sample_and_accumulate(handle, MC_NSTEPS, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS), "sample_and_accumulate");
    return Fetch();
}

Estimators GpuEnsembleSystem::ParallelUpdateExpectationValuesForGivenSamples(const std::vector<double>& uR,
                                                                             const std::vector<double>& uI, double phiR,
                                                                             double phiI, double time)
{
    Check(tdvmc_gpu_set_params(handle, uR.data(), uI.data(), phiR, phiI, time), "set_params");
    Check(tdvmc_gpu_reevaluate_stored(handle), "reevaluate_stored");
    return Fetch();
}

double GpuEnsembleSystem::GetExponent()
{
    double x = 0.0;
    Check(tdvmc_gpu_last_exponent(handle, &x), "last_exponent");
    return x;
}

} // namespace tdvmc_host
