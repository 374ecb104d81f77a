#include "GpuEnsembleSystem.h"

#include <cmath>
#include <cstring>
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
void PushRow(SystemTables& t, const std::vector<std::pair<int, double> >& row)
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

SystemTables MakeNUBosonsBulkPBBoxAndRadialTables(int N, double LBOX, int N_PARAM, const std::vector<double>& nodes,
                                                  const std::vector<std::vector<std::vector<double> > >& splineWeights,
                                                  const std::vector<double>& SYSTEM_PARAMS, int grBinCount)
{
    SystemTables t;
    t.system_kind = TDVMC_SYSTEM_BOX_RADIAL;
    t.n_particles = N;
    t.n_params = N_PARAM;
    t.lbox = LBOX;
    t.tail_param = -1;
    t.n_other = 3 + (grBinCount == 0 ? 400 : grBinCount); // NUBosonsBulkPBBoxAndRadial.cpp:72-78
    t.knots = nodes;
    t.spline_weights = FlattenWeights(splineWeights);
    t.system_params = SYSTEM_PARAMS;
    const int K = (int)nodes.size() - 4, PR = N_PARAM / 2;
    if (N_PARAM % 2 || K != PR + 3)
        throw std::runtime_error("NUBosonsBulkPBBoxAndRadial: numberOfSplines must equal N_PARAM / 2 + 3 (NUBosonsBulkPBBoxAndRadial.cpp:84-89)");
    t.n_ext = 2 * K; // [splineSumsRad | splineSums]
    // RefreshLocalOperators (:193-211)
    t.map_ptr.push_back(0);
    for (int i = 0; i < PR; i++)
    {
        if (i == 1) PushRow(t, { { 2, 1.0 }, { 0, 1.0 } });
        else if (i == PR - 1) PushRow(t, { { PR, 1.0 }, { K - 2, 1.0 / (-2.0) }, { K - 1, 1.0 } });
        else PushRow(t, { { i + 1, 1.0 } });
    }
    for (int i = 0; i < PR; i++)
    {
        if (i == 1) PushRow(t, { { K + 2, 1.0 }, { K, 1.0 } });
        else if (i == PR - 1) PushRow(t, { { K + PR, 1.0 }, { K + K - 2, 1.0 }, { K + K - 1, 1.0 } });
        else PushRow(t, { { K + i + 1, 1.0 } });
    }
    return t;
}

SystemTables MakeInhContactBosonsTables(int N, double LBOX, int N_PARAM, const std::vector<double>& SYSTEM_PARAMS,
                                        const SplinedFunctionTables& spf, const SplinedFunctionTables& pc)
{
    if (SYSTEM_PARAMS.size() != 4) throw std::runtime_error("InhContactBosons: the four-entry SYSTEM_PARAMS {range, strength, k, V0} is supported");
    const int K1 = (int)spf.nodes.size() - 4, K2 = (int)pc.nodes.size() - 4;
    if (N_PARAM != spf.np3 + pc.np3) throw std::runtime_error("InhContactBosons: wrong number of parameters (InhContactBosons.cpp:122-130)");
    SystemTables t;
    t.system_kind = TDVMC_SYSTEM_INH_CONTACT;
    t.dim = 1;
    t.n_particles = N;
    t.n_params = N_PARAM;
    t.lbox = LBOX;
    t.tail_param = -1;
    t.n_other = 9;
    t.n_splines_first = K1;
    t.n_ext = K1 + K2;
    t.system_params = SYSTEM_PARAMS;
    t.knots = spf.nodes;
    t.knots.insert(t.knots.end(), pc.nodes.begin(), pc.nodes.end());
    t.spline_weights = FlattenWeights(spf.splineWeights);
    const std::vector<double> w2 = FlattenWeights(pc.splineWeights);
    t.spline_weights.insert(t.spline_weights.end(), w2.begin(), w2.end());
    // RefreshLocalOperators (InhContactBosons.cpp:208-247)
    t.map_ptr.push_back(0);
    for (int i = 0; i < spf.np1; i++)
    {
        std::vector<std::pair<int, double> > e;
        for (int j = 0; j < 3; j++) e.push_back({ j, spf.bcFactorsStart[i][j] });
        for (int j = 0; j < 3; j++) e.push_back({ K1 - 3 + j, spf.bcFactorsEnd[i][j] });
        PushRow(t, e);
    }
    for (int i = spf.np1; i < spf.np2; i++) PushRow(t, { { 3 + (i - spf.np1), 1.0 } });
    for (int i = 0; i < pc.np1; i++)
    {
        std::vector<std::pair<int, double> > e;
        for (int j = 0; j < 3; j++) e.push_back({ K1 + j, pc.bcFactorsStart[i][j] });
        PushRow(t, e);
    }
    for (int i = pc.np1; i < pc.np2; i++) PushRow(t, { { K1 + 3 + (i - pc.np1), 1.0 } });
    for (int i = 0; i < pc.np3 - pc.np2; i++)
    {
        std::vector<std::pair<int, double> > e;
        for (int j = 0; j < 3; j++) e.push_back({ K1 + K2 - 3 + j, pc.bcFactorsEnd[i][j] });
        PushRow(t, e);
    }
    return t;
}

SystemTables MakeHeBulkTables(int N, double LBOX, int N_PARAM)
{
    SystemTables t;
    t.system_kind = TDVMC_SYSTEM_HE_BULK;
    t.n_particles = N;
    t.n_params = N_PARAM;
    t.lbox = LBOX;
    t.tail_param = -1;
    t.n_other = 3 + 100; // HeBulk.cpp:42-45
    const int K = N_PARAM - 1 + 3 + 3;
    const double rs = 1.95, h = (LBOX / 2.0 - rs) / (double)(K - 3.0);
    t.knots.assign(K + 4, 0.0); // unused by the He kernels (grid is (rs, h))
    t.n_ext = K + 3;
    const int MC = K;
    t.map_ptr.push_back(0);
    PushRow(t, { { MC, 1.0 }, { 0, 10.0 * h / std::pow(rs, 6.0) }, { 1, (-5.0 * h + 3.0 * rs) / (2.0 * std::pow(rs, 6.0)) } });
    PushRow(t, { { 2, 1.0 }, { 0, 1.0 }, { 1, -1.0 / 2.0 } });
    for (int i = 2; i < N_PARAM - 2; i++) PushRow(t, { { i + 1, 1.0 } });
    PushRow(t, { { K - 6, 1.0 }, { K - 5, -1.0 / 2.0 }, { K - 4, 1.0 } });
    PushRow(t, { { K - 5, -3.0 / 2.0 }, { K - 4, 0.0 } });
    t.map_const.assign(N_PARAM, 0.0);
    t.grad_const.assign(N_PARAM, 0.0);
    t.map_const[N_PARAM - 1] = 1.0;  // HeBulk.cpp:383
    t.grad_const[N_PARAM - 1] = 1.0; // HeBulk.cpp:351
    return t;
}

SystemTables MakeHeDropTables(int N, int N_PARAM)
{
    SystemTables t;
    t.system_kind = TDVMC_SYSTEM_HE_DROP;
    t.n_particles = N;
    t.n_params = N_PARAM;
    t.lbox = 0.0;
    t.tail_param = -1;
    t.n_other = 3 + 200 + 200; // HeDrop.cpp:75-79
    const double m = -4.7, rs = 3.0, hS = 0.1, hL = 0.5;
    const int nS = 70, K = N_PARAM + 1 + 2, nL = K - nS, P = N_PARAM;
    const double r2 = hS * (nS - 3.0) + rs, rt = hL * (nL - 3.0) + r2, d = 1.0 / (hS + hL);
    t.knots.assign(K + 4, 0.0);
    t.n_ext = K + 3;
    const int MC = K, CO = K + 1, LI = K + 2;
    t.map_ptr.push_back(0);
    PushRow(t, { { MC, 1.0 }, { 0, -2.0 * m * hS * std::pow(rs, m - 1.0) }, { 1, (m * hS + 3.0 * rs) * std::pow(rs, m - 1.0) / 2.0 } });
    PushRow(t, { { 2, 1.0 }, { 0, 1.0 }, { 1, -1.0 / 2.0 } });
    for (int i = 2; i < nS - 4; i++) PushRow(t, { { i + 1, 1.0 } });
    PushRow(t, { { nS - 3, 1.0 }, { nS - 1, (-hS + hL) * d }, { nS, (2.0 * hL) * d } });
    PushRow(t, { { nS - 2, 1.0 }, { nS - 1, (-4.0 * hS) * d }, { nS, (4.0 * hL) * d } });
    PushRow(t, { { nS + 1, 1.0 }, { nS - 1, (4.0 * hS) * d }, { nS, (-4.0 * hL) * d } });
    PushRow(t, { { nS + 2, 1.0 }, { nS - 1, (2.0 * hS) * d }, { nS, (hS - hL) * d } });
    for (int i = nS; i < P - 3; i++) PushRow(t, { { i + 3, 1.0 } });
    PushRow(t, { { K - 3, 1.0 }, { K - 2, -1.0 / 2.0 }, { K - 1, 1.0 } });
    PushRow(t, { { CO, 1.0 }, { K - 2, 3.0 / 2.0 }, { K - 1, 0.0 } });
    PushRow(t, { { LI, 1.0 }, { K - 2, 3.0 / 2.0 * rt - hL / 2.0 }, { K - 1, 2.0 * hL } });
    return t;
}

SystemTables MakeBosonMixtureClusterTables(int N, const std::vector<std::vector<int> >& correlationTypes,
                                           const std::vector<double>& hbarOver2mPerParticle,
                                           const std::vector<double>& massPerParticle,
                                           const std::vector<MixturePairType>& pairTypes, int numOfOtherExpectationValues,
                                           int splineOrder)
{
    if (splineOrder != 3 && splineOrder != 4) throw std::runtime_error("BosonMixtureCluster: spline order 3 or 4");
    SystemTables t;
    // paramOffset = 26 (BosonMixtureCluster.cpp:543); numberOfSplines = 26 cubic, 28 quartic (_4thorder.cpp:138-146)
    const int T = (int)pairTypes.size(), ord = splineOrder, nb = ord - 1, K = 26 + 2 * (ord - 3), EXT = K + 4;
    t.system_kind = TDVMC_SYSTEM_MIXTURE;
    t.spline_order = ord;
    t.n_particles = N;
    t.n_params = 26 * T;
    t.lbox = 0.0;
    t.tail_param = -1;
    t.n_other = numOfOtherExpectationValues;
    t.n_ext = T * EXT;
    t.n_pair_types = T;
    for (const auto& row : correlationTypes)
        for (int c : row) t.pair_type.push_back(c);
    t.hbar_over_2m = hbarOver2mPerParticle;
    t.mass = massPerParticle;
    t.map_ptr.push_back(0);
    for (int c = 0; c < T; c++)
    {
        const MixturePairType& p = pairTypes[c];
        std::vector<double> nodes = p.nodes;
        std::vector<double> w = FlattenWeights(p.splineWeights);
        if (ord == 4 && (int)nodes.size() == K + ord)
        {
            // 32 nodes carry 27 quartic splines; the reference's 28th never receives a term (see tdvmc_gpu.h): zero
            // spline behind one padding knot
            nodes.push_back(nodes.back() + 1.0);
            w.resize((size_t)K * (ord + 1) * (ord + 1), 0.0);
        }
        if ((int)nodes.size() != K + ord + 1 || (int)w.size() != K * (ord + 1) * (ord + 1))
            throw std::runtime_error("BosonMixtureCluster: 26 (cubic) / 28 (quartic) splines per pair type expected");
        t.type_knots.insert(t.type_knots.end(), nodes.begin(), nodes.end());
        t.type_weights.insert(t.type_weights.end(), w.begin(), w.end());
        t.type_mcmillan.push_back(p.mcMillanFactor);
        t.pair_potential.push_back(p.potential);
        const int b = c * EXT, MC = K, CO = K + 1, LI = K + 2, LG = K + 3;
        const auto& bc = p.bcFactors; // BosonMixtureCluster.cpp:636-645, BosonMixtureCluster_4thorder.cpp:641-650
        auto row = [&](int lead, int r, int first) {
            std::vector<std::pair<int, double> > e;
            e.push_back({ lead, 1.0 });
            for (int j = 0; j < nb; j++) e.push_back({ first + j, bc[r][j] });
            PushRow(t, e);
        };
        row(b + MC, 0, b);
        row(b + nb, 1, b);
        for (int i = 2; i < 22; i++) PushRow(t, { { b + i + nb - 1, 1.0 } });
        row(b + K - nb - 1, 2, b + K - nb);
        row(b + CO, 3, b + K - nb);
        row(b + LI, 4, b + K - nb);
        PushRow(t, { { b + LG, 1.0 } });
    }
    t.knots.assign(t.type_knots.begin(), t.type_knots.begin() + K + ord + 1);
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
    sd.dim = tb.dim;
    sd.n_params = tb.n_params;
    sd.n_splines = tb.system_kind == TDVMC_SYSTEM_MIXTURE ? 26 + 2 * (tb.spline_order - 3)
                   : (int32_t)tb.knots.size() - (tb.system_kind == TDVMC_SYSTEM_INH_CONTACT ? 8 : 4);
    sd.pair_rule = tb.pair_rule;
    sd.tail_param = tb.tail_param;
    sd.n_other = tb.n_other;
    sd.lbox = tb.lbox;
    sd.hbar2_2m = tb.hbar2_2m;
    sd.knots = tb.knots.data();
    sd.spline_weights = tb.spline_weights.empty() ? nullptr : tb.spline_weights.data();
    sd.map_ptr = tb.map_ptr.data();
    sd.map_col = tb.map_col.data();
    sd.map_val = tb.map_val.data();
    sd.system_params = tb.system_params.data();
    sd.n_system_params = (int32_t)tb.system_params.size();
    sd.system_kind = tb.system_kind;
    sd.n_ext = tb.n_ext > 0 ? tb.n_ext : sd.n_splines;
    sd.n_splines_first = tb.n_splines_first;
    sd.map_const = tb.map_const.empty() ? nullptr : tb.map_const.data();
    sd.grad_const = tb.grad_const.empty() ? nullptr : tb.grad_const.data();
    tdvmc_mixture_desc md;
    sd.mixture = nullptr;
    if (tb.system_kind == TDVMC_SYSTEM_MIXTURE)
    {
        md.n_pair_types = tb.n_pair_types;
        md.spline_order = tb.spline_order;
        md.pair_type = tb.pair_type.data();
        md.hbar_over_2m = tb.hbar_over_2m.data();
        md.mass = tb.mass.data();
        md.knots = tb.type_knots.data();
        md.spline_weights = tb.type_weights.data();
        md.mcmillan_factor = tb.type_mcmillan.data();
        md.potential = tb.pair_potential.data();
        sd.mixture = &md;
    }
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

void GpuEnsembleSystem::ResetCounters() { Check(tdvmc_gpu_reset_counters(handle), "reset_counters"); }

void GpuEnsembleSystem::SetMCStep(double MC_STEP) { Check(tdvmc_gpu_set_mc_step(handle, MC_STEP), "set_mc_step"); }

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

void GpuEnsembleSystem::SampleExpectationValues(const std::vector<double>& uR, const std::vector<double>& uI, double phiR,
                                                double phiI, int MC_NSTEPS, int MC_NTHERMSTEPS, int MC_NINITIALIZATIONSTEPS,
                                                double time)
{
    Check(tdvmc_gpu_set_params(handle, uR.data(), uI.data(), phiR, phiI, time), "set_params");
    Check(tdvmc_gpu_sample_and_accumulate(handle, MC_NSTEPS, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS), "sample_and_accumulate");
}

static tdvmc_solver_desc MakeSolverDesc(int IMAGINARY_TIME, int USE_PRECONDITIONING, int LINEAR_EQUATION_SOLVER_TYPE = 0)
{
    tdvmc_solver_desc sd;
    memset(&sd, 0, sizeof(sd));
    sd.struct_size = sizeof(sd);
    sd.imaginary_time = IMAGINARY_TIME;
    sd.use_preconditioning = USE_PRECONDITIONING;
    sd.solver_type = LINEAR_EQUATION_SOLVER_TYPE;
    sd.regularization = LINEAR_EQUATION_SOLVER_TYPE == 1 ? 0.002 : 0.001; // src/TDVMC.cpp:1770, :1737
    return sd;
}

bool GpuEnsembleSystem::SolveForParametersDot(std::vector<double>& uDotR, std::vector<double>& uDotI, double* phiDotR,
                                              double* phiDotI, int IMAGINARY_TIME, int USE_PRECONDITIONING)
{
    const tdvmc_solver_desc sd = MakeSolverDesc(IMAGINARY_TIME, USE_PRECONDITIONING, solverType);
    uDotR.assign(P, 0.0);
    uDotI.assign(P, 0.0);
    tdvmc_parameters_dot d;
    memset(&d, 0, sizeof(d));
    d.u_dot_r = uDotR.data();
    d.u_dot_i = uDotI.data();
    Check(tdvmc_gpu_solve_parameters_dot(handle, &sd, &d), "solve_parameters_dot");
    *phiDotR = d.phi_dot_r;
    *phiDotI = d.phi_dot_i;
    return d.not_positive_definite != 0;
}

bool GpuEnsembleSystem::CalculateNextParametersEuler(double dt, std::vector<double>& uR, std::vector<double>& uI, double* phiR,
                                                     double* phiI, int IMAGINARY_TIME, int USE_PRECONDITIONING, double time,
                                                     double* localEnergyR, double* localEnergyI)
{
    const tdvmc_solver_desc sd = MakeSolverDesc(IMAGINARY_TIME, USE_PRECONDITIONING, solverType);
    tdvmc_parameters_dot d;
    memset(&d, 0, sizeof(d));
    Check(tdvmc_gpu_euler_step(handle, &sd, dt, time, uR.data(), uI.data(), phiR, phiI, &d), "euler_step");
    if (localEnergyR) *localEnergyR = d.local_energy_r;
    if (localEnergyI) *localEnergyI = d.local_energy_i;
    return d.not_positive_definite != 0;
}

GpuEnsembleSystem::Dot GpuEnsembleSystem::SolveNow(int IMAGINARY_TIME, int USE_PRECONDITIONING)
{
    Dot d;
    d.notPD = SolveForParametersDot(d.uR, d.uI, &d.phiR, &d.phiI, IMAGINARY_TIME, USE_PRECONDITIONING);
    return d;
}

// one intermediate stage: estimators at the given parameters (fresh sampling, or the stored samples if mcCounts is null),
// then the solve on the device
GpuEnsembleSystem::Dot GpuEnsembleSystem::Stage(const std::vector<double>& uR, const std::vector<double>& uI, double phiR,
                                                 double phiI, const int* mcCounts, int IMAGINARY_TIME, int USE_PRECONDITIONING,
                                                 double time)
{
    if (mcCounts) SampleExpectationValues(uR, uI, phiR, phiI, mcCounts[0], mcCounts[1], mcCounts[2], time);
    else
    {
        Check(tdvmc_gpu_set_params(handle, uR.data(), uI.data(), phiR, phiI, time), "set_params");
        Check(tdvmc_gpu_reevaluate_stored(handle), "reevaluate_stored");
    }
    return SolveNow(IMAGINARY_TIME, USE_PRECONDITIONING);
}

bool GpuEnsembleSystem::PredictorCorrector(double dt, std::vector<double>& uR, std::vector<double>& uI, double* phiR, double* phiI,
                                           int pcSteps, const int* mcCounts, int IMAGINARY_TIME, int USE_PRECONDITIONING,
                                           double time)
{
    const double dt_2 = dt / 2.0;
    const Dot d0 = SolveNow(IMAGINARY_TIME, USE_PRECONDITIONING);
    bool bad = d0.notPD;
    std::vector<double> tR(P), tI(P);
    for (int i = 0; i < P; i++)
    {
        tR[i] = uR[i] + d0.uR[i] * dt;
        tI[i] = uI[i] + d0.uI[i] * dt;
    }
    double tpR = *phiR + d0.phiR * dt, tpI = *phiI + d0.phiI * dt;
    for (int s = 0; s < pcSteps; s++)
    {
        const Dot d1 = Stage(tR, tI, tpR, tpI, mcCounts, IMAGINARY_TIME, USE_PRECONDITIONING, time);
        bad = bad || d1.notPD;
        for (int i = 0; i < P; i++)
        {
            tR[i] = uR[i] + (d0.uR[i] + d1.uR[i]) * dt_2;
            tI[i] = uI[i] + (d0.uI[i] + d1.uI[i]) * dt_2;
        }
        tpR = *phiR + (d0.phiR + d1.phiR) * dt_2;
        tpI = *phiI + (d0.phiI + d1.phiI) * dt_2;
    }
    uR = tR;
    uI = tI;
    *phiR = tpR;
    *phiI = tpI;
    Check(tdvmc_gpu_set_params(handle, uR.data(), uI.data(), *phiR, *phiI, time), "set_params");
    return bad;
}

bool GpuEnsembleSystem::RungeKutta4(double dt, std::vector<double>& uR, std::vector<double>& uI, double* phiR, double* phiI,
                                    const int* mcCounts, int IMAGINARY_TIME, int USE_PRECONDITIONING, double time)
{
    const double steps[3] = { dt / 2.0, dt / 2.0, dt };
    std::vector<Dot> d;
    d.push_back(SolveNow(IMAGINARY_TIME, USE_PRECONDITIONING));
    std::vector<double> tR(P), tI(P);
    for (int k = 0; k < 3; k++)
    {
        const Dot& c = d.back();
        for (int i = 0; i < P; i++)
        {
            tR[i] = uR[i] + c.uR[i] * steps[k];
            tI[i] = uI[i] + c.uI[i] * steps[k];
        }
        d.push_back(Stage(tR, tI, *phiR + c.phiR * steps[k], *phiI + c.phiI * steps[k], mcCounts, IMAGINARY_TIME,
                          USE_PRECONDITIONING, time));
    }
    bool bad = false;
    for (const Dot& x : d) bad = bad || x.notPD;
    for (int i = 0; i < P; i++) // src/TDVMC.cpp:2030-2031
    {
        uR[i] = uR[i] + ((d[0].uR[i] + d[1].uR[i] * 2.0 + d[2].uR[i] * 2.0 + d[3].uR[i]) / 6.0) * dt;
        uI[i] = uI[i] + ((d[0].uI[i] + d[1].uI[i] * 2.0 + d[2].uI[i] * 2.0 + d[3].uI[i]) / 6.0) * dt;
    }
    *phiR = *phiR + ((d[0].phiR + d[1].phiR * 2.0 + d[2].phiR * 2.0 + d[3].phiR) / 6.0) * dt;
    *phiI = *phiI + ((d[0].phiI + d[1].phiI * 2.0 + d[2].phiI * 2.0 + d[3].phiI) / 6.0) * dt;
    Check(tdvmc_gpu_set_params(handle, uR.data(), uI.data(), *phiR, *phiI, time), "set_params");
    return bad;
}

bool GpuEnsembleSystem::CalculateNextParametersPC(double dt, std::vector<double>& uR, std::vector<double>& uI, double* phiR,
                                                  double* phiI, int MC_NSTEPS, int MC_NTHERMSTEPS, int MC_NINITIALIZATIONSTEPS,
                                                  int IMAGINARY_TIME, int USE_PRECONDITIONING, double time)
{
    const int mc[3] = { MC_NSTEPS, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS };
    return PredictorCorrector(dt, uR, uI, phiR, phiI, 1, mc, IMAGINARY_TIME, USE_PRECONDITIONING, time);
}

bool GpuEnsembleSystem::CalculateNextParametersPCReuseSamples(double dt, std::vector<double>& uR, std::vector<double>& uI,
                                                              double* phiR, double* phiI, int IMAGINARY_TIME,
                                                              int USE_PRECONDITIONING, double time)
{
    return PredictorCorrector(dt, uR, uI, phiR, phiI, 6, nullptr, IMAGINARY_TIME, USE_PRECONDITIONING, time);
}

bool GpuEnsembleSystem::CalculateNextParametersRK4(double dt, std::vector<double>& uR, std::vector<double>& uI, double* phiR,
                                                   double* phiI, int MC_NSTEPS, int MC_NTHERMSTEPS, int MC_NINITIALIZATIONSTEPS,
                                                   int IMAGINARY_TIME, int USE_PRECONDITIONING, double time)
{
    const int mc[3] = { MC_NSTEPS, MC_NTHERMSTEPS, MC_NINITIALIZATIONSTEPS };
    return RungeKutta4(dt, uR, uI, phiR, phiI, mc, IMAGINARY_TIME, USE_PRECONDITIONING, time);
}

bool GpuEnsembleSystem::CalculateNextParametersRK4ReuseSamples(double dt, std::vector<double>& uR, std::vector<double>& uI,
                                                               double* phiR, double* phiI, int IMAGINARY_TIME,
                                                               int USE_PRECONDITIONING, double time)
{
    return RungeKutta4(dt, uR, uI, phiR, phiI, nullptr, IMAGINARY_TIME, USE_PRECONDITIONING, time);
}

ObservableTables MakePairDistributionGrid(double rMax, int numOfPairDistributionValues, double weight)
{
    ObservableTables t;
    t.grMax = rMax;
    t.grSpacing = rMax / numOfPairDistributionValues;
    t.grCount = (int)((rMax - 0.0) / t.grSpacing); // Grid.cpp:21 (truncates)
    t.grWeight = weight;
    t.grScaling.assign(t.grCount, 0.0);
    for (int i = 0; i < t.grCount; i++) t.grScaling[i] = 4.0 * M_PI * std::pow(t.grSpacing * (i + 1), 3.0) / 3.0;
    for (int i = t.grCount - 1; i > 0; i--) t.grScaling[i] = t.grScaling[i] - t.grScaling[i - 1];
    return t;
}

AdditionalObservables GpuEnsembleSystem::ParallelCalculateAdditionalSystemProperties(
    const std::vector<double>& uR, const std::vector<double>& uI, double phiR, double phiI, const ObservableTables& obs,
    int MC_NADDITIONALSTEPS, int MC_NADDITIONALTHERMSTEPS, int MC_NADDITIONALINITIALIZATIONSTEPS, double time)
{
    Check(tdvmc_gpu_set_params(handle, uR.data(), uI.data(), phiR, phiI, time), "set_params");
    std::vector<int32_t> ptr(1, 0);
    std::vector<double> kv;
    for (const auto& shell : obs.kValues)
    {
        for (const auto& k : shell) kv.insert(kv.end(), k.begin(), k.begin() + 3);
        ptr.push_back((int32_t)(kv.size() / 3));
    }
    tdvmc_observable_desc od;
    od.gr_count = obs.grCount;
    od.n_shells = (int32_t)obs.kValues.size();
    od.gr_spacing = obs.grSpacing;
    od.gr_max = obs.grMax;
    od.gr_weight = obs.grWeight;
    od.gr_scaling = obs.grScaling.data();
    od.shell_ptr = ptr.data();
    od.kvec = kv.data();
    AdditionalObservables out;
    out.pairDistribution.assign(od.gr_count, 0.0);
    out.structureFactor.assign(od.n_shells, 0.0);
    Check(tdvmc_gpu_sample_observables(handle, &od, MC_NADDITIONALSTEPS, MC_NADDITIONALTHERMSTEPS, MC_NADDITIONALINITIALIZATIONSTEPS,
                                       out.pairDistribution.data(), out.structureFactor.data()),
          "sample_observables");
    return out;
}

ClusterObservables GpuEnsembleSystem::ParallelCalculateAdditionalSystemPropertiesCluster(
    const std::vector<double>& uR, const std::vector<double>& uI, double phiR, double phiI, const ClusterObservableTables& obs,
    int MC_NADDITIONALSTEPS, int MC_NADDITIONALTHERMSTEPS, int MC_NADDITIONALINITIALIZATIONSTEPS, double time)
{
    Check(tdvmc_gpu_set_params(handle, uR.data(), uI.data(), phiR, phiI, time), "set_params");
    tdvmc_cluster_observable_desc od;
    od.n_angle = obs.angleCount;
    od.n_density = obs.densityCount;
    od.n_distance = obs.distanceCount;
    od.reserved = 0;
    od.angle_spacing = obs.angleSpacing;
    od.density_spacing = obs.densitySpacing;
    od.density_max = obs.densityMax;
    od.distance_spacing = obs.distanceSpacing;
    od.distance_max = obs.distanceMax;
    od.density_scaling = obs.densityScaling.data();
    ClusterObservables out;
    out.angularDistribution.assign((size_t)3 * od.n_angle, 0.0);
    out.densityFromCOM.assign((size_t)3 * od.n_density, 0.0);
    out.particleDistances.assign((size_t)3 * od.n_distance, 0.0);
    Check(tdvmc_gpu_sample_cluster_observables(handle, &od, MC_NADDITIONALSTEPS, MC_NADDITIONALTHERMSTEPS,
                                               MC_NADDITIONALINITIALIZATIONSTEPS, &out.r2, out.angularDistribution.data(),
                                               out.densityFromCOM.data(), out.particleDistances.data()),
          "sample_cluster_observables");
    return out;
}

void GpuEnsembleSystem::UpdateSamplesConsecutive(int nrOfSamplesToUpdate, const std::vector<double>& uR,
                                                 const std::vector<double>& uI, double phiR, double phiI, int MC_NTHERMSTEPS,
                                                 double time)
{
    Check(tdvmc_gpu_set_params(handle, uR.data(), uI.data(), phiR, phiI, time), "set_params");
    Check(tdvmc_gpu_update_stored(handle, nrOfSamplesToUpdate, MC_NTHERMSTEPS), "update_stored");
}

double GpuEnsembleSystem::GetExponent()
{
    double x = 0.0;
    Check(tdvmc_gpu_last_exponent(handle, &x), "last_exponent");
    return x;
}

} // namespace tdvmc_host
