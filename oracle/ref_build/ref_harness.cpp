/*
 * ref_harness — drives the UNMODIFIED reference (mathiasgartner/TDVMC) as a CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This translation unit textually includes the
 * reference's src/TDVMC.cpp (where it lies under /root/reference; nothing is
 * copied into this repository) with its main() renamed, and then calls the
 * reference's own functions:
 *
 *   InitializePhysicalSystem()  src/TDVMC.cpp:2786
 *   Init()                      src/TDVMC.cpp:514
 *   PostSystemInit()            src/TDVMC.cpp:616
 *   sys->CalculateWavefunction / CalculateExpectationValues / CalculateWFQuotient
 *                               src/PhysicalSystems/IPhysicalSystem.h:98-122
 *   DoMetropolisStep()          src/TDVMC.cpp:858
 *   VectorDisplacementNIC()     src/Utils.cpp:376
 *
 * Modes (argv[1]):
 *   eval  <case> <out>   fixed-configuration evaluation -> named arrays, %.17g
 *   mc    <case> <out>   fixed-parameter sampling run   -> estimators + series
 *   bench <case>         timed sampling run, prints one line "trials seconds samples"
 *   nic   <in> <out>     minimum-image displacement for a list of vector pairs
 *
 * `private`/`protected` are re-declared public for this TU only, so that the
 * harness can read the reference's internal tables (knots, spline weights,
 * basis sums).  The reference objects it links against are compiled unchanged.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#define private public
#define protected public
#define main tdvmc_reference_main
#include "src/TDVMC.cpp"
#undef main
#undef private
#undef protected

namespace
{

struct Case
{
    std::map<std::string, std::vector<double> > num;
    std::map<std::string, std::string> str;
    std::vector<std::vector<double> > moves; // particle, new x, y, z

    bool has(const std::string& k) const { return num.count(k) > 0; }
    double d(const std::string& k, double def = 0.0) const
    {
        auto it = num.find(k);
        return (it == num.end() || it->second.empty()) ? def : it->second[0];
    }
    int i(const std::string& k, int def = 0) const { return (int)std::llround(d(k, def)); }
    std::vector<double> v(const std::string& k) const
    {
        auto it = num.find(k);
        return it == num.end() ? std::vector<double>() : it->second;
    }
};

// case file: one record per line, "<key> <values...>"; "system <name>", "configdir <path>"
// and "move <p> <x> <y> <z>" are special.
Case ReadCase(const std::string& path)
{
    Case c;
    std::ifstream f(path);
    if (!f)
    {
        std::cerr << "cannot open case file " << path << std::endl;
        std::exit(2);
    }
    std::string line;
    while (std::getline(f, line))
    {
        std::istringstream ss(line);
        std::string key;
        if (!(ss >> key) || key[0] == '#') continue;
        if (key == "system" || key == "configdir")
        {
            std::string s;
            ss >> s;
            c.str[key] = s;
            continue;
        }
        std::vector<double> vals;
        std::string tok;
        while (ss >> tok) vals.push_back(std::strtod(tok.c_str(), nullptr));
        if (key == "move") c.moves.push_back(vals);
        else c.num[key] = vals;
    }
    return c;
}

struct Dump
{
    std::ofstream f;
    explicit Dump(const std::string& path) : f(path)
    {
        f << std::setprecision(17);
    }
    void scalar(const std::string& name, double x)
    {
        f << name << " 0\n" << x << "\n";
    }
    void vec(const std::string& name, const std::vector<double>& x)
    {
        f << name << " 1 " << x.size() << "\n";
        for (double y : x) f << y << " ";
        f << "\n";
    }
    void vecll(const std::string& name, const std::vector<long long>& x)
    {
        f << name << " 1 " << x.size() << "\n";
        for (long long y : x) f << y << " ";
        f << "\n";
    }
    void mat(const std::string& name, const std::vector<std::vector<double> >& x)
    {
        size_t c = x.empty() ? 0 : x[0].size();
        f << name << " 2 " << x.size() << " " << c << "\n";
        for (auto& r : x)
            for (double y : r) f << y << " ";
        f << "\n";
    }
    void ten(const std::string& name, const std::vector<std::vector<std::vector<double> > >& x)
    {
        size_t b = x.empty() ? 0 : x[0].size();
        size_t c = (x.empty() || x[0].empty()) ? 0 : x[0][0].size();
        f << name << " 3 " << x.size() << " " << b << " " << c << "\n";
        for (auto& m : x)
            for (auto& r : m)
                for (double y : r) f << y << " ";
        f << "\n";
    }
};

std::vector<std::vector<double> > R;
std::vector<double> uR, uI;
double phiR = 0, phiI = 0;

// Sets the reference's file-scope configuration globals (src/TDVMC.cpp:76-131) from the
// case, builds the system through the reference's own factory and initialises it.
void SetupReference(const Case& c)
{
    SYSTEM_TYPE = c.str.at("system");
    configDirectory = c.str.count("configdir") ? c.str.at("configdir") : std::string("./");
    N = c.i("N");
    DIM = c.i("DIM", 3);
    LBOX = c.d("LBOX");
    N_PARAM = c.i("N_PARAM");
    RHO = c.d("RHO", 0.0);
    RC = c.d("RC", 0.0);
    MC_STEP = c.d("MC_STEP", 0.5);
    MC_NSTEPS = c.i("MC_NSTEPS", 1);
    MC_NTHERMSTEPS = c.i("MC_NTHERMSTEPS", 1);
    MC_NINITIALIZATIONSTEPS = c.i("MC_NINITIALIZATIONSTEPS", 0);
    IMAGINARY_TIME = c.i("IMAGINARY_TIME", 1);
    GR_BIN_COUNT = c.i("GR_BIN_COUNT", 100);
    RHO_BIN_COUNT = c.i("RHO_BIN_COUNT", 100);
    USE_NURBS = c.i("USE_NURBS", 0);
    NURBS_GRID = c.v("NURBS_GRID");
    SYSTEM_PARAMS = c.v("SYSTEM_PARAMS");
    PARTICLE_TYPES.clear();
    for (double t : c.v("PARTICLE_TYPES")) PARTICLE_TYPES.push_back((int)t);
    WRITE_EVERY_NTH_STEP_TO_FILE = 1;
    MC_NSTEP_MULTIPLICATION_FACTOR_FOR_WRITE_DATA = 1;
    UPDATE_SAMPLES_EVERY_NTH_STEP = 0;
    processRank = c.i("rank", 0);
    numOfProcesses = 1;
    isRootRank = false; // keeps the reference's progress output quiet

    if (!InitializePhysicalSystem())
    {
        std::cerr << "unknown system type " << SYSTEM_TYPE << std::endl;
        std::exit(2);
    }
    // Init() prints a few lines to stdout; silence them.
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    Init();
    std::cout.rdbuf(old);

    std::vector<double> flat = c.v("R");
    R.assign(N, std::vector<double>(DIM, 0.0));
    for (int n = 0; n < N; n++)
        for (int a = 0; a < DIM; a++) R[n][a] = flat[(size_t)n * DIM + a];
    uR = c.v("uR");
    uI = c.v("uI");
    uR.resize(N_PARAM, 0.0);
    uI.resize(N_PARAM, 0.0);
    phiR = c.d("phiR", 0.0);
    phiI = c.d("phiI", 0.0);

    std::cout.rdbuf(sink.rdbuf());
    sys->InitSystem();
    PostSystemInit();
    std::cout.rdbuf(old);
    sys->SetTime(c.d("time", 0.0));
}

template <class S>
void DumpSplineInternals(Dump& d, S* s)
{
    d.vec("knots", s->nodes);
    d.ten("spline_weights", s->splineWeights);
    d.vec("spline_sums", s->splineSums);
    d.scalar("outer_sum", s->outerSum);
    d.scalar("max_distance", s->maxDistance);
    d.ten("sD", s->splineSumsD);
    d.mat("sD2", s->splineSumsD2);
    d.vec("other_local_operators", s->otherLocalOperators);
}

// BosonMixtureCluster and BosonMixtureCluster_4thorder share their member names (per-pair-type CorrelationFunctionData)
template <class S>
void DumpMixtureInternals(Dump& d, S* s)
{
        std::vector<double> pt, ct, mass, hb;
        for (int t : s->particleTypes) pt.push_back(t);
        for (auto& row : s->correlationTypes)
            for (int t : row) ct.push_back(t);
        for (auto& pp : s->particleProperties)
        {
            mass.push_back(pp.mass);
            hb.push_back(pp.hbarOver2m);
        }
        d.vec("particle_types", pt);
        d.vec("correlation_types", ct);
        d.vec("type_mass", mass);
        d.vec("type_hbar_over_2m", hb);
        d.scalar("n_pair_types", (double)s->corrFuncData.size());
        for (size_t c = 0; c < s->corrFuncData.size(); c++)
        {
            auto& f = s->corrFuncData[c];
            std::string q = "_" + std::to_string(c);
            d.vec("knots" + q, f.nodes);
            d.ten("spline_weights" + q, f.splineWeights);
            d.mat("bc_factors" + q, f.bcFactors);
            d.vec("spline_sums" + q, f.splineSums);
            d.vec("extras" + q, { f.mcMillanSum, f.constSum, f.linearSum, f.logSum, f.rijSplit, f.rijTail, f.mcMillanFactor });
            d.ten("sD" + q, f.splineSumsD);
            d.mat("sD2" + q, f.splineSumsD2);
            d.mat("mcmillan_sum_d" + q, f.mcMillanSumD);
            d.vec("mcmillan_sum_d2" + q, f.mcMillanSumD2);
            d.mat("linear_sum_d" + q, f.linearSumD);
            d.vec("linear_sum_d2" + q, f.linearSumD2);
            d.mat("log_sum_d" + q, f.logSumD);
            d.vec("log_sum_d2" + q, f.logSumD2);
        }
    }

void DumpSystemInternals(Dump& d)
{
    if (auto s = dynamic_cast<PhysicalSystems::BosonsBulk*>(sys))
    {
        DumpSplineInternals(d, s);
        d.mat("bc_start", s->bcFactorsStart);
        d.mat("bc_end", s->bcFactorsEnd);
    }
    else if (auto s = dynamic_cast<PhysicalSystems::NUBosonsBulkPB*>(sys))
    {
        DumpSplineInternals(d, s);
    }
    else if (auto s = dynamic_cast<PhysicalSystems::NUBosonsBulkPBBoxAndRadial*>(sys))
    {
        // two bases on the same knots: radial splines in r_ij and "box" splines in |x_ij|, |y_ij|, |z_ij|
        d.vec("knots", s->nodes);
        d.vec("knots_rad", s->nodesRad);
        d.ten("spline_weights", s->splineWeights);
        d.ten("spline_weights_rad", s->splineWeightsRad);
        d.vec("spline_sums", s->splineSums);
        d.vec("spline_sums_rad", s->splineSumsRad);
        d.scalar("max_distance_rad", s->maxDistanceRad);
        d.scalar("half_length", s->halfLength);
        d.ten("sD", s->splineSumsD);
        d.mat("sD2", s->splineSumsD2);
        d.ten("sD_rad", s->splineSumsDRad);
        d.mat("sD2_rad", s->splineSumsD2Rad);
        d.vec("other_local_operators", s->otherLocalOperators);
        d.vec("gr_bins", s->grBins);
        d.vec("gr_bin_volumes", s->grBinVolumes);
        d.scalar("gr_node_point_spacing", s->grNodePointSpacing);
    }
    else if (auto s = dynamic_cast<PhysicalSystems::InhContactBosons*>(sys))
    {
        // one-dimensional: single-particle function (spf) on [0, L] + pair correlation (pc) on [0, L/2]
        auto part = [&](const std::string& q, WFParts::SplinedFunction& f) {
            d.vec("knots_" + q, f.nodes);
            d.ten("spline_weights_" + q, f.splineWeights);
            d.mat("bc_start_" + q, f.bcFactorsStart);
            d.mat("bc_end_" + q, f.bcFactorsEnd);
            d.vec("np_" + q, { (double)f.np1, (double)f.np2, (double)f.np3, (double)f.numberOfSplines });
            d.scalar("node_spacing_" + q, f.nodeSpacing);
            d.vec("spline_sums_" + q, f.splineSums);
            d.ten("sD_" + q, f.splineSumsD);
            d.mat("sD2_" + q, f.splineSumsD2);
        };
        part("spf", s->spf);
        part("pc", s->pc);
        d.scalar("gamma", s->gamma);
        d.scalar("max_distance", s->maxDistance);
        d.vec("other_local_operators", s->otherLocalOperators);
    }
    else if (auto s = dynamic_cast<PhysicalSystems::BosonMixtureCluster*>(sys))
    {
        DumpMixtureInternals(d, s);
    }
    else if (auto s = dynamic_cast<PhysicalSystems::BosonMixtureCluster_4thorder*>(sys))
    {
        DumpMixtureInternals(d, s);
    }
    else if (auto s = dynamic_cast<PhysicalSystems::HeDrop*>(sys))
    {
        d.vec("spline_sums", s->splineSums);
        d.scalar("mcmillan_sum", s->mcMillanSum);
        d.scalar("const_sum", s->constSum);
        d.scalar("linear_sum", s->linearSum);
        d.ten("sD", s->splineSumsD);
        d.mat("sD2", s->splineSumsD2);
        d.mat("mcmillan_sum_d", s->mcMillanSumD);
        d.vec("mcmillan_sum_d2", s->mcMillanSumD2);
        d.mat("linear_sum_d", s->linearSumD);
        d.vec("linear_sum_d2", s->linearSumD2);
        d.scalar("rij_split", s->rijSplit);
        d.scalar("rij_spline_split", s->rijSplineSplit);
        d.scalar("rij_tail", s->rijTail);
        d.vec("bc_factors", { s->factorFirstSpline1, s->factorFirstSpline2, s->factorSecondSpline1, s->factorSecondSpline2,
                              s->factorSecondLastSpline, s->factorSecondLastSplineConst, s->factorSecondLastSplineLinear,
                              s->factorLastSpline, s->factorLastSplineConst, s->factorLastSplineLinear,
                              s->factorSecondLastShort, s->factorSecondLastLarge, s->factorLastShort, s->factorLastLarge,
                              s->factorFirstShort, s->factorFirstLarge, s->factorSecondShort, s->factorSecondLarge });
    }
    else if (auto s = dynamic_cast<PhysicalSystems::HeBulk*>(sys))
    {
        d.vec("spline_sums", s->splineSums);
        d.scalar("mcmillan_sum", s->mcMillanSum);
        d.ten("sD", s->splineSumsD);
        d.mat("sD2", s->splineSumsD2);
        d.mat("mcmillan_sum_d", s->mcMillanSumD);
        d.vec("mcmillan_sum_d2", s->mcMillanSumD2);
        d.scalar("rij_split", s->rijSplit);
        d.scalar("node_point_spacing", s->nodePointSpacing);
        d.scalar("max_distance", s->maxDistance);
        d.vec("bc_factors", { s->factorFirstSpline1, s->factorFirstSpline2, s->factorSecondSpline1, s->factorSecondSpline2,
                              s->factorSecondLastSpline, s->factorLastSpline, s->factorSecondLastSplinePhi, s->factorLastSplinePhi });
    }
}

int ModeEval(const Case& c, const std::string& out)
{
    SetupReference(c);
    Dump d(out);

    sys->CalculateWavefunction(R, uR, uI, phiR, phiI);
    d.scalar("exponent_wf", sys->GetExponent());
    sys->CalculateExpectationValues(R, uR, uI, phiR, phiI);

    d.scalar("exponent", sys->GetExponent());
    d.scalar("wf", sys->GetWf());
    d.scalar("local_energy_r", sys->GetLocalEnergyR());
    d.scalar("local_energy_i", sys->GetLocalEnergyI());
    d.vec("local_operators", sys->GetLocalOperators());
    d.vec("local_operator_energy_r", sys->GetLocalOperatorlocalEnergyR());
    d.vec("local_operator_energy_i", sys->GetLocalOperatorlocalEnergyI());
    d.vec("other_expectation_values", sys->GetOtherExpectationValues());
    // one row and the diagonal of O (x) O are enough to pin the layout of the P x P matrix
    std::vector<std::vector<double> > M = sys->GetLocalOperatorsMatrix();
    std::vector<double> diag(M.size()), row3;
    for (size_t k = 0; k < M.size(); k++) diag[k] = M[k][k];
    if (M.size() > 3) row3 = M[3];
    d.vec("local_operators_matrix_diag", diag);
    d.vec("local_operators_matrix_row3", row3);
    DumpSystemInternals(d);

    // scripted single-particle moves: quotient for the proposal, state left untouched
    std::vector<double> q, en;
    for (auto& m : c.moves)
    {
        int p = (int)m[0];
        std::vector<double> oldPosition(R[p]);
        for (int a = 0; a < DIM; a++) R[p][a] = m[1 + a];
        q.push_back(sys->CalculateWFQuotient(R, uR, uI, phiR, phiI, p, oldPosition));
        en.push_back(sys->GetExponentNew());
        R[p] = oldPosition;
    }
    d.vec("move_quotient", q);
    d.vec("move_exponent_new", en);
    return 0;
}

// Fixed-parameter sampling, mirroring UpdateExpectationValues (src/TDVMC.cpp:1038-1150)
// but keeping the per-sample series so that error bars can be attached.
int ModeMC(const Case& c, const std::string& out, bool benchOnly)
{
    SetupReference(c);
    int nSamples = c.i("MC_NSTEPS", 1);
    int nTherm = c.i("MC_NTHERMSTEPS", 1);
    int nInit = c.i("MC_NINITIALIZATIONSTEPS", 0);
    int seed = c.i("seed", 1);
    generator = std::mt19937_64((unsigned long long)seed);

    std::vector<double> O(N_PARAM, 0.0), OER(N_PARAM, 0.0), OEI(N_PARAM, 0.0), erSeries, eiSeries;
    std::vector<std::vector<double> > S(N_PARAM, std::vector<double>(N_PARAM, 0.0));
    std::vector<double> other;
    double er = 0, ei = 0;

    sys->CalculateWavefunction(R, uR, uI, phiR, phiI);
    for (int i = 0; i < nInit; i++) DoMetropolisStep(R, uR, uI, phiR, phiI);
    nTrials = 0;
    nAcceptances = 0;

    auto t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < nSamples; s++)
    {
        for (int t = 0; t < nTherm; t++) DoMetropolisStep(R, uR, uI, phiR, phiI);
        sys->CalculateExpectationValues(R, uR, uI, phiR, phiI);
        if (benchOnly)
        {
            // the reference's estimator accumulation (src/TDVMC.cpp:1103-1109) is part of the path
            localOperators += sys->GetLocalOperators() / (double)nSamples;
            localEnergyR += sys->GetLocalEnergyR() / (double)nSamples;
            localEnergyI += sys->GetLocalEnergyI() / (double)nSamples;
            localOperatorsMatrix += sys->GetLocalOperatorsMatrix() / (double)nSamples;
            localOperatorlocalEnergyR += sys->GetLocalOperatorlocalEnergyR() / (double)nSamples;
            localOperatorlocalEnergyI += sys->GetLocalOperatorlocalEnergyI() / (double)nSamples;
            otherExpectationValues += sys->GetOtherExpectationValues() / (double)nSamples;
            continue;
        }
        std::vector<double> o = sys->GetLocalOperators();
        double e1 = sys->GetLocalEnergyR(), e2 = sys->GetLocalEnergyI();
        erSeries.push_back(e1);
        eiSeries.push_back(e2);
        er += e1;
        ei += e2;
        for (int k = 0; k < N_PARAM; k++)
        {
            O[k] += o[k];
            OER[k] += o[k] * e1;
            OEI[k] += o[k] * e2;
            for (int j = 0; j < N_PARAM; j++) S[k][j] += o[k] * o[j];
        }
        std::vector<double> oth = sys->GetOtherExpectationValues();
        if (other.empty()) other.assign(oth.size(), 0.0);
        for (size_t k = 0; k < oth.size(); k++) other[k] += oth[k];
    }
    auto t1 = std::chrono::steady_clock::now();
    double seconds = std::chrono::duration<double>(t1 - t0).count();

    if (benchOnly)
    {
        std::printf("%lld %.6f %d %.10g\n", nTrials, seconds, nSamples, localEnergyR);
        return 0;
    }

    double inv = 1.0 / nSamples;
    for (int k = 0; k < N_PARAM; k++)
    {
        O[k] *= inv;
        OER[k] *= inv;
        OEI[k] *= inv;
        for (int j = 0; j < N_PARAM; j++) S[k][j] *= inv;
    }
    for (auto& x : other) x *= inv;
    Dump d(out);
    d.scalar("local_energy_r", er * inv);
    d.scalar("local_energy_i", ei * inv);
    d.vec("local_operators", O);
    d.vec("local_operator_energy_r", OER);
    d.vec("local_operator_energy_i", OEI);
    d.mat("local_operators_matrix", S);
    d.vec("other_expectation_values", other);
    d.vec("energy_r_series", erSeries);
    d.vec("energy_i_series", eiSeries);
    d.scalar("n_trials", (double)nTrials);
    d.scalar("n_acceptances", (double)nAcceptances);
    d.scalar("seconds", seconds);
    std::vector<double> flat;
    for (auto& r : R)
        for (double x : r) flat.push_back(x);
    d.vec("R_final", flat);
    return 0;
}

// input: "L dim" then lines "i1 i2 i3 j1 j2 j3"; output per line: norm, d1, d2, d3
int ModeNic(const std::string& in, const std::string& out)
{
    std::ifstream f(in);
    std::ofstream o(out);
    o << std::setprecision(17);
    std::string line;
    while (std::getline(f, line))
    {
        std::istringstream ss(line);
        double L;
        int dim;
        if (!(ss >> L >> dim)) continue;
        LBOX = L;
        LBOX_R = 1.0 / LBOX;
        LBOX_2 = LBOX / 2.0;
        DIM = dim;
        std::vector<double> a(dim), b(dim), r(dim);
        for (int k = 0; k < dim; k++) ss >> a[k];
        for (int k = 0; k < dim; k++) ss >> b[k];
        double n = VectorDisplacementNIC(a, b, r);
        o << n;
        for (int k = 0; k < dim; k++) o << " " << r[k];
        o << "\n";
    }
    return 0;
}

} // namespace

// Additional observables (BosonsBulk.cpp:474-520, NUBosonsBulkPB.cpp:597-639): g(r) and S(k) of the case's
// configuration, then the reference's own sampling loop for them (src/TDVMC.cpp:1332-1388) with the
// MC_NADDITIONAL* counts of the case.
template <class S>
void DumpObservableSetup(Dump& d, S* s)
{
    d.scalar("gr_spacing", s->pairDistribution.grid.spacing);
    d.scalar("gr_max", s->pairDistribution.grid.max);
    d.scalar("gr_count", (double)s->pairDistribution.grid.count);
    d.vec("gr_scaling", s->pairDistribution.scalingGrid);
    d.vec("gr_fixed", s->pairDistribution.observablesV[0].values);
    d.vec("sk_fixed", s->structureFactor.observablesV[0].values);
    d.vec("k_norms", s->kNorms);
    std::vector<double> flat, sizes;
    for (int k = 0; k < s->numOfkValues; k++)
    {
        sizes.push_back((double)s->kValues[k].size());
        for (auto& v : s->kValues[k])
            for (double x : v) flat.push_back(x);
    }
    d.vec("k_shell_sizes", sizes);
    d.vec("k_vectors", flat);
}

void DumpOnGrid(Dump& d, const std::string& name, Observables::ObservableVsOnGrid& o, const std::string& suffix)
{
    std::vector<std::vector<double> > v;
    for (auto& ov : o.observablesV) v.push_back(ov.values);
    d.mat(name + suffix, v);
}

int ModeObs(const Case& c, const std::string& out)
{
    SetupReference(c);
    Dump d(out);
    sys->CalculateWavefunction(R, uR, uI, phiR, phiI);
    sys->CalculateAdditionalSystemProperties(R, uR, uI, phiR, phiI);
    auto mix = dynamic_cast<PhysicalSystems::BosonMixtureCluster*>(sys);
    if (auto s = dynamic_cast<PhysicalSystems::BosonsBulk*>(sys)) DumpObservableSetup(d, s);
    else if (auto s = dynamic_cast<PhysicalSystems::NUBosonsBulkPB*>(sys)) DumpObservableSetup(d, s);
    else if (dynamic_cast<PhysicalSystems::HeBulk*>(sys) || dynamic_cast<PhysicalSystems::HeDrop*>(sys))
    {
        // HeBulk.cpp:413-448 / HeDrop.cpp:655-702: [otherExpectationValues | S(k) per shell | (HeDrop) r2]
        d.vec("additional_fixed", sys->GetAdditionalSystemProperties());
        d.scalar("n_other", (double)sys->GetNumOfOtherExpectationValues());
        std::vector<double> flat, sizes;
        auto dumpK = [&](const std::vector<std::vector<std::vector<double> > >& kv, int nk) {
            for (int k = 0; k < nk; k++)
            {
                sizes.push_back((double)kv[k].size());
                for (auto& v : kv[k])
                    for (double x : v) flat.push_back(x);
            }
        };
        if (auto s = dynamic_cast<PhysicalSystems::HeBulk*>(sys)) dumpK(s->kValues, s->numOfkValues);
        if (auto s = dynamic_cast<PhysicalSystems::HeDrop*>(sys)) dumpK(s->kValues, s->numOfkValues);
        d.vec("k_shell_sizes", sizes);
        d.vec("k_vectors", flat);
        return 0;
    }
    else if (mix)
    {
        // BosonMixtureCluster.cpp:327-345, 680-741: r2, corner angles, density from the centre of mass, pair distances
        d.scalar("r2_fixed", mix->r2.value);
        DumpOnGrid(d, "angle", mix->angularDistribution, "_fixed");
        DumpOnGrid(d, "density", mix->densityFromCOM, "_fixed");
        DumpOnGrid(d, "distance", mix->particleDistances, "_fixed");
        d.vec("angle_grid", { (double)mix->angularDistribution.grid.count, mix->angularDistribution.grid.spacing, mix->angularDistribution.grid.max });
        d.vec("density_grid", { (double)mix->densityFromCOM.grid.count, mix->densityFromCOM.grid.spacing, mix->densityFromCOM.grid.max });
        d.vec("distance_grid", { (double)mix->particleDistances.grid.count, mix->particleDistances.grid.spacing, mix->particleDistances.grid.max });
        d.vec("density_scaling", mix->densityFromCOM.scalingGrid);
    }
    else
    {
        std::cerr << "obs: system without additional observables in this harness" << std::endl;
        return 2;
    }
    MC_NADDITIONALSTEPS = c.i("MC_NADDITIONALSTEPS", 0);
    MC_NADDITIONALTHERMSTEPS = c.i("MC_NADDITIONALTHERMSTEPS", 1);
    MC_NADDITIONALINITIALIZATIONSTEPS = c.i("MC_NADDITIONALINITIALIZATIONSTEPS", 0);
    mc_nadditionalsteps = MC_NADDITIONALSTEPS;
    if (MC_NADDITIONALSTEPS > 0)
    {
        generator = std::mt19937_64((unsigned long long)c.i("seed", 1));
        nTrials = 0;
        nAcceptances = 0;
        CalculateAdditionalSystemProperties(R, uR, uI, phiR, phiI);
        if (mix)
        {
            d.scalar("r2_mean", dynamic_cast<Observables::Observable*>(additionalObservablesMean.observables[0])->value);
            DumpOnGrid(d, "angle", *dynamic_cast<Observables::ObservableVsOnGrid*>(additionalObservablesMean.observables[1]), "_mean");
            DumpOnGrid(d, "density", *dynamic_cast<Observables::ObservableVsOnGrid*>(additionalObservablesMean.observables[2]), "_mean");
            DumpOnGrid(d, "distance", *dynamic_cast<Observables::ObservableVsOnGrid*>(additionalObservablesMean.observables[3]), "_mean");
            d.scalar("acceptance", (double)nAcceptances / (double)nTrials);
            return 0;
        }
        auto gr = dynamic_cast<Observables::ObservableVsOnGrid*>(additionalObservablesMean.observables[0]);
        auto sk = dynamic_cast<Observables::ObservableVsOnGrid*>(additionalObservablesMean.observables[1]);
        d.vec("gr_mean", gr->observablesV[0].values);
        d.vec("sk_mean", sk->observablesV[0].values);
        d.scalar("acceptance", (double)nAcceptances / (double)nTrials);
    }
    return 0;
}

// Imaginary- or real-time evolution with the reference's own time-step pieces: ParallelUpdateExpectationValues
// (src/TDVMC.cpp:1152-1220), SolveForParametersDot with the Cholesky path (:1713-1763: IncludePhi system, scaling
// preconditioner, +0.001 regularisation) and CalculateNextParametersEuler (:1834-1853).  Records the parameters and
// energies after every step, and the first step's estimators with the resulting derivatives.
int ModeEvolve(const Case& c, const std::string& out)
{
    SetupReference(c);
    IMAGINARY_TIME = c.i("IMAGINARY_TIME", 1);
    LINEAR_EQUATION_SOLVER_TYPE = c.i("LINEAR_EQUATION_SOLVER_TYPE", 0); // 1: the Eigen FullPivHouseholderQR branch (:1763-1827)
    USE_PRECONDITIONING = c.i("USE_PRECONDITIONING", 1);
    USE_PARAM_START = 0;
    USE_PARAM_END = 0;
    USED_PARAM_COUNT = N_PARAM;
    const int nSteps = c.i("time_steps", 10);
    const double dt = c.d("TIMESTEP", 1e-3);
    generator = std::mt19937_64((unsigned long long)c.i("seed", 1));
    Dump d(out);
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    isRootRank = true; // CalculateNextParametersEuler works on the root rank only (:1836)
    MPIMethods::isRootRank = true; // as mainMPI sets them (:3019-3022)
    MPIMethods::numOfProcesses = 1;
    MPIMethods::processRank = 0;
    MPIMethods::rootRank = 0;
    sys->CalculateWavefunction(R, uR, uI, phiR, phiI);
    for (int i = 0; i < c.i("equilibration_steps", 0); i++) DoMetropolisStep(R, uR, uI, phiR, phiI);
    std::vector<std::vector<double> > uRt, uIt;
    std::vector<double> eR, eI, phiRt;
    for (int step = 0; step < nSteps; step++)
    {
        ParallelUpdateExpectationValues(R, uR, uI, phiR, phiI, true);
        eR.push_back(localEnergyR);
        eI.push_back(localEnergyI);
        if (step == 0)
        {
            d.vec("first_O", localOperators);
            d.mat("first_S", localOperatorsMatrix);
            d.vec("first_OER", localOperatorlocalEnergyR);
            d.vec("first_OEI", localOperatorlocalEnergyI);
            d.scalar("first_ER", localEnergyR);
            d.scalar("first_EI", localEnergyI);
            std::vector<double> uDotR, uDotI;
            double phiDotR = 0, phiDotI = 0;
            SolveForParametersDot(uDotR, uDotI, &phiDotR, &phiDotI);
            d.vec("first_uDotR", uDotR);
            d.vec("first_uDotI", uDotI);
            d.scalar("first_phiDotR", phiDotR);
            d.scalar("first_phiDotI", phiDotI);
        }
        CalculateNextParametersEuler(dt, uR, uI, &phiR, &phiI);
        uRt.push_back(uR);
        uIt.push_back(uI);
        phiRt.push_back(phiR);
    }
    std::cout.rdbuf(old);
    isRootRank = false;
    d.mat("uR_t", uRt);
    d.mat("uI_t", uIt);
    d.vec("phiR_t", phiRt);
    d.vec("energy_r_t", eR);
    d.vec("energy_i_t", eI);
    d.scalar("acceptance", (double)nAcceptances / (double)nTrials);
    d.scalar("cholesky_failed", doNotAcceptStep ? 1.0 : 0.0);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 3)
    {
        std::cerr << "usage: ref_harness eval|mc|obs|evolve <case> <out> | bench <case> | nic <in> <out>" << std::endl;
        return 2;
    }
    std::string mode = argv[1];
    if (mode == "eval" && argc >= 4) return ModeEval(ReadCase(argv[2]), argv[3]);
    if (mode == "mc" && argc >= 4) return ModeMC(ReadCase(argv[2]), argv[3], false);
    if (mode == "obs" && argc >= 4) return ModeObs(ReadCase(argv[2]), argv[3]);
    if (mode == "evolve" && argc >= 4) return ModeEvolve(ReadCase(argv[2]), argv[3]);
    if (mode == "bench") return ModeMC(ReadCase(argv[2]), "", true);
    if (mode == "nic" && argc >= 4) return ModeNic(argv[2], argv[3]);
    std::cerr << "bad arguments" << std::endl;
    return 2;
}
