// Minimal C++ driver over GpuEnsembleSystem: what the reference's time loop does around one
// estimator evaluation (src/TDVMC.cpp:3437-3763), for a BosonsBulk system whose knots and spline
// table are read from a text file (the reference driver would pass its own SplineFactory output).
//
//   example_driver <tables.txt>      tables.txt: N LBOX N_PARAM n_knots, knots..., K*16 weights...
//   example_driver --map hebulk N LBOX N_PARAM | --map hedrop N N_PARAM | --map boxradial N LBOX N_PARAM | --map file <f>
//                                    prints the parameter map the adapter builds (no GPU needed)
//
// Exit code 0, one line "E_R=... acceptance=..." and one line "EULER ..." (parameters after one device-solved Euler step) on success; without a CUDA device the library
// refuses to run (no CPU fallback) and the driver prints the error and exits with code 3.
#include "GpuEnsembleSystem.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>

using namespace tdvmc_host;

int main(int argc, char** argv)
{
    if (argc < 2)
    {
        std::fprintf(stderr, "usage: example_driver <tables.txt>\n");
        return 2;
    }
    if (!std::strcmp(argv[1], "--map") && argc >= 4)
    {
        const bool bulk = !std::strcmp(argv[2], "hebulk");
        SystemTables t;
        if (!std::strcmp(argv[2], "boxradial") && argc >= 6)
        {
            // uniform grid of N_PARAM/2 + 1 points on [0, LBOX/2], mirrored by three knots on either side (SetNodes,
            // NUBosonsBulkPBBoxAndRadial.cpp:36-62); the table itself does not enter the map
            const int P = std::atoi(argv[5]), n = P / 2 + 1;
            const double L = std::atof(argv[4]);
            std::vector<double> grid(n), nodes;
            for (int i = 0; i < n; i++) grid[i] = (L / 2) * i / (n - 1);
            nodes = grid;
            for (int i = 0; i < 3; i++)
            {
                nodes.insert(nodes.begin(), -grid[i + 1]);
                nodes.push_back(2.0 * grid[n - 1] - grid[n - 2 - i]);
            }
            const int K = (int)nodes.size() - 4;
            std::vector<std::vector<std::vector<double> > > w(K, std::vector<std::vector<double> >(4, std::vector<double>(4, 0.0)));
            t = MakeNUBosonsBulkPBBoxAndRadialTables(std::atoi(argv[3]), L, P, nodes, w, { 0.1, 50.0 }, 400);
        }
        else if (!std::strcmp(argv[2], "file") && argc >= 4)
        {
            // builders that need the reference's own boundary factors: read them from a text file
            //   mixture <order> <N> <T> / N*N correlationTypes / per type: nk knots..., 5 x (order-1) bcFactors, mcMillan, potential
            //   inhcontact <N> <LBOX> <N_PARAM> / 4 SYSTEM_PARAMS / spf then pc: nk knots..., 3x3 bcStart rows count + values,
            //              bcEnd rows count + values, np1 np2 np3
            std::ifstream in(argv[3]);
            std::string what;
            in >> what;
            if (what == "mixture")
            {
                int order, N, T;
                in >> order >> N >> T;
                std::vector<std::vector<int> > ct(N, std::vector<int>(N));
                for (auto& row : ct)
                    for (int& c : row) in >> c;
                std::vector<MixturePairType> types(T);
                for (auto& p : types)
                {
                    int nk;
                    in >> nk;
                    p.nodes.resize(nk);
                    for (double& x : p.nodes) in >> x;
                    const int K = nk - order - 1;
                    p.splineWeights.assign(K, std::vector<std::vector<double> >(order + 1, std::vector<double>(order + 1, 0.0)));
                    p.bcFactors.assign(5, std::vector<double>(order - 1));
                    for (auto& row : p.bcFactors)
                        for (double& x : row) in >> x;
                    in >> p.mcMillanFactor >> p.potential;
                }
                t = MakeBosonMixtureClusterTables(N, ct, std::vector<double>(N, 1.0), std::vector<double>(N, 1.0), types, 403, order);
            }
            else
            {
                int N, P;
                double L;
                in >> N >> L >> P;
                std::vector<double> sp(4);
                for (double& x : sp) in >> x;
                SplinedFunctionTables part[2];
                for (auto& f : part)
                {
                    int nk, ns, ne;
                    in >> nk;
                    f.nodes.resize(nk);
                    for (double& x : f.nodes) in >> x;
                    f.splineWeights.assign(nk - 4, std::vector<std::vector<double> >(4, std::vector<double>(4, 0.0)));
                    in >> ns;
                    f.bcFactorsStart.assign(ns, std::vector<double>(3));
                    for (auto& row : f.bcFactorsStart)
                        for (double& x : row) in >> x;
                    in >> ne;
                    f.bcFactorsEnd.assign(ne, std::vector<double>(3));
                    for (auto& row : f.bcFactorsEnd)
                        for (double& x : row) in >> x;
                    in >> f.np1 >> f.np2 >> f.np3;
                }
                t = MakeInhContactBosonsTables(N, L, P, sp, part[0], part[1]);
            }
            if (!in) return 2;
        }
        else
            t = bulk ? MakeHeBulkTables(std::atoi(argv[3]), std::atof(argv[4]), std::atoi(argv[5]))
                     : MakeHeDropTables(std::atoi(argv[3]), std::atoi(argv[4]));
        std::printf("%d %d %d %d %d\n", t.system_kind, t.n_params, t.n_ext, t.n_other,
                    t.system_kind == TDVMC_SYSTEM_MIXTURE ? 26 + 2 * (t.spline_order - 3)
                                                          : (int)t.knots.size() - (t.system_kind == TDVMC_SYSTEM_INH_CONTACT ? 8 : 4));
        for (int p : t.map_ptr) std::printf("%d ", p);
        std::printf("\n");
        for (int c : t.map_col) std::printf("%d ", c);
        std::printf("\n");
        for (double v : t.map_val) std::printf("%.17g ", v);
        std::printf("\n");
        for (double v : t.map_const) std::printf("%.17g ", v);
        std::printf("\n");
        for (double v : t.grad_const) std::printf("%.17g ", v);
        std::printf("\n");
        return 0;
    }
    std::ifstream f(argv[1]);
    int N, P, nk;
    double L;
    if (!(f >> N >> L >> P >> nk)) return 2;
    std::vector<double> nodes(nk);
    for (double& x : nodes) f >> x;
    const int K = nk - 4;
    std::vector<std::vector<std::vector<double> > > w(K, std::vector<std::vector<double> >(4, std::vector<double>(4)));
    for (auto& s : w)
        for (auto& p : s)
            for (double& c : p) f >> c;
    if (!f) return 2;

    try
    {
        SystemTables t = MakeBosonsBulkTables(N, L, P, nodes, w, { 1.0, 1.0 });
        const int walkers = 64, MC_NSTEPS = 2, MC_NTHERMSTEPS = N, MC_NINIT = 10 * N;
        GpuEnsembleSystem gpu(t, walkers, 0.5, MC_NSTEPS, 0, 1ull, 0, 1, 0);
        // start-up lattice (src/TDVMC.cpp:727-739 shape)
        const int m = (int)std::lround(std::cbrt((double)N));
        std::vector<std::vector<std::vector<double> > > R(walkers, std::vector<std::vector<double> >(N, std::vector<double>(3)));
        for (int wk = 0; wk < walkers; wk++)
            for (int n = 0; n < N; n++)
            {
                R[wk][n][0] = ((n % m) + 0.5 + 0.01 * wk / walkers) * L / m - L / 2;
                R[wk][n][1] = (((n / m) % m) + 0.5) * L / m - L / 2;
                R[wk][n][2] = ((n / (m * m)) + 0.5) * L / m - L / 2;
            }
        gpu.SetPositions(R);
        std::vector<double> uR(P), uI(P, 0.0);
        for (int k = 0; k < P; k++) uR[k] = -0.5 * std::exp(-std::pow(k * (L / 2) / (P - 1) / 0.8, 2));
        gpu.MoveCoordinatesToFirstCell();
        Estimators e = gpu.ParallelUpdateExpectationValues(uR, uI, 0.0, 0.0, MC_NSTEPS, MC_NTHERMSTEPS, MC_NINIT, 0.0);
        std::printf("E_R=%.12g acceptance=%.4f samples=%lld exponent=%.10g\n", e.localEnergyR,
                    (double)e.nAcceptances / (double)e.nTrials, e.nSamples, gpu.GetExponent());
        // one imaginary-time Euler step solved on the device from the estimators that are still resident there
        // (CalculateNextParametersEuler, src/TDVMC.cpp:1834-1853)
        double phiR = 0.0, phiI = 0.0, eR = 0.0, eI = 0.0;
        const bool notPD = gpu.CalculateNextParametersEuler(1e-4, uR, uI, &phiR, &phiI, 1, 1, 1e-4, &eR, &eI);
        std::printf("EULER uR0=%.17g uRlast=%.17g phiR=%.17g E_R=%.17g notPD=%d\n", uR[0], uR[P - 1], phiR, eR, notPD ? 1 : 0);
        // and one Runge-Kutta step (CalculateNextParametersRK4, src/TDVMC.cpp:1969-2035): four device stages, each a
        // sampling pass followed by the device solve
        gpu.SampleExpectationValues(uR, uI, phiR, phiI, MC_NSTEPS, MC_NTHERMSTEPS, 0, 1e-4);
        const bool notPD4 = gpu.CalculateNextParametersRK4(1e-4, uR, uI, &phiR, &phiI, MC_NSTEPS, MC_NTHERMSTEPS, 0, 1, 1, 2e-4);
        std::printf("RK4 uR0=%.17g uRlast=%.17g phiR=%.17g notPD=%d\n", uR[0], uR[P - 1], phiR, notPD4 ? 1 : 0);
    }
    catch (const std::exception& ex)
    {
        std::fprintf(stderr, "example_driver: %s\n", ex.what());
        return 3;
    }
    return 0;
}
