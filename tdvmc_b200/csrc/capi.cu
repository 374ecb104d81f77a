// C ABI of libtdvmc_b200.so (include/tdvmc_gpu.h): handle, host-side table preparation, launch
// sequencing, NCCL all-reduce.  No CPU fallback: every entry point needs a CUDA device.
#include "../../include/tdvmc_gpu.h"
#include "kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

using namespace tdvmc;

namespace
{

std::string g_create_error;

// ---- NCCL, bound at run time so that single-GPU use needs no NCCL at all ----
struct UniqueId
{
    char internal[TDVMC_GPU_UNIQUE_ID_BYTES];
};
struct NcclApi
{
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /* ncclUniqueId by value */ UniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int kNcclDouble = 8; // ncclFloat64
constexpr int kNcclSum = 0;    // ncclSum

bool load_nccl(std::string& err)
{
    if (g_nccl.AllReduce) return true;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib)
    {
        err = std::string("cannot load libnccl.so.2: ") + dlerror();
        return false;
    }
    g_nccl.lib = lib;
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, UniqueId, int))dlsym(lib, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    {
        err = "libnccl.so.2 lacks a required symbol";
        g_nccl = NcclApi();
        return false;
    }
    return true;
}

template <class T>
struct DevBuf
{
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count)
    {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e != cudaSuccess) p = nullptr;
        return e;
    }
    cudaError_t ensure(size_t count)
    {
        if (count <= n && p) return cudaSuccess;
        return alloc(count);
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

struct TimedLaunch
{
    int kernel;
    cudaEvent_t e0, e1;
};

} // namespace

struct tdvmc_gpu_handle
{
    std::string error;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;

    // system (host copies)
    int N = 0, Np = 0, P = 0, K = 0, pair_rule = 0, tail_param = 0, n_other = 9, kind = 0, n_ext = 0, gr_bins = 0;
    double he_rs = 0, core_m = 0, he_hl = 0, r_split2 = 1e300, r_tail = 1e300, gr_max = 0, u_core = 0, u_const = 0, u_lin = 0;
    int n_short = 0, potential = 0, rho_bins = 0, use_phi = 0, periodic = 1, dim = 3;
    // BosonMixtureCluster
    int n_types = 0, mix_order = 3;
    std::vector<int> mix_pair_type, mix_pot;
    std::vector<double> mix_hbar, mix_mass, mix_knots, mix_weights, mix_mcm;
    DevBuf<int> d_mix_pair_type, d_mix_pot;
    DevBuf<double> d_mix_hbar, d_mix_mass, d_mix_knots, d_mix_weights, d_mix_mcm, d_mix_cub;
    std::vector<double> map_const, grad_const;
    double L = 0, hbar = 1.0;
    std::vector<double> knots, weights, map_val, sys_params, uR, uI;
    std::vector<int> map_ptr, map_col;
    double phiR = 0, phiI = 0, time = 0;
    bool params_set = false;
    int first_bin = 3, nbins = 0, ncell = 0, uniform = 0;
    double bin_guard = 0.0; // see SysDev::bin_guard
    double h = 0;

    // ensemble
    int W = 0, first_walker = 0, max_samples = 1, keep_positions = 0;
    uint64_t seed = 1;
    double mc_step = 0.5;
    uint64_t step_counter = 0; // Metropolis steps done per walker (identical for all walkers)
    uint64_t trials_local = 0; // proposals on this rank since creation
    int wpb = 8, npp = 0, resident_per_sm = 0;
    bool large_sweep = false;   // walker too large for the one-warp sweep: always sweep_split_kernel with 8 warps, 4 table replicas

    // device tables
    DevBuf<double> d_knots, d_rec, d_cub, d_map_val, d_uR, d_uI, d_utR, d_utI, d_map_const;
    DevBuf<unsigned short> d_lut;
    DevBuf<int> d_map_ptr, d_map_col;
    // walkers and samples
    DevBuf<double> d_pos, d_aos, d_A, d_other, d_exponent, d_samp_pos, d_est, d_scratch, d_eval_slab;
    DevBuf<unsigned long long> d_accepted;
    DevBuf<double> d_T, d_vint, d_tab_e; // K3/K4 exhibit buffers, allocated on demand
    int lda = 0, ldc = 0;
    long long rows_cap = 0;    // padded row capacity of d_A
    long long rows_used = 0;   // samples of the last accumulation
    int stored_samples = 0;    // samples per walker kept in d_samp_pos
    int update_cursor = 0;     // currentSampleIndexForUpdate (src/TDVMC.cpp:975-983)
    double* h_est = nullptr;   // pinned
    size_t est_len = 0;
    bool est_valid = false;
    bool est_reduced = false;  // d_est already holds the sum over all ranks
    DevBuf<double> d_ugR, d_ugI, d_gr_vol; // NUBosonsBulkPBBoxAndRadial: drift-side u~, g(r) shell volumes
    double gr_spacing = 0;
    DevBuf<double> d_sol, d_solve_L, d_est_fixed; // solve.cu: result, global factor scratch, caller-given estimators
    double* h_sol = nullptr;   // pinned, 2P + 5
    int smem_optin = 48 * 1024;

    // communicator
    void* comm = nullptr;
    int rank = 0, n_ranks = 1;

    // profiling
    bool profiling = false;
    std::vector<TimedLaunch> pending;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
    DevBuf<unsigned char> d_flush;
    long long n_launched = 0; // kernels of this library launched on the handle
    long long launches[TDVMC_KERNEL_COUNT] = { 0 };
    double total_ms[TDVMC_KERNEL_COUNT] = { 0 };

    SysDev sysdev() const;
};

namespace
{

#define CK(call)                                                                                        \
    do                                                                                                  \
    {                                                                                                   \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess)                                                                          \
        {                                                                                               \
            h->error = std::string(#call) + ": " + cudaGetErrorString(_e);                              \
            return (int)_e ? (int)_e : -1;                                                              \
        }                                                                                               \
    } while (0)

int fail(tdvmc_gpu_handle* h, const std::string& msg, int code = -1)
{
    h->error = msg;
    return code;
}

cudaEvent_t get_event(tdvmc_gpu_handle* h)
{
    if (!h->event_pool.empty())
    {
        cudaEvent_t e = h->event_pool.back();
        h->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct Timed
{
    tdvmc_gpu_handle* h;
    int kernel;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    Timed(tdvmc_gpu_handle* h_, int k, int n_kernels = 1) : h(h_), kernel(k)
    {
        h->launches[k]++;
        h->n_launched += n_kernels;
        if (h->profiling)
        {
            e0 = get_event(h);
            e1 = get_event(h);
            cudaEventRecord(e0, h->stream);
        }
    }
    ~Timed()
    {
        if (e0)
        {
            cudaEventRecord(e1, h->stream);
            h->pending.push_back({ kernel, e0, e1 });
        }
    }
};

void collect_timings(tdvmc_gpu_handle* h)
{
    if (h->pending.empty()) return;
    cudaStreamSynchronize(h->stream);
    for (auto& t : h->pending)
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) h->total_ms[t.kernel] += ms;
        h->event_pool.push_back(t.e0);
        h->event_pool.push_back(t.e1);
    }
    h->pending.clear();
}

template <class T>
cudaError_t upload(DevBuf<T>& b, const std::vector<T>& v, cudaStream_t st)
{
    cudaError_t e = b.ensure(v.size());
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
}

// time-switched square well (BosonsBulk.cpp:237-243, NUBosonsBulkPB.cpp:300-307)
void potential_ab(const tdvmc_gpu_handle* h, double& a, double& b)
{
    a = h->sys_params.size() > 0 ? h->sys_params[0] : 0.0;
    b = h->sys_params.size() > 1 ? h->sys_params[1] : 0.0;
    if (h->sys_params.size() > 2 && h->time >= h->sys_params[2])
    {
        a = h->sys_params[3];
        b = h->sys_params[4];
    }
}

// Static tables: per-interval records of the spline pieces, interval lookup grid.
int build_static_tables(tdvmc_gpu_handle* h)
{
    const int K = h->K;
    if (h->kind == TDVMC_SYSTEM_MIXTURE)
    {
        h->periodic = 0;
        h->use_phi = 1;
        h->first_bin = h->mix_order;
        h->nbins = K - 2 * h->mix_order + 3;
        h->uniform = 0;
        h->ncell = 1;
        h->h = 1.0;
        CK(upload(h->d_mix_pair_type, h->mix_pair_type, h->stream));
        CK(upload(h->d_mix_pot, h->mix_pot, h->stream));
        CK(upload(h->d_mix_hbar, h->mix_hbar, h->stream));
        CK(upload(h->d_mix_mass, h->mix_mass, h->stream));
        CK(upload(h->d_mix_knots, h->mix_knots, h->stream));
        CK(upload(h->d_mix_weights, h->mix_weights, h->stream));
        CK(upload(h->d_mix_mcm, h->mix_mcm, h->stream));
        CK(upload(h->d_map_ptr, h->map_ptr, h->stream));
        CK(upload(h->d_map_col, h->map_col, h->stream));
        CK(upload(h->d_map_val, h->map_val, h->stream));
        CK(upload(h->d_map_const, h->map_const, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (h->kind == TDVMC_SYSTEM_HE_BULK || h->kind == TDVMC_SYSTEM_HE_DROP)
    {
        std::vector<unsigned short> lut(1, 0);
        if (h->kind == TDVMC_SYSTEM_HE_BULK)
        {
            // HeBulk::InitSystem (HeBulk.cpp:40-70): uniform grid of K - 3 intervals from rijSplit to L/2
            h->he_rs = 1.95;
            h->core_m = -5.0;
            h->gr_bins = 100;
            h->n_short = K;
            h->nbins = K - 3;
            h->uniform = 1;
            h->ncell = 1;
            h->h = (h->L / 2.0 - h->he_rs) / (double)(K - 3.0);
            h->he_hl = h->h;
            h->gr_max = h->L / 2.0;
            h->potential = 0;
            h->periodic = 1;
        }
        else
        {
            // HeDrop::InitSystem (HeDrop.cpp:71-97): 70 splines of spacing 0.1 from rijSplit = 3, then spacing 0.5
            h->he_rs = 3.0;
            h->core_m = -4.7;
            h->gr_bins = 200;
            h->rho_bins = 200;
            h->n_short = 70;
            h->h = 0.1;
            h->he_hl = 0.5;
            const int n_large = K - h->n_short;
            if (n_large < 4) return fail(h, "HeDrop needs N_PARAM >= 71");
            h->r_split2 = h->h * (h->n_short - 3.0) + h->he_rs;
            h->r_tail = h->he_hl * (n_large - 3.0) + h->r_split2;
            h->gr_max = h->r_tail * 2.0; // HeDrop.cpp:222
            h->nbins = (h->n_short - 3) + (n_large - 3);
            h->uniform = 0;
            h->potential = 1;
            h->use_phi = 1;
            h->periodic = 0;
            // interval lookup over [0, r_tail - rs] in cells of half the short spacing
            const double span = h->r_tail - h->he_rs;
            h->ncell = (int)std::ceil(2.0 * span / h->h);
            lut.assign(h->ncell, 0);
            const double cw = span / h->ncell;
            for (int c = 0; c < h->ncell; c++)
            {
                const double x = c * cw;
                int jj = x < (h->r_split2 - h->he_rs) ? (int)std::floor(x / h->h)
                                                       : (h->n_short - 3) + (int)std::floor((x - (h->r_split2 - h->he_rs)) / h->he_hl);
                if (jj > 0) jj--; // start one interval low: the device loop only walks upwards
                lut[c] = (unsigned short)std::min(jj, h->nbins - 1);
            }
        }
        h->first_bin = 0;
        CK(upload(h->d_lut, lut, h->stream));
        CK(upload(h->d_map_ptr, h->map_ptr, h->stream));
        CK(upload(h->d_map_col, h->map_col, h->stream));
        CK(upload(h->d_map_val, h->map_val, h->stream));
        CK(upload(h->d_map_const, h->map_const, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return 0;
    }
    CK(upload(h->d_map_const, h->map_const, h->stream));
    if (h->kind == TDVMC_SYSTEM_INH_CONTACT)
    {
        // two knot vectors and the raw spline table; interval lookup is a binary search in the kernels (GetBinIndex)
        h->periodic = 1;
        h->use_phi = 1;
        h->first_bin = 3;
        h->nbins = K - 6;
        h->uniform = 0;
        h->ncell = 1;
        h->h = 1.0;
        CK(upload(h->d_knots, h->knots, h->stream));
        CK(upload(h->d_rec, h->weights, h->stream));
        CK(upload(h->d_map_ptr, h->map_ptr, h->stream));
        CK(upload(h->d_map_col, h->map_col, h->stream));
        CK(upload(h->d_map_val, h->map_val, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return 0;
    }
    const std::vector<double>& t = h->knots;
    // first interval a distance can fall in: knots[first_bin] <= 0 < knots[first_bin + 1]
    int fb = 0;
    while (fb + 1 < K && t[fb + 1] <= 0.0) fb++;
    h->first_bin = fb;
    h->nbins = K - fb;
    if (h->nbins <= 0 || fb < 3) return fail(h, "knot vector must have three knots below the first non-negative one");
    const double rmax = t[K];
    // uniform?
    const double h0 = (t[K] - t[fb]) / h->nbins;
    bool uni = (t[fb] == 0.0);
    double min_sp = rmax;
    for (int j = fb; j < K; j++)
    {
        const double sp = t[j + 1] - t[j];
        if (!(sp > 0.0)) return fail(h, "knots must be strictly increasing on [0, r_max]");
        min_sp = std::min(min_sp, sp);
        if (std::fabs(sp - h0) > 1e-9 * h0) uni = false;
    }
    h->uniform = uni ? 1 : 0;
    h->h = h0;
    h->bin_guard = 0.0;
    if (uni)
    {
        // largest deviation of a stored knot from the exact grid, in units of the spacing (the reference's knots are
        // (i * L / 2) / (P - 1), BosonsBulk.cpp:61-67: a few ulp)
        double dev = 0.0;
        for (int j = fb; j <= K; j++) dev = std::max(dev, std::fabs(t[j] - (j - fb) * h0) / h0);
        if (dev < 1e-7) h->bin_guard = 2.0 * dev + 1e-10;
    }
    h->ncell = uni ? h->nbins : (int)std::min(8192.0, std::ceil(2.0 * rmax / min_sp));
    if (h->ncell < 1) h->ncell = 1;
    std::vector<unsigned short> lut(h->ncell);
    const double cw = rmax / h->ncell;
    int j = fb;
    for (int c = 0; c < h->ncell; c++)
    {
        const double x = c * cw;
        while (j + 1 < K && t[j + 1] <= x) j++;
        lut[c] = (unsigned short)j;
    }
    std::vector<double> rec((size_t)h->nbins * kRecStride, 0.0);
    for (int b = fb; b < K; b++)
        for (int p = 0; p < 4; p++)
            for (int c = 0; c < 4; c++)
                rec[(size_t)(b - fb) * kRecStride + p * 4 + c] = h->weights[((size_t)(b - p) * 4 + p) * 4 + c];
    CK(upload(h->d_knots, h->knots, h->stream));
    CK(upload(h->d_rec, rec, h->stream));
    CK(upload(h->d_lut, lut, h->stream));
    CK(upload(h->d_map_ptr, h->map_ptr, h->stream));
    CK(upload(h->d_map_col, h->map_col, h->stream));
    CK(upload(h->d_map_val, h->map_val, h->stream));
    if (h->kind == TDVMC_SYSTEM_BOX_RADIAL)
    {
        // NUBosonsBulkPBBoxAndRadial::InitSystem :146-169: g(r) over (0, halfLength), bins weighted by 1 / shell volume
        h->gr_bins = h->n_other - 3;
        h->gr_max = h->L / 2.0;
        h->gr_spacing = h->gr_max / (double)h->gr_bins;
        std::vector<double> vol(h->gr_bins);
        for (int i = 0; i < h->gr_bins; i++)
            vol[i] = h->dim == 3 ? 4.0 * M_PI * pow(h->gr_spacing * (i + 1), 3.0) / 3.0
                                 : (h->dim == 2 ? M_PI * pow(h->gr_spacing * (i + 1), 2.0) : 2.0 * (h->gr_spacing * (i + 1)));
        for (int i = h->gr_bins - 1; i > 0; i--) vol[i] = vol[i] - vol[i - 1];
        h->use_phi = 1;
        CK(upload(h->d_gr_vol, vol, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// Parameter-dependent tables: u~ = M^T u and the per-interval cubic of the sweep.
int build_param_tables(tdvmc_gpu_handle* h)
{
    const int K = h->K, P = h->P, fb = h->first_bin;
    std::vector<double> utR(h->n_ext, 0.0), utI(h->n_ext, 0.0);
    for (int p = 0; p < P; p++)
        for (int j = h->map_ptr[p]; j < h->map_ptr[p + 1]; j++)
        {
            utR[h->map_col[j]] += h->uR[p] * h->map_val[j];
            utI[h->map_col[j]] += h->uI[p] * h->map_val[j];
        }
    // u(r) on interval b in the local coordinate s = r - t_b, expanded exactly (long double) from the
    if (h->kind == TDVMC_SYSTEM_MIXTURE)
    {
        // per pair type and knot interval: u(r) = sum_p u~[t][b-p] piece_p(b-p)(r) re-expanded around the left knot
        // (Taylor shift of the degree-`ord` monomial sum, long double); intervals ord .. K - (ord - 3) - 1
        const int ord = h->mix_order, np = ord + 1, nk = K + ord + 1, EXT = K + 4, nb = K - 2 * ord + 3, stride = ord + 3;
        std::vector<double> mc((size_t)h->n_types * nb * stride, 0.0);
        for (int t = 0; t < h->n_types; t++)
            for (int b = ord; b < ord + nb; b++)
            {
                long double C[5] = { 0, 0, 0, 0, 0 };
                for (int p = 0; p < np; p++)
                    for (int c = 0; c < np; c++)
                        C[c] += (long double)utR[t * EXT + b - p] * (long double)h->mix_weights[(((size_t)t * K + (b - p)) * np + p) * np + c];
                const long double t0 = h->mix_knots[(size_t)t * nk + b];
                double* q = &mc[((size_t)t * nb + (b - ord)) * stride];
                for (int j = 0; j <= ord; j++) // q_j = sum_{c >= j} binom(c, j) C_c t0^(c-j), Horner in t0
                {
                    long double acc = 0;
                    for (int c = ord; c >= j; c--)
                    {
                        long double binom = 1;
                        for (int i = 0; i < j; i++) binom = binom * (c - i) / (i + 1);
                        acc = acc * t0 + binom * C[c];
                    }
                    q[j] = (double)acc;
                }
                q[ord + 1] = h->mix_knots[(size_t)t * nk + b];
                q[ord + 2] = h->mix_knots[(size_t)t * nk + b + 1];
            }
        CK(upload(h->d_uR, h->uR, h->stream));
        CK(upload(h->d_uI, h->uI, h->stream));
        CK(upload(h->d_utR, utR, h->stream));
        CK(upload(h->d_utI, utI, h->stream));
        CK(upload(h->d_mix_cub, mc, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (h->kind == TDVMC_SYSTEM_INH_CONTACT)
    {
        // contact term of the exponent, -2 gamma h_pc ss_pc[0] (InhContactBosons.cpp:764, gradient / Laplacian :595-602):
        // a parameter-free coefficient of the first pair spline, real part only
        const int K1 = h->n_short, K2 = K - K1;
        const double* k1 = h->knots.data();
        const double* k2 = k1 + K1 + 4;
        const double gamma = (h->sys_params[0] == 0.0 ? h->sys_params[1] : 0.0) * (h->sys_params[2] * M_PI); // :25-29
        const double h_pc = k2[4] - k2[3]; // pc.nodeSpacing of the default grid (:91-95)
        h->u_const = -2.0 * gamma * h_pc;
        h->u_core = gamma;
        utR[K1] += h->u_const;
        std::vector<double> cub((size_t)(K - 6) * 6, 0.0);
        for (int part = 0; part < 2; part++)
        {
            const int Kp = part ? K2 : K1, off = part ? K1 : 0;
            const double* kn = part ? k2 : k1;
            double* c = cub.data() + (size_t)(part ? (K1 - 3) * 6 : 0);
            for (int b = 3; b < Kp; b++)
            {
                long double C[4] = { 0, 0, 0, 0 };
                for (int p = 0; p < 4; p++)
                    for (int q = 0; q < 4; q++)
                        C[q] += (long double)utR[off + b - p] * (long double)h->weights[((size_t)(off + b - p) * 4 + p) * 4 + q];
                const long double t0 = kn[b];
                double* o = c + (size_t)(b - 3) * 6;
                o[0] = (double)(C[0] + t0 * (C[1] + t0 * (C[2] + t0 * C[3])));
                o[1] = (double)(C[1] + t0 * (2 * C[2] + 3 * t0 * C[3]));
                o[2] = (double)(C[2] + 3 * t0 * C[3]);
                o[3] = (double)C[3];
                o[4] = kn[b];
                o[5] = kn[b + 1];
            }
        }
        CK(upload(h->d_uR, h->uR, h->stream));
        CK(upload(h->d_uI, h->uI, h->stream));
        CK(upload(h->d_utR, utR, h->stream));
        CK(upload(h->d_utI, utI, h->stream));
        CK(upload(h->d_cub, cub, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (h->kind == TDVMC_SYSTEM_BOX_RADIAL)
    {
        // two plane sets on the same knots: radial (u~[0..K)) then box (u~[K..2K)); record nbins of either is the zero
        // tail (no radial term beyond maxDistanceRad, :640; a box argument never exceeds knots[K] = L/2)
        const int nrec = h->nbins + 1;
        std::vector<double> cub((size_t)2 * nrec * 6, 0.0);
        for (int set = 0; set < 2; set++)
        {
            double* c = cub.data() + (size_t)set * nrec * 6;
            const double* ut = utR.data() + (size_t)set * K;
            for (int rec = 0; rec < h->nbins; rec++)
            {
                const int b = fb + rec;
                long double C[4] = { 0, 0, 0, 0 };
                for (int p = 0; p < 4; p++)
                    for (int q = 0; q < 4; q++)
                        C[q] += (long double)ut[b - p] * (long double)h->weights[((size_t)(b - p) * 4 + p) * 4 + q];
                const long double t0 = h->knots[b];
                c[(size_t)rec * 2 + 0] = (double)(C[0] + t0 * (C[1] + t0 * (C[2] + t0 * C[3])));
                c[(size_t)rec * 2 + 1] = (double)(C[1] + t0 * (2 * C[2] + 3 * t0 * C[3]));
                c[(size_t)(nrec + rec) * 2 + 0] = (double)(C[2] + 3 * t0 * C[3]);
                c[(size_t)(nrec + rec) * 2 + 1] = (double)C[3];
                c[(size_t)(2 * nrec + rec) * 2 + 0] = h->knots[b];
                c[(size_t)(2 * nrec + rec) * 2 + 1] = h->knots[b + 1];
            }
            c[(size_t)(2 * nrec + h->nbins) * 2 + 0] = h->knots[K];
            c[(size_t)(2 * nrec + h->nbins) * 2 + 1] = 1e300;
        }
        // the drift's view of the parameters: sD[K-1] (box) stands in for sDRad[K-1] (NUBosonsBulkPBBoxAndRadial.cpp:493-497)
        std::vector<double> ugR(utR), ugI(utI);
        ugR[2 * K - 1] += ugR[K - 1];
        ugI[2 * K - 1] += ugI[K - 1];
        ugR[K - 1] = 0.0;
        ugI[K - 1] = 0.0;
        CK(upload(h->d_uR, h->uR, h->stream));
        CK(upload(h->d_uI, h->uI, h->stream));
        CK(upload(h->d_utR, utR, h->stream));
        CK(upload(h->d_utI, utI, h->stream));
        CK(upload(h->d_ugR, ugR, h->stream));
        CK(upload(h->d_ugI, ugI, h->stream));
        CK(upload(h->d_cub, cub, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return 0;
    }
    // caller's monomial table: u(r) = sum_p u~[b-p] * piece_p(b-p)(r)
    // planes: [c0,c1] | [c2,c3] | [t_lo,t_hi], nbins + 1 records each; record nbins is the constant tail
    // uR[tail_param] every pair beyond r_max contributes (BosonsBulk.cpp:532-534)
    const int nrec = h->nbins + 1;
    std::vector<double> cub((size_t)nrec * 6, 0.0);
    const bool he = h->kind != TDVMC_SYSTEM_SPLINE_TABLE;
    if (he)
    {
        // u(r) on interval j of a uniform grid: sum_b u~[b0+b] beta_b(res), res = (r - t_lo)/h, with the cubic B-spline
        // pieces of HeBulk.cpp:483-486 as polynomials in res; stored in powers of s = res h; knots relative to rijSplit
        static const long double beta[4][4] = { { 1.0L / 6, -3.0L / 6, 3.0L / 6, -1.0L / 6 },
                                                { 4.0L / 6, 0.0L, -6.0L / 6, 3.0L / 6 },
                                                { 1.0L / 6, 3.0L / 6, 3.0L / 6, -3.0L / 6 },
                                                { 0.0L, 0.0L, 0.0L, 1.0L / 6 } };
        const int n_short_iv = h->n_short - 3;
        for (int rec = 0; rec < h->nbins; rec++)
        {
            const bool lng = h->kind == TDVMC_SYSTEM_HE_DROP && rec >= n_short_iv;
            const long double hh = lng ? h->he_hl : h->h;
            const int b0 = lng ? h->n_short + (rec - n_short_iv) : rec;
            const double t_lo = lng ? (h->r_split2 - h->he_rs) + (rec - n_short_iv) * h->he_hl : rec * h->h;
            long double C[4] = { 0, 0, 0, 0 };
            for (int b = 0; b < 4; b++)
                for (int c = 0; c < 4; c++) C[c] += (long double)utR[b0 + b] * beta[b][c];
            cub[(size_t)rec * 2 + 0] = (double)C[0];
            cub[(size_t)rec * 2 + 1] = (double)(C[1] / hh);
            cub[(size_t)(nrec + rec) * 2 + 0] = (double)(C[2] / (hh * hh));
            cub[(size_t)(nrec + rec) * 2 + 1] = (double)(C[3] / (hh * hh * hh));
            cub[(size_t)(2 * nrec + rec) * 2 + 0] = t_lo;
            cub[(size_t)(2 * nrec + rec) * 2 + 1] = t_lo + (double)hh;
        }
        h->u_core = utR[K];
        h->u_const = utR[K + 1];
        h->u_lin = utR[K + 2];
    }
    for (int rec = 0; !he && rec < h->nbins; rec++)
    {
        const int b = fb + rec;
        long double C[4] = { 0, 0, 0, 0 };
        for (int p = 0; p < 4; p++)
            for (int c = 0; c < 4; c++)
                C[c] += (long double)utR[b - p] * (long double)h->weights[((size_t)(b - p) * 4 + p) * 4 + c];
        const long double t0 = h->knots[b];
        cub[(size_t)rec * 2 + 0] = (double)(C[0] + t0 * (C[1] + t0 * (C[2] + t0 * C[3])));
        cub[(size_t)rec * 2 + 1] = (double)(C[1] + t0 * (2 * C[2] + 3 * t0 * C[3]));
        cub[(size_t)(nrec + rec) * 2 + 0] = (double)(C[2] + 3 * t0 * C[3]);
        cub[(size_t)(nrec + rec) * 2 + 1] = (double)C[3];
        cub[(size_t)(2 * nrec + rec) * 2 + 0] = h->knots[b];
        cub[(size_t)(2 * nrec + rec) * 2 + 1] = h->knots[b + 1];
    }
    cub[(size_t)h->nbins * 2 + 0] = h->tail_param >= 0 ? h->uR[h->tail_param] : 0.0;
    cub[(size_t)(2 * nrec + h->nbins) * 2 + 0] = he ? (h->kind == TDVMC_SYSTEM_HE_BULK ? h->L / 2.0 : h->r_tail) - h->he_rs : h->knots[K];
    cub[(size_t)(2 * nrec + h->nbins) * 2 + 1] = 1e300;
    CK(upload(h->d_uR, h->uR, h->stream));
    CK(upload(h->d_uI, h->uI, h->stream));
    CK(upload(h->d_utR, utR, h->stream));
    CK(upload(h->d_utI, utI, h->stream));
    CK(upload(h->d_cub, cub, h->stream));
    CK(cudaStreamSynchronize(h->stream)); // host vectors go out of scope
    return 0;
}

int need_params(tdvmc_gpu_handle* h)
{
    if (!h->params_set) return fail(h, "tdvmc_gpu_set_params must be called first");
    return 0;
}

} // namespace

SysDev tdvmc_gpu_handle::sysdev() const
{
    SysDev s;
    memset(&s, 0, sizeof(s));
    s.N = N; s.Np = Np; s.P = P; s.K = K;
    s.pair_rule = pair_rule; s.tail_param = tail_param; s.n_other = n_other;
    s.first_bin = first_bin; s.nbins = nbins; s.ncell = ncell; s.uniform = uniform;
    s.L = L; s.Linv = L > 0.0 ? 1.0 / L : 0.0; s.Lhalf = L / 2.0; // src/TDVMC.cpp:535-536
    s.kind = kind; s.n_ext = n_ext; s.gr_bins = gr_bins;
    s.dim = dim; s.dm1 = dim - 1.0;
    s.bin_guard = bin_guard;
    s.periodic = periodic; s.n_short = n_short; s.potential = potential; s.rho_bins = rho_bins; s.use_phi = use_phi;
    s.rmax = kind == TDVMC_SYSTEM_HE_BULK ? L / 2.0 : ((kind != TDVMC_SYSTEM_SPLINE_TABLE && kind != TDVMC_SYSTEM_BOX_RADIAL) ? 1e300 : knots[K]); // HeBulk.cpp:54
    s.n_types = n_types;
    s.spline_order = mix_order;
    s.pair_type = d_mix_pair_type.p; s.hbar_n = d_mix_hbar.p; s.mass_n = d_mix_mass.p; s.t_knots = d_mix_knots.p;
    s.t_weights = d_mix_weights.p; s.t_mcm = d_mix_mcm.p; s.t_pot = d_mix_pot.p; s.t_cub = d_mix_cub.p;
    s.hbar = hbar;
    s.r0 = he_rs; s.core_m = core_m; s.r_split2 = r_split2; s.r_tail = r_tail; s.h_large = he_hl; s.gr_max = gr_max;
    s.u_core = u_core; s.u_const = u_const; s.u_lin = u_lin;
    s.g0R = 0.0; s.g0I = 0.0;
    if (params_set)
        for (int p = 0; p < P; p++)
        {
            s.g0R += uR[p] * grad_const[p];
            s.g0I += uI[p] * grad_const[p];
        }
    if (kind == TDVMC_SYSTEM_SPLINE_TABLE || kind == TDVMC_SYSTEM_BOX_RADIAL) potential_ab(this, s.pot_a, s.pot_b);
    s.ugR = d_ugR.p; s.ugI = d_ugI.p; s.gr_vol = d_gr_vol.p; s.gr_spacing = gr_spacing;
    if (kind == TDVMC_SYSTEM_INH_CONTACT)
    {
        s.rmax = knots[(size_t)n_short + 4 + (K - n_short)]; // pc.nodes[size - 4] (InhContactBosons.cpp:100)
        s.pot_a = sys_params[0];
        s.pot_b = sys_params[1];
        s.ext_k = sys_params[2];
        s.ext_v0 = sys_params[3];
        s.gamma = u_core;
        s.exp_const = u_const;
    }
    s.phiR = phiR;
    s.inv_cell = kind == TDVMC_SYSTEM_HE_DROP ? ncell / (r_tail - he_rs) : ncell / s.rmax;
    s.h = h; s.inv_h = 1.0 / h;
    s.u_tail = (params_set && tail_param >= 0) ? uR[tail_param] : 0.0;
    if (tail_param < 0) s.tail_param = 0; // kernels index uR[tail_param]; the tail count is zero for these systems
    s.knots = d_knots.p; s.rec = d_rec.p; s.cub = d_cub.p; s.lut = d_lut.p;
    s.map_ptr = d_map_ptr.p; s.map_col = d_map_col.p; s.map_val = d_map_val.p;
    s.uR = d_uR.p; s.uI = d_uI.p; s.utR = d_utR.p; s.utI = d_utI.p; s.map_const = d_map_const.p;
    return s;
}

extern "C" {

int tdvmc_gpu_abi_version(void) { return TDVMC_GPU_ABI_VERSION; }

int tdvmc_gpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* tdvmc_gpu_last_error(const tdvmc_gpu_handle* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int tdvmc_gpu_create(const tdvmc_system_desc* sd, const tdvmc_ensemble_desc* ed, tdvmc_gpu_handle** out)
{
    if (!sd || !ed || !out)
    {
        g_create_error = "null argument";
        return -1;
    }
    *out = nullptr;
    if (sd->struct_size != sizeof(tdvmc_system_desc) || ed->struct_size != sizeof(tdvmc_ensemble_desc))
    {
        g_create_error = "struct_size mismatch (ABI version)";
        return -1;
    }
    if ((sd->dim != 3 && !(sd->dim == 1 && sd->system_kind == TDVMC_SYSTEM_INH_CONTACT) &&
         !((sd->dim == 1 || sd->dim == 2) && (sd->system_kind == TDVMC_SYSTEM_SPLINE_TABLE || sd->system_kind == TDVMC_SYSTEM_BOX_RADIAL))) ||
        sd->n_particles < 2 || sd->n_params < 1 || sd->n_splines < 4 || ed->n_walkers < 1 || sd->n_other < 3 ||
        sd->tail_param < -1 || sd->tail_param >= sd->n_params || (!(sd->lbox > 0.0) && sd->system_kind != TDVMC_SYSTEM_HE_DROP && sd->system_kind != TDVMC_SYSTEM_MIXTURE) || sd->n_ext < sd->n_splines ||
        (sd->system_kind != TDVMC_SYSTEM_SPLINE_TABLE && sd->system_kind != TDVMC_SYSTEM_HE_BULK &&
         sd->system_kind != TDVMC_SYSTEM_HE_DROP && sd->system_kind != TDVMC_SYSTEM_MIXTURE &&
         sd->system_kind != TDVMC_SYSTEM_BOX_RADIAL && sd->system_kind != TDVMC_SYSTEM_INH_CONTACT) ||
        (sd->system_kind == TDVMC_SYSTEM_INH_CONTACT &&
         (sd->dim != 1 || !sd->knots || !sd->spline_weights || sd->n_ext != sd->n_splines || sd->n_splines_first < 4 ||
          sd->n_splines - sd->n_splines_first < 4 || sd->n_system_params != 4 || sd->n_other != 9 || sd->n_particles > 32 ||
          sd->n_ext > 96)) ||
        (sd->system_kind == TDVMC_SYSTEM_BOX_RADIAL &&
         (!sd->knots || !sd->spline_weights || sd->n_ext != 2 * sd->n_splines || (sd->n_params & 1) ||
          sd->n_splines != sd->n_params / 2 + 3 || sd->n_other < 4 || sd->n_system_params < 2)) ||
        (sd->system_kind == TDVMC_SYSTEM_MIXTURE &&
         (!sd->mixture || sd->mixture->n_pair_types < 1 || sd->n_ext != sd->mixture->n_pair_types * (sd->n_splines + 4) ||
          sd->n_ext > 96 || sd->n_particles > 8 || sd->n_other < 6)) ||
        (sd->system_kind == TDVMC_SYSTEM_SPLINE_TABLE && (!sd->knots || !sd->spline_weights || sd->n_other < 9)) ||
        (sd->system_kind == TDVMC_SYSTEM_HE_BULK && (sd->n_ext != sd->n_splines + 3 || sd->n_other != 103)) ||
        (sd->system_kind == TDVMC_SYSTEM_HE_DROP && (sd->n_ext != sd->n_splines + 3 || sd->n_other != 403)))
    {
        g_create_error = "invalid system/ensemble description";
        return -1;
    }
    if (sd->n_params + 3 > 8 * kAccMaxTiles)
    {
        g_create_error = "N_PARAM + 3 exceeds the 208 columns of the accumulation kernel";
        return -1;
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev <= 0)
    {
        g_create_error = std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0") +
                         " (this library has no CPU fallback)";
        return -2;
    }
    tdvmc_gpu_handle* h = new tdvmc_gpu_handle();
    auto bail = [&](int code) {
        g_create_error = h->error;
        tdvmc_gpu_destroy(h);
        return code;
    };
    h->device = ed->device;
    if (cudaSetDevice(h->device) != cudaSuccess)
    {
        h->error = "cudaSetDevice failed";
        return bail(-2);
    }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    if (cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device) != cudaSuccess) h->smem_optin = 48 * 1024;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
    {
        h->error = "cudaStreamCreate failed";
        return bail(-2);
    }
    h->N = sd->n_particles;
    h->Np = (h->N + 3) & ~3;
    h->P = sd->n_params;
    h->K = sd->n_splines;
    h->pair_rule = sd->pair_rule;
    h->tail_param = sd->tail_param;
    h->n_other = sd->n_other;
    h->L = sd->lbox;
    h->hbar = sd->hbar2_2m;
    h->kind = sd->system_kind;
    h->dim = sd->dim;
    h->n_ext = sd->n_ext;
    if (sd->knots) h->knots.assign(sd->knots, sd->knots + h->K + (sd->system_kind == TDVMC_SYSTEM_INH_CONTACT ? 8 : 4));
    if (sd->system_kind == TDVMC_SYSTEM_INH_CONTACT) h->n_short = sd->n_splines_first;
    if (sd->spline_weights) h->weights.assign(sd->spline_weights, sd->spline_weights + (size_t)h->K * 16);
    if (h->kind == TDVMC_SYSTEM_MIXTURE)
    {
        const tdvmc_mixture_desc* m = sd->mixture;
        const int T = m->n_pair_types, N = h->N, K = h->K;
        if (m->spline_order != 0 && m->spline_order != 3 && m->spline_order != 4)
        {
            h->error = "mixture: spline_order must be 3 or 4";
            return bail(-1);
        }
        h->mix_order = m->spline_order == 4 ? 4 : 3;
        const int ord = h->mix_order;
        if (K < 2 * ord)
        {
            h->error = "mixture: too few splines for the spline order";
            return bail(-1);
        }
        h->n_types = T;
        h->mix_pair_type.assign(m->pair_type, m->pair_type + (size_t)N * N);
        h->mix_pot.assign(m->potential, m->potential + T);
        h->mix_hbar.assign(m->hbar_over_2m, m->hbar_over_2m + N);
        h->mix_mass.assign(m->mass, m->mass + N);
        h->mix_knots.assign(m->knots, m->knots + (size_t)T * (K + ord + 1));
        h->mix_weights.assign(m->spline_weights, m->spline_weights + (size_t)T * K * (ord + 1) * (ord + 1));
        h->mix_mcm.assign(m->mcmillan_factor, m->mcmillan_factor + T);
        for (int t : h->mix_pair_type)
            if (t < 0 || t >= T)
            {
                h->error = "pair type out of range";
                return bail(-1);
            }
    }
    h->map_const.assign(h->P, 0.0);
    h->grad_const.assign(h->P, 0.0);
    if (sd->map_const) h->map_const.assign(sd->map_const, sd->map_const + h->P);
    if (sd->grad_const) h->grad_const.assign(sd->grad_const, sd->grad_const + h->P);
    h->map_ptr.assign(sd->map_ptr, sd->map_ptr + h->P + 1);
    const int nnz = h->map_ptr[h->P];
    h->map_col.assign(sd->map_col, sd->map_col + nnz);
    h->map_val.assign(sd->map_val, sd->map_val + nnz);
    for (int c : h->map_col)
        if (c < 0 || c >= h->n_ext)
        {
            h->error = "boundary map column out of range";
            return bail(-1);
        }
    h->sys_params.assign(sd->system_params, sd->system_params + sd->n_system_params);
    if (h->kind != TDVMC_SYSTEM_INH_CONTACT && h->sys_params.size() > 2 && h->sys_params.size() < 5)
    {
        h->error = "SYSTEM_PARAMS needs 2 or 5 entries";
        return bail(-1);
    }
    h->uR.assign(h->P, 0.0);
    h->uI.assign(h->P, 0.0);

    h->W = ed->n_walkers;
    h->first_walker = ed->first_walker;
    h->max_samples = std::max(1, ed->max_samples_per_walker);
    h->keep_positions = ed->keep_sample_positions;
    h->seed = ed->seed;
    h->mc_step = ed->mc_step;

    int rc = build_static_tables(h);
    if (rc) return bail(rc);

    // sample store: augmented rows [O | E^R | E^I | 1], see accumulate.cu
    const int ncols = h->P + 3;
    h->ldc = 16 * ((ncols + 15) / 16);
    h->lda = h->ldc + 4;
    const long long rows = (long long)h->W * h->max_samples;
    h->rows_cap = ((rows + kAccChunkRows - 1) / kAccChunkRows) * kAccChunkRows;
    h->est_len = (size_t)h->P * h->P + 3 * (size_t)h->P + 2 + h->n_other + 3;
    const size_t scratch = (size_t)h->sm_count * h->ldc * h->ldc + (size_t)h->ldc * h->ldc + (size_t)148 * h->n_other + 2; // (+ the acceptance total)
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t x) {
        if (e == cudaSuccess) e = x;
    };
    A(h->d_pos.alloc((size_t)h->W * 3 * h->Np));
    A(h->d_aos.alloc((size_t)h->W * 3 * h->N));
    A(h->d_accepted.alloc(h->W));
    A(h->d_A.alloc((size_t)h->rows_cap * h->lda));
    A(h->d_other.alloc((size_t)h->rows_cap * h->n_other));
    A(h->d_exponent.alloc((size_t)h->rows_cap));
    A(h->d_est.alloc(h->est_len));
    A(h->d_scratch.alloc(scratch));
    if (h->keep_positions) A(h->d_samp_pos.alloc((size_t)rows * 3 * h->Np));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&h->h_est, h->est_len * sizeof(double));
    if (e != cudaSuccess)
    {
        h->error = std::string("device allocation failed: ") + cudaGetErrorString(e);
        return bail(-3);
    }
    cudaMemsetAsync(h->d_pos.p, 0, h->d_pos.n * sizeof(double), h->stream);
    cudaMemsetAsync(h->d_accepted.p, 0, h->d_accepted.n * sizeof(unsigned long long), h->stream);
    cudaMemsetAsync(h->d_A.p, 0, h->d_A.n * sizeof(double), h->stream);
    cudaMemsetAsync(h->d_other.p, 0, h->d_other.n * sizeof(double), h->stream);
    cudaMemsetAsync(h->d_est.p, 0, h->d_est.n * sizeof(double), h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess)
    {
        h->error = "initialisation failed";
        return bail(-3);
    }

    if (h->kind == TDVMC_SYSTEM_MIXTURE || h->kind == TDVMC_SYSTEM_INH_CONTACT)
    {
        h->npp = (h->N + 1) & ~1;
        h->wpb = 1;
        h->resident_per_sm = 1024; // one thread per walker
        *out = h;
        return 0;
    }
    if (h->kind == TDVMC_SYSTEM_BOX_RADIAL)
    {
        // one warp per walker, 8 (or as many as fit) walkers per block, three blocks per SM
        h->npp = (h->N + 1) & ~1;
        h->wpb = 8;
        const SysDev sb = h->sysdev();
        while (h->wpb > 1 && !sweep_br_fits(sb, h->wpb, h->npp, h->smem_optin)) h->wpb /= 2;
        if (!sweep_br_fits(sb, h->wpb, h->npp, h->smem_optin))
        {
            h->error = "system does not fit the sweep kernel's shared memory";
            return bail(-3);
        }
        h->resident_per_sm = 24;
        *out = h;
        return 0;
    }
    // sweep geometry: as many walkers (warps) per block as keep >= 2 blocks per SM resident
    h->npp = (h->N + 1) & ~1;
    SysDev s = h->sysdev();
    if (h->kind == TDVMC_SYSTEM_SPLINE_TABLE && evaluate_smem_bytes(s) > (size_t)h->smem_optin)
    {
        // (tables + per-warp histograms alone exceed shared memory: thousands of splines; the configuration itself moves to
        // global memory for large N, see evaluate_scratch_doubles)
        char msg[256];
        snprintf(msg, sizeof(msg), "%d splines need %zu bytes of shared memory in the evaluation kernel, the device offers %d: not supported",
                 h->K, evaluate_smem_bytes(s), h->smem_optin);
        h->error = msg;
        return bail(-3);
    }
    // Walkers (warps) per block: as many as fit, but a multiple of 4 per SM -- the four SM sub-partitions
    // each run resident/4 warps and the slowest one sets the pace (measured: 20 walkers/SM beat 21 by 13 %).
    int best_wpb = 1, best_res = 0, best_score = -1;
    for (int wpb = 1; wpb <= sweep_max_threads(s) / 32; wpb++)
    {
        const int res = sweep_blocks_per_sm(s, wpb, h->npp) * wpb; // warps resident per SM
        const int score = res >= 4 ? (res / 4) * 4 * 2 + (res % 4 == 0 ? 1 : 0) : res;
        if (score >= best_score && res > 0) // ties: the larger block shares one copy of the coefficient planes
        {
            best_score = score;
            best_res = res;
            best_wpb = wpb;
        }
    }
    if (const char* e = getenv("TDVMC_SWEEP_WPB")) // tuning knob
    {
        const int wpb = atoi(e);
        if (wpb >= 1 && wpb <= sweep_max_threads(s) / 32 && sweep_blocks_per_sm(s, wpb, h->npp) > 0)
        {
            best_wpb = wpb;
            best_res = sweep_blocks_per_sm(s, wpb, h->npp) * wpb;
        }
    }
    h->resident_per_sm = best_res * sweep_walkers_per_warp(s); // walkers
    if (best_res == 0 && h->kind == TDVMC_SYSTEM_SPLINE_TABLE && s.dim == 3 && sweep_large_fits(s, h->npp, h->smem_optin))
    {
        // one walker's positions alone nearly fill an SM (N = 8000: 192 KB): eight warps share the walker, four instead of
        // eight replicas of the coefficient planes make room (sweep_split_kernel<..., 8, 4>)
        h->large_sweep = true;
        h->resident_per_sm = 1;
        best_res = 1;
        best_wpb = 1;
    }
    if (best_res == 0)
    {
        char msg[256];
        snprintf(msg, sizeof(msg), "N = %d particles: one walker (%zu bytes of positions + the coefficient planes) does not fit the sweep "
                 "kernel's shared memory (%d bytes): not supported", h->N, (size_t)3 * h->npp * sizeof(double), h->smem_optin);
        h->error = msg;
        return bail(-3);
    }
    h->wpb = best_wpb;
    *out = h;
    return 0;
}

void tdvmc_gpu_destroy(tdvmc_gpu_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    for (auto& t : h->pending)
    {
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
    }
    for (auto e : h->event_pool) cudaEventDestroy(e);
    if (h->timer0) cudaEventDestroy(h->timer0);
    if (h->timer1) cudaEventDestroy(h->timer1);
    if (h->h_est) cudaFreeHost(h->h_est);
    if (h->h_sol) cudaFreeHost(h->h_sol);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int tdvmc_gpu_set_positions(tdvmc_gpu_handle* h, const double* R, int32_t first, int32_t n)
{
    if (!h || !R || first < 0 || n < 0 || first + n > h->W) return h ? fail(h, "set_positions: bad range") : -1;
    CK(cudaSetDevice(h->device));
    const size_t cnt = (size_t)n * h->N * 3;
    CK(cudaMemcpyAsync(h->d_aos.p, R, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(launch_transpose_in(h->d_aos.p, h->d_pos.p + (size_t)first * 3 * h->Np, n, h->N, h->Np, h->stream));
    h->n_launched++;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int tdvmc_gpu_get_positions(tdvmc_gpu_handle* h, double* R, int32_t first, int32_t n)
{
    if (!h || !R || first < 0 || n < 0 || first + n > h->W) return h ? fail(h, "get_positions: bad range") : -1;
    CK(cudaSetDevice(h->device));
    const size_t cnt = (size_t)n * h->N * 3;
    CK(launch_transpose_out(h->d_pos.p + (size_t)first * 3 * h->Np, h->d_aos.p, n, h->N, h->Np, h->stream));
    h->n_launched++;
    CK(cudaMemcpyAsync(R, h->d_aos.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int tdvmc_gpu_set_params(tdvmc_gpu_handle* h, const double* uR, const double* uI, double phiR, double phiI, double time)
{
    if (!h || !uR || !uI) return h ? fail(h, "set_params: null argument") : -1;
    CK(cudaSetDevice(h->device));
    h->uR.assign(uR, uR + h->P);
    h->uI.assign(uI, uI + h->P);
    h->phiR = phiR;
    h->phiI = phiI;
    h->time = time;
    h->params_set = true;
    return build_param_tables(h);
}

int tdvmc_gpu_wrap_positions(tdvmc_gpu_handle* h)
{
    if (!h) return -1;
    CK(cudaSetDevice(h->device));
    Timed t(h, TDVMC_KERNEL_OTHER);
    if (h->kind == TDVMC_SYSTEM_MIXTURE) CK(launch_com_mix(h->sysdev(), h->d_pos.p, h->W, h->stream));
    else CK(launch_wrap(h->sysdev(), h->d_pos.p, h->W, h->stream));
    return 0;
}

static int do_sweep(tdvmc_gpu_handle* h, long long n_steps, double* pos = nullptr)
{
    if (n_steps <= 0) return 0;
    SweepArgs a;
    a.s = h->sysdev();
    a.pos = pos ? pos : h->d_pos.p;
    a.accepted = h->d_accepted.p;
    a.W = h->W;
    a.first_walker = h->first_walker;
    a.wpb = h->wpb;
    a.npp = h->npp;
    a.pos_offset = 0;
    a.seed = h->seed;
    a.first_step = h->step_counter;
    a.n_steps = n_steps;
    a.mc_step = h->mc_step;
    {
        Timed t(h, TDVMC_KERNEL_SWEEP);
        const int split = h->large_sweep ? -8
                          : h->kind == TDVMC_SYSTEM_SPLINE_TABLE ? sweep_split_warps(a.s, h->W, h->sm_count, h->resident_per_sm) : 1;
        CK(h->kind == TDVMC_SYSTEM_MIXTURE ? launch_sweep_mix(a, h->stream)
           : h->kind == TDVMC_SYSTEM_INH_CONTACT ? launch_sweep_inh(a, h->stream)
           : h->kind == TDVMC_SYSTEM_BOX_RADIAL ? launch_sweep_br(a, h->stream)
           : split < 0 ? launch_sweep_split(a, -split, h->sm_count, h->smem_optin, h->stream, 4)
           : split > 1 ? launch_sweep_split(a, split, h->sm_count, h->smem_optin, h->stream)
           : (h->kind == TDVMC_SYSTEM_SPLINE_TABLE && sweep_queue_wanted(a.s, h->W, h->sm_count, h->resident_per_sm, n_steps))
               ? launch_sweep_queue(a, h->sm_count, h->smem_optin, h->stream)
               : launch_sweep(a, h->stream));
    }
    h->step_counter += (uint64_t)n_steps;
    h->trials_local += (uint64_t)n_steps * (uint64_t)h->W;
    return 0;
}

int tdvmc_gpu_reset_counters(tdvmc_gpu_handle* h)
{
    if (!h) return -1;
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(h->d_accepted.p, 0, h->d_accepted.n * sizeof(unsigned long long), h->stream));
    h->trials_local = 0;
    return 0;
}

int tdvmc_gpu_set_mc_step(tdvmc_gpu_handle* h, double mc_step)
{
    if (!h) return -1;
    if (!(mc_step > 0.0) || !std::isfinite(mc_step)) return fail(h, "set_mc_step: MC_STEP must be positive and finite");
    h->mc_step = mc_step;
    return 0;
}

int tdvmc_gpu_sweep(tdvmc_gpu_handle* h, int64_t n_steps)
{
    if (!h) return -1;
    if (int rc = need_params(h)) return rc;
    CK(cudaSetDevice(h->device));
    return do_sweep(h, n_steps);
}

static cudaError_t launch_evaluate_any(int kind, const EvalArgs& a, cudaStream_t st)
{
    if (kind == TDVMC_SYSTEM_MIXTURE) return launch_evaluate_mix(a, st);
    if (kind == TDVMC_SYSTEM_BOX_RADIAL) return launch_evaluate_br(a, st);
    if (kind == TDVMC_SYSTEM_INH_CONTACT) return launch_evaluate_inh(a, st);
    return kind != TDVMC_SYSTEM_SPLINE_TABLE ? launch_evaluate_he(a, st) : launch_evaluate(a, st);
}

// systems too large for one SM's shared memory: the evaluation kernel keeps the configuration in flight in a per-block slab
static int eval_slab(tdvmc_gpu_handle* h, EvalArgs& a)
{
    a.scratch = nullptr;
    if (h->kind != TDVMC_SYSTEM_SPLINE_TABLE) return 0;
    const size_t n = evaluate_scratch_doubles(a.s, h->sm_count);
    if (n == 0) return 0;
    CK(h->d_eval_slab.ensure(n));
    a.scratch = h->d_eval_slab.p;
    return 0;
}

static int do_evaluate_walkers(tdvmc_gpu_handle* h, const double* pos, int n_cfg, long long row0)
{
    EvalArgs a;
    memset(&a, 0, sizeof(a));
    a.s = h->sysdev();
    a.pos = pos;
    a.n_cfg = n_cfg;
    a.A = h->d_A.p;
    a.lda = h->lda;
    a.row0 = row0;
    a.row_stride = 1;
    a.other = h->d_other.p;
    a.exponent = h->d_exponent.p;
    if (int rc = eval_slab(h, a)) return rc;
    Timed t(h, TDVMC_KERNEL_EVALUATE);
    CK(launch_evaluate_any(h->kind, a, h->stream));
    return 0;
}

static int do_accumulate(tdvmc_gpu_handle* h, const double* A, const double* other, long long M)
{
    const long long rows_pad = ((M + kAccChunkRows - 1) / kAccChunkRows) * kAccChunkRows;
    AccArgs a;
    a.A = A;
    a.lda = h->lda;
    a.ncols = h->P + 3;
    a.n_chunks = rows_pad / kAccChunkRows;
    a.n_cta = (int)std::max(1ll, std::min((long long)h->sm_count, a.n_chunks / 4));
    a.partial = h->d_scratch.p;
    a.ldc = h->ldc;
    AccFinishArgs f;
    f.partial = h->d_scratch.p;
    f.n_cta = a.n_cta;
    f.ldc = h->ldc;
    f.P = h->P;
    f.other = other;
    f.M = M;
    f.n_other = h->n_other;
    f.accepted = h->d_accepted.p;
    f.W = h->W;
    f.n_trials = (double)h->trials_local;
    f.est = h->d_est.p;
    {
        Timed t(h, TDVMC_KERNEL_ACCUMULATE);
        CK(launch_accumulate(a, h->stream));
    }
    {
        Timed t(h, TDVMC_KERNEL_OTHER, 4);
        CK(launch_acc_finish(f, h->stream));
    }
    h->rows_used = M;
    h->est_valid = true;
    h->est_reduced = false;
    return 0;
}

int tdvmc_gpu_sample_and_accumulate(tdvmc_gpu_handle* h, int32_t n_samples, int32_t n_therm, int32_t n_init)
{
    if (!h) return -1;
    if (int rc = need_params(h)) return rc;
    if (n_samples < 1 || n_samples > h->max_samples || n_therm < 0 || n_init < 0)
        return fail(h, "sample_and_accumulate: n_samples exceeds max_samples_per_walker or negative step count");
    CK(cudaSetDevice(h->device));
    const long long M = (long long)n_samples * h->W;
    const long long rows_pad = ((M + kAccChunkRows - 1) / kAccChunkRows) * kAccChunkRows;
    if (rows_pad > M) // rows beyond M must be zero for the padded SYRK chunks
        CK(cudaMemsetAsync(h->d_A.p + (size_t)M * h->lda, 0, (size_t)(rows_pad - M) * h->lda * sizeof(double), h->stream));
    if (int rc = do_sweep(h, n_init)) return rc; // MC_NINITIALIZATIONSTEPS, src/TDVMC.cpp:1068-1071
    // One block evaluates one configuration, so a launch over W configurations runs in ceil(W / blocks per wave) waves:
    // 512 walkers of N = 1728 (one block per SM) are 3.46 waves, 14 % of the last one idle.  Where that loss exceeds a few per
    // cent - or the sample positions are kept anyway - the positions are snapshot per sample and all n_samples x W
    // configurations go through ONE launch after the last sweep (same rows, same arithmetic).
    bool batch = false;
    if (h->kind == TDVMC_SYSTEM_SPLINE_TABLE && n_samples > 1)
    {
        const double waves = (double)h->W / (double)(h->sm_count * evaluate_blocks_per_sm(h->sysdev()));
        batch = h->keep_positions || std::ceil(waves) / waves > 1.04;
        if (batch && !h->d_samp_pos.p)
        {
            if (h->d_samp_pos.alloc((size_t)h->rows_cap * 3 * h->Np) != cudaSuccess)
            {
                cudaGetLastError();
                batch = false; // no room for the snapshots: evaluate sample by sample
            }
        }
    }
    for (int m = 0; m < n_samples; m++)
    {
        if (int rc = do_sweep(h, n_therm)) return rc; // src/TDVMC.cpp:1075-1078
        if (!batch)
            if (int rc = do_evaluate_walkers(h, h->d_pos.p, h->W, (long long)m * h->W)) return rc; // :1080
        if (h->keep_positions || batch)                                                        // :1082-1091
            CK(cudaMemcpyAsync(h->d_samp_pos.p + (size_t)m * h->W * 3 * h->Np, h->d_pos.p, (size_t)h->W * 3 * h->Np * sizeof(double),
                               cudaMemcpyDeviceToDevice, h->stream));
    }
    if (batch)
        if (int rc = do_evaluate_walkers(h, h->d_samp_pos.p, (int)M, 0)) return rc;
    if (h->keep_positions) h->stored_samples = n_samples; // the update cursor keeps running, as :556 sets it only once
    return do_accumulate(h, h->d_A.p, h->d_other.p, M); // :1103-1109
}

int tdvmc_gpu_reevaluate_stored(tdvmc_gpu_handle* h)
{
    if (!h) return -1;
    if (int rc = need_params(h)) return rc;
    if (!h->keep_positions || h->stored_samples < 1) return fail(h, "reevaluate_stored: no stored samples (keep_sample_positions)");
    CK(cudaSetDevice(h->device));
    const long long M = (long long)h->stored_samples * h->W;
    if (int rc = do_evaluate_walkers(h, h->d_samp_pos.p, (int)M, 0)) return rc; // src/TDVMC.cpp:1244-1262
    return do_accumulate(h, h->d_A.p, h->d_other.p, M);
}

int tdvmc_gpu_update_stored(tdvmc_gpu_handle* h, int32_t n_update, int32_t n_therm)
{
    if (!h) return -1;
    if (int rc = need_params(h)) return rc;
    if (!h->keep_positions || h->stored_samples < 1) return fail(h, "update_stored: no stored samples (keep_sample_positions)");
    if (n_update < 0 || n_therm < 0) return fail(h, "update_stored: negative count");
    CK(cudaSetDevice(h->device));
    for (int i = 0; i < n_update; i++) // UpdateSamplesConsecutive, src/TDVMC.cpp:975-983
    {
        const int slot = h->update_cursor % h->stored_samples;
        // UpdateSample (:948-961): MC_NTHERMSTEPS Metropolis steps on the stored configuration at the current
        // parameters; the stored tables are recomputed from R by reevaluate_stored, so nothing else is refreshed here
        if (int rc = do_sweep(h, n_therm, h->d_samp_pos.p + (size_t)slot * h->W * 3 * h->Np)) return rc;
        h->update_cursor = (slot + 1) % h->stored_samples;
    }
    return 0;
}

// One packed in-place all-reduce per accumulation (an in-place sum must not be applied twice): fetch and solve share it.
static int ensure_reduced(tdvmc_gpu_handle* h)
{
    if (h->comm && !h->est_reduced)
    {
        int rc = g_nccl.AllReduce(h->d_est.p, h->d_est.p, h->est_len, kNcclDouble, kNcclSum, h->comm, h->stream);
        if (rc != 0) return fail(h, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"), rc);
    }
    h->est_reduced = true;
    return 0;
}

int tdvmc_gpu_allreduce_and_fetch(tdvmc_gpu_handle* h, tdvmc_estimators* out)
{
    if (!h || !out) return h ? fail(h, "allreduce_and_fetch: null argument") : -1;
    if (!h->est_valid) return fail(h, "allreduce_and_fetch: nothing accumulated yet");
    CK(cudaSetDevice(h->device));
    if (int rc = ensure_reduced(h)) return rc;
    CK(cudaMemcpyAsync(h->h_est, h->d_est.p, h->est_len * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const int P = h->P;
    const double* S = h->h_est;
    const double* FR = S + (size_t)P * P;
    const double* FI = FR + P;
    const double* O = FI + P;
    const double* E = O + P;
    const double* oth = E + 2;
    const double* cnt = oth + h->n_other;
    const double n = cnt[2];
    if (!(n > 0.0)) return fail(h, "allreduce_and_fetch: zero samples");
    const double inv = 1.0 / n; // ReduceToAverage: sum over ranks / numOfProcesses (src/MPIMethods.h:199-203)
    if (out->local_operators_matrix)
        for (size_t i = 0; i < (size_t)P * P; i++) out->local_operators_matrix[i] = S[i] * inv;
    for (int k = 0; k < P; k++)
    {
        if (out->local_operator_energy_r) out->local_operator_energy_r[k] = FR[k] * inv;
        if (out->local_operator_energy_i) out->local_operator_energy_i[k] = FI[k] * inv;
        if (out->local_operators) out->local_operators[k] = O[k] * inv;
    }
    if (out->local_energy_r) *out->local_energy_r = E[0] * inv;
    if (out->local_energy_i) *out->local_energy_i = E[1] * inv;
    if (out->other_expectation_values)
        for (int k = 0; k < h->n_other; k++) out->other_expectation_values[k] = oth[k] * inv;
    out->n_acceptances = (int64_t)llround(cnt[0]);
    out->n_trials = (int64_t)llround(cnt[1]);
    out->n_samples = (int64_t)llround(n);
    return 0;
}

// ---- parameter derivatives on the device (solve.cu) ----

static int do_solve(tdvmc_gpu_handle* h, const tdvmc_solver_desc* sd, const double* d_est, tdvmc_parameters_dot* out)
{
    if (sd->struct_size != sizeof(tdvmc_solver_desc)) return fail(h, "solver desc: struct_size mismatch");
    if (sd->imaginary_time < -1 || sd->imaginary_time > 1) return fail(h, "solver: IMAGINARY_TIME must be -1, 0 or 1");
    if (sd->solver_type != 0 && sd->solver_type != 1) return fail(h, "solver: LINEAR_EQUATION_SOLVER_TYPE must be 0 (Cholesky) or 1 (QR)");
    if (h->P > 1024) return fail(h, "solver: N_PARAM > 1024");
    const int P = h->P;
    CK(h->d_sol.ensure(2 * (size_t)P + 5));
    CK(h->d_solve_L.ensure(sd->solver_type == 1 ? (size_t)P * P : (size_t)P * (P + 1) / 2));
    if (!h->h_sol) CK(cudaMallocHost((void**)&h->h_sol, (2 * (size_t)P + 5) * sizeof(double)));
    SolveArgs a;
    memset(&a, 0, sizeof(a));
    a.est = d_est;
    a.cnt_offset = (int)(h->est_len - 3);
    a.P = P;
    a.imaginary_time = sd->imaginary_time;
    a.use_preconditioning = sd->use_preconditioning;
    a.regularization = sd->regularization;
    a.min_scaling = sd->min_scaling;
    a.L_global = h->d_solve_L.p;
    a.force_global = sd->force_global_scratch;
    a.out = h->d_sol.p;
    {
        Timed t(h, TDVMC_KERNEL_SOLVE);
        if (sd->solver_type == 1) CK(launch_solve_qr(a, h->stream));
        else CK(launch_solve(a, h->smem_optin, h->stream));
    }
    CK(cudaMemcpyAsync(h->h_sol, h->d_sol.p, (2 * (size_t)P + 5) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (out)
    {
        if (out->u_dot_r) memcpy(out->u_dot_r, h->h_sol, P * sizeof(double));
        if (out->u_dot_i) memcpy(out->u_dot_i, h->h_sol + P, P * sizeof(double));
        const double* tail = h->h_sol + 2 * (size_t)P;
        out->phi_dot_r = tail[0];
        out->phi_dot_i = tail[1];
        out->not_positive_definite = tail[2] != 0.0 ? 1 : 0;
        out->local_energy_r = tail[3];
        out->local_energy_i = tail[4];
    }
    return 0;
}

int tdvmc_gpu_solve_parameters_dot(tdvmc_gpu_handle* h, const tdvmc_solver_desc* sd, tdvmc_parameters_dot* out)
{
    if (!h || !sd || !out) return h ? fail(h, "solve_parameters_dot: null argument") : -1;
    if (!h->est_valid) return fail(h, "solve_parameters_dot: nothing accumulated yet");
    CK(cudaSetDevice(h->device));
    if (int rc = ensure_reduced(h)) return rc;
    return do_solve(h, sd, h->d_est.p, out);
}

int tdvmc_gpu_euler_step(tdvmc_gpu_handle* h, const tdvmc_solver_desc* sd, double dt, double time, double* uR, double* uI,
                         double* phiR, double* phiI, tdvmc_parameters_dot* dot)
{
    if (!h || !sd || !uR || !uI || !phiR || !phiI) return h ? fail(h, "euler_step: null argument") : -1;
    if (!h->est_valid) return fail(h, "euler_step: nothing accumulated yet");
    CK(cudaSetDevice(h->device));
    if (int rc = ensure_reduced(h)) return rc;
    tdvmc_parameters_dot local;
    memset(&local, 0, sizeof(local));
    tdvmc_parameters_dot* d = dot ? dot : &local;
    if (int rc = do_solve(h, sd, h->d_est.p, d)) return rc;
    const double* x = h->h_sol;
    for (int i = 0; i < h->P; i++) // uR = uR + uDotR * dt (src/TDVMC.cpp:1846-1847)
    {
        uR[i] = uR[i] + x[i] * dt;
        uI[i] = uI[i] + x[h->P + i] * dt;
    }
    *phiR = *phiR + d->phi_dot_r * dt;
    *phiI = *phiI + d->phi_dot_i * dt;
    return tdvmc_gpu_set_params(h, uR, uI, *phiR, *phiI, time);
}

int tdvmc_gpu_solve_fixed(tdvmc_gpu_handle* h, const tdvmc_solver_desc* sd, const tdvmc_estimators* est, tdvmc_parameters_dot* out)
{
    if (!h || !sd || !est || !out) return h ? fail(h, "solve_fixed: null argument") : -1;
    if (!est->local_operators || !est->local_operators_matrix || !est->local_operator_energy_r || !est->local_operator_energy_i ||
        !est->local_energy_r || !est->local_energy_i)
        return fail(h, "solve_fixed: estimators incomplete");
    CK(cudaSetDevice(h->device));
    const int P = h->P;
    // the packed layout with a sample count of one: sums == averages, the kernel's division by n is exact
    std::vector<double> e(h->est_len, 0.0);
    memcpy(e.data(), est->local_operators_matrix, (size_t)P * P * sizeof(double));
    double* p = e.data() + (size_t)P * P;
    memcpy(p, est->local_operator_energy_r, P * sizeof(double));
    memcpy(p + P, est->local_operator_energy_i, P * sizeof(double));
    memcpy(p + 2 * P, est->local_operators, P * sizeof(double));
    p[3 * P] = *est->local_energy_r;
    p[3 * P + 1] = *est->local_energy_i;
    e[h->est_len - 1] = 1.0;
    CK(upload(h->d_est_fixed, e, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return do_solve(h, sd, h->d_est_fixed.p, out);
}

int tdvmc_gpu_last_exponent(tdvmc_gpu_handle* h, double* exponent)
{
    if (!h || !exponent) return -1;
    if (h->rows_used < 1) return fail(h, "last_exponent: no sample evaluated yet");
    CK(cudaSetDevice(h->device));
    const long long row = h->rows_used - h->W; // first local walker, last sample
    CK(cudaMemcpyAsync(exponent, h->d_exponent.p + row, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int tdvmc_gpu_comm_unique_id(uint8_t id[TDVMC_GPU_UNIQUE_ID_BYTES])
{
    std::string err;
    if (!load_nccl(err))
    {
        g_create_error = err;
        return -1;
    }
    UniqueId u;
    memset(&u, 0, sizeof(u));
    int rc = g_nccl.GetUniqueId(&u);
    if (rc != 0)
    {
        g_create_error = "ncclGetUniqueId failed";
        return rc;
    }
    memcpy(id, u.internal, TDVMC_GPU_UNIQUE_ID_BYTES);
    return 0;
}

int tdvmc_gpu_comm_init(tdvmc_gpu_handle* h, const uint8_t id[TDVMC_GPU_UNIQUE_ID_BYTES], int32_t rank, int32_t n_ranks)
{
    if (!h || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return h ? fail(h, "comm_init: bad arguments") : -1;
    CK(cudaSetDevice(h->device));
    h->rank = rank;
    h->n_ranks = n_ranks;
    if (n_ranks == 1) return 0;
    std::string err;
    if (!load_nccl(err)) return fail(h, err);
    UniqueId u;
    memcpy(u.internal, id, TDVMC_GPU_UNIQUE_ID_BYTES);
    int rc = g_nccl.CommInitRank(&h->comm, n_ranks, u, rank);
    if (rc != 0)
    {
        h->comm = nullptr;
        return fail(h, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"), rc);
    }
    return 0;
}

// ---- fixed-configuration entry points ----

int tdvmc_gpu_evaluate_fixed(tdvmc_gpu_handle* h, const double* R, int32_t n_cfg, double* e_r, double* e_i, double* O,
                             double* other, double* exponent, double* drift_r, double* drift_i, double* spline_sums,
                             double* outer)
{
    if (!h || !R || n_cfg < 1) return h ? fail(h, "evaluate_fixed: bad arguments") : -1;
    if (int rc = need_params(h)) return rc;
    CK(cudaSetDevice(h->device));
    const int N = h->N, P = h->P, K = h->n_ext;
    DevBuf<double> aos, pos, A, oth, ex, dr, di, ss, out;
    CK(aos.alloc((size_t)n_cfg * N * 3));
    CK(pos.alloc((size_t)n_cfg * 3 * h->Np));
    CK(A.alloc((size_t)n_cfg * h->lda));
    CK(oth.alloc((size_t)n_cfg * h->n_other));
    CK(ex.alloc(n_cfg));
    CK(dr.alloc((size_t)n_cfg * N * 3));
    CK(di.alloc((size_t)n_cfg * N * 3));
    CK(ss.alloc((size_t)n_cfg * K));
    CK(out.alloc(n_cfg));
    CK(cudaMemsetAsync(pos.p, 0, pos.n * sizeof(double), h->stream));
    CK(cudaMemsetAsync(oth.p, 0, oth.n * sizeof(double), h->stream));
    CK(cudaMemcpyAsync(aos.p, R, aos.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(launch_transpose_in(aos.p, pos.p, n_cfg, N, h->Np, h->stream));
    EvalArgs a;
    memset(&a, 0, sizeof(a));
    a.s = h->sysdev();
    a.pos = pos.p;
    a.n_cfg = n_cfg;
    a.A = A.p;
    a.lda = h->lda;
    a.row0 = 0;
    a.row_stride = 1;
    a.other = oth.p;
    a.exponent = ex.p;
    a.drift_r = dr.p;
    a.drift_i = di.p;
    a.ss_out = ss.p;
    a.outer_out = out.p;
    if (int rc = eval_slab(h, a)) return rc;
    {
        Timed t(h, TDVMC_KERNEL_EVALUATE);
        CK(launch_evaluate_any(h->kind, a, h->stream));
    }
    std::vector<double> hA((size_t)n_cfg * h->lda);
    CK(cudaMemcpyAsync(hA.data(), A.p, hA.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (other) CK(cudaMemcpyAsync(other, oth.p, oth.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (exponent) CK(cudaMemcpyAsync(exponent, ex.p, ex.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (drift_r) CK(cudaMemcpyAsync(drift_r, dr.p, dr.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (drift_i) CK(cudaMemcpyAsync(drift_i, di.p, di.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (spline_sums) CK(cudaMemcpyAsync(spline_sums, ss.p, ss.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (outer) CK(cudaMemcpyAsync(outer, out.p, out.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int c = 0; c < n_cfg; c++)
    {
        const double* row = &hA[(size_t)c * h->lda];
        if (O) memcpy(O + (size_t)c * P, row, sizeof(double) * P);
        if (e_r) e_r[c] = row[P];
        if (e_i) e_i[c] = row[P + 1];
    }
    return 0;
}

int tdvmc_gpu_quotient_fixed(tdvmc_gpu_handle* h, const double* R, const double* moves, int32_t n_moves, double* quotient,
                             double* delta)
{
    if (!h || !R || !moves || n_moves < 1) return h ? fail(h, "quotient_fixed: bad arguments") : -1;
    if (int rc = need_params(h)) return rc;
    CK(cudaSetDevice(h->device));
    for (int m = 0; m < n_moves; m++)
        if (moves[4 * m] < 0 || moves[4 * m] >= h->N) return fail(h, "quotient_fixed: particle index out of range");
    DevBuf<double> aos, pos, mv, dl;
    CK(aos.alloc((size_t)h->N * 3));
    CK(pos.alloc((size_t)3 * h->Np));
    CK(mv.alloc((size_t)n_moves * 4));
    CK(dl.alloc(n_moves));
    CK(cudaMemsetAsync(pos.p, 0, pos.n * sizeof(double), h->stream));
    CK(cudaMemcpyAsync(aos.p, R, aos.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(mv.p, moves, mv.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(launch_transpose_in(aos.p, pos.p, 1, h->N, h->Np, h->stream));
    QuotientArgs a;
    a.s = h->sysdev();
    a.pos = pos.p;
    a.moves = mv.p;
    a.n_moves = n_moves;
    a.delta = dl.p;
    CK(h->kind == TDVMC_SYSTEM_MIXTURE ? launch_quotient_mix(a, h->stream)
       : h->kind == TDVMC_SYSTEM_INH_CONTACT ? launch_quotient_inh(a, h->stream)
       : h->kind == TDVMC_SYSTEM_BOX_RADIAL ? launch_quotient_br(a, h->stream) : launch_quotient(a, h->stream));
    std::vector<double> d(n_moves);
    CK(cudaMemcpyAsync(d.data(), dl.p, sizeof(double) * n_moves, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int m = 0; m < n_moves; m++)
    {
        if (delta) delta[m] = d[m];
        if (quotient) quotient[m] = exp(2.0 * d[m]); // BosonsBulk.cpp:652
    }
    return 0;
}

static int run_tables(tdvmc_gpu_handle* h, const double* pos, int n_cfg)
{
    if (h->kind != TDVMC_SYSTEM_SPLINE_TABLE) return fail(h, "the table kernels cover the spline-table systems only");
    CK(h->d_T.ensure((size_t)n_cfg * h->N * h->K * 4));
    CK(h->d_vint.ensure(n_cfg));
    CK(h->d_tab_e.ensure((size_t)n_cfg * 7));
    TableArgs a;
    a.s = h->sysdev();
    a.pos = pos;
    a.n_cfg = n_cfg;
    a.T = h->d_T.p;
    a.v_int = h->d_vint.p;
    Timed t(h, TDVMC_KERNEL_TABLES);
    CK(launch_tables(a, h->stream));
    return 0;
}

static int run_contract(tdvmc_gpu_handle* h, int n_cfg)
{
    ContractArgs a;
    a.s = h->sysdev();
    a.T = h->d_T.p;
    a.v_int = h->d_vint.p;
    a.n_cfg = n_cfg;
    a.e_r = h->d_tab_e.p;
    a.e_i = h->d_tab_e.p + n_cfg;
    a.sums = h->d_tab_e.p + 2 * (size_t)n_cfg;
    Timed t(h, TDVMC_KERNEL_CONTRACT);
    CK(launch_contract(a, h->stream));
    return 0;
}

int tdvmc_gpu_tables_fixed(tdvmc_gpu_handle* h, const double* R, double* sD, double* sD2)
{
    if (!h || !R) return h ? fail(h, "tables_fixed: bad arguments") : -1;
    if (int rc = need_params(h)) return rc;
    CK(cudaSetDevice(h->device));
    const int N = h->N, K = h->K;
    DevBuf<double> aos, pos;
    CK(aos.alloc((size_t)N * 3));
    CK(pos.alloc((size_t)3 * h->Np));
    CK(cudaMemsetAsync(pos.p, 0, pos.n * sizeof(double), h->stream));
    CK(cudaMemcpyAsync(aos.p, R, aos.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(launch_transpose_in(aos.p, pos.p, 1, N, h->Np, h->stream));
    if (int rc = run_tables(h, pos.p, 1)) return rc;
    std::vector<double> T((size_t)N * K * 4);
    CK(cudaMemcpyAsync(T.data(), h->d_T.p, T.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int n = 0; n < N; n++)
        for (int k = 0; k < K; k++)
        {
            const double* t = &T[((size_t)n * K + k) * 4];
            if (sD)
                for (int a = 0; a < 3; a++) sD[((size_t)k * N + n) * 3 + a] = t[a]; // reference layout [k][n][a]
            if (sD2) sD2[(size_t)k * N + n] = t[3];
        }
    return 0;
}

int tdvmc_gpu_tables_resident(tdvmc_gpu_handle* h, int32_t n_walkers)
{
    if (!h || n_walkers < 1 || n_walkers > h->W) return h ? fail(h, "tables_resident: bad walker count") : -1;
    if (int rc = need_params(h)) return rc;
    CK(cudaSetDevice(h->device));
    return run_tables(h, h->d_pos.p, n_walkers);
}

int tdvmc_gpu_contract_resident(tdvmc_gpu_handle* h, int32_t n_walkers, double* e_r, double* e_i)
{
    if (!h || n_walkers < 1 || n_walkers > h->W) return h ? fail(h, "contract_resident: bad walker count") : -1;
    if (int rc = need_params(h)) return rc;
    if (!h->d_T.p || h->d_T.n < (size_t)n_walkers * h->N * h->K * 4) return fail(h, "contract_resident: tables not built");
    CK(cudaSetDevice(h->device));
    if (int rc = run_contract(h, n_walkers)) return rc;
    if (e_r) CK(cudaMemcpyAsync(e_r, h->d_tab_e.p, sizeof(double) * n_walkers, cudaMemcpyDeviceToHost, h->stream));
    if (e_i) CK(cudaMemcpyAsync(e_i, h->d_tab_e.p + n_walkers, sizeof(double) * n_walkers, cudaMemcpyDeviceToHost, h->stream));
    if (e_r || e_i) CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int tdvmc_gpu_min_image(tdvmc_gpu_handle* h, double lbox, const double* a, const double* b, int32_t n, double* norm,
                        double* disp)
{
    if (!h || !a || !b || !norm || !disp || n < 1) return h ? fail(h, "min_image: bad arguments") : -1;
    CK(cudaSetDevice(h->device));
    DevBuf<double> da, db, dn, dd;
    CK(da.alloc((size_t)n * 3));
    CK(db.alloc((size_t)n * 3));
    CK(dn.alloc(n));
    CK(dd.alloc((size_t)n * 3));
    CK(cudaMemcpyAsync(da.p, a, da.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(db.p, b, db.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(launch_min_image(h->sysdev(), lbox, da.p, db.p, n, dn.p, dd.p, h->stream));
    CK(cudaMemcpyAsync(norm, dn.p, dn.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(disp, dd.p, dd.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int tdvmc_gpu_accumulate_fixed(tdvmc_gpu_handle* h, const double* O, const double* e_r, const double* e_i, int64_t M,
                               double* S, double* f_r, double* f_i, double* o)
{
    if (!h || !O || !e_r || !e_i || M < 1) return h ? fail(h, "accumulate_fixed: bad arguments") : -1;
    CK(cudaSetDevice(h->device));
    const int P = h->P;
    const long long rows_pad = ((M + kAccChunkRows - 1) / kAccChunkRows) * kAccChunkRows;
    DevBuf<double> dO, dER, dEI, A, oth;
    CK(dO.alloc((size_t)M * P));
    CK(dER.alloc(M));
    CK(dEI.alloc(M));
    CK(A.alloc((size_t)rows_pad * h->lda));
    CK(oth.alloc((size_t)M * h->n_other));
    CK(cudaMemsetAsync(A.p, 0, A.n * sizeof(double), h->stream));
    CK(cudaMemsetAsync(oth.p, 0, oth.n * sizeof(double), h->stream));
    CK(cudaMemcpyAsync(dO.p, O, dO.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dER.p, e_r, dER.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dEI.p, e_i, dEI.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(launch_fill_rows(A.p, h->lda, P, dO.p, dER.p, dEI.p, M, h->stream));
    const long long saved_rows = h->rows_used;
    if (int rc = do_accumulate(h, A.p, oth.p, M)) return rc;
    h->rows_used = saved_rows;
    h->est_valid = false;
    std::vector<double> est(h->est_len);
    CK(cudaMemcpyAsync(est.data(), h->d_est.p, est.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (S) memcpy(S, est.data(), sizeof(double) * (size_t)P * P);
    if (f_r) memcpy(f_r, est.data() + (size_t)P * P, sizeof(double) * P);
    if (f_i) memcpy(f_i, est.data() + (size_t)P * P + P, sizeof(double) * P);
    if (o) memcpy(o, est.data() + (size_t)P * P + 2 * P, sizeof(double) * P);
    return 0;
}

// ---- additional observables (observables.cu) ----
namespace
{
struct ObsDevice
{
    DevBuf<int> shell_ptr;
    DevBuf<double> kvec;
    int n_kvec = 0;
};

int check_observable_desc(tdvmc_gpu_handle* h, const tdvmc_observable_desc* od)
{
    if (!od) return fail(h, "observables: null description");
    if (h->kind == TDVMC_SYSTEM_MIXTURE) return fail(h, "observables: g(r)/S(k) are defined for the one-species systems");
    if (od->gr_count < 0 || od->n_shells < 0 || (od->gr_count == 0 && od->n_shells == 0))
        return fail(h, "observables: empty description");
    if (od->gr_count > 0 && (!(od->gr_spacing > 0.0) || !(od->gr_max > 0.0) || !od->gr_scaling))
        return fail(h, "observables: bad g(r) grid");
    if (od->n_shells > 0)
    {
        if (!od->shell_ptr || !od->kvec || od->shell_ptr[0] != 0) return fail(h, "observables: bad wave-vector shells");
        for (int k = 0; k < od->n_shells; k++)
            if (od->shell_ptr[k + 1] <= od->shell_ptr[k]) return fail(h, "observables: empty wave-vector shell");
    }
    return 0;
}

int upload_observable_desc(tdvmc_gpu_handle* h, const tdvmc_observable_desc* od, ObsDevice& d)
{
    d.n_kvec = od->n_shells > 0 ? od->shell_ptr[od->n_shells] : 0;
    CK(d.shell_ptr.alloc((size_t)od->n_shells + 1));
    CK(d.kvec.alloc((size_t)3 * d.n_kvec + 1));
    if (od->n_shells > 0)
    {
        CK(cudaMemcpyAsync(d.shell_ptr.p, od->shell_ptr, ((size_t)od->n_shells + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(d.kvec.p, od->kvec, (size_t)3 * d.n_kvec * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    return 0;
}

ObsArgs make_obs_args(tdvmc_gpu_handle* h, const tdvmc_observable_desc* od, const ObsDevice& d)
{
    ObsArgs a;
    memset(&a, 0, sizeof(a));
    a.s = h->sysdev();
    a.gr_count = od->gr_count;
    a.gr_spacing = od->gr_spacing;
    a.gr_max = od->gr_max;
    a.n_shells = od->n_shells;
    a.n_kvec = d.n_kvec;
    a.shell_ptr = d.shell_ptr.p;
    a.kvec = d.kvec.p;
    return a;
}
} // namespace

int tdvmc_gpu_observables_fixed(tdvmc_gpu_handle* h, const tdvmc_observable_desc* od, const double* R, int32_t n_cfg, double* gr,
                                double* sk)
{
    if (h && h->dim != 3) return fail(h, "observables: the g(r) / S(k) pass is offered for DIM = 3 only");
    if (!h || !R || n_cfg < 1) return h ? fail(h, "observables_fixed: bad arguments") : -1;
    if (int rc = check_observable_desc(h, od)) return rc;
    CK(cudaSetDevice(h->device));
    ObsDevice d;
    if (int rc = upload_observable_desc(h, od, d)) return rc;
    DevBuf<double> aos, pos, skr;
    DevBuf<unsigned long long> grr;
    CK(aos.alloc((size_t)n_cfg * h->N * 3));
    CK(pos.alloc((size_t)n_cfg * 3 * h->Np));
    CK(grr.alloc((size_t)n_cfg * od->gr_count + 1));
    CK(skr.alloc((size_t)n_cfg * od->n_shells + 1));
    CK(cudaMemsetAsync(pos.p, 0, pos.n * sizeof(double), h->stream));
    CK(cudaMemcpyAsync(aos.p, R, aos.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(launch_transpose_in(aos.p, pos.p, n_cfg, h->N, h->Np, h->stream));
    ObsArgs a = make_obs_args(h, od, d);
    a.pos = pos.p;
    a.n_cfg = n_cfg;
    a.accumulate = 0;
    a.gr_rows = grr.p;
    a.sk_rows = skr.p;
    {
        Timed t(h, TDVMC_KERNEL_OTHER);
        CK(launch_observables(a, h->stream));
    }
    std::vector<unsigned long long> cnt((size_t)n_cfg * od->gr_count);
    if (!cnt.empty()) CK(cudaMemcpyAsync(cnt.data(), grr.p, cnt.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    if (sk && od->n_shells > 0)
        CK(cudaMemcpyAsync(sk, skr.p, (size_t)n_cfg * od->n_shells * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (gr)
        for (int c = 0; c < n_cfg; c++)
            for (int b = 0; b < od->gr_count; b++) // ObservableVsOnGridWithScaling.cpp:51, once per counted pair
                gr[(size_t)c * od->gr_count + b] = (double)cnt[(size_t)c * od->gr_count + b] * (od->gr_weight / od->gr_scaling[b]);
    return 0;
}

int tdvmc_gpu_sample_observables(tdvmc_gpu_handle* h, const tdvmc_observable_desc* od, int32_t n_samples, int32_t n_therm,
                                 int32_t n_init, double* gr, double* sk)
{
    if (h && h->dim != 3) return fail(h, "observables: the g(r) / S(k) pass is offered for DIM = 3 only");
    if (!h) return -1;
    if (int rc = need_params(h)) return rc;
    if (int rc = check_observable_desc(h, od)) return rc;
    if (n_samples < 1 || n_therm < 0 || n_init < 0) return fail(h, "sample_observables: bad step counts");
    CK(cudaSetDevice(h->device));
    ObsDevice d;
    if (int rc = upload_observable_desc(h, od, d)) return rc;
    DevBuf<double> skr, sum;
    DevBuf<unsigned long long> grr;
    const int ncol = od->gr_count + od->n_shells;
    CK(grr.alloc((size_t)h->W * od->gr_count + 1));
    CK(skr.alloc((size_t)h->W * od->n_shells + 1));
    CK(sum.alloc((size_t)ncol + 1));
    CK(cudaMemsetAsync(grr.p, 0, grr.n * sizeof(unsigned long long), h->stream));
    CK(cudaMemsetAsync(skr.p, 0, skr.n * sizeof(double), h->stream));
    ObsArgs a = make_obs_args(h, od, d);
    a.pos = h->d_pos.p;
    a.n_cfg = h->W;
    a.accumulate = 1;
    a.gr_rows = grr.p;
    a.sk_rows = skr.p;
    if (int rc = do_sweep(h, n_init)) return rc; // src/TDVMC.cpp:1338-1341
    for (int m = 0; m < n_samples; m++)
    {
        if (int rc = do_sweep(h, n_therm)) return rc; // :1344-1347
        Timed t(h, TDVMC_KERNEL_OTHER);
        CK(launch_observables(a, h->stream)); // :1349
    }
    CK(launch_obs_reduce(grr.p, skr.p, h->W, od->gr_count, od->n_shells, sum.p, h->stream));
    const double n_local = (double)n_samples * (double)h->W;
    CK(cudaMemcpyAsync(sum.p + ncol, &n_local, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (h->comm) // MPIMethods::ReduceToAverage(additionalObservablesMean), src/TDVMC.cpp:1443
    {
        int rc = g_nccl.AllReduce(sum.p, sum.p, (size_t)ncol + 1, kNcclDouble, kNcclSum, h->comm, h->stream);
        if (rc != 0) return fail(h, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"), rc);
    }
    std::vector<double> hs((size_t)ncol + 1);
    CK(cudaMemcpyAsync(hs.data(), sum.p, hs.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const double inv = 1.0 / hs[ncol]; // equal weights: the running mean of :1358-1361 and the mean over ranks
    if (gr)
        for (int b = 0; b < od->gr_count; b++) gr[b] = hs[b] * (od->gr_weight / od->gr_scaling[b]) * inv;
    if (sk)
        for (int k = 0; k < od->n_shells; k++) sk[k] = hs[od->gr_count + k] * inv;
    return 0;
}

// ---- cluster observables ----
namespace
{
int check_cluster_desc(tdvmc_gpu_handle* h, const tdvmc_cluster_observable_desc* od)
{
    if (!od) return fail(h, "cluster observables: null description");
    if (h->kind != TDVMC_SYSTEM_MIXTURE || h->N != 3)
        return fail(h, "cluster observables: defined for the three-particle mixture cluster (BosonMixtureCluster.cpp:331-340)");
    if (od->n_angle < 1 || od->n_density < 1 || od->n_distance < 1 || !(od->angle_spacing > 0.0) || !(od->density_spacing > 0.0) ||
        !(od->distance_spacing > 0.0) || !od->density_scaling)
        return fail(h, "cluster observables: bad grids");
    return 0;
}

ClusterObsArgs make_cluster_args(tdvmc_gpu_handle* h, const tdvmc_cluster_observable_desc* od)
{
    ClusterObsArgs a;
    memset(&a, 0, sizeof(a));
    a.s = h->sysdev();
    a.n_angle = od->n_angle;
    a.n_density = od->n_density;
    a.n_distance = od->n_distance;
    a.angle_spacing = od->angle_spacing;
    a.density_spacing = od->density_spacing;
    a.density_max = od->density_max;
    a.distance_spacing = od->distance_spacing;
    a.distance_max = od->distance_max;
    return a;
}

// counts -> the reference's histogram values, scaled by 1/n_cfg_total
void unpack_cluster_hist(const tdvmc_cluster_observable_desc* od, const double* cnt, double inv, double* angle, double* density,
                         double* distance)
{
    const int na = od->n_angle, nd = od->n_density, np = od->n_distance;
    if (angle)
        for (int i = 0; i < 3 * na; i++) angle[i] = cnt[i] * inv;
    if (density)
        for (int i = 0; i < 3; i++)
            for (int b = 0; b < nd; b++) density[i * nd + b] = cnt[3 * na + i * nd + b] * (1.0 / od->density_scaling[b]) * inv;
    if (distance)
        for (int i = 0; i < 3 * np; i++) distance[i] = cnt[3 * na + 3 * nd + i] * inv;
}
} // namespace

int tdvmc_gpu_cluster_observables_fixed(tdvmc_gpu_handle* h, const tdvmc_cluster_observable_desc* od, const double* R, int32_t n_cfg,
                                        double* r2, double* angle, double* density, double* distance)
{
    if (!h || !R || n_cfg < 1) return h ? fail(h, "cluster_observables_fixed: bad arguments") : -1;
    if (int rc = check_cluster_desc(h, od)) return rc;
    CK(cudaSetDevice(h->device));
    const int nh = 3 * (od->n_angle + od->n_density + od->n_distance);
    DevBuf<double> aos, pos, r2r;
    DevBuf<unsigned long long> hist;
    CK(aos.alloc((size_t)n_cfg * 9));
    CK(pos.alloc((size_t)n_cfg * 3 * h->Np));
    CK(r2r.alloc(n_cfg));
    CK(hist.alloc((size_t)n_cfg * nh));
    CK(cudaMemsetAsync(pos.p, 0, pos.n * sizeof(double), h->stream));
    CK(cudaMemcpyAsync(aos.p, R, aos.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(launch_transpose_in(aos.p, pos.p, n_cfg, h->N, h->Np, h->stream));
    ClusterObsArgs a = make_cluster_args(h, od);
    a.pos = pos.p;
    a.n_cfg = n_cfg;
    a.accumulate = 0;
    a.per_cfg_hist = 1;
    a.r2_rows = r2r.p;
    a.hist = hist.p;
    {
        Timed t(h, TDVMC_KERNEL_OTHER);
        CK(launch_cluster_observables(a, h->stream));
    }
    std::vector<unsigned long long> cnt((size_t)n_cfg * nh);
    CK(cudaMemcpyAsync(cnt.data(), hist.p, cnt.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    if (r2) CK(cudaMemcpyAsync(r2, r2r.p, (size_t)n_cfg * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    std::vector<double> c(nh);
    for (int k = 0; k < n_cfg; k++)
    {
        for (int i = 0; i < nh; i++) c[i] = (double)cnt[(size_t)k * nh + i];
        unpack_cluster_hist(od, c.data(), 1.0, angle ? angle + (size_t)k * 3 * od->n_angle : nullptr,
                            density ? density + (size_t)k * 3 * od->n_density : nullptr,
                            distance ? distance + (size_t)k * 3 * od->n_distance : nullptr);
    }
    return 0;
}

int tdvmc_gpu_sample_cluster_observables(tdvmc_gpu_handle* h, const tdvmc_cluster_observable_desc* od, int32_t n_samples,
                                         int32_t n_therm, int32_t n_init, double* r2, double* angle, double* density,
                                         double* distance)
{
    if (!h) return -1;
    if (int rc = need_params(h)) return rc;
    if (int rc = check_cluster_desc(h, od)) return rc;
    if (n_samples < 1 || n_therm < 0 || n_init < 0) return fail(h, "sample_cluster_observables: bad step counts");
    CK(cudaSetDevice(h->device));
    const int nh = 3 * (od->n_angle + od->n_density + od->n_distance);
    DevBuf<double> r2r, sum;
    DevBuf<unsigned long long> hist;
    CK(r2r.alloc(h->W));
    CK(hist.alloc(nh));
    CK(sum.alloc((size_t)nh + 2));
    CK(cudaMemsetAsync(r2r.p, 0, r2r.n * sizeof(double), h->stream));
    CK(cudaMemsetAsync(hist.p, 0, hist.n * sizeof(unsigned long long), h->stream));
    ClusterObsArgs a = make_cluster_args(h, od);
    a.pos = h->d_pos.p;
    a.n_cfg = h->W;
    a.accumulate = 1;
    a.per_cfg_hist = 0;
    a.r2_rows = r2r.p;
    a.hist = hist.p;
    if (int rc = do_sweep(h, n_init)) return rc; // src/TDVMC.cpp:1338-1341
    for (int m = 0; m < n_samples; m++)
    {
        if (int rc = do_sweep(h, n_therm)) return rc; // :1344-1347
        Timed t(h, TDVMC_KERNEL_OTHER);
        CK(launch_cluster_observables(a, h->stream)); // :1349
    }
    // pack [counts as doubles | sum r2 | number of configurations] for one all-reduce
    std::vector<unsigned long long> cnt(nh);
    CK(launch_sum_rows(r2r.p, h->W, sum.p + nh, h->stream));
    CK(cudaMemcpyAsync(cnt.data(), hist.p, (size_t)nh * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    std::vector<double> hs((size_t)nh + 2);
    for (int i = 0; i < nh; i++) hs[i] = (double)cnt[i];
    hs[nh + 1] = (double)n_samples * (double)h->W;
    CK(cudaMemcpyAsync(&hs[nh], sum.p + nh, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->comm) // MPIMethods::ReduceToAverage(additionalObservablesMean), src/TDVMC.cpp:1443
    {
        CK(cudaMemcpyAsync(sum.p, hs.data(), hs.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        int rc = g_nccl.AllReduce(sum.p, sum.p, hs.size(), kNcclDouble, kNcclSum, h->comm, h->stream);
        if (rc != 0) return fail(h, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"), rc);
        CK(cudaMemcpyAsync(hs.data(), sum.p, hs.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    const double inv = 1.0 / hs[nh + 1];
    if (r2) *r2 = hs[nh] * inv;
    unpack_cluster_hist(od, hs.data(), inv, angle, density, distance);
    return 0;
}

int tdvmc_gpu_proposals(tdvmc_gpu_handle* h, int32_t global_walker, int64_t first_step, int32_t n, int32_t* particle,
                        double* disp, double* log_u)
{
    if (!h || n < 1 || !particle || !disp || !log_u) return h ? fail(h, "proposals: bad arguments") : -1;
    CK(cudaSetDevice(h->device));
    DevBuf<int> dp;
    DevBuf<double> dd, dl;
    CK(dp.alloc(n));
    CK(dd.alloc((size_t)n * 3));
    CK(dl.alloc(n));
    CK(launch_proposals(h->seed, (uint32_t)global_walker, (uint64_t)first_step, n, h->N, h->mc_step, dp.p, dd.p, dl.p, h->stream));
    CK(cudaMemcpyAsync(particle, dp.p, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(disp, dd.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(log_u, dl.p, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- measurement hooks ----

int tdvmc_gpu_profile(tdvmc_gpu_handle* h, int32_t enable, int32_t reset)
{
    if (!h) return -1;
    CK(cudaSetDevice(h->device));
    collect_timings(h);
    if (reset)
        for (int k = 0; k < TDVMC_KERNEL_COUNT; k++)
        {
            h->launches[k] = 0;
            h->total_ms[k] = 0.0;
        }
    h->profiling = enable != 0;
    return 0;
}

int tdvmc_gpu_kernel_stats(tdvmc_gpu_handle* h, int32_t kernel_id, int64_t* launches, double* total_ms)
{
    if (!h || kernel_id < 0 || kernel_id >= TDVMC_KERNEL_COUNT) return h ? fail(h, "kernel_stats: bad id") : -1;
    CK(cudaSetDevice(h->device));
    collect_timings(h);
    if (launches) *launches = h->launches[kernel_id];
    if (total_ms) *total_ms = h->total_ms[kernel_id];
    return 0;
}

int tdvmc_gpu_synchronize(tdvmc_gpu_handle* h)
{
    if (!h) return -1;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int tdvmc_gpu_timer_start(tdvmc_gpu_handle* h)
{
    if (!h) return -1;
    CK(cudaSetDevice(h->device));
    if (!h->timer0)
    {
        CK(cudaEventCreate(&h->timer0));
        CK(cudaEventCreate(&h->timer1));
    }
    CK(cudaEventRecord(h->timer0, h->stream));
    return 0;
}

int tdvmc_gpu_timer_stop(tdvmc_gpu_handle* h, double* elapsed_ms)
{
    if (!h || !elapsed_ms || !h->timer0) return h ? fail(h, "timer_stop: timer not started") : -1;
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->timer1, h->stream));
    CK(cudaEventSynchronize(h->timer1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->timer0, h->timer1));
    *elapsed_ms = ms;
    return 0;
}

int tdvmc_gpu_launch_count(tdvmc_gpu_handle* h, int64_t* n)
{
    if (!h || !n) return -1;
    *n = h->n_launched;
    return 0;
}

int tdvmc_gpu_flush_l2(tdvmc_gpu_handle* h, int64_t n_bytes)
{
    if (!h || n_bytes < 1) return h ? fail(h, "flush_l2: bad size") : -1;
    CK(cudaSetDevice(h->device));
    CK(h->d_flush.ensure((size_t)n_bytes));
    CK(cudaMemsetAsync(h->d_flush.p, 0, (size_t)n_bytes, h->stream));
    return 0;
}

int tdvmc_gpu_resident_walkers(tdvmc_gpu_handle* h, int32_t* per_sm, int32_t* sm_count)
{
    if (!h) return -1;
    if (per_sm) *per_sm = h->resident_per_sm;
    if (sm_count) *sm_count = h->sm_count;
    return 0;
}

int tdvmc_gpu_measure_fp64_peak(tdvmc_gpu_handle* h, double* dfma_tflops, double* dmma_tflops)
{
    if (!h || !dfma_tflops || !dmma_tflops) return -1;
    CK(cudaSetDevice(h->device));
    CK(measure_fp64(dfma_tflops, dmma_tflops, h->stream));
    return 0;
}

} // extern "C"
