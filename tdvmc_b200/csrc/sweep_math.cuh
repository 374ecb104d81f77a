// Arithmetic shared by the Metropolis sweep kernels (sweep.cu, boxradial.cu): fast wrap into the first cell,
// select-free minimum image of wrapped points, light square root.  See sweep.cu for the measurements behind them.
#pragma once

#include "common.cuh"

namespace tdvmc
{

constexpr double kMagic = 6755399441055744.0; // 2^52 + 2^51: (x + kMagic) - kMagic rounds x to nearest

// x - L * rint(x / L): into [-L/2, L/2]
__device__ __forceinline__ double wrap_fast(double d, double L, double Linv)
{
    double k = fma(d, Linv, kMagic) - kMagic;
    return fma(-k, L, d);
}

// squared minimum-image distance of two points that both lie in the first cell (|d| <= L per coordinate):
// min(|d|, L - |d|) = L/2 - ||d| - L/2|, two FP64 adds with free |.| modifiers and no select
__device__ __forceinline__ double mi2_wrapped(double dx, double dy, double dz, double Lhalf)
{
    const double mx = Lhalf - fabs(fabs(dx) - Lhalf);
    const double my = Lhalf - fabs(fabs(dy) - Lhalf);
    const double mz = Lhalf - fabs(fabs(dz) - Lhalf);
    return fma(mz, mz, fma(my, my, mx * mx));
}

// squared pair distance: minimum image of wrapped points, or the plain difference for open boundaries
template <bool OPEN>
__device__ __forceinline__ double dist2(double dx, double dy, double dz, double Lhalf)
{
    if (OPEN) return fma(dz, dz, fma(dy, dy, dx * dx));
    return mi2_wrapped(dx, dy, dz, Lhalf);
}

// sqrt(x) for the sampler: MUFU.RSQ64H seed y and ONE Newton step  t + (x - t^2) y/2,  t = x y.  Three FP64 instructions;
// y/2 is an exponent decrement on the integer pipe (the seed has an empty low word and is never subnormal for a pair
// distance).  Relative error -1.5 delta^2 with delta the seed's error: the seed carries the 20 mantissa bits of its high
// word, delta <= 2^-20, so <= 1.4e-12 (tests/test_device_arithmetic_mirrors.py); on the device the exponent change of a proposal
// comes out within 3e-13 of exact arithmetic (test_sampler_exponent_change_against_exact_arithmetic).  The error is a
// smooth, deterministic function of x - the chain samples |psi|^2 of distances stretched by ~1e-12, far below the 1e-7 the
// move ratio is held to against the reference - and E_L, O_k and the drift are evaluated with the exact distance.
// x == 0 gives NaN, which the callers discard (it only happens for the moved particle against itself).
// (TDVMC_SQRT_3RD_ORDER: the third-order step of round 1, ~2 ulp, five FP64 instructions.)
__device__ __forceinline__ double sqrt_fast(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;               // ~ sqrt(x)
#ifdef TDVMC_SQRT_3RD_ORDER
    const double e = fma(-t, y, 1.0);     // 1 - x y^2
    const double p = fma(e, 0.375, 0.5);
    return fma(t * e, p, t);              // t (1 + e/2 + 3 e^2/8)
#else
    const double hy = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
    return fma(fma(-t, t, x), hy, t);
#endif
}

} // namespace tdvmc
