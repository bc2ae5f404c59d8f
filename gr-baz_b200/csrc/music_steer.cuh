// music_steer.cuh - device-side steering-table builder.
//
// Restates /root/reference/python/music_doa_helper.py:29-46 (unit_vect + calculate_antenna_array_response)
// followed by the complex128 -> complex64 rounding the SWIG typemap applies (swig/baz_swig.i:564), one
// thread per (angle step, antenna):
//     theta = (k * 360.0 / K) * (pi / 180)                                   :34  (IEEE mul, div, mul)
//     u     = [cos theta, sin theta]                                         :29-30
//     d     = inner(p, u) / lambda                                           :40  (numpy.inner on 2 elements =
//             BLAS ddot = fma(p1, u1, p0 * u0))
//     a     = exp(-1j * 2 * pi * d) = (cos phi, sin phi), phi = 0 + -(2 pi * d) :41
//     table[k][ant] = ((float)Re a, (float)Im a)
//
// Every step but the four sin/cos is a correctly rounded IEEE operation and is reproduced exactly.  CUDA's
// fp64 sincos is within 2 ulp of the true value, the host libm the reference runs on within ~0.55 ulp, so
// the two fp64 values can differ in the last bits; that only matters when the value sits within that
// distance of a float32 rounding boundary.  The kernel therefore reports every component closer than
// 2^-40 * max(1, |phi|) to a boundary (about 3e-5 of the entries, plus the exact zeros of cos/sin) and the
// host re-evaluates just those with libm in the same literal order (steer_entry_host), so the table is the
// one the Python helper builds on this host, bit for bit.
#pragma once
#include <cmath>
#include <cstdint>

namespace music {

constexpr double STEER_DEG2RAD = 3.14159265358979323846 / 180.0;  // numpy.pi / 180.0
constexpr double STEER_TWO_PI = 2.0 * 3.14159265358979323846;
constexpr double STEER_GUARD = 9.094947017729282e-13;              // 2^-40
constexpr int STEER_GUARD_CAP = 65536;

// true when v is so close to the midpoint between two adjacent floats that a few-ulp change of v could
// change (float)v
__device__ __forceinline__ bool steer_ambiguous(double v, double guard)
{
    const float f = __double2float_rn(v);
    const float up = nextafterf(f, INFINITY), dn = nextafterf(f, -INFINITY);
    const double m_up = 0.5 * ((double)f + (double)up), m_dn = 0.5 * ((double)f + (double)dn);  // exact
    return fabs(v - m_up) < guard || fabs(v - m_dn) < guard;
}

__global__ void steer_table_kernel(const double *__restrict__ pos /*[M][2]*/, double lambda, int K, int M,
                                   float2 *__restrict__ table, unsigned *__restrict__ guard_count,
                                   int *__restrict__ guard_list)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * M) return;
    const int k = idx / M, ant = idx - k * M;
    const double theta = __dmul_rn(__ddiv_rn(__dmul_rn((double)k, 360.0), (double)K), STEER_DEG2RAD);
    double s, c;
    sincos(theta, &s, &c);
    const double inner = __fma_rn(pos[2 * ant + 1], s, __dmul_rn(pos[2 * ant], c));
    const double phi = __dadd_rn(0.0, -__dmul_rn(STEER_TWO_PI, __ddiv_rn(inner, lambda)));  // 0 + (-0) = +0, as in the complex product
    double sp, cp;
    sincos(phi, &sp, &cp);
    table[idx] = make_float2(__double2float_rn(cp), __double2float_rn(sp));
    const double guard = STEER_GUARD * fmax(1.0, fabs(phi));
    // phi == 0 (an element at the origin) gives exactly (1, +0) on every libm: nothing to guard
    if (phi != 0.0 && (!(fabs(phi) < 1e300) || steer_ambiguous(cp, guard) || steer_ambiguous(sp, guard))) {
        const unsigned slot = atomicAdd(guard_count, 1u);
        if (slot < (unsigned)STEER_GUARD_CAP) guard_list[slot] = idx;
    }
}

// scatter of the host-evaluated entries into the table
__global__ void steer_patch_kernel(float2 *__restrict__ table, const int *__restrict__ idx, const float2 *__restrict__ val, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) table[idx[i]] = val[i];
}

// host libm evaluation of one entry, same order of operations as the reference helper
inline void steer_entry_host(const double *pos, double lambda, int K, int k, int ant, float *re, float *im)
{
    volatile double t0 = (double)k * 360.0;
    volatile double t1 = t0 / (double)K;
    volatile double theta = t1 * STEER_DEG2RAD;
    volatile double c = std::cos(theta), s = std::sin(theta);
    volatile double p0c = pos[2 * ant] * c;
    volatile double inner = std::fma(pos[2 * ant + 1], s, p0c);
    volatile double d = inner / lambda;
    volatile double phi = 0.0 + -(STEER_TWO_PI * d);  // the complex product (0 - 2 pi j) * d adds a +0: -0 becomes +0
    *re = (float)std::cos(phi);
    *im = (float)std::sin(phi);
}

}  // namespace music
