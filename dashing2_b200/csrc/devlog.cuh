// devlog.cuh -- ref_log(): double-precision natural log that is bit-identical to the host libm the
// reference links against (glibc >= 2.28 `log`, FMA ifunc variant selected on every AVX2+FMA x86-64).
//
// The reference's registers for Full SetSketch (/root/reference/src/setsketch.h:390,419), ProbMinHash /
// BagMinHash (bonsai/hll/include/sketch/bmh.h:175) and its Mash distance (src/cmp_core.cpp:361) are
// produced by that libm `log`.  CUDA's log() is within 1 ulp but not bit-identical, so the published
// algorithm (ARM optimized-routines / glibc sysdeps/ieee754/dbl-64/e_log.c: 128-entry table,
// log(x) = k ln2 + log(c) + log1p(z/c - 1)) is evaluated here with exactly the fused/unfused operation
// order of the compiled glibc 2.39 routine.  Constants: devlog_table.h (generated from the system libm
// by scripts/extract_glibc_log_table.py).  Every operation uses an explicit rounding intrinsic so
// neither nvcc nor the host compiler can re-associate or contract it.
#pragma once
#include <stdint.h>
#include <string.h>
#include "devlog_table.h"
#if !defined(__CUDA_ARCH__)
#include <cmath>
#endif

namespace d2g {

struct LogTabEntry { double invc, logc; };

#if defined(__CUDACC__)
__device__ __constant__ const double kLogA_dev[5] = D2G_LOG_POLY_A;
__device__ __constant__ const double kLogB_dev[11] = D2G_LOG_POLY_B;
__device__ const LogTabEntry kLogTab_dev[128] = D2G_LOG_TAB;
#endif
static const double kLogA_host[5] = D2G_LOG_POLY_A;
static const double kLogB_host[11] = D2G_LOG_POLY_B;
static const LogTabEntry kLogTab_host[128] = D2G_LOG_TAB;

#if defined(__CUDA_ARCH__)
#define D2G_FMA(a, b, c) __fma_rn((a), (b), (c))
#define D2G_ADD(a, b) __dadd_rn((a), (b))
#define D2G_MUL(a, b) __dmul_rn((a), (b))
#define D2G_LOGA kLogA_dev
#define D2G_LOGB kLogB_dev
#define D2G_LOGT kLogTab_dev
#else
#define D2G_FMA(a, b, c) std::fma((a), (b), (c))
#define D2G_ADD(a, b) ((a) + (b))
#define D2G_MUL(a, b) ((a) * (b))
#define D2G_LOGA kLogA_host
#define D2G_LOGB kLogB_host
#define D2G_LOGT kLogTab_host
#endif

#if defined(__CUDACC__)
__host__ __device__
#endif
inline double ref_log(double x) {
    uint64_t ix; memcpy(&ix, &x, 8);
    const double *A = D2G_LOGA, *B = D2G_LOGB;
    // |x - 1| small: 1 - 2^-4 <= x < 1 + 0x1.09p-4
    if (ix - 0x3fee000000000000ULL < 0x3090000000000ULL) {
        if (ix == 0x3ff0000000000000ULL) return 0.;
        const double r = D2G_ADD(x, -1.0);
        const double r2 = D2G_MUL(r, r), r3 = D2G_MUL(r, r2);
        const double q1 = D2G_FMA(r2, B[3], D2G_FMA(r, B[2], B[1]));
        const double q2 = D2G_FMA(r2, B[6], D2G_FMA(r, B[5], B[4]));
        const double q3 = D2G_FMA(r3, B[10], D2G_FMA(r2, B[9], D2G_FMA(r, B[8], B[7])));
        const double br = D2G_FMA(D2G_FMA(q3, r3, q2), r3, q1);
        const double rhi = D2G_FMA(-0x1p27, r, D2G_FMA(r, 0x1p27, r));
        const double rlo = D2G_ADD(r, -rhi);
        const double rr = D2G_MUL(rhi, rhi);
        const double hi = D2G_FMA(rr, B[0], r);
        double lo = D2G_FMA(rr, B[0], D2G_ADD(r, -hi));
        lo = D2G_FMA(D2G_MUL(B[0], rlo), D2G_ADD(r, rhi), lo);
        const double y = D2G_FMA(br, r3, lo);
        return D2G_ADD(hi, y);
    }
    uint32_t top = (uint32_t)(ix >> 48);
    if (top - 0x0010u >= 0x7ff0u - 0x0010u) {
        if ((ix << 1) == 0) return -__builtin_huge_val();                                  // log(+-0) = -inf
        if (ix == 0x7ff0000000000000ULL) return x;                                       // log(inf) = inf
        if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return __builtin_nan("");     // x < 0 or NaN
        const double xs = D2G_MUL(x, 0x1p52);                                            // subnormal
        memcpy(&ix, &xs, 8);
        ix -= 52ULL << 52;
    }
    const uint64_t tmp = ix - 0x3fe6000000000000ULL;
    const int i = (int)((tmp >> 45) & 127);
    const int64_t k = (int64_t)tmp >> 52;
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ULL);
    double z; memcpy(&z, &iz, 8);
    const double invc = D2G_LOGT[i].invc, logc = D2G_LOGT[i].logc;
    const double kd = (double)(int32_t)k;
    const double w = D2G_FMA(kd, D2G_LOG_LN2HI, logc);
    const double r = D2G_FMA(z, invc, -1.0);
    const double hi = D2G_ADD(r, w);
    const double r2 = D2G_MUL(r, r);
    const double lo = D2G_FMA(kd, D2G_LOG_LN2LO, D2G_ADD(D2G_ADD(w, -hi), r));
    const double r3 = D2G_MUL(r, r2);
    const double p = D2G_FMA(D2G_FMA(r, A[4], A[3]), r2, D2G_FMA(r, A[2], A[1]));
    const double y = D2G_FMA(r3, p, D2G_FMA(r2, A[0], lo));
    return D2G_ADD(y, hi);
}

} // namespace d2g
