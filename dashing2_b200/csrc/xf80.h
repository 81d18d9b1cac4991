// xf80.h -- software x87 extended precision (64-bit significand, round-to-nearest-even) for device code.
//
// Why: the reference finalises every pairwise comparison in `long double`
// (/root/reference/src/cmp_core.cpp:359-361,461-494,505-516) and then narrows to float32.  To emit a
// bit-identical float32 matrix from the GPU for *any* sketch size and measure, the handful of
// add/sub/mul/div operations per pair are replayed here with integer arithmetic.  The header is
// host+device so that tests can check every operation against the CPU's native long double.
//
// Representation: value = (-1)^sign * sig * 2^(exp-63), sig normalised (bit 63 set) for finite
// non-zero values.  Exponent range is int32 (x87 has 15 bits; the pair finalisation never gets near
// either limit because its inputs are doubles and at most ~6 operations deep), so x87 overflow,
// underflow and denormals are not modelled.  inf / nan are carried as classes.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define XF_HD __host__ __device__ __forceinline__
#else
#define XF_HD inline
#endif

namespace xf {

enum : uint32_t { FINITE = 0, INF = 1, NAN_ = 2 };

struct f80 {
    uint64_t sig;
    int32_t exp;
    uint16_t sign;
    uint16_t cls;
};

XF_HD f80 make(uint64_t sig, int32_t exp, uint32_t sign, uint32_t cls = FINITE) {
    f80 r; r.sig = sig; r.exp = exp; r.sign = (uint16_t)sign; r.cls = (uint16_t)cls; return r;
}
XF_HD f80 zero(uint32_t sign = 0) { return make(0, 0, sign); }
XF_HD f80 inf(uint32_t sign = 0) { return make(0, 0, sign, INF); }
XF_HD f80 nan() { return make(0, 0, 0, NAN_); }
XF_HD bool is_zero(const f80 &a) { return a.cls == FINITE && a.sig == 0; }
XF_HD bool is_nan(const f80 &a) { return a.cls == NAN_; }
XF_HD bool is_inf(const f80 &a) { return a.cls == INF; }

XF_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
XF_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

XF_HD f80 from_u64(uint64_t v) {
    if (!v) return zero();
    const int lz = clz64(v);
    return make(v << lz, 63 - lz, 0);
}

XF_HD f80 from_double(double d) {
    uint64_t b; memcpy(&b, &d, 8);
    const uint32_t sign = (uint32_t)(b >> 63);
    const int32_t e = (int32_t)((b >> 52) & 0x7ff);
    uint64_t frac = b & 0xfffffffffffffULL;
    if (e == 0x7ff) return frac ? nan() : inf(sign);
    if (e == 0) {
        if (!frac) return zero(sign);
        const int lz = clz64(frac);
        return make(frac << lz, -1022 - 52 + (63 - lz), sign);
    }
    return make((frac | (1ULL << 52)) << 11, e - 1023, sign);
}

XF_HD f80 from_float(float f) {
    uint32_t b; memcpy(&b, &f, 4);
    const uint32_t sign = b >> 31;
    const int32_t e = (int32_t)((b >> 23) & 0xff);
    uint64_t frac = b & 0x7fffffu;
    if (e == 0xff) return frac ? nan() : inf(sign);
    if (e == 0) {
        if (!frac) return zero(sign);
        const int lz = clz64(frac);
        return make(frac << lz, -126 - 23 + (63 - lz), sign);
    }
    return make((frac | (1ULL << 23)) << 40, e - 127, sign);
}

// Round a 128-bit significand hi:lo (hi normalised: bit 63 set) + extra sticky to 64 bits, RNE.
XF_HD f80 round_pack(uint32_t sign, int32_t exp, uint64_t hi, uint64_t lo, bool sticky) {
    const bool half = (lo >> 63) != 0;
    const bool rest = ((lo << 1) != 0) || sticky;
    if (half && (rest || (hi & 1))) {
        ++hi;
        if (hi == 0) { hi = 1ULL << 63; ++exp; }
    }
    return make(hi, exp, sign);
}

XF_HD f80 neg(f80 a) { a.sign ^= 1; return a; }

// magnitude add of finite non-zero a, b with result sign `sign`
XF_HD f80 add_mag(const f80 &a, const f80 &b, uint32_t sign) {
    const f80 &x = (a.exp >= b.exp) ? a : b;
    const f80 &y = (a.exp >= b.exp) ? b : a;
    const uint32_t d = (uint32_t)(x.exp - y.exp);
    uint64_t yh, yl; bool sticky = false;
    if (d == 0) { yh = y.sig; yl = 0; }
    else if (d < 64) { yh = y.sig >> d; yl = y.sig << (64 - d); }
    else if (d == 64) { yh = 0; yl = y.sig; }
    else if (d < 128) { yh = 0; yl = y.sig >> (d - 64); sticky = (y.sig << (128 - d)) != 0; }
    else { yh = 0; yl = 0; sticky = true; }
    uint64_t hi = x.sig + yh;
    const bool carry = hi < x.sig;
    uint64_t lo = yl;
    int32_t exp = x.exp;
    if (carry) {
        sticky = sticky || (lo & 1);
        lo = (lo >> 1) | (hi << 63);
        hi = (hi >> 1) | (1ULL << 63);
        ++exp;
    }
    return round_pack(sign, exp, hi, lo, sticky);
}

// magnitude subtract |a| - |b| of finite non-zero values, |a| > |b| guaranteed by caller
XF_HD f80 sub_mag(const f80 &a, const f80 &b, uint32_t sign) {
    const uint32_t d = (uint32_t)(a.exp - b.exp);
    uint64_t bh, bl; bool sticky = false;
    if (d == 0) { bh = b.sig; bl = 0; }
    else if (d < 64) { bh = b.sig >> d; bl = b.sig << (64 - d); }
    else if (d == 64) { bh = 0; bl = b.sig; }
    else if (d < 128) { bh = 0; bl = b.sig >> (d - 64); sticky = (b.sig << (128 - d)) != 0; }
    else { bh = 0; bl = 0; sticky = true; }
    // (a.sig:0) - (bh:bl) - (sticky ? tiny : 0): a sticky remainder makes the true difference slightly
    // smaller; model it by borrowing one ulp of the 128-bit value and keeping sticky set.
    uint64_t lo = 0 - bl;
    uint64_t borrow = (bl != 0);
    uint64_t hi = a.sig - bh - borrow;
    if (sticky) {
        if (lo == 0) --hi;
        --lo;
    }
    if (hi == 0 && lo == 0) return zero(0); // exact cancellation: +0 under round-to-nearest
    int32_t exp = a.exp;
    if (hi == 0) { hi = lo; lo = 0; exp -= 64; }
    const int lz = clz64(hi);
    if (lz) {
        hi = (hi << lz) | (lo >> (64 - lz));
        lo <<= lz;
        exp -= lz;
    }
    return round_pack(sign, exp, hi, lo, sticky);
}

XF_HD int cmp_mag(const f80 &a, const f80 &b) { // finite, possibly zero
    if (a.sig == 0 || b.sig == 0) return (a.sig != 0) - (b.sig != 0);
    if (a.exp != b.exp) return a.exp > b.exp ? 1 : -1;
    return a.sig == b.sig ? 0 : (a.sig > b.sig ? 1 : -1);
}

XF_HD f80 add(const f80 &a, const f80 &b) {
    if (a.cls | b.cls) {
        if (is_nan(a) || is_nan(b)) return nan();
        if (is_inf(a) && is_inf(b)) return a.sign == b.sign ? a : nan();
        return is_inf(a) ? a : b;
    }
    if (a.sig == 0) return b.sig == 0 ? zero(a.sign & b.sign) : b;
    if (b.sig == 0) return a;
    if (a.sign == b.sign) return add_mag(a, b, a.sign);
    const int c = cmp_mag(a, b);
    if (c == 0) return zero(0);
    return c > 0 ? sub_mag(a, b, a.sign) : sub_mag(b, a, b.sign);
}
XF_HD f80 sub(const f80 &a, const f80 &b) { return add(a, neg(b)); }

XF_HD f80 mul(const f80 &a, const f80 &b) {
    const uint32_t sign = a.sign ^ b.sign;
    if (a.cls | b.cls) {
        if (is_nan(a) || is_nan(b)) return nan();
        if (is_zero(a) || is_zero(b)) return nan();
        return inf(sign);
    }
    if (a.sig == 0 || b.sig == 0) return zero(sign);
    uint64_t hi = mulhi64(a.sig, b.sig), lo = a.sig * b.sig;
    int32_t exp = a.exp + b.exp + 1; // product of two [1,2) significands is in [1,4)
    if (!(hi >> 63)) { hi = (hi << 1) | (lo >> 63); lo <<= 1; --exp; }
    return round_pack(sign, exp, hi, lo, false);
}

// 128/64 -> 64 division (Hacker's Delight divlu), requires u1 < v and v normalised (bit 63 set).
XF_HD uint64_t div128by64(uint64_t u1, uint64_t u0, uint64_t v, uint64_t *rem) {
    const uint64_t b = 1ULL << 32;
    const uint64_t vn1 = v >> 32, vn0 = v & 0xffffffffULL;
    const uint64_t un1 = u0 >> 32, un0 = u0 & 0xffffffffULL;
    uint64_t q1 = u1 / vn1, rhat = u1 - q1 * vn1;
    while (q1 >= b || q1 * vn0 > b * rhat + un1) { --q1; rhat += vn1; if (rhat >= b) break; }
    const uint64_t un21 = u1 * b + un1 - q1 * v;
    uint64_t q0 = un21 / vn1; rhat = un21 - q0 * vn1;
    while (q0 >= b || q0 * vn0 > b * rhat + un0) { --q0; rhat += vn1; if (rhat >= b) break; }
    *rem = un21 * b + un0 - q0 * v;
    return q1 * b + q0;
}

XF_HD f80 div(const f80 &a, const f80 &b) {
    const uint32_t sign = a.sign ^ b.sign;
    if (a.cls | b.cls) {
        if (is_nan(a) || is_nan(b)) return nan();
        if (is_inf(a)) return is_inf(b) ? nan() : inf(sign);
        return zero(sign); // finite / inf
    }
    if (b.sig == 0) return a.sig == 0 ? nan() : inf(sign);
    if (a.sig == 0) return zero(sign);
    uint64_t q, r; int32_t exp;
    if (a.sig >= b.sig) { q = div128by64(a.sig >> 1, a.sig << 63, b.sig, &r); exp = a.exp - b.exp; }
    else { q = div128by64(a.sig, 0, b.sig, &r); exp = a.exp - b.exp - 1; }
    // round to nearest even on the remainder: compare 2r with b.sig without overflow
    if (r) {
        const uint64_t other = b.sig - r;
        if (r > other || (r == other && (q & 1))) { ++q; if (q == 0) { q = 1ULL << 63; ++exp; } }
    }
    return make(q, exp, sign);
}

// three-way compare for non-NaN values: -1, 0, 1
XF_HD int cmp(const f80 &a, const f80 &b) {
    if (is_inf(a) || is_inf(b)) {
        const int av = is_inf(a) ? (a.sign ? -2 : 2) : 0, bv = is_inf(b) ? (b.sign ? -2 : 2) : 0;
        return av == bv ? 0 : (av > bv ? 1 : -1);
    }
    const bool az = a.sig == 0, bz = b.sig == 0;
    if (az && bz) return 0;
    if (az) return b.sign ? 1 : -1;
    if (bz) return a.sign ? -1 : 1;
    if (a.sign != b.sign) return a.sign ? -1 : 1;
    const int m = cmp_mag(a, b);
    return a.sign ? -m : m;
}
XF_HD bool lt(const f80 &a, const f80 &b) { return !(is_nan(a) || is_nan(b)) && cmp(a, b) < 0; }
XF_HD bool le(const f80 &a, const f80 &b) { return !(is_nan(a) || is_nan(b)) && cmp(a, b) <= 0; }
// std::max(a,b) = (a < b) ? b : a ; std::min(a,b) = (b < a) ? b : a
XF_HD f80 max_std(const f80 &a, const f80 &b) { return lt(a, b) ? b : a; }
XF_HD f80 min_std(const f80 &a, const f80 &b) { return lt(b, a) ? b : a; }

// narrow to IEEE binary32 / binary64, round-to-nearest-even, with overflow to inf and gradual underflow
XF_HD uint64_t narrow_bits(const f80 &a, int mant_bits, int exp_bits) {
    const int bias = (1 << (exp_bits - 1)) - 1, emax = bias, emin = 1 - bias;
    const uint64_t signbit = (uint64_t)a.sign << (mant_bits + exp_bits);
    const uint64_t expmask = ((1ULL << exp_bits) - 1) << mant_bits;
    if (is_nan(a)) return signbit | expmask | (1ULL << (mant_bits - 1));
    if (is_inf(a)) return signbit | expmask;
    if (a.sig == 0) return signbit;
    int32_t e = a.exp;
    int shift = 63 - mant_bits; // bits to drop for a normal result
    if (e < emin) { // subnormal target: drop more bits
        const int64_t extra = (int64_t)emin - e;
        if (extra > 64) return signbit; // far below the smallest subnormal
        shift += (int)extra;
        e = emin;
    }
    uint64_t kept, half, rest;
    if (shift >= 64) { kept = 0; half = (shift == 64) ? (a.sig >> 63) : 0; rest = (shift == 64) ? (a.sig << 1) : a.sig; }
    else { kept = a.sig >> shift; half = (a.sig >> (shift - 1)) & 1; rest = a.sig & ((1ULL << (shift - 1)) - 1); }
    if (half && (rest || (kept & 1))) ++kept;
    // kept holds an integer significand with the hidden bit at position mant_bits (if normal)
    uint64_t biased;
    if (kept >> (mant_bits + 1)) { kept >>= 1; ++e; }
    if (kept >> mant_bits) { // normal (or rounded up into normal)
        if (e > emax) return signbit | expmask;
        biased = (uint64_t)(e + bias);
        return signbit | (biased << mant_bits) | (kept & ((1ULL << mant_bits) - 1));
    }
    return signbit | kept; // subnormal
}
XF_HD float to_float(const f80 &a) {
    const uint32_t b = (uint32_t)narrow_bits(a, 23, 8); float f; memcpy(&f, &b, 4); return f;
}
XF_HD double to_double(const f80 &a) {
    const uint64_t b = narrow_bits(a, 52, 11); double d; memcpy(&d, &b, 8); return d;
}

#if defined(__x86_64__) || defined(__i386__)
// host-only bridges to the native x87 type, used to build constants and by the tests
inline f80 from_long_double(long double v) {
    unsigned char raw[16] = {0}; memcpy(raw, &v, 10);
    uint64_t sig; uint16_t se; memcpy(&sig, raw, 8); memcpy(&se, raw + 8, 2);
    const uint32_t sign = se >> 15; const int32_t e = se & 0x7fff;
    if (e == 0x7fff) return (sig << 1) ? nan() : inf(sign);
    if (e == 0 && sig == 0) return zero(sign);
    if (e == 0) { const int lz = __builtin_clzll(sig); return make(sig << lz, -16382 - lz, sign); }
    return make(sig, e - 16383, sign);
}
inline long double to_long_double(const f80 &a) {
    if (is_nan(a)) return __builtin_nanl("");
    if (is_inf(a)) return a.sign ? -__builtin_infl() : __builtin_infl();
    if (a.sig == 0) return a.sign ? -0.0L : 0.0L;
    unsigned char raw[16] = {0};
    const uint16_t se = (uint16_t)((a.sign << 15) | (uint16_t)(a.exp + 16383));
    memcpy(raw, &a.sig, 8); memcpy(raw + 8, &se, 2);
    long double v; memcpy(&v, raw, 10 > sizeof(v) ? sizeof(v) : 10);
    return v;
}
#endif

} // namespace xf
