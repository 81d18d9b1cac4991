// common.cuh -- integer hashes and small device utilities shared by the sketch / compare kernels.
// Each function cites the reference arithmetic it reproduces (paths under /root/reference).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#define D2G_HD __host__ __device__ __forceinline__

namespace d2g {

// Thomas Wang 64-bit mix -- bonsai/hll/include/sketch/hash.h:42-62 (WangHash)
D2G_HD uint64_t wang64(uint64_t key) {
    key = (~key) + (key << 21);
    key ^= key >> 24;
    key = (key + (key << 3)) + (key << 8);
    key ^= key >> 14;
    key = (key + (key << 2)) + (key << 4);
    key ^= key >> 28;
    key += key << 31;
    return key;
}

// Inverse of wang64 (WangHash::inverse): multiplicative inverses mod 2^64 are compile-time constants.
__host__ __device__ constexpr uint64_t modinv64(uint64_t a) {
    uint64_t x = a;
    for (int i = 0; i < 6; ++i) x *= 2 - a * x;
    return x;
}
D2G_HD uint64_t wang64_inv(uint64_t h) {
    constexpr uint64_t i1 = modinv64((1ULL << 31) + 1), i21 = modinv64(21), i265 = modinv64(265),
                       i2m = modinv64((1ULL << 21) - 1);
    h *= i1;
    h ^= h >> 28; h ^= h >> 56;
    h *= i21;
    h ^= (h >> 14) ^ (h >> 28) ^ (h >> 42) ^ (h >> 56);
    h *= i265;
    h ^= (h >> 24) ^ (h >> 48);
    return (h + 1) * i2m;
}

constexpr uint64_t CE_XOR1 = 0x533f8c2151b20f97ULL;
constexpr uint64_t CE_MUL = 0x9a98567ed20c127dULL | 1ULL;
constexpr uint64_t CE_XOR2 = 0x691a9d706391077aULL;

// FRev64: minimizer ordering key -- bonsai/include/bonsai/encoder.h:47,59 (lex_score)
D2G_HD uint64_t frev64(uint64_t x) {
    x ^= CE_XOR1; x *= CE_MUL; x = (x << 31) | (x >> 33);
    return x ^ CE_XOR2;
}
D2G_HD uint64_t frev64_inv(uint64_t s) {
    constexpr uint64_t imul = modinv64(CE_MUL);
    s ^= CE_XOR2; s = (s >> 31) | (s << 33); s *= imul;
    return s ^ CE_XOR1;
}
// CEHasher -- hash.h:858
D2G_HD uint64_t cehash(uint64_t x) { x ^= CE_XOR1; x *= CE_MUL; return x ^ CE_XOR2; }

// wyhash64_stateless -- bonsai/hll/include/aesctr/wy.h:45-59
D2G_HD uint64_t wymum(uint64_t x, uint64_t y) {
#if defined(__CUDA_ARCH__)
    return (x * y) ^ __umul64hi(x, y);
#else
    unsigned __int128 l = (unsigned __int128)x * y;
    return (uint64_t)l ^ (uint64_t)(l >> 64);
#endif
}
D2G_HD uint64_t wyhash64(uint64_t &state) {
    state += 0x60bee2bee120fc15ULL;
    return wymum(state ^ 0xe7037ed1a0b428dbULL, state);
}

// DHasher of the one-permutation sketch -- src/oph.h:44-71.
// seed_ = std::mt19937_64(0x321b919a61cb41f7)() = 0x8f1896f3f85ef4a3 (libstdc++; fixed constant).
constexpr uint64_t OPH_SEED = 0x8f1896f3f85ef4a3ULL;
constexpr uint64_t OPH_PREXOR = OPH_SEED ^ CE_XOR1;
D2G_HD uint64_t dhash(uint64_t x) { return wang64(x ^ OPH_PREXOR); }
D2G_HD uint64_t dhash_inv(uint64_t h) { return wang64_inv(h) ^ OPH_PREXOR; }

// reverse complement of a 2-bit packed k-mer -- bonsai/include/bonsai/kmerutil.h:83-90
__device__ __forceinline__ uint64_t revcomp(uint64_t x, int k) {
    x = __brevll(x);                                                     // reverse all 64 bits
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1); // restore bit order inside each base
    return (~x) >> (64 - 2 * k);
}
__device__ __forceinline__ uint64_t canonical(uint64_t x, int k) {
    const uint64_t rc = revcomp(x, k);
    return x < rc ? x : rc;
}

// exact x % d for 32-bit operands with a precomputed magic (Lemire fastmod; the reference's
// Schismatic<uint32_t>::mod, bonsai/hll/include/sketch/div.h:256-262, is also exact)
struct FastMod32 {
    uint64_t M; uint32_t d; uint32_t pow2mask; // pow2mask != 0 => d is a power of two
};
inline FastMod32 make_fastmod32(uint32_t d) {
    FastMod32 f; f.d = d; f.M = d > 1 ? (~0ULL / d + 1) : 0; f.pow2mask = (d & (d - 1)) == 0 ? d - 1 : 0;
    return f;
}
__device__ __forceinline__ uint32_t fastmod32(uint32_t a, const FastMod32 &f) {
    if (f.pow2mask) return a & f.pow2mask;
    const uint64_t low = f.M * a;
    return (uint32_t)__umul64hi(low, f.d);
}

// order-preserving u64 key of a double (and back)
__host__ __device__ __forceinline__ uint64_t dkey(double d) {   // order-preserving u64 key of a double
    uint64_t b;
#if defined(__CUDA_ARCH__)
    b = (uint64_t)__double_as_longlong(d);
#else
    memcpy(&b, &d, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__host__ __device__ __forceinline__ double dunkey(uint64_t k) {
    const uint64_t b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d; memcpy(&d, &b, 8); return d;
#endif
}

} // namespace d2g
