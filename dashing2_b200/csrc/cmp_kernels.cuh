// cmp_kernels.cuh -- K7: tiled all-pairs register comparison + exact float32 finalisation.
//
// Replaces compare() (/root/reference/src/cmp_core.cpp:349-575; inner loops count_gtlt<double>
// bonsai/hll/include/sketch/count_eq.h:412-445 and count_eq :40-56) and the row/column orderings of
// emit_rectangular (src/emitrect.cpp:229-326).
//
// A CTA owns a 64 x 64 tile of (row sketch, column sketch) pairs.  The S registers are streamed
// through shared memory in chunks of CMP_SC; each thread keeps a 4 x 4 micro-tile of (gt, lt)
// counters in registers, so every register loaded from shared memory is used 4 times and every
// register loaded from HBM/L2 is used 64 times.  The per-pair epilogue replays the reference's
// long-double arithmetic with xf80 (software x87) so the float32 written is bit-identical.
#pragma once
#include "common.cuh"
#include "devlog.cuh"
#include "xf80.h"

namespace d2g {

constexpr int CMP_T = 64;        // tile edge (pairs)
constexpr int CMP_SC = 32;       // registers per shared-memory chunk
constexpr int CMP_LD = CMP_SC + 1; // padded leading dimension (doubles): conflict-free column reads
constexpr int CMP_THREADS = 256; // 16 x 16 threads, 4 x 4 pairs each

enum { MEAS_SIM = 0, MEAS_CONTAIN = 1, MEAS_SYMCONTAIN = 2, MEAS_LLR = 3, MEAS_ISZ = 4, MEAS_USZ = 5 };

struct CmpConsts {
    xf::f80 invdenom;      // 1.L / S             (cmp_core.cpp:360)
    xf::f80 eps;           // 1e-15L              (cmp_core.cpp:476)
    double poisson_mult;   // -1. / max(1, k)     (cmp_core.cpp:361)
    uint32_t S;
    int measure;
    int cmp_kind;          // 0 gt/lt, 1 equality, 2 log-quantised registers (gt/lt through g_b), 3 b-bit signatures
    int fast_sim;          // S power of two and measure == SIMILARITY: sim = (S-gt-lt)/S exactly
    const float *eq_llr_lut; // [S+1] equality-branch POISSON_LLR values (host long double logl); kind 3: by equal count;
                             // kind 2: triangular [(gt, lt), gt + lt <= S], see tri_index
    const xf::f80 *lut80;    // kind 2: g_b(b, c/S) for c = 0..S (powl on the host); kind 3: the collision-corrected similarity by equal count
};
__host__ __device__ __forceinline__ uint64_t tri_index(uint64_t gt, uint64_t lt, uint64_t S) { return gt * (S + 1) - gt * (gt - 1) / 2 + lt; }

struct CmpArgs {
    const double *regs;    // [n][S]
    const double *cards;   // [n]
    uint64_t n;            // total sketches
    uint64_t row0, row1;   // rows (global sketch ids) this launch computes
    uint64_t col0, col1;   // columns (global sketch ids) this launch computes
    uint64_t out_row0;     // first row held by the packed output buffer
    uint64_t col_base;     // output geometry: first column sketch index (PANEL: n - nq, else 0)
    uint64_t ncols;        // output geometry: number of column sketches
    int shape;             // 0 symmetric, 1 asymmetric, 2 panel
    float *out;            // packed output for rows >= out_row0
    uint32_t *c0_out, *c1_out; // optional raw counts (row-major rows x cols), out may be null then
    uint64_t tiles_j;      // number of column tiles
    const int *use_flag;   // device flag: the launch only runs when *use_flag == want (nullptr = always)
    int want;
    CmpConsts c;
};

// float32 result of one pair from its integer counts; mirrors cmp_core.cpp:458-517 + :573
__device__ __forceinline__ float finalize_pair(const CmpConsts &c, uint32_t c0, uint32_t c1, double lhc, double rhc) {
    using namespace xf;
    if (c.fast_sim) {
        const int32_t e = (int32_t)c.S - (int32_t)c0 - (int32_t)c1;
        return e <= 0 ? 0.f : (float)e / (float)c.S;
    }
    const f80 lhcard = from_double(lhc), rhcard = from_double(rhc);
    const f80 one = from_u64(1), two = from_u64(2);
    f80 ret;
    if (c.cmp_kind == 3) {                       // cmp_core.cpp:406-424
        if (c.measure == MEAS_LLR) return c.eq_llr_lut[c0];
        ret = c.lut80[c0];
        if (c.measure != MEAS_SIM) {
            const f80 mu = max_std(div(add(lhcard, rhcard), sub(two, sub(one, ret))), zero());
            switch (c.measure) {
                case MEAS_ISZ: ret = mu; break;
                case MEAS_USZ: ret = sub(add(lhcard, rhcard), mu); break;
                case MEAS_CONTAIN: ret = div(mul(mu, ret), lhcard); break;
                case MEAS_SYMCONTAIN: ret = div(mul(mu, ret), min_std(lhcard, rhcard)); break;
                default: break;
            }
        }
    } else if (c.cmp_kind == 2) {                // cmp_core.cpp:425-448
        if (c.measure == MEAS_LLR) return c.eq_llr_lut[tri_index(c0, c1, c.S)];
        const f80 alpha = c.lut80[c0], beta = c.lut80[c1], ab = add(alpha, beta);
        const f80 mu = le(one, ab) ? add(lhcard, rhcard) : max_std(div(add(lhcard, rhcard), sub(sub(two, alpha), beta)), zero());
        ret = max_std(sub(one, ab), zero());
        switch (c.measure) {
            case MEAS_ISZ: ret = mul(ret, mu); break;
            case MEAS_USZ: ret = sub(add(lhcard, rhcard), mul(ret, mu)); break;
            case MEAS_CONTAIN: ret = div(mul(ret, mu), lhcard); break;
            case MEAS_SYMCONTAIN: ret = div(mul(ret, mu), min_std(lhcard, rhcard)); break;
            default: break;
        }
    } else if (c.cmp_kind == 0) {
        const f80 alpha = mul(from_u64(c0), c.invdenom), beta = mul(from_u64(c1), c.invdenom);
        f80 eq = sub(sub(one, alpha), beta);
        const f80 ucard = max_std(div(add(lhcard, rhcard), sub(sub(two, alpha), beta)), zero());
        if (le(eq, zero())) return c.measure != MEAS_LLR ? 0.f : __int_as_float(0x7f800000);
        if (le(eq, c.eps)) eq = zero();
        const float isz = to_float(mul(ucard, eq)), sim = to_float(eq);
        switch (c.measure) {
            case MEAS_SIM: return sim;
            case MEAS_ISZ: ret = from_float(isz); break;
            case MEAS_CONTAIN: ret = div(from_float(isz), rhcard); break;
            case MEAS_SYMCONTAIN: ret = div(from_float(isz), min_std(lhcard, rhcard)); break;
            case MEAS_LLR: {
                if (sim == 0.f) return __int_as_float(0x7f800000);
                const double x = (double)sim;
                ret = from_double(ref_log(2. * x / (1. + x)) * c.poisson_mult);
            } break;
            case MEAS_USZ: ret = sub(add(lhcard, rhcard), from_float(isz)); break;
            default: ret = from_float(-1.f);
        }
    } else {
        if (c.measure == MEAS_LLR) return c.eq_llr_lut[c0];
        ret = mul(c.invdenom, from_u64(c0));
        const f80 mu = max_std(div(add(lhcard, rhcard), add(one, ret)), zero());
        switch (c.measure) {
            case MEAS_ISZ: ret = mul(ret, mu); break;
            case MEAS_SYMCONTAIN: ret = mul(ret, div(mu, min_std(lhcard, rhcard))); break;
            case MEAS_CONTAIN: ret = mul(ret, div(mu, lhcard)); break;
            case MEAS_USZ: ret = sub(add(lhcard, rhcard), mul(ret, mu)); break;
            default: break;
        }
    }
    if (is_nan(ret) || is_inf(ret)) return __int_as_float(0x7f800000); // LDBL_MAX -> float = inf
    return to_float(ret);
}

// position of pair (i, j) in the packed output of rows >= out_row0
__device__ __forceinline__ uint64_t out_index(const CmpArgs &a, uint64_t i, uint64_t j) {
    if (a.shape == 0) { // condensed upper triangle: row i holds j = i+1..n-1
        const uint64_t base = i * a.n - i * (i + 1) / 2 - (a.out_row0 * a.n - a.out_row0 * (a.out_row0 + 1) / 2);
        return base + (j - i - 1);
    }
    return (i - a.out_row0) * a.ncols + (j - a.col_base);
}

template <int KIND>
static __global__ void __launch_bounds__(CMP_THREADS)
cmp_tile_kernel(const CmpArgs a) {
    __shared__ double sA[CMP_T * CMP_LD];
    __shared__ double sB[CMP_T * CMP_LD];
    const uint64_t tj = blockIdx.x % a.tiles_j, ti = blockIdx.x / a.tiles_j;
    const uint64_t i0 = a.row0 + ti * CMP_T;          // first row sketch of the tile
    const uint64_t j0 = a.col0 + tj * CMP_T;          // first column sketch of the tile
    const uint64_t jend = a.col1;
    if (a.use_flag && *a.use_flag != a.want) return;
    if (a.shape == 0 && j0 + CMP_T <= i0 + 1) return; // tile entirely on/below the diagonal
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const uint32_t S = a.c.S;
    uint32_t gt[4][4], lt[4][4];
    #pragma unroll
    for (int u = 0; u < 4; ++u)
        #pragma unroll
        for (int v = 0; v < 4; ++v) gt[u][v] = lt[u][v] = 0;

    for (uint32_t r0 = 0; r0 < S; r0 += CMP_SC) {
        __syncthreads();
        // 64 sketches x 32 registers per operand: 2048 doubles, 8 per thread; 32 consecutive lanes read
        // one sketch's 256-byte run
        #pragma unroll
        for (int it = 0; it < (CMP_T * CMP_SC) / CMP_THREADS; ++it) {
            const int e = it * CMP_THREADS + threadIdx.x;
            const int row = e / CMP_SC, r = e % CMP_SC;
            const uint64_t gi = i0 + row, gj = j0 + row;
            double va = 0., vb = 0.;
            if (r0 + r < S) {
                if (gi < a.row1) va = __ldg(a.regs + gi * S + r0 + r);
                if (gj < jend) vb = __ldg(a.regs + gj * S + r0 + r);
            }
            sA[row * CMP_LD + r] = va;
            sB[row * CMP_LD + r] = vb;
        }
        __syncthreads();
        #pragma unroll 4
        for (int r = 0; r < CMP_SC; ++r) {
            double av[4], bv[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) { av[u] = sA[(ty + 16 * u) * CMP_LD + r]; bv[u] = sB[(tx + 16 * u) * CMP_LD + r]; }
            #pragma unroll
            for (int u = 0; u < 4; ++u)
                #pragma unroll
                for (int v = 0; v < 4; ++v) {
                    if (KIND == 0) { gt[u][v] += av[u] > bv[v]; lt[u][v] += av[u] < bv[v]; }
                    else gt[u][v] += !(av[u] == bv[v]);   // count_eq<double> is the IEEE `==` (count_eq.h:40-45), also over k-mer ids viewed as doubles: NaN patterns never match, -0 == +0
                }
        }
    }
    #pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint64_t i = i0 + ty + 16 * u;
        if (i >= a.row1) continue;
        const double lhc = a.cards ? __ldg(a.cards + i) : 0.;
        #pragma unroll
        for (int v = 0; v < 4; ++v) {
            const uint64_t j = j0 + tx + 16 * v;
            if (j >= jend) continue;
            if (a.shape == 0 && j <= i) continue;
            const uint32_t c0 = KIND == 0 ? gt[u][v] : S - gt[u][v];
            const uint32_t c1 = lt[u][v];
            const uint64_t oi = out_index(a, i, j);
            if (a.c0_out) { a.c0_out[oi] = c0; if (a.c1_out) a.c1_out[oi] = c1; }
            if (a.out) a.out[oi] = finalize_pair(a.c, c0, c1, lhc, __ldg(a.cards + j));
        }
    }
}

// ---- densify (src/cmp_core.cpp:577-613): one thread per empty register --------------------------
static __global__ void densify_kernel(double *sig, uint64_t *kmers, uint64_t n, uint32_t S, double *tmp, uint64_t *ktmp) {
    const uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (e >= n * S) return;
    const uint64_t g = e / S; const uint32_t i = (uint32_t)(e % S);
    const double *row = sig + g * S;
    double v = row[i];
    uint64_t kv = kmers ? kmers[g * S + i] : 0;
    if (v == 0.) {
        // all-empty sketches are left untouched (cmp_core.cpp:586-588): detect by bounded probing of
        // the deterministic sequence; a sketch with no non-empty register would loop forever
        uint64_t rng = i + 0x5bf2b8bdf07c06cULL;
        bool any = false;
        for (uint32_t q = 0; q < S; ++q) if (row[q] != 0.) { any = true; break; }
        if (any) {
            uint64_t j;
            do { j = wyhash64(rng) % S; } while (row[j] == 0.);
            v = row[j];
            if (kmers) kv = kmers[g * S + j];
        }
    }
    tmp[e] = v;
    if (kmers) ktmp[e] = kv;
}

} // namespace d2g
