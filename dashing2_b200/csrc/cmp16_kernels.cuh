// cmp16_kernels.cuh -- K7b: all-pairs register comparison on 16-bit order codes.
//
// compare() (/root/reference/src/cmp_core.cpp:458-517) only needs, per pair of sketches, HOW MANY of
// the S register positions hold a larger / smaller / different value (count_gtlt<double> and count_eq,
// bonsai/hll/include/sketch/count_eq.h:412-445,40-56).  Those counts are invariant under any strictly
// increasing map applied to one register position of all sketches at once.  So, per comparison job,
// every register position (column of the n x S register matrix) is replaced by the DENSE RANK of the
// value among the sketches of the job, and the rank is stored as the rank-th smallest finite IEEE
// binary16 value (63 487 distinct finite halfs: -65504 .. -2^-24, +0 .. +65504).  Two registers
// compare exactly like their codes, and the SM compares two codes per instruction (HSET2 on a packed
// half2).  Per two register pairs the inner loop issues 2 HSET2 (mask output) + 1 IADD3 (two masks
// accumulated at once) instead of 4 DSETP + 4 IADD of the f64 kernel, on a quarter of the shared-memory
// bytes.  The counts (and therefore the float32 results, finalised by the same finalize_pair) are
// bit-identical to the f64 kernel's; jobs the codes cannot express (NaN registers, more than 63 487
// sketches per block pair) run on cmp_tile_kernel.
//
// Layout in HBM.  codes: uint32 words [code block][k-pair][64]; code block b holds sketches
// 64b..64b+63 of the job, word (b, kp, r) = code(register 2kp) | code(register 2kp+1) << 16 of sketch
// 64b + r.  A chunk of C16_KC k-pairs of one block is one contiguous 8 KiB run, fetched with a single
// bulk asynchronous copy (cp.async.bulk -> UBLKCP) into a C16_STAGES-deep shared-memory ring guarded by
// mbarriers.  A CTA owns 128 x 64 pairs (two row blocks x one column block); a thread owns 8 x 4 pairs.
#pragma once
#include <cuda_fp16.h>
#include "cmp_kernels.cuh"
#include "async_copy.cuh"

namespace d2g {

constexpr int C16_BLK = 64;          // sketches per code block
constexpr int C16_KC = 32;           // k-pairs per chunk (64 registers)
constexpr int C16_STAGES = 4;
constexpr int C16_THREADS = 256;
constexpr int C16_TM = 128, C16_TN = 64;
constexpr uint32_t C16_MAXRANK = 63487;   // distinct finite binary16 values (one zero)
constexpr int C16_CHUNK_WORDS = C16_KC * C16_BLK;                 // 2048 words = 8 KiB
constexpr size_t C16_SMEM = (size_t)C16_STAGES * 3 * C16_CHUNK_WORDS * 4 + C16_STAGES * 8 + 64;

// rank (0-based, < C16_MAXRANK) -> bit pattern of the rank-th smallest finite half
__host__ __device__ __forceinline__ uint16_t rank_to_half(uint32_t r) {
    return r < 31743u ? (uint16_t)(0xfbffu - r) : (uint16_t)(r - 31743u);
}

struct C16Args {
    const uint32_t *codes;   // [nblocks][KP][64]
    uint32_t KP;             // k-pairs per sketch, multiple of C16_KC (registers >= S hold code 0 everywhere)
    uint32_t a_blk0, b_blk0; // first code block of the row / column operand
    uint32_t n_a, n_b;       // row / column sketches of this job
    uint64_t gi0, gj0;       // global sketch index of local row 0 / local column 0
    uint32_t tiles_j;
    const int *use_flag;     // device flag: run only when *use_flag == want (nullptr = always)
    int want;
    uint32_t one;            // = 1; multiplier of the IMAD accumulation (a kernel parameter so that ptxas keeps the IMAD)
    int ne_is_gt;            // MODE 1 on gt/lt registers with power-of-two S: pass (ne, 0) as (gt, lt), see cmp16_tile_kernel
    CmpArgs o;               // output mapping + finalisation constants (regs unused)
};

__device__ __forceinline__ uint32_t hgt2m(uint32_t a, uint32_t b) { uint32_t d; asm("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hlt2m(uint32_t a, uint32_t b) { uint32_t d; asm("set.lt.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hne2m(uint32_t a, uint32_t b) { uint32_t d; asm("set.ne.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

// Sum (mod 2^32) of masks whose 16-bit lanes are 0x0000 / 0xffff -> number of set lanes.
// sum = (L - H) * 2^16 - L with L, H the low / high lane counts (each < 2^16).
__device__ __forceinline__ uint32_t mask_sum_count(uint32_t s) {
    const uint32_t L = (0u - s) & 0xffffu;
    const uint32_t H = (L - ((s + L) >> 16)) & 0xffffu;
    return L + H;
}

// acc += m on the FMA pipe (IMAD) instead of the ALU pipe (IADD3): HSET2 executes on the ALU pipe, which
// is the bound of this kernel (ncu: pipe_alu 98 %, pipe_fma 0.5 % with IADD3 accumulation).
__device__ __forceinline__ void acc_imad(uint32_t &acc, uint32_t m, uint32_t one) {
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(m), "r"(one));
}

// MODE 0: count (a > b, a < b); MODE 1: count (a != b).
// MODE 1 also serves gt/lt registers when S is a power of two (ne_is_gt): gt/S, lt/S and every sum the
// reference forms from them (cmp_core.cpp:461-470: 1 - a - b, 2 - a - b) are then exact in long double, so the
// result depends on gt + lt only and (ne, 0) finalises to the same float as (gt, lt).
// ACC: 1 = IMAD accumulation (FMA pipe; the shipped path), 0 = IADD3 (two masks per instruction, ALU pipe; kept as the
// measured alternative: 17.6 ms vs 13.1 ms gt/lt, 9.0 ms vs 7.1 ms != on 10 000 x 4096, profiles/r1b_*).
template <int MODE, int ACC>
__global__ void __launch_bounds__(C16_THREADS, 2)
cmp16_tile_kernel(const C16Args a) {
    extern __shared__ __align__(128) unsigned char c16_smem[];
    if (a.use_flag && *a.use_flag != a.want) return;
    uint32_t *sA = reinterpret_cast<uint32_t *>(c16_smem);                      // [STAGES][2][KC][64]
    uint32_t *sB = sA + C16_STAGES * 2 * C16_CHUNK_WORDS;                        // [STAGES][KC][64]
    uint64_t *bar = reinterpret_cast<uint64_t *>(sB + C16_STAGES * C16_CHUNK_WORDS);
    const uint32_t tj = blockIdx.x % a.tiles_j, ti = blockIdx.x / a.tiles_j;
    const uint32_t li0 = ti * C16_TM, lj0 = tj * C16_TN;                         // local first row / column
    const uint64_t i0 = a.gi0 + li0, j0 = a.gj0 + lj0;
    if (a.o.shape == 0 && j0 + C16_TN <= i0 + 1) return;                          // tile entirely on/below the diagonal
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const uint32_t nchunks = a.KP / C16_KC;
    const uint32_t *gA = a.codes + (uint64_t)(a.a_blk0 + 2 * ti) * a.KP * C16_BLK;
    const uint32_t *gB = a.codes + (uint64_t)(a.b_blk0 + tj) * a.KP * C16_BLK;
    const uint64_t blk_stride = (uint64_t)a.KP * C16_BLK;                        // words between code blocks

    if (tid == 0) {
        #pragma unroll
        for (int s = 0; s < C16_STAGES; ++s) mbar_init(bar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint32_t st, uint32_t chunk) {
        mbar_expect_tx(bar + st, 3 * C16_CHUNK_WORDS * 4);
        const uint64_t off = (uint64_t)chunk * C16_CHUNK_WORDS;
        bulk_g2s(sA + (st * 2 + 0) * C16_CHUNK_WORDS, gA + off, C16_CHUNK_WORDS * 4, bar + st);
        bulk_g2s(sA + (st * 2 + 1) * C16_CHUNK_WORDS, gA + blk_stride + off, C16_CHUNK_WORDS * 4, bar + st);
        bulk_g2s(sB + st * C16_CHUNK_WORDS, gB + off, C16_CHUNK_WORDS * 4, bar + st);
    };
    if (tid == 0)
        for (uint32_t s = 0; s < (uint32_t)C16_STAGES && s < nchunks; ++s) issue(s, s);

    uint32_t acc0[8][4], acc1[8][4];
    #pragma unroll
    for (int u = 0; u < 8; ++u)
        #pragma unroll
        for (int v = 0; v < 4; ++v) { acc0[u][v] = 0; acc1[u][v] = 0; }

    // thread's operands inside a stage: rows ty*8..+7 -> block ty>>3, offset (ty&7)*8; columns tx*4..+3
    const uint32_t one = a.one;
    const uint32_t aoff = (ty >> 3) * C16_CHUNK_WORDS + (ty & 7) * 8;
    const uint32_t boff = tx * 4;
    for (uint32_t c = 0; c < nchunks; ++c) {
        const uint32_t st = c % C16_STAGES, ph = (c / C16_STAGES) & 1;
        mbar_wait(bar + st, ph);
        const uint32_t *pa = sA + st * 2 * C16_CHUNK_WORDS + aoff;
        const uint32_t *pb = sB + st * C16_CHUNK_WORDS + boff;
        if (ACC == 1) {
            // one k-pair per step: every mask goes straight into an IMAD, operands stay at 12 registers
            #pragma unroll 4
            for (int kp = 0; kp < C16_KC; ++kp) {
                const uint4 x0 = *reinterpret_cast<const uint4 *>(pa + kp * C16_BLK);
                const uint4 x1 = *reinterpret_cast<const uint4 *>(pa + kp * C16_BLK + 4);
                const uint4 y = *reinterpret_cast<const uint4 *>(pb + kp * C16_BLK);
                const uint32_t av[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                const uint32_t bv[4] = {y.x, y.y, y.z, y.w};
                #pragma unroll
                for (int u = 0; u < 8; ++u)
                    #pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        if (MODE == 0) { acc_imad(acc0[u][v], hgt2m(av[u], bv[v]), one); acc_imad(acc1[u][v], hlt2m(av[u], bv[v]), one); }
                        else acc_imad(acc0[u][v], hne2m(av[u], bv[v]), one);
                    }
            }
        } else {
        #pragma unroll 2
        for (int kp = 0; kp < C16_KC; kp += 2) {
            uint32_t av[2][8], bv[2][4];
            #pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint4 x0 = *reinterpret_cast<const uint4 *>(pa + (kp + h) * C16_BLK);
                const uint4 x1 = *reinterpret_cast<const uint4 *>(pa + (kp + h) * C16_BLK + 4);
                const uint4 y = *reinterpret_cast<const uint4 *>(pb + (kp + h) * C16_BLK);
                av[h][0] = x0.x; av[h][1] = x0.y; av[h][2] = x0.z; av[h][3] = x0.w;
                av[h][4] = x1.x; av[h][5] = x1.y; av[h][6] = x1.z; av[h][7] = x1.w;
                bv[h][0] = y.x; bv[h][1] = y.y; bv[h][2] = y.z; bv[h][3] = y.w;
            }
            #pragma unroll
            for (int u = 0; u < 8; ++u)
                #pragma unroll
                for (int v = 0; v < 4; ++v) {
                    if (MODE == 0) {
                        acc0[u][v] += hgt2m(av[0][u], bv[0][v]) + hgt2m(av[1][u], bv[1][v]);
                        acc1[u][v] += hlt2m(av[0][u], bv[0][v]) + hlt2m(av[1][u], bv[1][v]);
                    } else {
                        acc0[u][v] += hne2m(av[0][u], bv[0][v]) + hne2m(av[1][u], bv[1][v]);
                    }
                }
        }
        }
        __syncthreads();                                   // every thread is done reading stage st
        if (tid == 0 && c + C16_STAGES < nchunks) issue(st, c + C16_STAGES);
    }

    // Epilogue through shared memory (the ring is free now: the last iteration ended with a barrier and no copy is
    // in flight): counts are parked as (c0 | c1 << 16), then consecutive threads finalise consecutive columns of a
    // row -- one copy of finalize_pair, coalesced stores.
    uint32_t *sC = sA;                                         // [C16_TM][C16_TN]
    #pragma unroll
    for (int u = 0; u < 8; ++u)
        #pragma unroll
        for (int v = 0; v < 4; ++v) {
            const uint32_t n0 = mask_sum_count(acc0[u][v]);
            const uint32_t n1 = MODE == 0 ? mask_sum_count(acc1[u][v]) : 0;
            sC[(ty * 8 + u) * C16_TN + tx * 4 + v] = n0 | (n1 << 16);
        }
    __syncthreads();
    const uint32_t S = a.o.c.S;
    for (int e = tid; e < C16_TM * C16_TN; e += C16_THREADS) {
        const uint32_t li = li0 + e / C16_TN, lj = lj0 + e % C16_TN;
        if (li >= a.n_a || lj >= a.n_b) continue;
        const uint64_t i = a.gi0 + li, j = a.gj0 + lj;
        if (i < a.o.row0 || i >= a.o.row1 || j < a.o.col0 || j >= a.o.col1) continue;   // operand blocks may overhang the launch's rows / columns
        if (a.o.shape == 0 && j <= i) continue;
        const uint32_t pk = sC[e], n0 = pk & 0xffffu;
        const uint32_t c0 = (MODE == 0 || a.ne_is_gt) ? n0 : S - n0;
        const uint32_t c1 = pk >> 16;
        const uint64_t oi = out_index(a.o, i, j);
        if (a.o.c0_out) { a.o.c0_out[oi] = c0; if (a.o.c1_out) a.o.c1_out[oi] = c1; }
        if (a.o.out) a.o.out[oi] = finalize_pair(a.o.c, c0, c1, a.o.cards ? __ldg(a.o.cards + i) : 0., a.o.cards ? __ldg(a.o.cards + j) : 0.);
    }
}

// ---- building the codes ---------------------------------------------------------------------------
// (1) keys: transposes f64 registers of the job's sketches into column-major order-preserving u64 keys.
//     Job sketch u (0 <= u < U) is global sketch (u < nA ? gA0 + u : gB0 + u - nA).
//     GTLT: key = dkey(value) with -0 folded onto +0 (they compare equal); a NaN raises *nan_flag (NaN is
//     unordered: neither >, < nor expressible as a rank).  EQ: key = the raw bit pattern (cmp_kernels.cuh
//     counts bitwise-different registers).
struct C16Job {
    const double *regs; uint32_t S;
    uint64_t gA0, gB0; uint32_t nA, nB;      // two ranges of global sketches; nB may be 0
    uint32_t posB0;                          // position (in sketches, multiple of 64) of the first column sketch in code space
    uint32_t KP;
    uint32_t s_begin, s_count;               // register positions [s_begin, s_begin + s_count) handled by this launch (sort buffers hold s_count columns)
};
__device__ __forceinline__ uint64_t job_sketch(const C16Job &j, uint32_t u) { return u < j.nA ? j.gA0 + u : j.gB0 + (u - j.nA); }
__device__ __forceinline__ uint32_t job_pos(const C16Job &j, uint32_t u) { return u < j.nA ? u : j.posB0 + (u - j.nA); }

template <int KIND>
__global__ void __launch_bounds__(256)
c16_keys_kernel(const C16Job j, uint64_t *keysT, uint32_t *idxT, int *nan_flag) {
    __shared__ uint64_t t[32][33];
    const uint32_t U = j.nA + j.nB;
    const uint32_t u0 = blockIdx.x * 32, s0 = blockIdx.y * 32;      // s: column index inside this launch's group
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;   // 32 x 8
    bool nan = false;
    #pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t u = u0 + ly + 8 * r, s = s0 + lx;
        uint64_t key = 0;
        if (u < U && s < j.s_count) {
            const double d = __ldg(j.regs + job_sketch(j, u) * j.S + j.s_begin + s);
            if (KIND == 0) { nan |= d != d; key = dkey(d == 0. ? 0. : d); }
            else { nan |= d != d; key = (uint64_t)__double_as_longlong(d == 0. ? 0. : d); }   // IEEE ==: -0 folds onto +0, NaN never matches (f64 kernel)
        }
        t[ly + 8 * r][lx] = key;
    }
    if (nan) *nan_flag = 1;
    __syncthreads();
    #pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t s = s0 + ly + 8 * r, u = u0 + lx;
        if (u < U && s < j.s_count) { keysT[(uint64_t)s * U + u] = t[lx][ly + 8 * r]; idxT[(uint64_t)s * U + u] = u; }
    }
}

// (2) after the per-column sort: dense ranks -> half codes, scattered into the blocked code layout.
//     One CTA per register position; codes16 is the uint16 view of the code words.
// grank != nullptr: the job is the whole sketch range of a multi-job comparison; write the dense ranks themselves
// (u32 [S][U]) instead of codes -- c16_local_codes_kernel turns them into per-job codes without sorting again.
__global__ void __launch_bounds__(256)
c16_rank_kernel(const C16Job j, const uint64_t *keys_sorted, const uint32_t *idx_sorted, uint16_t *codes16, uint32_t *grank, int *overflow_flag) {
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t carry_s;
    const uint32_t U = j.nA + j.nB, sl = blockIdx.x, s = j.s_begin + sl;
    const uint64_t *k = keys_sorted + (uint64_t)sl * U;
    const uint32_t *ix = idx_sorted + (uint64_t)sl * U;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < U; base += 256) {
        const uint32_t i = base + threadIdx.x;
        uint32_t f = 0;
        if (i < U && i > 0) f = k[i] != k[i - 1];
        uint32_t x = f;                                   // inclusive scan of the "new value" flags
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        uint32_t pre = carry_s;
        for (int q = 0; q < w; ++q) pre += wsum[q];
        const uint32_t rank = pre + x;
        if (i < U) {
            if (grank) grank[(uint64_t)s * U + ix[i]] = rank;
            else {
                if (rank >= C16_MAXRANK) *overflow_flag = 1;
                const uint32_t pos = job_pos(j, ix[i]);
                const uint64_t word = ((uint64_t)(pos / C16_BLK) * j.KP + (s >> 1)) * C16_BLK + (pos % C16_BLK);
                codes16[word * 2 + (s & 1)] = rank_to_half(rank);
            }
        }
        __syncthreads();
        if (threadIdx.x == 255) carry_s = rank;
        __syncthreads();
    }
}

// (2b) jobs that only need to know WHICH registers differ (MODE 1) and are small enough for a shared-memory table skip
//      the sort: any injective code per register position will do, and the slot a value lands in in an open-addressing
//      table of the position's values is one.  A CTA handles four adjacent register positions one after the other (one
//      32-byte sector of a sketch row serves all four), 64-bit keys, linear probing, code = rank_to_half(slot).
constexpr uint32_t C16_HASH_MAX_SKETCHES = 16384;     // table of <= 24576 slots (192 KiB) at load <= 2/3
constexpr uint64_t C16_HASH_EMPTY = ~0ULL;            // KIND 0: an unused NaN image; KIND 1: handled as a value of its own
template <int KIND>
__global__ void __launch_bounds__(512)
c16_hash_codes_kernel(const C16Job j, uint32_t TS, uint16_t *codes16, int *nan_flag) {
    extern __shared__ unsigned long long hc_tab[];
    const uint32_t U = j.nA + j.nB;
    bool nan = false;
    for (uint32_t s = blockIdx.x * 4; s < min(j.S, blockIdx.x * 4 + 4); ++s) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < TS; i += blockDim.x) hc_tab[i] = C16_HASH_EMPTY;
        __syncthreads();
        for (uint32_t u = threadIdx.x; u < U; u += blockDim.x) {
            const double d = __ldg(j.regs + job_sketch(j, u) * j.S + s);
            uint64_t key;
            if (KIND == 0) { nan |= d != d; key = dkey(d == 0. ? 0. : d); }
            else { nan |= d != d; key = (uint64_t)__double_as_longlong(d == 0. ? 0. : d); }   // IEEE ==: -0 folds onto +0, NaN never matches (f64 kernel)
            uint32_t slot;
            if (key == C16_HASH_EMPTY) slot = TS;                        // cannot live in the table: its own code
            else {
                uint64_t h = key * 0x9E3779B97F4A7C15ULL; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL;
                slot = (uint32_t)(((h >> 32) * TS) >> 32);
                for (;;) {
                    const unsigned long long old = atomicCAS(hc_tab + slot, (unsigned long long)C16_HASH_EMPTY, (unsigned long long)key);
                    if (old == C16_HASH_EMPTY || old == key) break;
                    if (++slot == TS) slot = 0;
                }
            }
            const uint32_t pos = job_pos(j, u);
            const uint64_t word = ((uint64_t)(pos / C16_BLK) * j.KP + (s >> 1)) * C16_BLK + (pos % C16_BLK);
            codes16[word * 2 + (s & 1)] = rank_to_half(slot);
        }
    }
    if (nan) *nan_flag = 1;
}

// (3) multi-job comparisons: global dense ranks (u32 [S][N], sketch g at column g - g0) -> codes of one job.
//     One CTA per register position: presence bitmap of the ranks the job's sketches hold, exclusive prefix popcount,
//     local rank = number of present ranks below.  Shared memory: 2 * ceil(N / 32) words.
__global__ void __launch_bounds__(256)
c16_local_codes_kernel(const C16Job j, const uint32_t *grank, uint64_t g0, uint32_t N, uint16_t *codes16) {
    extern __shared__ uint32_t lc_smem[];
    __shared__ uint32_t part[256];
    const uint32_t W = (N + 31) / 32, U = j.nA + j.nB, s = blockIdx.x;
    uint32_t *bits = lc_smem, *pre = lc_smem + W;
    const uint32_t *gr = grank + (uint64_t)s * N;
    for (uint32_t i = threadIdx.x; i < W; i += 256) bits[i] = 0;
    __syncthreads();
    for (uint32_t u = threadIdx.x; u < U; u += 256) {
        const uint32_t r = gr[job_sketch(j, u) - g0];
        atomicOr(bits + (r >> 5), 1u << (r & 31));
    }
    __syncthreads();
    const uint32_t per = (W + 255) / 256, w0 = threadIdx.x * per, w1 = min(W, w0 + per);
    uint32_t sum = 0;
    for (uint32_t i = w0; i < w1; ++i) sum += __popc(bits[i]);
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t acc = 0; for (int t = 0; t < 256; ++t) { const uint32_t v = part[t]; part[t] = acc; acc += v; } }
    __syncthreads();
    uint32_t acc = part[threadIdx.x];
    for (uint32_t i = w0; i < w1; ++i) { pre[i] = acc; acc += __popc(bits[i]); }
    __syncthreads();
    for (uint32_t u = threadIdx.x; u < U; u += 256) {
        const uint32_t r = gr[job_sketch(j, u) - g0];
        const uint32_t local = pre[r >> 5] + __popc(bits[r >> 5] & ((1u << (r & 31)) - 1u));
        const uint32_t pos = job_pos(j, u);
        const uint64_t word = ((uint64_t)(pos / C16_BLK) * j.KP + (s >> 1)) * C16_BLK + (pos % C16_BLK);
        codes16[word * 2 + (s & 1)] = rank_to_half(local);
    }
}

} // namespace d2g
