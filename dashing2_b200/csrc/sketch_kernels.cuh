// sketch_kernels.cuh -- K1/K2/K3: sequence tile -> 2-bit pack -> k-mer / minimizer -> hash -> sketch update.
//
// One CTA owns a contiguous span of START POSITIONS of the concatenated sequence buffer and walks the
// records that overlap it (k-mers never cross a record; windows reset per record --
// /root/reference/bonsai/include/bonsai/encoder.h:201-206).  Per tile of SK_TILE start positions:
//   1. coalesced 16-byte loads of ASCII bases -> 2 bits/base + 1 invalid bit/base in shared memory
//      (alphabet: bonsai alphabet.h:128 DNA4, case-insensitive; anything else invalid);
//   2. each thread rolls SK_PPT consecutive k-mers (forward and reverse-complement) out of the packed
//      tile (encoder.h:241-272; kmerutil.h:83-90,137-140);
//   3. windowed mode: per-position keys FRev64(canonical k-mer) go to shared memory and every thread
//      takes the minimum over its SK_PPT windows (encoder.h:212-217, qmap.h:79-87).  FRev64 is a
//      bijection, so ordering by (score, k-mer) equals ordering by score, and the k-mer is recovered
//      from the winning score with the inverse permutation -- no (score,k-mer) pairs, no tree;
//   4. maskfn (src/enums.h:136-140) and the sketch update (a Consumer) against registers held in
//      shared memory; the CTA merges its registers into the per-entity registers in HBM when the
//      entity changes.
// Consumers: OpmhConsumer (src/oph.h:176-211), FssBootConsumer / FssMainConsumer
// (src/setsketch.h:369-423, see fss_kernels.cuh).
#pragma once
#include "common.cuh"

namespace d2g {

constexpr int SK_THREADS = 256;
constexpr int SK_PPT = 8;                       // start positions per thread per tile
constexpr int SK_TILE = SK_THREADS * SK_PPT;    // 2048 start positions per tile
constexpr int SK_MAX_W = 1024;                  // largest window (bases) the tile halo supports
constexpr int SK_NWORDS = (SK_TILE + SK_MAX_W + 16 + 31) / 32 + 2;
// Window keys live in shared memory at index q + q/8: thread t owns keys 8t..8t+7 (+ window tail), so a
// warp's simultaneous accesses are 9 u64 apart instead of 8 -- the minimum two wavefronts instead of 16.
__host__ __device__ constexpr int sk_pad(int q) { return q + (q >> 3); }
constexpr int SK_SCORE_SLOTS = sk_pad(SK_TILE + SK_MAX_W) + 1;

struct SketchArgs {
    const uint8_t *seq;          // concatenated record bytes (device), 16-byte aligned
    const uint64_t *rec_off;     // [n_rec + 1]
    const uint32_t *rec_entity;  // [n_rec]
    uint64_t n_rec;
    uint64_t total_len;
    uint64_t span;               // start positions per CTA (multiple of SK_TILE)
    int k, w, canon;
    uint64_t xormask;
    uint32_t m;                  // registers per entity
    uint32_t tile_stride;        // process only tiles whose global index % tile_stride == 0 (sampling); 1 = all
    uint32_t score_slots;        // shared-memory window-key slots (windowed mode): sk_pad(SK_TILE + w - k + 1) + 1
};

// ---- ASCII -> packed codes ---------------------------------------------------------------------
// 4 ASCII bytes (first base in the low byte) -> 8 bits of codes (first base in the two MSBs) and a
// 4-bit invalid mask (first base in bit 3).  A0 C1 G2 T3.
__device__ __forceinline__ void decode4(uint32_t v, uint32_t &codes, uint32_t &inv) {
    uint32_t x = (v >> 1) & 0x03030303u;               // A0 C1 T2 G3
    x ^= (x >> 1) & 0x01010101u;                        // A0 C1 G2 T3
    codes = (x * 0x40100401u) >> 24;
    const uint32_t u = v & 0xDFDFDFDFu;                 // fold case
    uint32_t z, nz;
    z = u ^ 0x41414141u; nz = (((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z);
    z = u ^ 0x43434343u; nz &= (((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z);
    z = u ^ 0x47474747u; nz &= (((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z);
    z = u ^ 0x54545454u; nz &= (((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z);
    inv = ((((nz & 0x80808080u) >> 7) * 0x08040201u) >> 24) & 0xFu;
}

// loads bytes [o, o + 16*n16) (o 16-byte aligned) of the sequence into the packed tile.
// codes32: as uint64 words the first base of each 32 sits in the MSBs; inv16 likewise for uint32 words.
__device__ __forceinline__ void load_tile(const SketchArgs &a, uint32_t *codes32, uint16_t *inv16, uint64_t o, int n16) {
    const uint64_t lim16 = (a.total_len + 15) >> 4;
    const uint4 *src = reinterpret_cast<const uint4 *>(a.seq);
    for (int j = threadIdx.x; j < n16; j += SK_THREADS) {
        const uint64_t g16 = (o >> 4) + j;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (g16 < lim16) v = __ldg(src + g16);
        uint32_t c0, c1, c2, c3, i0, i1, i2, i3;
        decode4(v.x, c0, i0); decode4(v.y, c1, i1); decode4(v.z, c2, i2); decode4(v.w, c3, i3);
        codes32[j ^ 1] = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
        inv16[j ^ 1] = (uint16_t)((i0 << 12) | (i1 << 8) | (i2 << 4) | i3);
    }
}

__device__ __forceinline__ uint64_t tile_kmer(const uint64_t *W, int b, int k) {
    const int i = b >> 5, s = (b & 31) * 2;
    const uint64_t hi = W[i], lo = W[i + 1];
    const uint64_t x = s ? ((hi << s) | (lo >> (64 - s))) : hi;
    return x >> (64 - 2 * k);
}
__device__ __forceinline__ uint32_t tile_code(const uint64_t *W, int b) {
    return (uint32_t)(W[b >> 5] >> (62 - 2 * (b & 31))) & 3u;
}
// non-zero iff any of the k bases starting at local index b is invalid
__device__ __forceinline__ uint32_t tile_invalid(const uint32_t *M, int b, int k) {
    const int i = b >> 5, s = b & 31;
    return __funnelshift_l(M[i + 1], M[i], s) >> (32 - k);
}

// ---- one-permutation MinHash consumer (src/oph.h:176-211) ---------------------------------------
struct OpmhConsumer {
    struct Params { uint64_t *regs; FastMod32 fm; uint32_t m; };   // regs [n_entities][m], initialised to ~0
    static __host__ __device__ size_t smem_bytes(uint32_t m) { return (size_t)m * 8; }
    uint64_t *sreg; Params p;
    __device__ __forceinline__ void init(unsigned char *smem, const Params &pp) {
        p = pp; sreg = reinterpret_cast<uint64_t *>(smem);
        for (uint32_t i = threadIdx.x; i < p.m; i += SK_THREADS) sreg[i] = ~0ULL;
    }
    static constexpr bool kEveryWindow = false;   // set semantics: consecutive equal minimizers feed the sketch once
    __device__ __forceinline__ void begin_entity(uint32_t, uint64_t) {}
    __device__ __forceinline__ void consume(uint64_t hv) {
        const uint64_t id = dhash(hv);                       // oph.h:178
        const uint32_t idx = fastmod32((uint32_t)id, p.fm);  // oph.h:184 (32-bit truncation, div.h:256-262)
        if (id < sreg[idx]) atomicMin(reinterpret_cast<unsigned long long *>(sreg + idx), (unsigned long long)id);
    }
    __device__ __forceinline__ void end_tile(uint32_t) {}
    // all threads; flushes the CTA-local registers of entity `ent` to HBM and clears them
    __device__ __forceinline__ void flush(uint32_t ent) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < p.m; i += SK_THREADS) {
            const uint64_t v = sreg[i];
            if (v != ~0ULL) { atomicMin(reinterpret_cast<unsigned long long *>(p.regs + (uint64_t)ent * p.m + i), (unsigned long long)v); sreg[i] = ~0ULL; }
        }
        __syncthreads();
    }
};

// ---- the kernel ---------------------------------------------------------------------------------
template <bool WINDOWED, class Consumer>
__global__ void __launch_bounds__(SK_THREADS)
sketch_kernel(const SketchArgs a, const typename Consumer::Params cp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = a.k;
    const int need = WINDOWED ? a.w : a.k;           // bases a start position needs to its right
    const int wsz = WINDOWED ? (a.w - a.k + 1) : 1;  // k-mers per window
    uint64_t *W = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *M = reinterpret_cast<uint32_t *>(W + SK_NWORDS);
    uint64_t *score = reinterpret_cast<uint64_t *>(M + SK_NWORDS + (SK_NWORDS & 1));
    unsigned char *csmem = reinterpret_cast<unsigned char *>(score + (WINDOWED ? a.score_slots : 0));
    Consumer cons;
    cons.init(csmem, cp);

    const uint64_t span_lo = (uint64_t)blockIdx.x * a.span;
    const uint64_t span_hi = min(span_lo + a.span, a.total_len);
    if (span_lo >= span_hi) return;

    // first record whose end lies beyond span_lo (records are sorted by offset)
    uint64_t lo = 0, hi = a.n_rec;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (a.rec_off[mid + 1] > span_lo) hi = mid; else lo = mid + 1; }
    uint32_t cur_ent = 0xFFFFFFFFu;
    const uint64_t kmask = k < 32 ? ((1ULL << (2 * k)) - 1) : ~0ULL;
    __syncthreads();

    for (uint64_t r = lo; r < a.n_rec; ++r) {
        const uint64_t rs = a.rec_off[r], re = a.rec_off[r + 1];
        if (rs >= span_hi) break;
        if (re - rs < (uint64_t)need) continue;
        const uint64_t p0 = max(span_lo, rs);
        const uint64_t p1 = min(span_hi, re - need + 1);   // owned, usable start positions [p0, p1)
        if (p0 >= p1) continue;
        const uint32_t ent = a.rec_entity[r];
        if (ent != cur_ent) {
            if (cur_ent != 0xFFFFFFFFu) cons.flush(cur_ent);
            cur_ent = ent;
            cons.begin_entity(ent, p0 - span_lo);
        }
        for (uint64_t t0 = p0; t0 < p1; t0 += SK_TILE) {
            if (a.tile_stride > 1 && ((t0 / SK_TILE) % a.tile_stride) != 0) continue;
            const uint64_t o = t0 & ~15ULL;
            const int off = (int)(t0 - o);
            const int nstart = (int)min((uint64_t)SK_TILE, p1 - t0);  // start positions in this tile
            const int nbases = off + nstart + need - 1;
            __syncthreads();                                          // previous tile fully consumed
            load_tile(a, reinterpret_cast<uint32_t *>(W), reinterpret_cast<uint16_t *>(M), o, (nbases + 15) >> 4);
            __syncthreads();
            if (!WINDOWED) {
                const int j0 = threadIdx.x * SK_PPT;
                if (j0 < nstart) {
                    int b = off + j0;
                    uint64_t fw = tile_kmer(W, b, k);
                    uint64_t rc = revcomp(fw, k);
                    const int jn = min(SK_PPT, nstart - j0);
                    for (int j = 0; j < jn; ++j, ++b) {
                        if (j) {
                            const uint64_t c = tile_code(W, b + k - 1);
                            fw = ((fw << 2) | c) & kmask;
                            rc = (rc >> 2) | ((3ULL - c) << (2 * k - 2));
                        }
                        if (tile_invalid(M, b, k)) continue;          // encoder.h:254 -- k-mers holding a non-ACGT base are skipped
                        const uint64_t km = a.canon ? (fw < rc ? fw : rc) : fw;
                        cons.consume(wang64(km ^ a.xormask));         // maskfn, src/enums.h:136-140
                    }
                }
            } else {
                // per-position keys for k-mer positions [0, nstart + wsz - 1) of the tile
                const int npos = nstart + wsz - 1;
                for (int q0 = threadIdx.x * SK_PPT; q0 < npos; q0 += SK_TILE) {
                    int b = off + q0;
                    uint64_t fw = tile_kmer(W, b, k);
                    uint64_t rc = revcomp(fw, k);
                    const int qn = min(SK_PPT, npos - q0);
                    for (int j = 0; j < qn; ++j, ++b) {
                        if (j) {
                            const uint64_t c = tile_code(W, b + k - 1);
                            fw = ((fw << 2) | c) & kmask;
                            rc = (rc >> 2) | ((3ULL - c) << (2 * k - 2));
                        }
                        // canonical windowed path: a k-mer holding an invalid base enters the window as
                        // k-mer 0 (encoder.h:568-571 + kmerutil.h:137-140)
                        const uint64_t km = tile_invalid(M, b, k) ? 0ULL : (fw < rc ? fw : rc);
                        score[sk_pad(q0 + j)] = frev64(km);
                    }
                }
                __syncthreads();
                const int j0 = threadIdx.x * SK_PPT;
                if (j0 < nstart) {
                    const int jn = min(SK_PPT, nstart - j0);
                    // window j0+j covers score[j0+j .. j0+j+wsz-1]; consecutive windows that share their
                    // minimizer feed the (idempotent) set sketches once
                    uint64_t prev = 0; bool have_prev = false;
                    if (wsz >= SK_PPT) {
                        uint64_t common = ~0ULL;                       // entries shared by all windows of this thread
                        for (int q = j0 + jn - 1; q <= j0 + wsz - 1; ++q) common = min(common, score[sk_pad(q)]);
                        uint64_t left[SK_PPT];                          // suffix minima of the leading entries
                        uint64_t run = ~0ULL;
                        #pragma unroll
                        for (int j = SK_PPT - 1; j >= 0; --j) {
                            if (j < jn - 1) run = min(run, score[sk_pad(j0 + j)]);
                            left[j] = run;
                        }
                        run = ~0ULL;
                        #pragma unroll
                        for (int j = 0; j < SK_PPT; ++j) {
                            if (j < jn) {
                                if (j) run = min(run, score[sk_pad(j0 + wsz - 1 + j)]);
                                const uint64_t mn = min(min(left[j], common), run);
                                if (Consumer::kEveryWindow || !have_prev || mn != prev) {
                                    const uint64_t km = frev64_inv(mn);
                                    if (km != ~0ULL) cons.consume(wang64(km ^ a.xormask));
                                }
                                prev = mn; have_prev = true;
                            }
                        }
                    } else {
                        for (int j = 0; j < jn; ++j) {
                            uint64_t mn = ~0ULL;
                            for (int q = 0; q < wsz; ++q) mn = min(mn, score[sk_pad(j0 + j + q)]);
                            if (Consumer::kEveryWindow || !have_prev || mn != prev) {
                                const uint64_t km = frev64_inv(mn);
                                if (km != ~0ULL) cons.consume(wang64(km ^ a.xormask));
                            }
                            prev = mn; have_prev = true;
                        }
                    }
                }
            }
            cons.end_tile(cur_ent);
        }
    }
    if (cur_ent != 0xFFFFFFFFu) cons.flush(cur_ent);
}

inline uint32_t sketch_score_slots(int k, int w) { return w > k ? (uint32_t)sk_pad(SK_TILE + (w - k + 1)) + 2 : 0; }
template <class Consumer>
inline size_t sketch_smem_bytes(uint32_t m, uint32_t score_slots) {
    size_t b = (size_t)SK_NWORDS * 8 + (size_t)(SK_NWORDS + (SK_NWORDS & 1)) * 4;
    return b + (size_t)score_slots * 8 + Consumer::smem_bytes(m);
}

} // namespace d2g
