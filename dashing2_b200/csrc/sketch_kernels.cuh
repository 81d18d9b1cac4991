// sketch_kernels.cuh -- K1/K2/K3: sequence tile -> 2-bit pack -> k-mer / minimizer -> hash -> sketch update.
//
// One CTA owns a contiguous span of START POSITIONS of the concatenated sequence buffer and walks the
// records that overlap it (k-mers never cross a record; windows reset per record --
// /root/reference/bonsai/include/bonsai/encoder.h:201-206).  Per tile of SK_TILE start positions:
//   1. the packed sequence (2 bits/base + 1 invalid bit/base, pack_kernels.cuh) of the tile goes to shared memory;
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
#include "pack_kernels.cuh"

namespace d2g {

constexpr int SK_THREADS = 256;
constexpr int SK_PPT = 8;                       // start positions per thread per tile
constexpr int SK_TILE = SK_THREADS * SK_PPT;    // 2048 start positions per tile
constexpr int SK_MAX_W = 1024;                  // largest window (bases) the tile halo supports
constexpr int SK_NWORDS = (SK_TILE + SK_MAX_W + 128 + 31) / 32 + 4;   // a tile starts at a multiple of 128 bases (16-byte aligned in both packed arrays)
// Window keys live in shared memory at index q + q/8: thread t owns keys 8t..8t+7 (+ window tail), so a
// warp's simultaneous accesses are 9 u64 apart instead of 8 -- the minimum two wavefronts instead of 16.
__host__ __device__ constexpr int sk_pad(int q) { return q + (q >> 3); }
constexpr int SK_SCORE_SLOTS = sk_pad(SK_TILE + SK_MAX_W) + 1;

struct SketchArgs {
    PackedSeq seq;               // packed bases of the whole batch buffer (device), padded (pack_kernels.cuh)
    const uint64_t *rec_off;     // [n_rec + 1]
    const uint32_t *rec_entity;  // [n_rec]
    uint64_t n_rec;
    uint64_t pos_base, pos_end;  // this launch covers start positions [pos_base, pos_end); rec_off values are absolute
    uint32_t ent_base;           // subtracted from rec_entity: the consumer's registers start at this entity
    const uint32_t *ent_state;   // optional [entities of this launch]: only records whose entity has ent_state == want_state are processed
    uint32_t want_state;
    uint64_t span;               // start positions per CTA (multiple of SK_TILE)
    int k, w, canon;
    uint64_t xormask;
    uint32_t m;                  // registers per entity
    uint32_t tile_stride;        // process only tiles whose global index % tile_stride == 0 (sampling); 1 = all
    uint32_t score_slots;        // shared-memory window-key slots (windowed mode): sk_pad(SK_TILE + w - k + 1) + 1
    uint32_t keymask;            // fast windowed kernel: bits of the 32-bit window key that take part (all; fewer only to provoke ties in tests)
    const uint64_t *filter;      // --filterset (src/fastxsketch.cpp:385-388, src/filterset.h): sorted hashed k-mers that never reach the sketch
    uint64_t filter_n;           // 0 = no filter set; kernels are instantiated with FILTER = true only when it is set
};

// FilterSet::in_set, the sorted-hash-set flavour (src/filterset.h:207-213)
__device__ __forceinline__ bool sk_filtered(const SketchArgs &a, uint64_t hv) {
    uint64_t lo = 0, hi = a.filter_n;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (__ldg(a.filter + mid) < hv) lo = mid + 1; else hi = mid; }
    return lo < a.filter_n && __ldg(a.filter + lo) == hv;
}

// ---- packed tile -> shared memory ----------------------------------------------------------------------
// copies words [o/32, o/32 + nw) of the packed batch (o a multiple of 32 bases) into the tile.
// W: as uint64 words the first base of each 32 sits in the MSBs; M likewise for uint32 words.
__device__ __forceinline__ void load_tile(const SketchArgs &a, uint64_t *W, uint32_t *M, uint64_t o, int nw) {
    const uint64_t w0 = o >> 5;
    for (int j = threadIdx.x; j < 2 * nw; j += SK_THREADS) {
        if (j < nw) W[j] = __ldg(a.seq.codes + w0 + j);
        else M[j - nw] = __ldg(a.seq.mask + w0 + (j - nw));
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
    static __host__ __device__ size_t smem_bytes(uint32_t m, bool) { return (size_t)m * 8; }
    uint64_t *sreg; Params p;
    __device__ __forceinline__ void init(unsigned char *smem, const Params &pp, bool) {
        p = pp; sreg = reinterpret_cast<uint64_t *>(smem);
        for (uint32_t i = threadIdx.x; i < p.m; i += SK_THREADS) sreg[i] = ~0ULL;
    }
    static constexpr bool kEveryWindow = false;   // set semantics: consecutive equal minimizers feed the sketch once
    static constexpr int kMinBlocks = 4;          // 64 registers/thread: four CTAs per SM when the registers fit in shared memory (S <= 4096 windowed)
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

// ---- eight consecutive k-mers out of the packed tile ------------------------------------------------
// km[j] = canonical (or forward) k-mer starting at local base b + j, j = 0..7, rolled with fixed shifts
// (encoder.h:241-272; kmerutil.h:83-90,137-140).  The reverse complement gains its new base at bit 2k-2;
// for k >= 17 that is the high word and the insertion is one IMAD.  Returns a bitmask of the j whose
// k-mer holds a non-ACGT base (0 in the common case, tested once for all eight).
__device__ __forceinline__ uint32_t kmers8(const uint64_t *W, const uint32_t *M, int b, int k, uint64_t kmask, bool canon, uint64_t km[8]) {
    uint64_t fw = tile_kmer(W, b, k);
    uint64_t rc = revcomp(fw, k);
    const uint32_t nx = (uint32_t)tile_kmer(W, b + k, 7);          // the next seven base codes, first in bits 13:12
    const int sh = 2 * k - 2;
    const uint32_t mlo = (uint32_t)kmask, mhi = (uint32_t)(kmask >> 32);
    uint32_t flo = (uint32_t)fw, fhi = (uint32_t)(fw >> 32), rlo = (uint32_t)rc, rhi = (uint32_t)(rc >> 32);
    km[0] = canon ? (fw < rc ? fw : rc) : fw;
    if (sh >= 32) {
        const uint32_t ph = 1u << (sh - 32);
        #pragma unroll
        for (int j = 1; j < 8; ++j) {
            const uint32_t c = (nx >> (14 - 2 * j)) & 3u;
            fhi = __funnelshift_l(flo, fhi, 2) & mhi; flo = ((flo << 2) | c) & mlo;
            rlo = __funnelshift_r(rlo, rhi, 2); rhi = (c ^ 3u) * ph + (rhi >> 2);
            const uint64_t f = ((uint64_t)fhi << 32) | flo, r = ((uint64_t)rhi << 32) | rlo;
            km[j] = canon ? (f < r ? f : r) : f;
        }
    } else {
        #pragma unroll
        for (int j = 1; j < 8; ++j) {
            const uint32_t c = (nx >> (14 - 2 * j)) & 3u;
            fhi = __funnelshift_l(flo, fhi, 2) & mhi; flo = ((flo << 2) | c) & mlo;
            rlo = __funnelshift_r(rlo, rhi, 2) | ((c ^ 3u) << sh); rhi = rhi >> 2;
            const uint64_t f = ((uint64_t)fhi << 32) | flo, r = ((uint64_t)rhi << 32) | rlo;
            km[j] = canon ? (f < r ? f : r) : f;
        }
    }
    // invalid bases among b .. b+k+6 (at most 39): none in the common case
    const int n = k + 7;
    uint32_t any = tile_invalid(M, b, n < 32 ? n : 32);
    if (n > 32) any |= tile_invalid(M, b + 32, n - 32);
    if (any == 0) return 0;
    uint32_t bad = 0;
    #pragma unroll
    for (int j = 0; j < 8; ++j) bad |= (tile_invalid(M, b + j, k) ? 1u : 0u) << j;
    return bad;
}

// ---- the kernel ---------------------------------------------------------------------------------
constexpr int SK_SCAP = 512;   // staged window minima per tile (expected ~2/(wsz+1) of the tile); the rest take the direct path

// ---- one span of start positions [span_lo, span_hi) -----------------------------------------------------
// All threads of the CTA; the consumer is initialised by the caller and is flushed at the end of the span.
template <bool WINDOWED, class Consumer, bool FILTER = false, int K_T = 0>
__device__ __forceinline__ void sketch_span(const SketchArgs &a, Consumer &cons, uint64_t *W, uint32_t *M, uint64_t *score, uint64_t *stage, int *scount,
                                            const uint64_t span_lo, const uint64_t span_hi) {
    const int k = K_T ? K_T : a.k;                   // K_T: k as a compile-time constant (masks and shift counts fold)
    const int need = WINDOWED ? a.w : k;           // bases a start position needs to its right
    const int wsz = WINDOWED ? (a.w - a.k + 1) : 1;  // k-mers per window
    // first record whose end lies beyond span_lo (records are sorted by offset)
    uint64_t lo = 0, hi = a.n_rec;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (a.rec_off[mid + 1] > span_lo) hi = mid; else lo = mid + 1; }
    uint32_t cur_ent = 0xFFFFFFFFu;
    const uint64_t kmask = k < 32 ? ((1ULL << (2 * k)) - 1) : ~0ULL;
    const bool canon = a.canon != 0;
    const int lane = threadIdx.x & 31;
    auto feed = [&](uint64_t hv) { if (FILTER && sk_filtered(a, hv)) return; cons.consume(hv); };
    __syncthreads();

    // Windowed mode stages the minimizers of a tile (as window keys) and hashes them densely at the start of the
    // next tile: frev64_inv + maskfn + the consumer run on full warps instead of on the ~10 % of lanes whose
    // window changed its minimizer.  All threads; the caller guarantees a barrier since the last staging.
    auto drain_stage = [&]() {
        if (!WINDOWED) return;
        const int n = min(*scount, SK_SCAP);
        for (int i = threadIdx.x; i < n; i += SK_THREADS) {
            const uint64_t km = frev64_inv(stage[i]);
            if (km != ~0ULL) feed(wang64(km ^ a.xormask));
        }
    };

    for (uint64_t r = lo; r < a.n_rec; ++r) {
        const uint64_t rs = a.rec_off[r], re = a.rec_off[r + 1];
        if (rs >= span_hi) break;
        if (re - rs < (uint64_t)need) continue;
        const uint64_t p0 = max(span_lo, rs);
        const uint64_t p1 = min(span_hi, re - need + 1);   // owned, usable start positions [p0, p1)
        if (p0 >= p1) continue;
        const uint32_t ent = a.rec_entity[r] - a.ent_base;
        if (a.ent_state && a.ent_state[ent] != a.want_state) continue;
        if (ent != cur_ent) {
            if (cur_ent != 0xFFFFFFFFu) {
                __syncthreads();
                drain_stage();
                __syncthreads();
                if (WINDOWED && threadIdx.x == 0) *scount = 0;
                cons.flush(cur_ent);
            }
            cur_ent = ent;
            cons.begin_entity(ent, p0 - span_lo);   // offset inside this CTA's span
        }
        // Windowed mode: window keys live at slot OFF + (key - E0), E0 = first NEW key of the tile; the wsz-1 keys
        // before it are carried over from the previous tile (copied to slots [OFF-(wsz-1), OFF)), so only the first
        // tile of a record segment computes them.  A window is identified by its LAST key: thread t owns the
        // windows ending at its own eight keys.
        const int OFF = (wsz - 1 + 7) & ~7;
        bool first = true;
        for (uint64_t t0 = p0; t0 < p1; t0 += SK_TILE) {
            if (a.tile_stride > 1 && ((t0 / SK_TILE) % a.tile_stride) != 0) { first = true; continue; }
            const int nstart = (int)min((uint64_t)SK_TILE, p1 - t0);  // start positions (= windows = new keys) in this tile
            const int npre = (WINDOWED && first) ? wsz - 1 : 0;        // keys before E0 this tile has to compute itself
            const uint64_t kb = WINDOWED ? (first ? t0 : t0 + (uint64_t)(wsz - 1)) : t0;   // first key / k-mer whose bases are needed
            const uint64_t o = kb & ~127ULL;
            const int off = (int)(kb - o);
            const int nbases = off + npre + nstart + k - 1;
            __syncthreads();                                          // previous tile fully consumed, its minimizers staged
            drain_stage();
            cons.end_tile(cur_ent);
            if (WINDOWED && !first)                                   // carry the last wsz-1 keys of the previous (full) tile
                for (int i = threadIdx.x; i < wsz - 1; i += SK_THREADS)
                    score[sk_pad(OFF - (wsz - 1) + i)] = score[sk_pad(OFF + SK_TILE - (wsz - 1) + i)];
            load_tile(a, W, M, o, ((nbases + 31) >> 5) + 2);   // + two words: tile_kmer / tile_invalid read past the last base
            __syncthreads();
            if (WINDOWED && threadIdx.x == 0) *scount = 0;
            if (!WINDOWED) {
                const int j0 = threadIdx.x * SK_PPT;
                if (j0 < nstart) {
                    uint64_t km[8];
                    const uint32_t bad = kmers8(W, M, off + j0, k, kmask, canon, km);
                    const int jn = min(SK_PPT, nstart - j0);
                    #pragma unroll
                    for (int j = 0; j < SK_PPT; ++j)
                        if (j < jn && !((bad >> j) & 1u))             // encoder.h:254 -- k-mers holding a non-ACGT base are skipped
                            feed(wang64(km[j] ^ a.xormask));  // maskfn, src/enums.h:136-140
                }
            } else {
                // canonical windowed path: a k-mer holding an invalid base enters the window as k-mer 0
                // (encoder.h:568-571 + kmerutil.h:137-140)
                if (first)
                    for (int i = threadIdx.x * SK_PPT; i < npre; i += SK_TILE) {
                        uint64_t km[8];
                        const uint32_t bad = kmers8(W, M, off + i, k, kmask, true, km);
                        #pragma unroll
                        for (int j = 0; j < SK_PPT; ++j)
                            if (i + j < npre) score[sk_pad(OFF - npre + i + j)] = frev64(((bad >> j) & 1u) ? 0ULL : km[j]);
                    }
                const int j0 = threadIdx.x * SK_PPT;
                const int jn = max(0, min(SK_PPT, nstart - j0));
                if (jn > 0) {
                    uint64_t km[8];
                    const uint32_t bad = kmers8(W, M, off + npre + j0, k, kmask, true, km);
                    uint64_t *dst = score + sk_pad(OFF + j0);         // OFF + j0 is a multiple of 8: slots dst[0..7]
                    if (jn == SK_PPT && bad == 0) {                     // the common case: a full chunk of valid k-mers
                        #pragma unroll
                        for (int j = 0; j < SK_PPT; ++j) dst[j] = frev64(km[j]);
                    } else {
                        #pragma unroll
                        for (int j = 0; j < SK_PPT; ++j)
                            if (j < jn) dst[j] = frev64(((bad >> j) & 1u) ? 0ULL : km[j]);
                    }
                }
                first = false;
                __syncthreads();
                // window j (ending at own key j) covers slots [OFF+j0+j-(wsz-1), OFF+j0+j]; consecutive windows that share
                // their minimizer feed the (idempotent) set sketches once -- also across the threads of a warp
                uint64_t mn[SK_PPT];
                if (jn > 0) {
                    const int B0 = OFF + j0;
                    const uint64_t *own = score + sk_pad(B0);
                    if (wsz >= SK_PPT) {
                        // slots [B0 + 7 - (wsz-1), B0) are shared by all eight windows of this thread.  B0 is a multiple of 8, so
                        // walking backwards they are whole padded groups of eight (consecutive words) and one partial group.
                        uint64_t common = ~0ULL;
                        int rem = wsz - 8;
                        const uint64_t *g = own - 9;                   // group of slots B0-8 .. B0-1
                        for (; rem >= 8; rem -= 8, g -= 9) {
                            #pragma unroll
                            for (int i = 0; i < 8; ++i) common = min(common, g[i]);
                        }
                        #pragma unroll
                        for (int i = 1; i < 8; ++i) if (i >= 8 - rem) common = min(common, g[i]);
                        uint64_t left[SK_PPT];                          // suffix minima of the slots only the earlier windows reach
                        uint64_t run = ~0ULL;
                        left[SK_PPT - 1] = run;
                        #pragma unroll
                        for (int j = SK_PPT - 2; j >= 0; --j) {
                            run = min(run, score[sk_pad(B0 + j - (wsz - 1))]);
                            left[j] = run;
                        }
                        run = ~0ULL;                                    // prefix minima of the own keys
                        if (jn == SK_PPT) {
                            #pragma unroll
                            for (int j = 0; j < SK_PPT; ++j) { run = min(run, own[j]); mn[j] = min(min(left[j], common), run); }
                        } else {
                            #pragma unroll
                            for (int j = 0; j < SK_PPT; ++j) {
                                if (j < jn) { run = min(run, own[j]); mn[j] = min(min(left[j], common), run); }
                                else mn[j] = mn[j - 1];                 // j >= jn >= 1: repeat the last window (never emits)
                            }
                        }
                    } else {
                        #pragma unroll
                        for (int j = 0; j < SK_PPT; ++j) {
                            uint64_t v = ~0ULL;
                            if (j < jn) for (int q = 0; q < wsz; ++q) v = min(v, score[sk_pad(B0 + j - q)]);
                            mn[j] = j < jn ? v : mn[j > 0 ? j - 1 : 0];
                        }
                    }
                }
                if (jn == 0) {
                    #pragma unroll
                    for (int j = 0; j < SK_PPT; ++j) mn[j] = 0;
                }
                // mn[j >= jn] repeats mn[jn-1], so mn[7] is the last window of the thread and needs no guard below
                uint64_t prev = __shfl_up_sync(0xffffffffu, mn[SK_PPT - 1], 1);
                if (lane == 0) prev = ~mn[0];                           // lane 0 (and so the first thread of a tile) always emits
                if (Consumer::kEveryWindow) {
                    #pragma unroll
                    for (int j = 0; j < SK_PPT; ++j)
                        if (j < jn) { const uint64_t km = frev64_inv(mn[j]); if (km != ~0ULL) feed(wang64(km ^ a.xormask)); }
                } else {
                    uint32_t emask = 0;
                    #pragma unroll
                    for (int j = 0; j < SK_PPT; ++j) { emask |= (uint32_t)(mn[j] != prev) << j; prev = mn[j]; }
                    if (jn == 0) emask = 0;
                    // one slot range per warp: exclusive scan of the per-thread counts, one atomic
                    const int cnt = __popc(emask);
                    int incl = cnt;
                    #pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    if (total) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(scount, total);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        int slot = base + incl - cnt;
                        if (base + total <= SK_SCAP) {                 // warp-uniform: plain predicated stores
                            #pragma unroll
                            for (int j = 0; j < SK_PPT; ++j) {
                                if ((emask >> j) & 1u) stage[slot] = mn[j];
                                slot += (emask >> j) & 1u;
                            }
                        } else {
                            #pragma unroll
                            for (int j = 0; j < SK_PPT; ++j)
                                if ((emask >> j) & 1u) {
                                    if (slot < SK_SCAP) stage[slot] = mn[j];
                                    else { const uint64_t km = frev64_inv(mn[j]); if (km != ~0ULL) feed(wang64(km ^ a.xormask)); }
                                    ++slot;
                                }
                        }
                    }
                }
            }
        }
    }
    if (cur_ent != 0xFFFFFFFFu) {
        __syncthreads();
        drain_stage();
        cons.flush(cur_ent);
    }
}

// ---- the kernels ---------------------------------------------------------------------------------
struct SketchSmem { uint64_t *W; uint32_t *M; uint64_t *score, *stage; int *scount; unsigned char *csmem; };
template <bool WINDOWED>
__device__ __forceinline__ SketchSmem sketch_smem_carve(unsigned char *smem_raw, const SketchArgs &a) {
    SketchSmem s;
    s.W = reinterpret_cast<uint64_t *>(smem_raw);
    s.M = reinterpret_cast<uint32_t *>(s.W + SK_NWORDS);
    s.score = reinterpret_cast<uint64_t *>(s.M + SK_NWORDS + (SK_NWORDS & 1));
    s.stage = s.score + (WINDOWED ? a.score_slots : 0);
    s.scount = reinterpret_cast<int *>(s.stage + (WINDOWED ? SK_SCAP : 0));
    s.csmem = reinterpret_cast<unsigned char *>(s.scount + (WINDOWED ? 2 : 0));
    return s;
}

template <bool WINDOWED, class Consumer, bool FILTER = false, int K_T = 0>
__global__ void __launch_bounds__(SK_THREADS, Consumer::kMinBlocks)
sketch_kernel(const SketchArgs a, const typename Consumer::Params cp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SketchSmem s = sketch_smem_carve<WINDOWED>(smem_raw, a);
    Consumer cons;
    cons.init(s.csmem, cp, WINDOWED);
    if (WINDOWED && threadIdx.x == 0) *s.scount = 0;
    const uint64_t span_lo = a.pos_base + (uint64_t)blockIdx.x * a.span;
    const uint64_t span_hi = min(span_lo + a.span, a.pos_end);
    if (span_lo >= span_hi) return;
    sketch_span<WINDOWED, Consumer, FILTER, K_T>(a, cons, s.W, s.M, s.score, s.stage, s.scount, span_lo, span_hi);
}

// The exact kernel over a LIST of tiles: tile_list[i] is the first start position of a tile of SK_TILE positions that the fast
// windowed kernel (sketch_fast.cuh) could not finish exactly (a 32-bit key tie between different k-mers, an event list overflow).
// Set sketches are idempotent minima, so recomputing all windows of those tiles on top of what the fast kernel delivered is exact.
// When the list overflowed (count > cap) every tile of [pos_base, pos_end) is recomputed.
template <class Consumer>
__global__ void __launch_bounds__(SK_THREADS, Consumer::kMinBlocks)
sketch_redo_kernel(const SketchArgs a, const typename Consumer::Params cp, const uint64_t *tile_list, const unsigned long long *tile_count, uint64_t cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned long long cnt = *tile_count;
    if (cnt == 0) return;
    const SketchSmem s = sketch_smem_carve<true>(smem_raw, a);
    Consumer cons;
    cons.init(s.csmem, cp, true);
    const bool all = cnt > cap;
    const uint64_t n = all ? (a.pos_end - a.pos_base + SK_TILE - 1) / SK_TILE : (uint64_t)cnt;
    for (uint64_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint64_t lo = all ? a.pos_base + i * SK_TILE : tile_list[i];
        const uint64_t hi = min(lo + (uint64_t)SK_TILE, a.pos_end);
        __syncthreads();
        if (threadIdx.x == 0) *s.scount = 0;
        __syncthreads();
        if (lo < hi) sketch_span<true>(a, cons, s.W, s.M, s.score, s.stage, s.scount, lo, hi);
    }
}

inline uint32_t sketch_score_slots(int k, int w) { return w > k ? (uint32_t)sk_pad(SK_TILE + (w - k + 1) + 8) + 2 : 0; }
template <class Consumer>
inline size_t sketch_smem_bytes(uint32_t m, uint32_t score_slots) {
    size_t b = (size_t)SK_NWORDS * 8 + (size_t)(SK_NWORDS + (SK_NWORDS & 1)) * 4;
    if (score_slots) b += (size_t)SK_SCAP * 8 + 8;      // windowed: staged minimizers + counter
    return b + (size_t)score_slots * 8 + Consumer::smem_bytes(m, score_slots != 0);
}

} // namespace d2g
