// sketch_fast.cuh -- the windowed-minimizer sketch kernel with 32-bit window keys (set sketches only).
//
// Same contract as sketch_kernel<true, Consumer> (sketch_kernels.cuh; reference: for_each_canon_windowed,
// /root/reference/bonsai/include/bonsai/encoder.h:212-217,622-628 and QueueMap::next_value, qmap.h:79-87): every window of
// w-k+1 consecutive canonical k-mers contributes its minimizer under the order (FRev64(k-mer), k-mer) to the consumer, a k-mer
// with an invalid base entering as k-mer 0.  FRev64 is a bijection, so the order is the order of the 64-bit score alone.
//
// What is different: the sliding minimum runs on the HIGH 32 BITS of the score (one VIMNMX / VIMNMX3 per comparison instead of
// a four-instruction 64-bit minimum, 16-byte shared-memory loads of four keys), and only windows whose minimizer CHANGES do
// any further work ("events", ~2/(w-k+2) of the windows):
//   type A  the newest key of the window is strictly below the minimum of the previous window: the newest k-mer is the
//           minimizer, no search, no tie possible;
//   type B  the previous minimizer left the window (the minimum rose), or the newest key EQUALS the previous minimum: the
//           window is scanned for the positions that hold its minimum.  One position: that k-mer is the minimizer.  Several
//           positions with the same k-mer (repeats): the same.  Several positions with DIFFERENT k-mers (a 32-bit tie,
//           probability ~ (w-k+1)^2 / 2^33 per window): the exact minimizer is chosen by the full score, and the tile (and its
//           successor) is put on the redo list, because such a tie can later change the minimizer without changing the
//           32-bit minimum.
// The event flags of a warp's 256 windows are compacted into the two event lists by the warp itself (type A in its lower half,
// type B in its upper half: same instructions, one shared-memory atomic per half-warp and tile); the events are hashed on dense
// warps during the next tile.  Tiles on the redo list are recomputed by the
// exact 64-bit kernel (sketch_redo_kernel): set sketches are idempotent minima, so the union of both passes is exact.  An
// event list that overflows (pathological repeats) also sends its tile to the redo list.  Everything the fast pass emits is
// a true minimizer; everything it might have missed lies in a listed tile.
//
// Sequence tiles arrive in shared memory through cp.async.bulk (one elected thread, mbarrier completion) one tile ahead of
// their use; the packed words of the previous tile stay resident for the k-mer lookups of its staged events.
#pragma once
#include "async_copy.cuh"
#include "sketch_kernels.cuh"

namespace d2g {

constexpr int SF_TILE = SK_TILE;                 // windows per tile (256 threads x 8)
constexpr int SF_OFF = 64;                       // key slots in front of the tile's first new key: the last keys of the previous tile
constexpr int SF_MAX_WSZ = 63;                   // k-mers per window the fast kernel takes (larger windows: exact kernel)
constexpr int SF_LCAP = 256;                     // capacity of each event list per tile
constexpr int SF_NW = 76;                        // packed words per tile buffer: (127 + 2048 + 62 + 31 + 31) / 32 + 2, rounded up to 4
constexpr int SF_WBYTES = SF_NW * 12;            // codes (u64) + mask (u32) of one tile buffer, a multiple of 16
__host__ __device__ constexpr int sf_pad(int i) { return i + ((i >> 5) << 2); }   // 16-byte loads of 4 keys by threads 8 keys apart: conflict free
constexpr int SF_KEYS = sf_pad(SF_OFF + SF_TILE) + 8;

struct FastAux {
    unsigned long long *redo_count;   // tiles appended so far (may exceed redo_cap: then every tile is redone)
    uint64_t *redo_list;              // first start position of each tile to recompute
    uint64_t redo_cap;
};

inline size_t sketch_fast_smem_bytes(size_t consumer_bytes) {
    return (size_t)3 * SF_WBYTES + 32 + (size_t)2 * SF_KEYS * 4 + (size_t)SF_LCAP * 2 * 2 + 32 + consumer_bytes;
}

// canonical k-mer at base offset b of a packed tile; 0 when one of its bases is invalid (encoder.h:568-571 + kmerutil.h:137-140)
__device__ __forceinline__ uint64_t tile_canonical_or_zero(const uint64_t *W, const uint32_t *M, int b, int k) {
    if (tile_invalid(M, b, k)) return 0;
    return canonical(tile_kmer(W, b, k), k);
}

// WSZ_T / K_T: window size (k-mers) and k as compile-time constants (0 = run time); the BASELINE shape k = 31, w = 51 gets both.
template <int WSZ_T, class Consumer, int K_T = 0>
__global__ void __launch_bounds__(SK_THREADS, Consumer::kMinBlocks)
sketch_fast_kernel(const SketchArgs a, const typename Consumer::Params cp, const FastAux fx) {
    static_assert(!Consumer::kEveryWindow, "the fast windowed kernel serves set sketches only");
    static_assert(SK_THREADS == 256 && SK_PPT == 8, "event compaction assumes 256 threads x 8 windows");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int k = K_T ? K_T : a.k;
    const int wsz = WSZ_T ? WSZ_T : (a.w - a.k + 1);
    const int need = a.w;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + 3 * SF_WBYTES);
    uint32_t *KH = reinterpret_cast<uint32_t *>(smem_raw + 3 * SF_WBYTES + 32);          // two key buffers of SF_KEYS words
    uint16_t *LA = reinterpret_cast<uint16_t *>(KH + 2 * SF_KEYS);                        // type A events: key slot
    uint16_t *LB = LA + SF_LCAP;                                                          // type B events: key slot of the window's last key
    uint32_t *ctl = reinterpret_cast<uint32_t *>(LB + SF_LCAP);                           // [par] nA, [2 + par] nB, [4 + par] redo flags; par = tile parity
    unsigned char *csmem = reinterpret_cast<unsigned char *>(ctl + 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Consumer cons;
    cons.init(csmem, cp, true);
    if (tid == 0) {
        mbar_init(bar + 0, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1);
        mbar_fence_init();
        for (int i = 0; i < 8; ++i) ctl[i] = 0;
    }
    const uint64_t span_lo = a.pos_base + (uint64_t)blockIdx.x * a.span;
    const uint64_t span_hi = min(span_lo + a.span, a.pos_end);
    if (span_lo >= span_hi) return;
    uint64_t lo = 0, hi = a.n_rec;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (a.rec_off[mid + 1] > span_lo) hi = mid; else lo = mid + 1; }
    uint32_t cur_ent = 0xFFFFFFFFu;
    const uint64_t kmask = k < 32 ? ((1ULL << (2 * k)) - 1) : ~0ULL;
    __syncthreads();

    // packed tile buffers rotate 0 -> 1 -> 2 -> 0; bit b of wphase is the mbarrier parity the next wait on buffer b expects
    uint32_t wb = 0, wphase = 0;
    auto next_buf = [](uint32_t b) { return b == 2 ? 0u : b + 1; };
    // thread 0: bulk copies of the packed words of the tile starting at start position t0 into buffer b
    auto issue_tile = [&](uint32_t b, uint64_t t0, int nstart) {
        const uint64_t o = t0 & ~127ULL;
        const int nbases = (int)(t0 - o) + nstart + wsz - 1 + k - 1;
        const uint32_t nw = (uint32_t)((((nbases + 31) >> 5) + 2 + 3) & ~3);
        unsigned char *dst = smem_raw + b * SF_WBYTES;
        mbar_expect_tx(bar + b, nw * 12);
        bulk_g2s(dst, a.seq.codes + (o >> 5), nw * 8, bar + b);
        bulk_g2s(dst + SF_NW * 8, a.seq.mask + (o >> 5), nw * 4, bar + b);
    };

    // ---- event flags of a warp's 256 windows -> the two event lists ----
    // flags = type A flags | type B flags << 8 of this thread's eight windows.  Lanes 0-15 build the A list, lanes 16-31 the B list: lane
    // l (mod 16) takes the flags of threads 2l and 2l+1 of the warp = 16 consecutive windows.  All 32 lanes of the warp call this.
    auto compact_events = [&](uint32_t flags, uint32_t par) {
        const int h = lane & 15;
        const uint32_t f0 = __shfl_sync(0xffffffffu, flags, 2 * h), f1 = __shfl_sync(0xffffffffu, flags, 2 * h + 1);
        const bool isB = lane >= 16;
        uint32_t m = isB ? ((f0 >> 8) | (f1 & 0xFF00u)) : ((f0 & 0xFFu) | ((f1 & 0xFFu) << 8));   // bit b: window 16h + b of the warp
        const int cnt = __popc(m);
        int incl = cnt;
        #pragma unroll
        for (int o = 1; o < 16; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o, 16); if (h >= o) incl += y; }
        const int total = __shfl_sync(0xffffffffu, incl, 15, 16);
        int base = 0;
        if (h == 0 && total) base = (int)atomicAdd(ctl + (isB ? 2 : 0) + par, (uint32_t)total);
        base = __shfl_sync(0xffffffffu, base, 0, 16);
        int at = base + incl - cnt;
        uint16_t *L = isB ? LB : LA;
        const int slot0 = SF_OFF + 256 * warp + 16 * h;
        for (; m; m &= m - 1, ++at) if (at < SF_LCAP) L[at] = (uint16_t)(slot0 + __ffs(m) - 1);
        if (h == 0 && base + total > SF_LCAP) atomicOr(ctl + 4 + par, 1u);          // a list overflowed: the exact kernel redoes this tile
    };

    // ---- staged events of one tile -> k-mers -> consumer (all threads; dense warps) ----
    // W / M / keys: the tile's packed words and key buffer; boff: base offset of the k-mer whose key sits in slot SF_OFF
    auto drain_events = [&](const uint64_t *W, const uint32_t *M, const uint32_t *keys, int boff, uint32_t par) {
        const int nA = min((int)ctl[par], SF_LCAP), nB = min((int)ctl[2 + par], SF_LCAP);
        for (int i = tid; i < nA; i += SK_THREADS) {
            const int e = LA[i];
            const uint64_t km = tile_canonical_or_zero(W, M, boff + (e - SF_OFF), k);
            cons.consume(wang64(km ^ a.xormask));                                   // maskfn, src/enums.h:136-140
        }
        for (int i = (tid - 96) & (SK_THREADS - 1); i < nB; i += SK_THREADS) {        // warps 3, 4, 5 first: warps 0-2 hash the A events
            const int e = LB[i];
            // the window's keys sit in slots e-(wsz-1) .. e; in the padded layout they are contiguous except for one possible 4-word pad
            const int x0 = e - (wsz - 1);
            const uint32_t *pa = keys + sf_pad(x0);
            const int thr = 32 - (x0 & 31);                                        // window elements d >= thr lie behind a pad (and d >= thr + 32 behind two)
            auto key_at = [&](int d) {
                if (WSZ_T && WSZ_T <= 33) return (d < thr ? pa : pa + 4)[d];          // at most one pad inside a window of <= 33 keys
                return pa[d + (d >= thr ? 4 : 0) + (d >= thr + 32 ? 4 : 0)];
            };
            uint32_t m = 0xFFFFFFFFu;
            int q = 0, cnt = 0;
            if (WSZ_T) {
                uint32_t v[WSZ_T ? WSZ_T : 1];
                #pragma unroll
                for (int d = 0; d < WSZ_T; ++d) { v[d] = key_at(d); m = min(m, v[d]); }
                #pragma unroll
                for (int d = 0; d < WSZ_T; ++d) if (v[d] == m) { q = d; ++cnt; }
            } else {
                for (int d = 0; d < wsz; ++d) m = min(m, key_at(d));
                for (int d = 0; d < wsz; ++d) if (key_at(d) == m) { q = d; ++cnt; }
            }
            uint64_t km = tile_canonical_or_zero(W, M, boff + (x0 + q - SF_OFF), k);
            if (cnt > 1) {                                                           // several positions hold the 32-bit minimum
                uint64_t best = frev64(km); bool differ = false;
                for (int d = 0; d < wsz; ++d)
                    if (d != q && key_at(d) == m) {
                        const uint64_t k2 = tile_canonical_or_zero(W, M, boff + (x0 + d - SF_OFF), k);
                        const uint64_t s2 = frev64(k2);
                        if (s2 != best) differ = true;
                        if (s2 < best) { best = s2; km = k2; }
                    }
                if (differ) atomicOr(ctl + 4 + par, 3u);                             // redo this tile and the next one
            }
            cons.consume(wang64(km ^ a.xormask));
        }
    };
    // thread 0, after the barrier that ends the drain of a tile: put it on the redo list when flagged
    auto settle_redo = [&](uint32_t par, uint64_t t0) {
        const uint32_t f = ctl[4 + par];
        if (f) {
            ctl[4 + par] = 0;
            const unsigned long long at = atomicAdd(fx.redo_count, (f & 2u) ? 2ULL : 1ULL);
            if (at < fx.redo_cap) fx.redo_list[at] = t0;
            if ((f & 2u) && at + 1 < fx.redo_cap) fx.redo_list[at + 1] = t0 + SF_TILE;
        }
    };

    for (uint64_t r = lo; r < a.n_rec; ++r) {
        const uint64_t rs = a.rec_off[r], re = a.rec_off[r + 1];
        if (rs >= span_hi) break;
        if (re - rs < (uint64_t)need) continue;
        const uint64_t p0 = max(span_lo, rs);
        const uint64_t p1 = min(span_hi, re - need + 1);   // owned, usable window start positions [p0, p1)
        if (p0 >= p1) continue;
        const uint32_t ent = a.rec_entity[r] - a.ent_base;
        if (a.ent_state && a.ent_state[ent] != a.want_state) continue;
        if (ent != cur_ent) {
            if (cur_ent != 0xFFFFFFFFu) cons.flush(cur_ent);          // nothing is pending here: every segment drains its last tile
            cur_ent = ent;
            cons.begin_entity(ent, p0 - span_lo);
        }
        // Pipeline over the tiles of the segment.  Tile n: phase 1 hashes the events of tile n-1 and computes the keys of tile n;
        // phase 2 computes the sliding minima and the event lists of tile n.
        bool first = true, have_prev = false;
        uint32_t par = 0;                                             // parity of the current tile within the segment
        uint32_t prev_wb = 0; int prev_boff = 0; uint64_t prev_t0 = 0;
        if (tid == 0) issue_tile(wb, p0, (int)min((uint64_t)SF_TILE, p1 - p0));
        for (uint64_t t0 = p0; t0 < p1; t0 += SF_TILE, par ^= 1u) {
            const int nstart = (int)min((uint64_t)SF_TILE, p1 - t0);
            const int boff = (int)(t0 & 127ULL) + wsz - 1;           // base offset of the first NEW key's k-mer (window of start position t0 ends there)
            uint32_t *keys = KH + par * SF_KEYS;
            const uint64_t *W = reinterpret_cast<const uint64_t *>(smem_raw + wb * SF_WBYTES);
            const uint32_t *M = reinterpret_cast<const uint32_t *>(smem_raw + wb * SF_WBYTES + SF_NW * 8);
            const int j0 = tid * SK_PPT;
            const int jn = max(0, min(SK_PPT, nstart - j0));
            // ---- phase 1 ----
            if (tid == 0) {
                if (t0 + SF_TILE < p1) issue_tile(next_buf(wb), t0 + SF_TILE, (int)min((uint64_t)SF_TILE, p1 - t0 - SF_TILE));
                ctl[par] = 0; ctl[2 + par] = 0;                       // this tile's event counters (last read two barriers ago)
            }
            if (have_prev)
                drain_events(reinterpret_cast<const uint64_t *>(smem_raw + prev_wb * SF_WBYTES),
                             reinterpret_cast<const uint32_t *>(smem_raw + prev_wb * SF_WBYTES + SF_NW * 8), KH + (par ^ 1u) * SF_KEYS, prev_boff, par ^ 1u);
            mbar_wait(bar + wb, (wphase >> wb) & 1u);
            wphase ^= 1u << wb;
            if (jn > 0) {
                uint64_t km[8];
                const uint32_t bad = kmers8(W, M, boff + j0, k, kmask, true, km);
                uint32_t kh[8];
                #pragma unroll
                for (int j = 0; j < SK_PPT; ++j) kh[j] = (uint32_t)(frev64(((bad >> j) & 1u) ? 0ULL : km[j]) >> 32) & a.keymask;
                uint32_t *dst = keys + sf_pad(SF_OFF + j0);                   // SF_OFF + j0 is a multiple of 8: two aligned groups of four
                if (jn == SK_PPT) {
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(kh[0], kh[1], kh[2], kh[3]);
                    *reinterpret_cast<uint4 *>(dst + 4) = make_uint4(kh[4], kh[5], kh[6], kh[7]);
                } else {
                    #pragma unroll
                    for (int j = 0; j < SK_PPT; ++j) if (j < jn) dst[j] = kh[j];
                }
            }
            if (first) {
                // the wsz-1 keys before the first window's last key; the slot before them is never a window member: largest key
                for (int i = tid; i < wsz - 1; i += SK_THREADS)
                    keys[sf_pad(SF_OFF - (wsz - 1) + i)] = (uint32_t)(frev64(tile_canonical_or_zero(W, M, boff - (wsz - 1) + i, k)) >> 32) & a.keymask;
                if (tid == 0) keys[sf_pad(SF_OFF - wsz)] = 0xFFFFFFFFu;
            } else {
                // carry the last wsz keys of the previous (full) tile
                const uint32_t *pk = KH + (par ^ 1u) * SF_KEYS;
                for (int i = tid; i < wsz; i += SK_THREADS) keys[sf_pad(SF_OFF - wsz + i)] = pk[sf_pad(SF_OFF + SF_TILE - wsz + i)];
            }
            __syncthreads();
            cons.end_tile(cur_ent);
            // ---- phase 2 ----
            if (tid == 0 && have_prev) settle_redo(par ^ 1u, prev_t0);
            uint32_t flags = 0;
            if (jn > 0) {
                const int B0 = SF_OFF + j0;
                uint32_t O[8], L[7], C, extra;
                uint32_t mn[8], prev0;
                if (WSZ_T == 21) {
                    // slots B0-20 .. B0+7 as seven aligned groups of four; in the padded layout the groups from index gc on sit 4 words further
                    const int u = B0 - 20;
                    const uint32_t *pa = keys + sf_pad(u);
                    const int gc = (32 - (u & 31)) >> 2;                                 // u & 31 is 4, 12, 20 or 28
                    auto grp = [&](int g) { return *reinterpret_cast<const uint4 *>((g < gc ? pa : pa + 4) + 4 * g); };
                    const uint4 v0 = grp(0), v1 = grp(1), v2 = grp(2), v3 = grp(3), v4 = grp(4), o0 = grp(5), o1 = grp(6);
                    extra = pa[-1];
                    O[0] = o0.x; O[1] = o0.y; O[2] = o0.z; O[3] = o0.w; O[4] = o1.x; O[5] = o1.y; O[6] = o1.z; O[7] = o1.w;
                    L[0] = v0.x; L[1] = v0.y; L[2] = v0.z; L[3] = v0.w; L[4] = v1.x; L[5] = v1.y; L[6] = v1.z;   // L0..L6 = B0-20 .. B0-14
                    C = min(min(v1.w, v2.x), v2.y);                                                             // common = B0-13 .. B0-1
                    C = min(min(C, v2.z), v2.w); C = min(min(C, v3.x), v3.y); C = min(min(C, v3.z), v3.w);
                    C = min(min(C, v4.x), v4.y); C = min(min(C, v4.z), v4.w);
                } else {
                    const uint4 o0 = *reinterpret_cast<const uint4 *>(keys + sf_pad(B0)), o1 = *reinterpret_cast<const uint4 *>(keys + sf_pad(B0 + 4));
                    O[0] = o0.x; O[1] = o0.y; O[2] = o0.z; O[3] = o0.w; O[4] = o1.x; O[5] = o1.y; O[6] = o1.z; O[7] = o1.w;
                    if (wsz >= SK_PPT) {
                        C = 0xFFFFFFFFu;
                        for (int i = 1; i <= wsz - SK_PPT; ++i) C = min(C, keys[sf_pad(B0 - i)]);      // slots B0-(wsz-8) .. B0-1 belong to all eight windows
                        #pragma unroll
                        for (int j = 0; j < 7; ++j) L[j] = keys[sf_pad(B0 + j - (wsz - 1))];
                        extra = keys[sf_pad(B0 - wsz)];
                    }
                }
                if (jn < SK_PPT) {
                    #pragma unroll
                    for (int j = 0; j < SK_PPT; ++j) if (j >= jn) O[j] = 0xFFFFFFFFu;
                }
                if (WSZ_T == 21 || wsz >= SK_PPT) {
                    uint32_t ls[7];                                                      // ls[j] = min(L[j..6])
                    ls[6] = L[6];
                    #pragma unroll
                    for (int j = 5; j >= 0; --j) ls[j] = min(L[j], ls[j + 1]);
                    prev0 = min(min(extra, ls[0]), C);
                    uint32_t cr = C;
                    #pragma unroll
                    for (int j = 0; j < SK_PPT; ++j) { cr = min(cr, O[j]); mn[j] = j < 7 ? min(ls[j], cr) : cr; }
                } else {
                    // short windows (2..7 k-mers): X[7 + j] = own key j, X[7 - d] = the d-th key before
                    uint32_t X[15];
                    #pragma unroll
                    for (int d = 1; d <= 7; ++d) X[7 - d] = d <= wsz ? keys[sf_pad(B0 - d)] : 0xFFFFFFFFu;
                    #pragma unroll
                    for (int j = 0; j < SK_PPT; ++j) X[7 + j] = O[j];
                    prev0 = 0xFFFFFFFFu;
                    #pragma unroll
                    for (int d = 1; d <= 7; ++d) if (d <= wsz) prev0 = min(prev0, X[7 - d]);
                    #pragma unroll
                    for (int j = 0; j < SK_PPT; ++j) {
                        uint32_t v = X[7 + j];
                        #pragma unroll
                        for (int d = 1; d < 7; ++d) if (d < wsz) v = min(v, X[7 + j - d]);
                        mn[j] = v;
                    }
                }
                // events
                uint32_t fa = 0, fb = 0;                                                 // bit j: window j is a type A / type B event
                #pragma unroll
                for (int j = 0; j < SK_PPT; ++j) {
                    const uint32_t pv = j ? mn[j - 1] : prev0;
                    fa |= (uint32_t)(O[j] < pv) << j;
                    fb |= (uint32_t)((O[j] == pv) | (mn[j] > pv)) << j;
                }
                if (first && tid == 0) { fa &= ~1u; fb |= 1u; }                          // the first window of a record segment has no predecessor
                if (jn < SK_PPT) { fa &= (1u << jn) - 1u; fb &= (1u << jn) - 1u; }
                flags = fa | (fb << 8);
            }
            compact_events(flags, par);
            __syncthreads();
            have_prev = true; prev_wb = wb; prev_boff = boff; prev_t0 = t0;
            wb = next_buf(wb);
            first = false;
        }
        // the last tile of the segment (it has the parity par ^ 1): its events, then its redo flag
        if (have_prev) {
            drain_events(reinterpret_cast<const uint64_t *>(smem_raw + prev_wb * SF_WBYTES),
                         reinterpret_cast<const uint32_t *>(smem_raw + prev_wb * SF_WBYTES + SF_NW * 8), KH + (par ^ 1u) * SF_KEYS, prev_boff, par ^ 1u);
            __syncthreads();
            cons.end_tile(cur_ent);
            if (tid == 0) settle_redo(par ^ 1u, prev_t0);
            __syncthreads();
        }
    }
    if (cur_ent != 0xFFFFFFFFu) cons.flush(cur_ent);
}

} // namespace d2g
