// stream_kernels.cuh -- the k-mer streams that are NOT the exact 2-bit canonical encoder of sketch_kernels.cuh:
//   * k > 32: RollingHasher<uint64_t> over CyclicHash (/root/reference/bonsai/include/bonsai/encoder.h:644-865,
//     rollinghash/cyclichash.h:101-121), dispatch src/fastxsketch.cpp:399-421;
//   * -C with a window (k <= 32): Encoder::for_each_uncanon, encoder.h:274-307;
//   * protein alphabets: the same rolling encode over alph::AMINO20 / SEB14 / SEB6 / SEB8 (alphabet.h:107-120, rhtraits.h:52-62).
// All three have state that the reference carries along the record (the jump over an N, the accumulator of `(min * mul) | code`,
// a window that is not reset at an invalid base), so they are produced in two steps:
//   produce : a mode-specific kernel writes the ELEMENTS of every record (pre-maskfn values, in the order the reference pushes them)
//             into the item region of the record, items[mult * rec_off[r] ... + item_cnt[r]);
//   consume : stream_kernel<Consumer> walks the item regions exactly as sketch_kernel walks sequence positions -- same spans, same
//             Consumer protocol -- taking the minimum by (FRev64 score, element) over w-k+1 consecutive items when windowed
//             (qmap.h:79-87; FRev64 is a bijection, so the score alone orders), one output per full window, and the single flush of
//             a window that never filled (encoder.h:304-305,795-796).
// Launchers see this as a PackedSeq whose `items` is set; rec_off then holds the item-region offsets and (k, w) = (1, w-k+1).
#pragma once
#include "common.cuh"
#include "sketch_kernels.cuh"

namespace d2g {

constexpr uint32_t STREAM_SKIP_ONES = 1u;      // a full window whose minimizer is ~0 emits nothing (encoder.h:299 `if(... != ERROR)`)
constexpr uint32_t STREAM_NOFLUSH_BIT = 0x80000000u;   // item_cnt[r]: the record ended where the reference returns without the tail flush

// ---- consume ----------------------------------------------------------------------------------------------------------------
constexpr int ST_SCORE_SLOTS = SK_TILE + SK_MAX_W + 8;

template <class Consumer, bool FILTER = false>
__global__ void __launch_bounds__(SK_THREADS, Consumer::kMinBlocks)
stream_kernel(const SketchArgs a, const typename Consumer::Params cp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const bool windowed = a.w > a.k;
    const int wsz = windowed ? a.w : 1;
    uint64_t *score = reinterpret_cast<uint64_t *>(smem_raw);
    Consumer cons;
    cons.init(smem_raw + (windowed ? (size_t)ST_SCORE_SLOTS * 8 : 0), cp, false);
    const uint64_t span_lo = a.pos_base + (uint64_t)blockIdx.x * a.span;
    const uint64_t span_hi = min(span_lo + a.span, a.pos_end);
    if (span_lo >= span_hi) return;
    uint64_t lo = 0, hi = a.n_rec;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (a.rec_off[mid + 1] > span_lo) hi = mid; else lo = mid + 1; }
    uint32_t cur_ent = 0xFFFFFFFFu;
    const uint64_t *items = a.seq.items;
    auto feed = [&](uint64_t hv) { if (FILTER && sk_filtered(a, hv)) return; cons.consume(hv); };
    __syncthreads();
    for (uint64_t r = lo; r < a.n_rec; ++r) {
        const uint64_t rs = a.rec_off[r];
        if (rs >= span_hi) break;
        const uint32_t cf = a.seq.item_cnt[r];
        const uint64_t cnt = cf & ~STREAM_NOFLUSH_BIT;
        if (cnt == 0) continue;
        const bool tail = cnt < (uint64_t)wsz;               // the window never fills: one flush of everything pushed
        if (tail && (cf & STREAM_NOFLUSH_BIT)) continue;
        const uint64_t nout = tail ? 1 : cnt - wsz + 1;
        const uint64_t p0 = max(span_lo, rs), p1 = min(span_hi, rs + nout);
        if (p0 >= p1) continue;
        const uint32_t ent = a.rec_entity[r] - a.ent_base;
        if (a.ent_state && a.ent_state[ent] != a.want_state) continue;
        if (ent != cur_ent) {
            if (cur_ent != 0xFFFFFFFFu) cons.flush(cur_ent);
            cur_ent = ent;
            cons.begin_entity(ent, p0 - span_lo);
        }
        for (uint64_t t0 = p0; t0 < p1; t0 += SK_TILE) {
            if (a.tile_stride > 1 && ((t0 / SK_TILE) % a.tile_stride) != 0) continue;
            const int nstart = (int)min((uint64_t)SK_TILE, p1 - t0);
            __syncthreads();
            cons.end_tile(cur_ent);
            if (!windowed) {
                for (int j = threadIdx.x; j < nstart; j += SK_THREADS) feed(wang64(items[t0 + j] ^ a.xormask));
                continue;
            }
            const int wl = tail ? (int)cnt : wsz;
            const int nload = nstart + wl - 1;
            for (int i = threadIdx.x; i < nload; i += SK_THREADS) score[i] = frev64(items[t0 + i]);
            __syncthreads();
            for (int j = threadIdx.x; j < nstart; j += SK_THREADS) {
                uint64_t mn = score[j];
                for (int q = 1; q < wl; ++q) mn = min(mn, score[j + q]);
                const uint64_t el = frev64_inv(mn);
                if (tail || !(a.seq.item_flags & STREAM_SKIP_ONES) || el != ~0ULL) feed(wang64(el ^ a.xormask));
            }
        }
    }
    if (cur_ent != 0xFFFFFFFFu) cons.flush(cur_ent);
}

template <class Consumer>
inline size_t stream_smem_bytes(uint32_t m, bool windowed) { return (windowed ? (size_t)ST_SCORE_SLOTS * 8 : 0) + Consumer::smem_bytes(m, false); }

#ifdef D2G_STREAM_PRODUCERS   // only api_stream.cu launches the producers
// ---- produce: shared helpers ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t seq_code(const uint64_t *codes, uint64_t p) { return (uint32_t)(codes[p >> 5] >> (62 - 2 * (int)(p & 31))) & 3u; }
__device__ __forceinline__ bool seq_invalid(const uint32_t *mask, uint64_t p) { return (mask[p >> 5] >> (31 - (int)(p & 31))) & 1u; }
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { r &= 63; return (x << r) | (x >> ((64 - r) & 63)); }
__device__ __forceinline__ uint64_t rotr64(uint64_t x, int r) { r &= 63; return (x >> r) | (x << ((64 - r) & 63)); }

// first invalid base in [from, end) (end if there is none); the whole warp calls with the same arguments
__device__ __forceinline__ uint64_t warp_next_invalid(const uint32_t *mask, uint64_t from, uint64_t end) {
    const int lane = threadIdx.x & 31;
    for (uint64_t w0 = from >> 5; w0 * 32 < end; w0 += 32) {
        const uint64_t wi = w0 + lane;
        uint32_t mw = wi * 32 < end ? mask[wi] : 0u;
        if (wi == (from >> 5)) mw &= 0xFFFFFFFFu >> (from & 31);
        const unsigned b = __ballot_sync(0xffffffffu, mw != 0);
        if (b) {
            const int f = __ffs((int)b) - 1;
            const uint32_t mf = __shfl_sync(0xffffffffu, mw, f);
            const uint64_t pos = (w0 + f) * 32 + __clz((int)mf);
            return pos < end ? pos : end;
        }
    }
    return end;
}

// ---- k > 32: cyclic-polynomial rolling hash ---------------------------------------------------------------------------------
// The forward hash of the window at p is XOR_j rotl(tf[c(p+j)], K-1-j).  The reference's reverse-complement hasher is seeded with K
// copies of the complement of the LAST base of the first window of a run (encoder.h:713: the index i - nfilled + k - 1 does not move
// while the window fills) and from there follows the proper reverse update, so its state at p is
//     RC(p) ^ rotr(E0, p - s0),   RC(p) = XOR_j rotl(tr[3 - c(p+j)], j),   E0 = q[c(s0+K-1)] ^ RC(s0)
// for the run that starts at s0 -- a pure function of the window and the run start.  Runs are found by one warp per record
// (roll_walk_kernel: an N at i makes the next examined base i+K+1, encoder.h:700-707,746-750); the hashes by one thread per 32
// consecutive start positions (roll_fill_kernel), rolled inside the chunk.
struct RollConsts { uint64_t tf[4], tr[4], q[4]; int K, canon, windowed; };
struct RollSeg { uint64_t s0, e, e0; uint64_t out_base; };   // windows start at s0 .. e-K; out_base = items of the record before this run

__device__ __forceinline__ uint64_t roll_seg_base(uint64_t rs, uint64_t r, int K) { return rs / (uint64_t)(K + 1) + r; }

static __global__ void roll_walk_kernel(const PackedSeq seq, const uint64_t *rec_off, uint64_t n_rec, const RollConsts rc, RollSeg *segs, uint32_t *nseg,
                                 uint32_t *item_cnt) {
    const uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rec) return;
    const int lane = threadIdx.x & 31;
    const uint64_t rs = rec_off[r], re = rec_off[r + 1], l = re - rs, K = (uint64_t)rc.K;
    const uint64_t mult = (rc.canon && rc.windowed) ? 2 : 1;
    RollSeg *out = segs + roll_seg_base(rs, r, rc.K);
    uint64_t cur = rs, items = 0; uint32_t ns = 0; bool noflush = false;
    if (l >= K) {
        for (;;) {
            const uint64_t ni = warp_next_invalid(seq.mask, cur, re);
            if (ni - cur >= K) {
                uint64_t e0 = 0;
                if (rc.canon) {
                    for (uint64_t j = lane; j < K; j += 32) e0 ^= rotl64(rc.tr[3u - seq_code(seq.codes, cur + j)], (int)(j & 63));
                    #pragma unroll
                    for (int o = 16; o > 0; o >>= 1) e0 ^= __shfl_xor_sync(0xffffffffu, e0, o);
                    e0 ^= rc.q[seq_code(seq.codes, cur + K - 1)];
                }
                if (lane == 0) out[ns] = RollSeg{cur, ni, e0, items};
                ++ns; items += mult * (ni - cur - K + 1);
            }
            if (ni >= re) { if (!rc.canon && ni - cur < K) noflush = true; break; }     // encoder.h:776 "All failed": no flush
            if (rc.canon && (ni - rs) + 2 * K >= l) break;                               // encoder.h:702,747
            cur = ni + K + 1;
            if (cur >= re) { if (!rc.canon) noflush = true; break; }
        }
    }
    if (lane == 0) { nseg[r] = ns; item_cnt[r] = (uint32_t)items | (noflush ? STREAM_NOFLUSH_BIT : 0u); }
}

constexpr int ROLL_CHUNK = 32;
static __global__ void roll_fill_kernel(const PackedSeq seq, const uint64_t *rec_off, uint64_t n_rec, uint64_t total_len, const RollConsts rc, const RollSeg *segs,
                                 const uint32_t *nseg, uint64_t *items) {
    const uint64_t pb = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * ROLL_CHUNK;
    if (pb >= total_len) return;
    const uint64_t pe = min(pb + ROLL_CHUNK, total_len);
    const int K = rc.K;
    const uint64_t mult = (rc.canon && rc.windowed) ? 2 : 1;
    uint64_t p = pb;
    while (p < pe) {
        // record and run holding p (or the next run start after it)
        uint64_t lo = 0, hi = n_rec;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (rec_off[mid + 1] > p) hi = mid; else lo = mid + 1; }
        if (lo >= n_rec) return;
        const uint64_t r = lo, rs = rec_off[r], re = rec_off[r + 1];
        const RollSeg *list = segs + roll_seg_base(rs, r, K);
        const uint32_t n = nseg[r];
        uint32_t a = 0, b = n;                                   // first run with s0 > p
        while (a < b) { const uint32_t mid = (a + b) >> 1; if (list[mid].s0 > p) b = mid; else a = mid + 1; }
        if (a == 0) { p = n ? list[0].s0 : re; continue; }
        const RollSeg sg = list[a - 1];
        const uint64_t last = sg.e - K;                          // last window start of the run
        if (p > last) { p = a < n ? list[a].s0 : re; continue; }
        const uint64_t stop = min(pe, last + 1);
        // direct evaluation at p, then rolled
        uint64_t fw = 0, rv = 0;
        for (int j = 0; j < K; ++j) {
            const uint32_t c = seq_code(seq.codes, p + j);
            fw ^= rotl64(rc.tf[c], K - 1 - j);
            if (rc.canon) rv ^= rotl64(rc.tr[3u - c], j);
        }
        rv ^= rotr64(sg.e0, (int)((p - sg.s0) & 63));
        uint64_t *dst = items + mult * rs + sg.out_base + mult * (p - sg.s0);
        for (;;) {
            if (mult == 2) { dst[0] = fw; dst[1] = rv; }
            else dst[0] = rc.canon ? min(fw, rv) : fw;
            dst += mult;
            if (++p >= stop) break;
            const uint32_t co = seq_code(seq.codes, p - 1), cn = seq_code(seq.codes, p + K - 1);
            fw = rotl64(fw, 1) ^ rotl64(rc.tf[co], K) ^ rc.tf[cn];                                   // cyclichash.h:101-108
            if (rc.canon) rv = rotr64(rv ^ rotl64(rc.tr[3u - cn], K) ^ rc.tr[3u - co], 1);           // cyclichash.h:110-116
        }
    }
}

// ---- -C with a window, k <= 32 (encoder.h:274-307) --------------------------------------------------------------------------------
// The accumulator `min = (min << 2) | code` restarts the k-mer run (not the window) when it becomes all ones: at an invalid base
// (code -1), and -- for k >= 31, where the unmasked accumulator spans 64 bits -- at every 32nd base of a run of T's.  Those bases are
// added to a copy of the invalid mask first (ncw_virtual_invalid_kernel); a k-mer is pushed iff its k bases hold none of them.
static __global__ void ncw_virtual_invalid_kernel(const PackedSeq seq, const uint64_t *rec_off, uint64_t n_rec, uint64_t total_len, uint32_t *vmask, uint64_t n_words) {
    const uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;       // one warp = one mask word
    bool v = true;                                                        // beyond the end: invalid, as in the packed mask
    if (q < total_len) {
        v = seq_invalid(seq.mask, q);
        if (!v && q >= 31) {
            bool allT = true;
            const uint64_t w = q >> 5; const int s = (int)(q & 31);
            // bases q-31 .. q: all T and none invalid
            const uint64_t hi = seq.codes[w], lo = w ? seq.codes[w - 1] : 0;
            const uint64_t win = s == 31 ? hi : ((lo << (2 * (s + 1))) | (hi >> (62 - 2 * s)));
            const uint32_t mh = seq.mask[w], ml = w ? seq.mask[w - 1] : 0;
            const uint32_t mwin = s == 31 ? mh : ((ml << (s + 1)) | (mh >> (31 - s)));
            allT = win == ~0ULL && mwin == 0;
            if (allT) {
                uint64_t lo2 = 0, hi2 = n_rec;
                while (lo2 < hi2) { const uint64_t mid = (lo2 + hi2) >> 1; if (rec_off[mid + 1] > q) hi2 = mid; else lo2 = mid + 1; }
                const uint64_t rs = rec_off[lo2];
                if (q - rs >= 31) {
                    uint64_t a = q - 31;                           // start of the run of valid T's
                    while (a > rs && seq_code(seq.codes, a - 1) == 3u && !seq_invalid(seq.mask, a - 1)) --a;
                    v = ((q - a + 1) & 31) == 0;
                }
            }
        }
    }
    const unsigned b = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && (q >> 5) < n_words) vmask[q >> 5] = __brev(b);
}

// flag[p] = 1 iff a k-mer is pushed for start position p
static __global__ void ncw_flag_kernel(const uint32_t *emask, const uint64_t *rec_off, uint64_t n_rec, uint64_t total_len, int k, uint32_t *flag) {
    const uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (p > total_len) return;
    uint32_t f = 0;
    if (p < total_len) {
        uint64_t lo = 0, hi = n_rec;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (rec_off[mid + 1] > p) hi = mid; else lo = mid + 1; }
        if (lo < n_rec && p + k <= rec_off[lo + 1]) {
            const uint64_t w = p >> 5; const int s = (int)(p & 31);
            const uint32_t m0 = __funnelshift_l(emask[w + 1], emask[w], s);                       // bases p .. p+31
            f = (m0 >> (32 - k)) == 0;
        }
    }
    flag[p] = f;
}
static __global__ void ncw_fill_kernel(const PackedSeq seq, const uint64_t *rec_off, uint64_t n_rec, uint64_t total_len, int k, const uint32_t *flag,
                                const uint32_t *excl, uint64_t *items, uint32_t *item_cnt) {
    const uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (p < n_rec) item_cnt[p] = excl[rec_off[p + 1]] - excl[rec_off[p]];
    if (p >= total_len || !flag[p]) return;
    uint64_t lo = 0, hi = n_rec;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (rec_off[mid + 1] > p) hi = mid; else lo = mid + 1; }
    const uint64_t rs = rec_off[lo];
    const uint64_t w = p >> 5; const int s = (int)(p & 31) * 2;
    const uint64_t hi64 = seq.codes[w], lo64 = seq.codes[w + 1];
    const uint64_t x = s ? ((hi64 << s) | (lo64 >> (64 - s))) : hi64;
    items[rs + (excl[p] - excl[rs])] = x >> (64 - 2 * k);
}

// ---- protein alphabets (non-canonical; alphabet.h:107-120, rhtraits.h:52-62, encoder.h:241-306) -------------------------------------
// `min = (min * mul) | code` then `min %= mul^k` (or `&=` a k-bit mask for the 3-bit alphabet) with the reduced value carried on: the OR
// makes the state depend on the whole run, so a record is encoded by one thread, sequentially, as the reference does.  Records are
// proteins (hundreds of residues), the parallelism is across them.
struct ProteinConsts { uint64_t mul, mask; int k, windowed, bitmask; };
static __global__ void protein_kernel(const uint8_t *seq, const uint64_t *rec_off, uint64_t n_rec, const int8_t *lut_d, const ProteinConsts pc, uint64_t *items,
                               uint32_t *item_cnt) {
    __shared__ int8_t lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = lut_d[i];
    __syncthreads();
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint64_t rs = rec_off[r], re = rec_off[r + 1];
    uint64_t *dst = items + rs;
    uint64_t kmer = 0, n = 0; int filled = 0;
    for (uint64_t pos = rs; pos < re; ++pos) {
        const int8_t nv = lut[seq[pos]];
        if (!pc.windowed) {
            if (nv < 0) { kmer = 0; filled = 0; continue; }
            kmer = (kmer * pc.mul) | (uint64_t)nv;
        } else {
            kmer = (kmer * pc.mul) | (uint64_t)(int64_t)nv;          // -1 sign-extends to all ones, encoder.h:285
            if (kmer == ~0ULL) { kmer = 0; filled = 0; continue; }
        }
        if (++filled == pc.k) {
            kmer = pc.bitmask ? (kmer & pc.mask) : (kmer % pc.mask);
            dst[n++] = kmer;
            --filled;
        }
    }
    item_cnt[r] = (uint32_t)n;
}

#endif // D2G_STREAM_PRODUCERS

} // namespace d2g
