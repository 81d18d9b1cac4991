// lsh_kernels.cuh -- K8: LSH candidate generation + bounded neighbour lists + refinement for --topk.
//
// Reference: SetSketchIndex (/root/reference/src/ssi.h:290-453), build_index + heap update
// (src/index_build.cpp:20-165), refine_results (src/refine.cpp:6-81), emit_neighbors (src/emitnn.cpp:12-52),
// table geometry src/cmp_core.cpp:757-772.  The reference's output depends on thread timing;
// the contract implemented here is its sequential (-p1) order, reproduced exactly:
//   keys    : 32-bit key of every (table, sketch): table type 0 hashes one register, type 1 two registers;
//   index   : per table, (key, id) sorted by key with ids ascending inside a bucket (= insertion order under -p1)
//             -- a segmented radix sort replaces 1.5*S hash maps;
//   query   : one warp per sketch walks the tables most-specific first; 32 tables are looked up at a time, their
//             buckets are then merged strictly in table order, stopping the moment `maxcand` distinct ids were seen;
//   arrivals: every candidate edge (q, pos) contributes (count, q) to list[cand] and (count, cand) to list[q]; a stable
//             sort by list keeps the (q, pos, side) order, and one thread per list replays the bounded-heap rule;
//   refine  : exact compare() of every survivor (one warp per edge), sort, drop zeros, keep top-k plus ties -> CSR.
#pragma once
#include "cmp_kernels.cuh"
#include "common.cuh"

namespace d2g {

// XXH64 (the published algorithm; the reference vendors xxHash) of four 64-bit words
__device__ __forceinline__ uint64_t xxh_round(uint64_t acc, uint64_t in) { acc += in * 0xC2B2AE3D27D4EB4FULL; acc = (acc << 31) | (acc >> 33); return acc * 0x9E3779B185EBCA87ULL; }
__device__ __forceinline__ uint64_t xxh_merge(uint64_t h, uint64_t v) { h ^= xxh_round(0, v); return h * 0x9E3779B185EBCA87ULL + 0x85EBCA77C2B2AE63ULL; }
__device__ __forceinline__ uint64_t xxh64_4words(uint64_t w0, uint64_t w1, uint64_t w2, uint64_t w3, uint64_t seed) {
    const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL;
    const uint64_t v1 = xxh_round(seed + P1 + P2, w0), v2 = xxh_round(seed + P2, w1), v3 = xxh_round(seed, w2), v4 = xxh_round(seed - P1, w3);
    uint64_t h = ((v1 << 1) | (v1 >> 63)) + ((v2 << 7) | (v2 >> 57)) + ((v3 << 12) | (v3 >> 52)) + ((v4 << 18) | (v4 >> 46));
    h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
    h += 32;
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}
// XXH64 over nw 64-bit words (nw >= 4)
__device__ __forceinline__ uint64_t xxh64_words(const uint64_t *w, int nw, uint64_t seed) {
    const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL, P4 = 0x85EBCA77C2B2AE63ULL;
    uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    int i = 0;
    for (; i + 4 <= nw; i += 4) { v1 = xxh_round(v1, w[i]); v2 = xxh_round(v2, w[i + 1]); v3 = xxh_round(v3, w[i + 2]); v4 = xxh_round(v4, w[i + 3]); }
    uint64_t h = ((v1 << 1) | (v1 >> 63)) + ((v2 << 7) | (v2 >> 57)) + ((v3 << 12) | (v3 >> 52)) + ((v4 << 18) | (v4 >> 46));
    h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
    h += (uint64_t)nw * 8;
    for (; i < nw; ++i) { h ^= xxh_round(0, w[i]); h = ((h << 27) | (h >> 37)) * P1 + P4; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}
// XXH3_64bits of W = 6 .. 16 whole words (48 .. 128 bytes), seed 0, default secret: XXH3_len_17to128_64b of xxHash 0.8.0 (the published
// algorithm; the reference reaches it through hashmem's default branch, ssi.h:352).  Secret = first 128 bytes of XXH3_kSecret, little endian.
__device__ __forceinline__ uint64_t xxh3_secret(int i) {
    constexpr uint64_t sec[16] = {
        0xbe4ba423396cfeb8ULL, 0x1cad21f72c81017cULL, 0xdb979083e96dd4deULL, 0x1f67b3b7a4a44072ULL, 0x78e5c0cc4ee679cbULL, 0x2172ffcc7dd05a82ULL,
        0x8e2443f7744608b8ULL, 0x4c263a81e69035e0ULL, 0xcb00c391bb52283cULL, 0xa32e531b8b65d088ULL, 0x4ef90da297486471ULL, 0xd8acdea946ef1938ULL,
        0x3f349ce33f76faa8ULL, 0x1d4f0bc7c7bbdcf9ULL, 0x3159b4cd4be0518aULL, 0x647378d9c97e9fc8ULL};
    return sec[i];
}
__device__ __forceinline__ uint64_t xxh3_mix16(const uint64_t *in, int s) { return wymum(in[0] ^ xxh3_secret(s), in[1] ^ xxh3_secret(s + 1)); }
__device__ __forceinline__ uint64_t xxh3_64_words(const uint64_t *in, int W) {
    const uint64_t len = (uint64_t)W * 8;
    uint64_t acc = len * 0x9E3779B185EBCA87ULL;
    if (len > 32) {
        if (len > 64) {
            if (len > 96) { acc += xxh3_mix16(in + 6, 12); acc += xxh3_mix16(in + W - 8, 14); }
            acc += xxh3_mix16(in + 4, 8); acc += xxh3_mix16(in + W - 6, 10);
        }
        acc += xxh3_mix16(in + 2, 4); acc += xxh3_mix16(in + W - 4, 6);
    }
    acc += xxh3_mix16(in, 0); acc += xxh3_mix16(in + W - 2, 2);
    acc ^= acc >> 37; acc *= 0x165667919E3779F9ULL; acc ^= acc >> 32;
    return acc;
}
// Table geometry of --nLSH L (src/cmp_core.cpp:757-770): type ty hashes 1, 2, 4, 6, 8, ... registers per key (2 * ty from type 3 on) and has
// S / nper (types 0, 1) or 8S / nper tables; tables of type ty are [start[ty], start[ty] + cnt[ty]) in the sorted-table arrays; the query
// scans the most specific type first (ssi.h:425).
constexpr int LSH_MAX_TYPES = 9;      // XXH3's 17..128-byte branch covers up to 16 registers per key
struct LshGeom { uint32_t ntypes, ntab; uint32_t cnt[LSH_MAX_TYPES], start[LSH_MAX_TYPES], scan0[LSH_MAX_TYPES]; };   // scan0[ty] = tables of more specific types
__host__ __device__ __forceinline__ uint32_t lsh_nper(uint32_t ty) { return ty < 3 ? (1u << ty) : 2u * ty; }
// key of table (type, j), hash_index ssi.h:355-392: one register hashmem64, two hashmem128, four hashmem256, more XXH3_64bits while
// nper (j + 1) <= S; else XXH64 seeded with (type << 32) | j over registers picked by wyhash64(seed) -- 32 bits of it -- mod S, eight per whole
// eight of nper and then nper more
__device__ __forceinline__ uint32_t lsh_key(const double *sig, uint32_t S, uint32_t type, uint64_t j) {
    if (type == 0) return (uint32_t)wang64((uint64_t)__double_as_longlong(sig[j]));
    if (type == 1) {
        const uint64_t v0 = wang64((uint64_t)__double_as_longlong(sig[2 * j]));
        const uint64_t v1 = wang64((uint64_t)__double_as_longlong(sig[2 * j + 1]) ^ v0);
        return (uint32_t)(v0 ^ v1);
    }
    if (type == 2) {
        uint64_t v[4];
        if ((j + 1) * 4 <= S) {
            #pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = (uint64_t)__double_as_longlong(sig[4 * j + r]);
            return (uint32_t)wang64(cehash(v[0]) ^ (cehash(v[1]) * cehash(v[2]) - v[3]));
        }
        uint64_t seed = ((uint64_t)type << 32) | j;
        const uint64_t seed0 = seed;
        #pragma unroll
        for (int r = 0; r < 4; ++r) { const uint32_t pick = (uint32_t)wyhash64(seed) % S; v[r] = (uint64_t)__double_as_longlong(sig[pick]); }
        return (uint32_t)xxh64_4words(v[0], v[1], v[2], v[3], seed0);
    }
    const int nper = (int)lsh_nper(type);
    uint64_t v[32];
    if ((j + 1) * (uint64_t)nper <= S) {
        for (int r = 0; r < nper; ++r) v[r] = (uint64_t)__double_as_longlong(sig[(uint64_t)nper * j + r]);
        return (uint32_t)xxh3_64_words(v, nper);
    }
    uint64_t seed = ((uint64_t)type << 32) | j;
    const uint64_t seed0 = seed;
    const int nw = nper + 8 * (nper / 8);
    for (int r = 0; r < nw; ++r) { const uint32_t pick = (uint32_t)wyhash64(seed) % S; v[r] = (uint64_t)__double_as_longlong(sig[pick]); }
    return (uint32_t)xxh64_words(v, nw, seed0);
}
__device__ __forceinline__ uint32_t lsh_key_of_table(const double *sig, uint32_t S, const LshGeom &g, uint32_t t) {
    uint32_t ty = 0;
    while (ty + 1 < g.ntypes && t >= g.start[ty + 1]) ++ty;
    return lsh_key(sig, S, ty, t - g.start[ty]);
}
// position o of the scan order -> table index
__device__ __forceinline__ uint32_t lsh_scan_table(const LshGeom &g, uint32_t o) {
    int ty = (int)g.ntypes - 1;
    while (ty > 0 && o >= g.scan0[ty] + g.cnt[ty]) --ty;
    return g.start[ty] + (o - g.scan0[ty]);
}

// keys[t][i], ids[t][i] = i ; table t < S: type 0 register t ; t >= S: type 1 registers 2(t-S), 2(t-S)+1
__global__ void lsh_keys_kernel(const double *regs, uint64_t n, uint32_t S, uint32_t t0, uint32_t nt, uint32_t *keys, uint32_t *ids, const LshGeom g) {
    const uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (e >= (uint64_t)nt * n) return;
    const uint32_t t = t0 + (uint32_t)(e / n); const uint64_t i = e % n;
    keys[e] = lsh_key_of_table(regs + i * S, S, g, t);
    ids[e] = (uint32_t)i;
}

// One warp per query.  skeys/sids: [ntab][n] sorted per table.  Outputs cand[q][maxcand], cnt[q][maxcand], ncand[q].
__global__ void lsh_query_kernel(const double *regs, uint64_t n, uint32_t S, const uint32_t *skeys, const uint32_t *sids,
                                 uint32_t maxcand, uint32_t *cand, uint32_t *cnt, uint32_t *ncand, const LshGeom g) {
    extern __shared__ uint32_t sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const uint64_t q = (uint64_t)blockIdx.x * wpb + wib;
    if (q >= n) return;
    uint32_t *set_id = sm + (size_t)wib * 2 * maxcand, *set_ct = set_id + maxcand;
    const double *sig = regs + q * S;
    uint32_t nset = 0;
    const uint32_t ntab = g.ntab;
    // scan order, most specific first (ssi.h:425): the tables of the last type, ..., type 1 tables j = 0..S/2-1, then type 0 tables j = 0..S-1
    for (uint32_t base = 0; base < ntab && nset < maxcand; base += 32) {
        const uint32_t o = base + lane;                  // position in scan order
        uint32_t lo = 0, hi = 0;
        if (o < ntab) {
            const uint32_t t = lsh_scan_table(g, o);
            const uint32_t key = lsh_key_of_table(sig, S, g, t);
            const uint32_t *K = skeys + (uint64_t)t * n;
            uint32_t a = 0, b = (uint32_t)n;
            while (a < b) { const uint32_t mid = (a + b) >> 1; if (K[mid] < key) a = mid + 1; else b = mid; }
            lo = a; b = (uint32_t)n;
            while (a < b) { const uint32_t mid = (a + b) >> 1; if (K[mid] <= key) a = mid + 1; else b = mid; }
            hi = a;
        }
        for (int l = 0; l < 32 && nset < maxcand; ++l) {     // merge the 32 buckets strictly in table order
            const uint32_t blo = __shfl_sync(0xffffffffu, lo, l), bhi = __shfl_sync(0xffffffffu, hi, l);
            const uint32_t oo = base + l;
            if (oo >= ntab) break;
            const uint32_t t = lsh_scan_table(g, oo);
            const uint32_t *I = sids + (uint64_t)t * n;
            for (uint32_t p = blo; p < bhi && nset < maxcand; p += 32) {
                const bool have = p + lane < bhi;
                const uint32_t id = have ? I[p + lane] : 0xFFFFFFFFu;
                int found = -1;
                if (have) for (uint32_t f = 0; f < nset; ++f) if (set_id[f] == id) { found = (int)f; break; }
                const uint32_t newmask = __ballot_sync(0xffffffffu, have && found < 0);
                const uint32_t room = maxcand - nset;
                // lane index at which the set becomes full (the room-th new id), if it does within this chunk
                uint32_t stop_lane = 32;
                if ((uint32_t)__popc(newmask) >= room) {          // 0-based lane of the room-th set bit
                    uint32_t mm = newmask;
                    for (uint32_t r = 1; r < room; ++r) mm &= mm - 1;
                    stop_lane = (uint32_t)__ffs((int)mm) - 1;
                }
                if (have && (uint32_t)lane <= stop_lane) {
                    if (found >= 0) ++set_ct[found];          // ids inside one bucket are distinct: no two lanes share an entry
                    else { const uint32_t pos = nset + __popc(newmask & ((1u << lane) - 1)); set_id[pos] = id; set_ct[pos] = 1; }
                }
                nset += min((uint32_t)__popc(newmask), room);
                __syncwarp();
            }
        }
    }
    for (uint32_t f = lane; f < nset; f += 32) { cand[q * maxcand + f] = set_id[f]; cnt[q * maxcand + f] = set_ct[f]; }
    if (lane == 0) ncand[q] = nset;
}

// arrival records in (q, pos, side) order: index a = (q*maxcand + pos)*2 + side ; key = destination list
// only arrivals whose destination list lies in [x0, x1) are kept (lists are sharded over GPUs; every GPU scans all queries)
__global__ void lsh_arrivals_kernel(const uint32_t *cand, const uint32_t *cnt, const uint32_t *ncand, uint64_t n, uint32_t maxcand,
                                    uint32_t x0, uint32_t x1, uint32_t *alist, uint64_t *apay) {
    const uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (e >= n * maxcand) return;
    const uint64_t q = e / maxcand; const uint32_t pos = (uint32_t)(e % maxcand);
    uint32_t l0 = 0xFFFFFFFFu, l1 = 0xFFFFFFFFu; uint64_t p0 = 0, p1 = 0;
    if (pos < ncand[q]) {
        const uint32_t oid = cand[e];
        if (oid != (uint32_t)q) {                                       // index_build.cpp:131
            const uint32_t cd = __float_as_uint(-(float)cnt[e]);
            l0 = oid; p0 = ((uint64_t)(uint32_t)q << 32) | cd;          // update(neighbor_lists[oid], {cd, id})
            l1 = (uint32_t)q; p1 = ((uint64_t)oid << 32) | cd;          // update(neighbor_lists[id], {cd, oid})
        }
    }
    if (l0 < x0 || l0 >= x1) l0 = 0xFFFFFFFFu;
    if (l1 < x0 || l1 >= x1) l1 = 0xFFFFFFFFu;
    alist[2 * e] = l0; apay[2 * e] = p0; alist[2 * e + 1] = l1; apay[2 * e + 1] = p1;
}

struct Nb { float d; uint32_t id; };
__device__ __forceinline__ bool nb_less(const Nb &a, const Nb &b) { return a.d < b.d || (a.d == b.d && a.id < b.id); }

// Replay of update() (index_build.cpp:20-44) over the arrivals [seg[x], seg[x+1]) of list x, in arrival order.
// What the rule needs from the list is its LARGEST element (the priority-queue top) and membership in the dedup set; the
// order of the other entries is irrelevant (refinement recomputes every value and the trim step sorts).
//
// lsh_replay_warp_kernel: one WARP per list: list (unsorted, as 64-bit keys in nb_less order; in shared memory up to
// LSH_REPLAY_LCAP arrivals, else in place in HBM) and dedup set (shared memory, k <= LSH_REPLAY_DCAP), membership /
// maximum / removal as warp-wide scans.  lsh_replay_kernel: one thread per list, list kept sorted in HBM (larger k).
__device__ __forceinline__ uint64_t nb_key(const Nb &e);
__device__ __forceinline__ Nb nb_unkey(uint64_t k);
constexpr int LSH_REPLAY_LCAP = 1024, LSH_REPLAY_DCAP = 512, LSH_REPLAY_WARPS = 4;

__global__ void lsh_replay_kernel(const uint64_t *apay, const uint32_t *seg, uint64_t n, uint32_t k, uint32_t min_arrivals, Nb *lst, uint32_t *dset, uint32_t *lsize) {
    const uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (x >= n) return;
    const uint32_t s0 = seg[x], s1 = seg[x + 1];
    if (s1 - s0 <= min_arrivals) return;
    Nb *L = lst + s0; uint32_t *D = dset + s0;
    uint32_t nl = 0, nd = 0;
    auto push = [&](Nb it) { uint32_t i = nl++; while (i && nb_less(it, L[i - 1])) { L[i] = L[i - 1]; --i; } L[i] = it; };
    for (uint32_t a = s0; a < s1; ++a) {
        const uint64_t pay = apay[a];
        const Nb it{__uint_as_float((uint32_t)pay), (uint32_t)(pay >> 32)};
        bool in = false;
        for (uint32_t f = 0; f < nd; ++f) if (D[f] == it.id) { in = true; break; }
        if (in) continue;
        if (nl < k) { D[nd++] = it.id; push(it); continue; }
        const Nb top = L[nl - 1];
        if (it.d <= top.d) {
            if (top.d != it.d) { for (uint32_t f = 0; f < nd; ++f) if (D[f] == top.id) { D[f] = D[--nd]; break; } --nl; }
            push(it);                                                      // not added to the dedup set (reference quirk)
        }
    }
    lsize[x] = nl;
}

__global__ void __launch_bounds__(LSH_REPLAY_WARPS * 32)
lsh_replay_warp_kernel(const uint64_t *apay, const uint32_t *seg, uint64_t n, uint32_t k, Nb *lst, uint32_t *lsize) {
    __shared__ uint64_t sL[LSH_REPLAY_WARPS][LSH_REPLAY_LCAP];
    __shared__ uint32_t sD[LSH_REPLAY_WARPS][LSH_REPLAY_DCAP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t x = blockIdx.x * (uint64_t)LSH_REPLAY_WARPS + wid;
    if (x >= n) return;
    const uint32_t s0 = seg[x], s1 = seg[x + 1];
    // longer lists keep their keys in place in HBM (the list region itself: an entry and its key are both 8 bytes)
    uint64_t *L = s1 - s0 <= LSH_REPLAY_LCAP ? sL[wid] : reinterpret_cast<uint64_t *>(lst + s0);
    uint32_t *D = sD[wid];
    uint32_t nl = 0, nd = 0;
    uint64_t topkey = 0; uint32_t toppos = 0;                              // largest key of the list and where it sits
    for (uint32_t a = s0; a < s1; ++a) {
        const uint64_t pay = apay[a];
        const Nb it{__uint_as_float((uint32_t)pay), (uint32_t)(pay >> 32)};
        bool in = false;
        for (uint32_t f = lane; f < nd; f += 32) in |= D[f] == it.id;
        if (__any_sync(0xffffffffu, in)) continue;
        const uint64_t key = nb_key(it);
        if (nl >= k) {
            const Nb top = nb_unkey(topkey);
            if (!(it.d <= top.d)) continue;
            if (top.d != it.d) {
                // pop the top: out of the dedup set (if it is there), out of the list, new maximum
                uint32_t where = 0xFFFFFFFFu;
                for (uint32_t f = lane; f < nd; f += 32) if (D[f] == top.id) where = f;
                const unsigned found = __ballot_sync(0xffffffffu, where != 0xFFFFFFFFu);
                if (found) { const uint32_t f = __shfl_sync(0xffffffffu, where, __ffs((int)found) - 1); if (lane == 0) D[f] = D[nd - 1]; --nd; }
                if (lane == 0) L[toppos] = L[nl - 1];
                --nl;
                __syncwarp();
                uint64_t mk = 0; uint32_t mp = 0;
                for (uint32_t i = lane; i < nl; i += 32) { const uint64_t v = L[i]; if (v >= mk) { mk = v; mp = i; } }
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const uint64_t ok = __shfl_xor_sync(0xffffffffu, mk, o); const uint32_t op = __shfl_xor_sync(0xffffffffu, mp, o);
                    if (ok > mk || (ok == mk && op > mp)) { mk = ok; mp = op; }
                }
                topkey = mk; toppos = mp;
            }
        } else {
            if (lane == 0) D[nd] = it.id;
            ++nd;
        }
        if (lane == 0) L[nl] = key;
        if (nl == 0 || key >= topkey) { topkey = key; toppos = nl; }
        ++nl;
        __syncwarp();
    }
    for (uint32_t i = lane; i < nl; i += 32) lst[s0 + i] = nb_unkey(L[i]);
    if (lane == 0) lsize[x] = nl;
}

// one warp per surviving edge: exact compare() of list owner x and neighbour id.  Consecutive warps work on the same list, so the owner's
// row comes out of L1 / L2 and HBM only streams the neighbours' rows: 210.8 M entries x 8 KiB in 309 ms at n = 10^6, S = 1024 = 5.6 TB/s,
// 86 % of the measured HBM peak (profiles/r2k_lsh_refine_per_entry_n100000_S1024.ncu.txt).  A one-warp-per-list variant that kept the owner's
// row in registers was measured and dropped: the same HBM bytes with fewer warps in flight, 348 ms.
// KIND 0: gt / lt counts (f64 registers, or log-quantised ones: cmp_kind 2); KIND 1: equal count (cmp_kind 1 and 3).  regs are the
// registers compare() sees -- the compressed ones under --fastcmp, whatever the index was built over (cmp_core.cpp:362-449).
template <int KIND>
__global__ void lsh_refine_kernel(const double *regs, const double *cards, uint64_t n, const uint32_t *seg, const uint32_t *lsize,
                                  Nb *lst, const CmpConsts c, float mult) {
    const uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
    // edges are addressed by arrival slot: slot e belongs to list x if seg[x] <= e < seg[x] + lsize[x]
    const uint32_t total = seg[n];
    if (w >= total) return;
    // find the list that owns slot w (seg is sorted)
    uint64_t lo = 0, hi = n;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (seg[mid + 1] > w) hi = mid; else lo = mid + 1; }
    const uint64_t x = lo;
    if (w - seg[x] >= lsize[x]) return;
    const uint32_t id = lst[w].id;
    const double *A = regs + x * c.S, *B = regs + (uint64_t)id * c.S;
    uint32_t g = 0, l = 0;
    #pragma unroll 8
    for (uint32_t r = lane; r < c.S; r += 32) {            // 8 independent 256-byte row segments in flight per warp
        const double a = __ldg(A + r), b = __ldg(B + r);
        if (KIND == 0) { g += a > b; l += a < b; } else g += !(a == b);
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { g += __shfl_xor_sync(0xffffffffu, g, o); l += __shfl_xor_sync(0xffffffffu, l, o); }
    if (lane == 0) lst[w].d = mult * finalize_pair(c, KIND == 0 ? g : c.S - g, l, cards[x], cards[id]);
}

// Sort by (d, id), drop zero similarities, keep top-k plus ties, undo the sign (refine.cpp:30-74).
// lsh_trim_kernel: one WARP per list of up to LSH_TRIM_CAP entries, bitonic network in shared memory on 64-bit keys
// (order-preserving image of d, then id).  lsh_trim_big_kernel: one CTA per longer list (the bounded-heap rule keeps
// ties, so a few lists grow past a thousand entries), same network over LSH_TRIM_BIG_CAP slots; beyond that one thread
// sorts by insertion.
constexpr int LSH_TRIM_CAP = 512, LSH_TRIM_WARPS = 4, LSH_TRIM_BIG_CAP = 8192, LSH_TRIM_BIG_THREADS = 256;
// key order == nb_less order: -0.0 and +0.0 compare equal there, so zeros share one key image and the sign of a zero
// travels in bit 0 of the low word, below the id (ids are < 2^31)
__device__ __forceinline__ uint64_t nb_key(const Nb &e) {
    const uint32_t raw = __float_as_uint(e.d);
    uint32_t f = e.d == 0.f ? 0u : raw;
    f = (f >> 31) ? ~f : (f | 0x80000000u);
    return ((uint64_t)f << 32) | ((uint64_t)e.id << 1) | (e.d == 0.f ? raw >> 31 : 0u);
}
__device__ __forceinline__ Nb nb_unkey(uint64_t k) {
    uint32_t f = (uint32_t)(k >> 32);
    f = (f >> 31) ? (f & 0x7fffffffu) : ~f;
    Nb e; e.id = (uint32_t)k >> 1;
    e.d = __uint_as_float(f);
    if (e.d == 0.f && (k & 1u)) e.d = -0.f;
    return e;
}
// NT threads (a warp or a CTA) sort B[0..P) ascending and trim; SYNC() separates the steps
template <int NT, class Sync>
__device__ __forceinline__ void trim_sorted_list(uint64_t *B, Nb *L, uint32_t nl, uint32_t topk, int is_dist, uint32_t *lsize_x, int t, uint32_t *scratch, Sync sync, int keep_zeros = 0) {
    uint32_t P = 32; while (P < nl) P <<= 1;
    for (uint32_t i = t; i < P; i += NT) B[i] = i < nl ? nb_key(L[i]) : ~0ULL;
    sync();
    for (uint32_t k = 2; k <= P; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = t; i < P; i += NT) {
                const uint32_t l = i ^ j;
                if (l > i) {
                    const uint64_t a = B[i], c = B[l];
                    const bool up = (i & k) == 0;
                    if ((a > c) == up) { B[i] = c; B[l] = a; }
                }
            }
            sync();
        }
    // sorted ascending.  Similarities are stored negated: zeros sort last and are dropped; distances keep theirs.
    // Both cut-offs are prefixes of the sorted list, so counting suffices.
    if (t == 0) { scratch[0] = 0; scratch[1] = 0; }
    sync();
    uint32_t nz = 0;
    if (!is_dist && !keep_zeros) { for (uint32_t i = t; i < nl; i += NT) nz += nb_unkey(B[i]).d != 0.f; if (nz) atomicAdd(scratch, nz); }
    sync();
    uint32_t keep = (is_dist || keep_zeros) ? nl : scratch[0];
    if (topk < keep) {                       // everything tied with the k-th entry stays (refine.cpp:39-42)
        const float bs = nb_unkey(B[topk - 1]).d;
        uint32_t le = 0;
        for (uint32_t i = t; i < keep; i += NT) le += !(nb_unkey(B[i]).d > bs);
        if (le) atomicAdd(scratch + 1, le);
        sync();
        keep = scratch[1];
    }
    for (uint32_t i = t; i < keep; i += NT) { Nb e = nb_unkey(B[i]); if (!is_dist) e.d = -e.d; L[i] = e; }
    if (t == 0) *lsize_x = keep;
}
__global__ void __launch_bounds__(LSH_TRIM_WARPS * 32)
lsh_trim_kernel(const uint32_t *seg, uint64_t n, uint32_t topk, int is_dist, Nb *lst, uint32_t *lsize, int keep_zeros = 0) {
    __shared__ uint64_t buf[LSH_TRIM_WARPS][LSH_TRIM_CAP];
    __shared__ uint32_t scr[LSH_TRIM_WARPS][2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t x = blockIdx.x * (uint64_t)LSH_TRIM_WARPS + wid;
    if (x >= n) return;
    const uint32_t nl = lsize[x];
    if (nl > LSH_TRIM_CAP) return;           // lsh_trim_big_kernel
    trim_sorted_list<32>(buf[wid], lst + seg[x], nl, topk, is_dist, lsize + x, lane, scr[wid], [] { __syncwarp(); }, keep_zeros);
}
__global__ void __launch_bounds__(LSH_TRIM_BIG_THREADS)
lsh_trim_big_kernel(const uint32_t *seg, uint64_t n, uint32_t topk, int is_dist, Nb *lst, uint32_t *lsize, int keep_zeros = 0) {
    extern __shared__ uint64_t bigbuf[];
    __shared__ uint32_t scr[2];
    const uint64_t x = blockIdx.x;
    uint32_t nl = lsize[x];
    if (nl <= LSH_TRIM_CAP) return;
    Nb *L = lst + seg[x];
    if (nl <= LSH_TRIM_BIG_CAP) { trim_sorted_list<LSH_TRIM_BIG_THREADS>(bigbuf, L, nl, topk, is_dist, lsize + x, (int)threadIdx.x, scr, [] { __syncthreads(); }, keep_zeros); return; }
    if (threadIdx.x == 0) {
        for (uint32_t i = 1; i < nl; ++i) { const Nb it = L[i]; uint32_t j = i; while (j && nb_less(it, L[j - 1])) { L[j] = L[j - 1]; --j; } L[j] = it; }
        if (!is_dist && !keep_zeros) { uint32_t j = 0; while (j < nl && L[j].d != 0.f) ++j; nl = j; }
        if (topk < nl) { const float bs = L[topk - 1].d; uint32_t j = topk; while (j < nl && !(L[j].d > bs)) ++j; nl = j; }
        if (!is_dist) for (uint32_t j = 0; j < nl; ++j) L[j].d = -L[j].d;
        lsize[x] = nl;
    }
}

__global__ void lsh_csr_kernel(const uint32_t *seg, const uint32_t *lsize, const uint64_t *indptr, uint64_t n, const Nb *lst, uint32_t *idx, float *val) {
    const uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (x >= n) return;
    const Nb *L = lst + seg[x]; const uint64_t o = indptr[x];
    for (uint32_t j = 0; j < lsize[x]; ++j) { idx[o + j] = L[j].id; val[o + j] = L[j].d; }
}

// seg[x] = first arrival slot whose list >= x (lists sorted ascending; 0xFFFFFFFF = unused slots at the end)
__global__ void lsh_segments_kernel(const uint32_t *alist_sorted, uint64_t na, uint64_t n, uint32_t *seg) {
    const uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (x > n) return;
    uint64_t lo = 0, hi = na;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (alist_sorted[mid] < (uint32_t)x) lo = mid + 1; else hi = mid; }
    seg[x] = (uint32_t)lo;
}


// ---- similarity-threshold graphs (--similarity-threshold x; src/index_build.cpp:26-31 with topk = -1, src/refine.cpp:43-68) ----------
// Lists are not capped: every candidate of every query joins both endpoints' lists unless it is there already, and keeps the hit count
// of its FIRST arrival.  Over the arrivals sorted by list (stable, so still in arrival order inside a list) that is: sort by
// (id, arrival rank) inside the list, keep the head of every id run, sort the heads by (-hits, id) -- two segmented radix sorts.
__global__ void lsh_thr_key1_kernel(const uint32_t *alist, const uint64_t *apay, const uint32_t *seg, uint32_t total, uint64_t *k1, uint32_t *cd) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= total) return;
    const uint64_t pay = apay[a];
    k1[a] = (pay & 0xFFFFFFFF00000000ULL) | (uint64_t)(a - seg[alist[a]]);
    cd[a] = (uint32_t)pay;
}
__global__ void lsh_thr_key2_kernel(const uint32_t *alist, const uint64_t *k1s, const uint32_t *cds, const uint32_t *seg, uint32_t total, uint64_t *k2) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= total) return;
    const uint32_t id = (uint32_t)(k1s[a] >> 32);
    const bool head = a == seg[alist[a]] || (uint32_t)(k1s[a - 1] >> 32) != id;
    k2[a] = head ? nb_key(Nb{__uint_as_float(cds[a]), id}) : ~0ULL;
}
// one thread per list: entries (heads first after the sort) from keys to Nb in place, list size
__global__ void lsh_thr_lists_kernel(const uint32_t *seg, uint64_t n, uint64_t *k2s_and_lst, uint32_t *lsize) {
    const uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (x >= n) return;
    const uint32_t s0 = seg[x], s1 = seg[x + 1];
    uint32_t nl = 0;
    Nb *L = reinterpret_cast<Nb *>(k2s_and_lst);
    for (uint32_t a = s0; a < s1; ++a) { const uint64_t k = k2s_and_lst[a]; if (k == ~0ULL) break; L[a] = nb_unkey(k); ++nl; }
    lsize[x] = nl;
}
// refine.cpp:45-68 over a list in (-hits, id) order whose d already holds mult * measure: keep what passes the threshold, give up after
// 20 consecutive failures.  One thread per list; survivors compacted in place (the trim kernels sort them afterwards).
__global__ void lsh_thr_filter_kernel(const uint32_t *seg, uint64_t n, double min_sim, int is_dist, Nb *lst, uint32_t *lsize) {
    const uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (x >= n) return;
    Nb *L = lst + seg[x];
    const uint32_t nl = lsize[x];
    uint32_t o = 0, failures = 0;
    for (uint32_t j = 0; j < nl; ++j) {
        const Nb e = L[j];
        const float v = is_dist ? e.d : -e.d;
        const bool pass = is_dist ? (double)v < min_sim : (double)v >= min_sim;
        if (pass) { L[o++] = e; failures = 0; }
        else if (++failures == 20) break;
    }
    lsize[x] = o;
}

} // namespace d2g
