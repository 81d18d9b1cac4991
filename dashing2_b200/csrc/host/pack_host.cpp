// pack_host.cpp -- host side of the packed sequence layout (csrc/pack_kernels.cuh): ASCII bases -> 2 bits per base
// + 1 invalid bit per base, written by the host threads while the previous chunk is on its way to the device.
// One byte per base over PCIe is what bounded the end-to-end sketch rate; packed it is a quarter (plus an eighth
// for the mask words, which are only uploaded for chunks that hold a non-ACGT base at all).
//
// Alphabet: A0 C1 G2 T3, case-insensitive, anything else invalid (bonsai/include/bonsai/alphabet.h:128 DNA4).
// Layout: codes word i = bases [32i, 32i+32), first base in bits 63:62; mask word i likewise, first base in bit 31.
#include "pack_host.h"

#include <immintrin.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace d2g_host {

// ---- a small persistent pool: parallel_for(n, fn) runs fn(0..n-1) on the workers + the caller ----------------
namespace {
struct Pool {
    std::vector<std::thread> th;
    std::mutex mu; std::condition_variable cv, done_cv;
    const std::function<void(size_t)> *fn = nullptr;
    size_t n = 0; std::atomic<size_t> next{0}; size_t active = 0; uint64_t gen = 0; bool stop = false;
    explicit Pool(unsigned nt) {
        for (unsigned t = 1; t < nt; ++t) th.emplace_back([this] { worker(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto &t : th) t.join();
    }
    void worker() {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(size_t)> *f; size_t cnt;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                if (!fn) continue;                         // woke after that run had already finished
                f = fn; cnt = n; ++active;
            }
            for (size_t i; (i = next.fetch_add(1)) < cnt;) (*f)(i);
            { std::lock_guard<std::mutex> lk(mu); --active; }
            done_cv.notify_all();
        }
    }
    void run(size_t cnt, const std::function<void(size_t)> &f) {
        if (th.empty() || cnt <= 1) { for (size_t i = 0; i < cnt; ++i) f(i); return; }
        {
            std::lock_guard<std::mutex> lk(mu);
            fn = &f; n = cnt; next.store(0); ++gen;
        }
        cv.notify_all();
        for (size_t i; (i = next.fetch_add(1)) < cnt;) f(i);
        std::unique_lock<std::mutex> lk(mu);
        // every worker that picked this generation up has left its loop; workers that never woke see next >= n
        done_cv.wait(lk, [&] { return active == 0; });
        fn = nullptr;
    }
};
std::mutex g_pool_mu;                // one parallel_for at a time (contexts on several devices share the pool)
Pool *g_pool = nullptr;
}  // namespace

unsigned host_threads() {
    static unsigned nt = [] {
        if (const char *ev = getenv("D2G_HOST_THREADS")) { const int v = atoi(ev); if (v > 0) return (unsigned)std::min(v, 256); }
        cpu_set_t set; CPU_ZERO(&set);
        unsigned n = 0;
        if (sched_getaffinity(0, sizeof set, &set) == 0) n = (unsigned)CPU_COUNT(&set);
        if (!n) n = std::max(1u, std::thread::hardware_concurrency());
        return std::min(n, 64u);
    }();
    return nt;
}

void parallel_for(size_t n, const std::function<void(size_t)> &fn) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!g_pool) g_pool = new Pool(host_threads());
    g_pool->run(n, fn);
}

// ---- 32 / 64 bases at a time ------------------------------------------------------------------------------------
namespace {

struct Lut { uint8_t code[256]; uint8_t inv[256]; };
const Lut &lut() {
    static const Lut L = [] {
        Lut l;
        for (int c = 0; c < 256; ++c) {
            const unsigned x = ((unsigned)c >> 1) & 3u;
            l.code[c] = (uint8_t)(x ^ (x >> 1));            // the same two bits of the byte the device decode takes
            const int u = c & 0xDF;
            l.inv[c] = !(u == 'A' || u == 'C' || u == 'G' || u == 'T');
        }
        return l;
    }();
    return L;
}

inline void pack32_scalar(const uint8_t *s, uint64_t &cw, uint32_t &mw) {
    const Lut &l = lut();
    uint64_t c = 0; uint32_t m = 0;
    for (int i = 0; i < 32; ++i) { c = (c << 2) | l.code[s[i]]; m = (m << 1) | l.inv[s[i]]; }
    cw = c; mw = m;
}

__attribute__((target("avx2,bmi2"))) inline void pack32_avx2(const uint8_t *s, uint64_t &cw, uint32_t &mw) {
    const __m256i rev = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s));
    v = _mm256_shuffle_epi8(v, rev);
    v = _mm256_permute2x128_si256(v, v, 1);                // byte j = base 31 - j: movemask bit 31 is the first base
    const __m256i u = _mm256_and_si256(v, _mm256_set1_epi8((char)0xDF));
    const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(u, _mm256_set1_epi8('C'))),
                                       _mm256_or_si256(_mm256_cmpeq_epi8(u, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(u, _mm256_set1_epi8('T'))));
    mw = ~(uint32_t)_mm256_movemask_epi8(ok);
    const uint32_t b2 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5));
    const uint32_t b1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6));
    cw = _pdep_u64(b1 ^ b2, 0x5555555555555555ULL) | _pdep_u64(b2, 0xAAAAAAAAAAAAAAAAULL);
}

__attribute__((target("avx512f,avx512bw,bmi2"))) inline void pack64_avx512(const uint8_t *s, uint64_t *cw, uint32_t *mw) {
    const __m512i rev = _mm512_broadcast_i32x4(_mm_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0));
    __m512i v = _mm512_loadu_si512(s);
    v = _mm512_shuffle_epi8(v, rev);
    v = _mm512_shuffle_i64x2(v, v, 0x1B);                  // 128-bit lanes reversed: byte j = base 63 - j
    const __m512i u = _mm512_and_si512(v, _mm512_set1_epi8((char)0xDF));
    const uint64_t ok = _mm512_cmpeq_epi8_mask(u, _mm512_set1_epi8('A')) | _mm512_cmpeq_epi8_mask(u, _mm512_set1_epi8('C')) |
                        _mm512_cmpeq_epi8_mask(u, _mm512_set1_epi8('G')) | _mm512_cmpeq_epi8_mask(u, _mm512_set1_epi8('T'));
    const uint64_t b2 = _mm512_test_epi8_mask(v, _mm512_set1_epi8(4)), b1 = _mm512_test_epi8_mask(v, _mm512_set1_epi8(2));
    const uint64_t lo = b1 ^ b2;
    cw[0] = _pdep_u64(lo >> 32, 0x5555555555555555ULL) | _pdep_u64(b2 >> 32, 0xAAAAAAAAAAAAAAAAULL);
    cw[1] = _pdep_u64(lo & 0xFFFFFFFFULL, 0x5555555555555555ULL) | _pdep_u64(b2 & 0xFFFFFFFFULL, 0xAAAAAAAAAAAAAAAAULL);
    mw[0] = ~(uint32_t)(ok >> 32); mw[1] = ~(uint32_t)ok;
}

int isa_level() {   // 2 = AVX-512BW, 1 = AVX2 + BMI2, 0 = scalar
    static const int lvl = [] {
        __builtin_cpu_init();
        int have = 0;
        if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("bmi2")) have = 2;
        else if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2")) have = 1;
        if (const char *ev = getenv("D2G_PACK_ISA")) have = std::max(0, std::min(have, atoi(ev)));   // test knob: force a lower path
        return have;
    }();
    return lvl;
}

__attribute__((target("avx512f,avx512bw,bmi2"))) uint64_t pack_words_avx512(const uint8_t *s, uint64_t nw, uint64_t *codes, uint32_t *mask) {
    uint64_t nz = 0, w = 0;
    for (; w + 2 <= nw; w += 2) { pack64_avx512(s + 32 * w, codes + w, mask + w); nz += (mask[w] != 0) + (mask[w + 1] != 0); }
    for (; w < nw; ++w) { pack32_scalar(s + 32 * w, codes[w], mask[w]); nz += mask[w] != 0; }
    return nz;
}
__attribute__((target("avx2,bmi2"))) uint64_t pack_words_avx2(const uint8_t *s, uint64_t nw, uint64_t *codes, uint32_t *mask) {
    uint64_t nz = 0;
    for (uint64_t w = 0; w < nw; ++w) { pack32_avx2(s + 32 * w, codes[w], mask[w]); nz += mask[w] != 0; }
    return nz;
}
uint64_t pack_words_scalar(const uint8_t *s, uint64_t nw, uint64_t *codes, uint32_t *mask) {
    uint64_t nz = 0;
    for (uint64_t w = 0; w < nw; ++w) { pack32_scalar(s + 32 * w, codes[w], mask[w]); nz += mask[w] != 0; }
    return nz;
}
// 1 iff one of the first `got` bases of the word is invalid (the zero fill past the end of the data is invalid by construction)
inline uint64_t real_invalid(uint32_t m, uint64_t got) { return got >= 32 ? m != 0 : (got ? (m >> (32 - got)) != 0 : 0); }
// nw whole words from contiguous bytes
uint64_t pack_words(const uint8_t *s, uint64_t nw, uint64_t *codes, uint32_t *mask) {
    switch (isa_level()) {
        case 2: return pack_words_avx512(s, nw, codes, mask);
        case 1: return pack_words_avx2(s, nw, codes, mask);
        default: return pack_words_scalar(s, nw, codes, mask);
    }
}

}  // namespace

// words [w0, w1) of the packed form of the n bases at seq; bases beyond n read as invalid
uint64_t pack_contiguous(const char *seq, uint64_t n, uint64_t w0, uint64_t w1, uint64_t *codes, uint32_t *mask) {
    if (w1 <= w0) return 0;
    const uint8_t *s = reinterpret_cast<const uint8_t *>(seq);
    const uint64_t full_end = std::min(w1, n / 32);          // words entirely inside the data
    const uint64_t nfull = full_end > w0 ? full_end - w0 : 0;
    std::atomic<uint64_t> nz{0};
    const uint64_t grain = 1u << 15;                          // 1 Mi bases per task
    const uint64_t ntasks = (nfull + grain - 1) / grain;
    if (ntasks > 1) {
        parallel_for(ntasks, [&](size_t t) {
            const uint64_t a = w0 + t * grain, b = std::min(w0 + nfull, a + grain);
            nz += pack_words(s + 32 * a, b - a, codes + (a - w0), mask + (a - w0));
        });
    } else if (nfull) nz += pack_words(s + 32 * w0, nfull, codes, mask);
    for (uint64_t w = std::max(w0, full_end); w < w1; ++w) {  // ragged end + padding
        uint8_t tmp[32];
        memset(tmp, 0, sizeof tmp);
        if (32 * w < n) memcpy(tmp, s + 32 * w, std::min<uint64_t>(32, n - 32 * w));
        pack32_scalar(tmp, codes[w - w0], mask[w - w0]);
        nz += real_invalid(mask[w - w0], 32 * w < n ? std::min<uint64_t>(32, n - 32 * w) : 0);
    }
    return nz.load();
}

// the concatenation of pieces; words [0, n_words)
uint64_t pack_pieces(const char *const *pieces, const uint64_t *piece_len, uint64_t n_pieces, uint64_t n_words, uint64_t *codes, uint32_t *mask) {
    std::vector<uint64_t> start(n_pieces + 1, 0);
    for (uint64_t i = 0; i < n_pieces; ++i) start[i + 1] = start[i] + piece_len[i];
    const uint64_t total = start[n_pieces];
    std::atomic<uint64_t> nz{0};
    const uint64_t grain = 1u << 15;
    const uint64_t ntasks = (n_words + grain - 1) / grain;
    auto task = [&](size_t t) {
        const uint64_t a = t * grain, b = std::min(n_words, a + grain);
        uint64_t w = a, cnt = 0;
        // piece holding base 32*w
        uint64_t pi = std::upper_bound(start.begin(), start.end(), 32 * w) - start.begin();
        pi = pi ? pi - 1 : 0;
        while (w < b) {
            const uint64_t pos = 32 * w;
            if (pos >= total) {                                            // padding
                for (; w < b; ++w) { codes[w] = 0; mask[w] = 0xFFFFFFFFu; }
                break;
            }
            while (pi + 1 < n_pieces && start[pi + 1] <= pos) ++pi;
            const uint64_t in_piece = start[pi + 1] - pos;                 // bases of this piece from pos on
            if (in_piece >= 32) {
                const uint64_t nw = std::min(b - w, in_piece / 32);
                cnt += pack_words(reinterpret_cast<const uint8_t *>(pieces[pi]) + (pos - start[pi]), nw, codes + w, mask + w);
                w += nw;
            } else {                                                       // a word that straddles pieces (or the end)
                uint8_t tmp[32];
                memset(tmp, 0, sizeof tmp);
                uint64_t q = pi, got = 0, p = pos;
                while (got < 32 && p < total) {
                    while (start[q + 1] <= p) ++q;
                    const uint64_t take = std::min<uint64_t>(32 - got, start[q + 1] - p);
                    memcpy(tmp + got, pieces[q] + (p - start[q]), take);
                    got += take; p += take;
                }
                pack32_scalar(tmp, codes[w], mask[w]);
                cnt += real_invalid(mask[w], got);
                ++w;
            }
        }
        nz += cnt;
    };
    if (ntasks > 1) parallel_for(ntasks, task); else if (ntasks) task(0);
    return nz.load();
}

const char *pack_isa_name() { return isa_level() == 2 ? "avx512bw" : isa_level() == 1 ? "avx2" : "scalar"; }

}  // namespace d2g_host
