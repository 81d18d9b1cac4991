/*
 * d2_oracle.c -- CPU restatement (plain C) of the dashing2 sketch / cmp hot paths.
 *
 * TEST INFRASTRUCTURE ONLY -- see d2_oracle.h.  Citations are file:line under /root/reference.
 * Written from the reference's *behaviour*; no reference code is included or linked.
 *
 * Third-party arithmetic this depends on (same as the reference): glibc libm log/logl of the host,
 * and the constant D2O_OPH_SEED = first output of libstdc++ std::mt19937_64(0x321b919a61cb41f7).
 */
#include "d2_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------------ */
/* hashes                                                                                       */
/* ------------------------------------------------------------------------------------------ */

/* Thomas Wang 64-bit mix. hash.h:42-62 */
uint64_t d2o_wang64(uint64_t key) {
    key = (~key) + (key << 21);
    key ^= key >> 24;
    key = (key + (key << 3)) + (key << 8);
    key ^= key >> 14;
    key = (key + (key << 2)) + (key << 4);
    key ^= key >> 28;
    key += key << 31;
    return key;
}

static uint64_t modinv64(uint64_t a) { /* a odd; Newton iteration mod 2^64 */
    uint64_t x = a;
    for (int i = 0; i < 6; ++i) x *= 2 - a * x;
    return x;
}

/* Inverse of the bijection above (any correct inverse is *the* inverse). Used by oph.h:81-83. */
uint64_t d2o_wang64_inv(uint64_t h) {
    h *= modinv64((1ULL << 31) + 1);
    h ^= h >> 28; h ^= h >> 56;
    h *= modinv64(21);
    h ^= (h >> 14) ^ (h >> 28) ^ (h >> 42) ^ (h >> 56);
    h *= modinv64(265);
    h ^= (h >> 24) ^ (h >> 48);
    /* forward: key = key*(2^21 - 1) - 1 */
    h = (h + 1) * modinv64((1ULL << 21) - 1);
    return h;
}

/* FRev64 = XOR c1, MUL (c2|1), ROTL 31, XOR c3. encoder.h:47; hash.h:688-709,764-826 */
uint64_t d2o_frev64(uint64_t x) {
    x ^= 0x533f8c2151b20f97ULL;
    x *= (0x9a98567ed20c127dULL | 1);
    x = (x << 31) ^ (x >> 33);
    return x ^ 0x691a9d706391077aULL;
}

/* CEHasher = XOR c1, MUL (c2|1), XOR c3. hash.h:858 */
uint64_t d2o_cehash(uint64_t x) {
    x ^= 0x533f8c2151b20f97ULL;
    x *= (0x9a98567ed20c127dULL | 1);
    return x ^ 0x691a9d706391077aULL;
}

/* wy.h:45-59 */
static inline uint64_t wymum(uint64_t x, uint64_t y) {
    u128 l = (u128)x * y;
    return (uint64_t)l ^ (uint64_t)(l >> 64);
}
uint64_t d2o_wyhash64(uint64_t *state) {
    *state += 0x60bee2bee120fc15ULL;
    return wymum(*state ^ 0xe7037ed1a0b428dbULL, *state);
}

/* kmerutil.h:83-90 : reverse the 2-bit groups, complement, right-align. */
uint64_t d2o_revcomp(uint64_t x, int k) {
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    x = (x >> 32) | (x << 32);
    return (~x) >> (64 - 2 * k);
}

static inline uint64_t canonical(uint64_t x, int k) { /* kmerutil.h:137-140 */
    uint64_t rc = d2o_revcomp(x, k);
    return x < rc ? x : rc;
}

/* src/enums.cpp:133-140 */
uint64_t d2o_xormask_for_seed(uint64_t seed) { return seed ? d2o_wang64(seed) : 0; }

/* ------------------------------------------------------------------------------------------ */
/* k-mer / minimizer stream                                                                     */
/* ------------------------------------------------------------------------------------------ */

/* alphabet.h:128 DNA4 ("A,C,G,T", case-insensitive): A0 C1 G2 T3, everything else invalid.
 * (The "U:T" alias is a no-op in the reference's table builder, alphabet.h:50-54.) */
static inline int dna_code(unsigned char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

uint64_t d2o_kmer_positions(uint64_t len, int k) { return len >= (uint64_t)k ? len - k + 1 : 0; }

typedef struct { uint64_t score, el; } elscore_t;
static inline int es_less(elscore_t a, elscore_t b) { /* qmap.h:23-25 */
    return a.score < b.score || (a.score == b.score && a.el < b.el);
}

/* Sliding window minimum by (score, el): equivalent to QueueMap::next_value (qmap.h:79-87), which
 * keeps the last wsz entries in a deque and reads the smallest key of an ordered multiset.
 * Implemented as a ring buffer + linear rescan only when the minimum leaves (oracle: clarity first). */
typedef struct {
    elscore_t *ring; uint32_t wsz, n, head; /* n entries, oldest at head */
} window_t;

static void win_init(window_t *w, uint32_t wsz) {
    w->ring = (elscore_t *)malloc(sizeof(elscore_t) * wsz); w->wsz = wsz; w->n = 0; w->head = 0;
}
static void win_reset(window_t *w) { w->n = 0; w->head = 0; }
static elscore_t win_min(const window_t *w) {
    elscore_t best = w->ring[w->head];
    for (uint32_t i = 1; i < w->n; ++i) {
        elscore_t e = w->ring[(w->head + i) % w->wsz];
        if (es_less(e, best)) best = e;
    }
    return best;
}
/* returns 1 and sets *out when the window is full after the push */
static int win_push(window_t *w, uint64_t el, uint64_t score, uint64_t *out) {
    if (w->n == w->wsz) { w->head = (w->head + 1) % w->wsz; --w->n; } /* pop_front once size > wsz */
    w->ring[(w->head + w->n) % w->wsz] = (elscore_t){score, el};
    ++w->n;
    if (w->n == w->wsz) { *out = win_min(w).el; return 1; }
    return 0;
}

#define EMIT(v) do { if (nout < cap) out[nout] = d2o_wang64((v) ^ xormask); ++nout; } while (0)

uint64_t d2o_hash_stream(const char *seq, uint64_t len, int k, int w, int canon, uint64_t xormask,
                         uint64_t *out, uint64_t cap) {
    uint64_t nout = 0;
    const uint64_t mask = k < 32 ? ((1ULL << (2 * k)) - 1) : ~0ULL; /* rhtraits.h:52-55 */
    const int windowed = w > k;   /* spacer.h:57 w_ = max(c_, w); unwindowed() iff w_ == k_ */
    if (!windowed) {
        /* encoder.h:241-272 (+ canonicalising wrapper :219-232): rolling encode, any invalid base
         * restarts the run, so k-mers containing it are skipped. */
        uint64_t kmer = 0; int filled = 0;
        for (uint64_t pos = 0; pos < len; ++pos) {
            int c = dna_code((unsigned char)seq[pos]);
            if (c < 0) { kmer = 0; filled = 0; continue; }
            kmer = ((kmer << 2) | (uint64_t)c) & mask;
            if (filled < k) ++filled; /* note: reference masks only once filled==k; same value */
            if (filled == k) {
                uint64_t v = canon ? canonical(kmer, k) : kmer;
                EMIT(v);
            }
        }
        return nout;
    }
    window_t win; win_init(&win, (uint32_t)(w - k + 1)); /* encoder.h:141 qmap_(w - c + 1) */
    if (canon) {
        /* encoder.h:212-217,622-628,547-592: every position is (re)encoded; a k-mer holding an
         * invalid base encodes as all-ones and canonicalises to min(~0, revcomp(~0)) = 0, so it
         * ENTERS the window as k-mer 0. One emission per full window. */
        if (len >= (uint64_t)k) {
            for (uint64_t pos = 0; pos + k <= len; ++pos) {
                uint64_t kmer = 0; int bad = 0;
                for (int i = 0; i < k; ++i) {
                    int c = dna_code((unsigned char)seq[pos + i]);
                    if (c < 0) { bad = 1; break; }
                    kmer = (kmer << 2) | (uint64_t)c;
                }
                if (bad) kmer = ~0ULL;
                kmer = canonical(kmer, k);
                uint64_t m;
                if (win_push(&win, kmer, d2o_frev64(kmer), &m) && m != ~0ULL) EMIT(m);
            }
        }
    } else {
        /* encoder.h:274-306: rolling encode; an invalid base (or an accumulator that becomes
         * all-ones) restarts the k-mer run but NOT the window; trailing partial window flushes once. */
        uint64_t kmer = 0; int filled = 0; uint64_t pos = 0;
        while (pos < len) {
            int restart = 0;
            while (filled < k && pos < len) {
                int c = dna_code((unsigned char)seq[pos++]);
                kmer <<= 2;
                kmer |= (uint64_t)(int64_t)c; /* -1 sign-extends to all ones */
                if (kmer == ~0ULL) { restart = 1; break; }
                ++filled;
            }
            if (restart) { kmer = 0; filled = 0; continue; }
            if (filled == k) {
                kmer &= mask;
                uint64_t m;
                if (win_push(&win, kmer, d2o_frev64(kmer), &m) && m != ~0ULL) EMIT(m);
                --filled;
            }
        }
        if (win.n > 0 && win.n < win.wsz) EMIT(win_min(&win).el); /* encoder.h:304-305 */
    }
    (void)win_reset;
    free(win.ring);
    return nout;
}
#undef EMIT

/* ------------------------------------------------------------------------------------------ */
/* Protein alphabets (--protein/--protein20, --protein14, --protein6, --protein8; src/options.h:328-331, canon = false).
 * Tables: bonsai alphabet.h:107-120 built by make_lut (:30-58): comma-separated groups get codes 0, 1, ..., both cases; the "OU:KC"
 * alias is a no-op because it indexes the table by CODE, not by character (same as "U:T" for DNA), so O and U stay invalid.
 * Rolling encode (encoder.h:241-306): min = (min * mul) | code -- an OR even when mul is not a power of two -- then
 * min %= mul^k (PROTEIN20 / 14 / 6, rhtraits.h:59-62) or min &= (1 << k) - 1 (the 3-bit alphabet: k bits, not 3k; rhtraits.h:57-58).
 * The reduced value is carried on as the rolling state. */
/* ------------------------------------------------------------------------------------------ */
static void alpha_lut(const char *groups, int8_t lut[256]) {
    memset(lut, -1, 256);
    int id = 0;
    for (const char *p = groups; *p; ++p) {
        if (*p == ',') { ++id; continue; }
        lut[(unsigned char)(*p | 32)] = (int8_t)id; lut[(unsigned char)(*p & 0xdf)] = (int8_t)id;
    }
}
/* alphabet: 20 = AMINO20, 14 = SE-B(14), 6 = SE-B(6), 8 = SE-B(8) with the 3-bit encoding */
uint64_t d2o_hash_stream_protein(const char *seq, uint64_t len, int k, int w, int alphabet, uint64_t xormask, uint64_t *out, uint64_t cap) {
#define EMIT(v) do { if (nout < cap) out[nout] = d2o_wang64((v) ^ xormask); ++nout; } while (0)
    uint64_t nout = 0;
    int8_t lut[256];
    alpha_lut(alphabet == 20 ? "A,C,D,E,F,G,H,I,K,L,M,N,P,Q,R,S,T,V,W,Y" : alphabet == 14 ? "A,C,D,EQ,FY,G,H,IV,KR,LM,N,P,ST,W"
              : alphabet == 6 ? "AST,CP,DHNEKQR,FWY,G,ILMV" : "AST,C,DHN,EKQR,FWY,G,ILMV,P", lut);
    const uint64_t mul = (uint64_t)alphabet;
    uint64_t mask;
    if (alphabet == 8) mask = ~0ULL >> (64 - k);
    else mask = (uint64_t)pow((double)alphabet, (double)k);            /* rhtraits.h:59-61: std::pow narrowed to the k-mer type */
    const int windowed = w > k;
    window_t win; win_init(&win, windowed ? (uint32_t)(w - k + 1) : 1);
    uint64_t kmer = 0, pos = 0; int filled = 0;
    while (pos < len) {
        int restart = 0;
        while (filled < k && pos < len) {
            const int8_t nv = lut[(unsigned char)seq[pos++]];
            if (!windowed) {
                if (nv == -1) { restart = 1; break; }
                kmer = (kmer * mul) | (uint64_t)nv;
            } else {
                kmer *= mul;
                kmer |= (uint64_t)(int64_t)nv;                         /* -1 sign-extends to all ones, encoder.h:285 */
                if (kmer == ~0ULL) { restart = 1; break; }
            }
            ++filled;
        }
        if (restart) { kmer = 0; filled = 0; continue; }
        if (filled == k) {
            if (alphabet == 8) kmer &= mask; else kmer %= mask;
            if (!windowed) EMIT(kmer);
            else { uint64_t m; if (win_push(&win, kmer, d2o_frev64(kmer), &m) && m != ~0ULL) EMIT(m); }
            --filled;
        }
    }
    if (windowed && win.n > 0 && win.n < win.wsz) EMIT(win_min(&win).el);
    free(win.ring);
    return nout;
#undef EMIT
}

/* ------------------------------------------------------------------------------------------ */
/* k > 32: RollingHasher<uint64_t> over CyclicHash (bonsai encoder.h:644-865, rollinghash/cyclichash.h,
 * rollinghash/characterhash.h).  Word size 64, so every rotation is a plain 64-bit rotate.  Character tables:
 * 256 draws of WyRand<uint64_t> seeded with (seed1 ^ seed2) for the forward hasher and (seed2 * seed1) ^ (seed2 ^ seed1)
 * for the reverse-complement hasher, truncated to 32 bits (CyclicHash::seed -> CharacterHash::seed(uint32_t)), with the
 * constructor defaults seed1 = 1337, seed2 = 137 (encoder.h:673, src/d2.h:136). */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t tab[256]; uint64_t h; int myr; } cyc_t;
static inline uint64_t rotl64(uint64_t x, int r) { r &= 63; return r ? (x << r) | (x >> (64 - r)) : x; }
static void cyc_init(cyc_t *c, int k, uint64_t s1, uint64_t s2) {
    uint64_t state = (uint32_t)(s1 ^ s2);          /* CyclicHash::seed: s1 ^= s2; CharacterHash::seed takes uint32_t */
    if (!state) state = 1337;                      /* WyRand(seed): seed ? seed : 1337 (wy.h:112) */
    for (int i = 0; i < 256; ++i) c->tab[i] = d2o_wyhash64(&state);   /* clear_hashvalues: next &= all-ones, never above maxval */
    c->h = 0; c->myr = k % 64;
}
static inline void cyc_eat(cyc_t *c, uint8_t in) { c->h = rotl64(c->h, 1) ^ c->tab[in]; }                                   /* cyclichash.h:118-121 */
static inline void cyc_update(cyc_t *c, uint8_t out, uint8_t in) { c->h = rotl64(c->h, 1) ^ rotl64(c->tab[out], c->myr) ^ c->tab[in]; } /* :101-108 */
static inline void cyc_reverse_update(cyc_t *c, uint8_t out, uint8_t in) {                                                  /* :110-116 */
    c->h ^= rotl64(c->tab[out], c->myr) ^ c->tab[in];
    c->h = (c->h >> 1) | ((c->h & 1) << 63);
}
static inline uint8_t rc_code(unsigned char ch) { const int c = dna_code(ch); return c < 0 ? (uint8_t)255 : (uint8_t)(3 - c); } /* cstr_rc_lut */

/* RollingHasher::for_each_hash of ONE record (encoder.h:692-797), each value through maskfn.  Canonical: min(forward, reverse
 * complement) per position, or -- windowed -- BOTH hashes pushed into the window of w-k+1 entries (two pushes per position, the
 * window is not reset at an N), tail flush of a partially filled window.  An N skips k further bases (the `fixup` jump) and
 * ends the record when fewer than 2k bases remain (canonical path only). */
#define EMIT(v) do { if (nout < cap) out[nout] = d2o_wang64((v) ^ xormask); ++nout; } while (0)
uint64_t d2o_hash_stream_rolling(const char *seq, uint64_t len, int k, int w, int canon, uint64_t xormask,
                                 uint64_t *out, uint64_t cap) {
    uint64_t nout = 0;
    const unsigned char *s = (const unsigned char *)seq;
    const uint64_t l = len, K = (uint64_t)k;
    if (l < K) return 0;
    cyc_t fw, rc;
    cyc_init(&fw, k, 1337, 137);
    cyc_init(&rc, k, 137ULL * 1337ULL, 137ULL ^ 1337ULL);
    const int windowed = w > k;
    window_t win; win_init(&win, windowed ? (uint32_t)(w - k + 1) : 1);
    uint64_t i = 0, nf = 0, mn;
    if (canon) {
        for (;;) {
            /* fill */
            int ended = 0;
            for (; nf < K && i < l; ++i) {
                const int v1 = dna_code(s[i]);
                if (v1 < 0) {
                    if (i + 2 * K >= l) { ended = 1; break; }
                    i += K; nf = 0; fw.h = 0; rc.h = 0;
                } else { cyc_eat(&fw, (uint8_t)v1); cyc_eat(&rc, rc_code(s[i - nf + K - 1])); ++nf; }
            }
            if (ended || nf < K) break;
            if (windowed) { if (win_push(&win, fw.h, d2o_frev64(fw.h), &mn)) EMIT(mn); if (win_push(&win, rc.h, d2o_frev64(rc.h), &mn)) EMIT(mn); }
            else EMIT(fw.h < rc.h ? fw.h : rc.h);
            int hitn = 0;
            for (; i < l; ++i) {
                const int v1 = dna_code(s[i]);
                if (v1 < 0) { hitn = 1; break; }
                cyc_update(&fw, (uint8_t)dna_code(s[i - K]), (uint8_t)v1);
                cyc_reverse_update(&rc, rc_code(s[i]), rc_code(s[i - K]));
                if (windowed) { if (win_push(&win, fw.h, d2o_frev64(fw.h), &mn)) EMIT(mn); if (win_push(&win, rc.h, d2o_frev64(rc.h), &mn)) EMIT(mn); }
                else EMIT(fw.h < rc.h ? fw.h : rc.h);
            }
            if (!hitn) break;
            /* goto fixup: the same jump as in the fill loop, then the fill loop's ++i */
            if (i + 2 * K >= l) break;
            i += K; nf = 0; fw.h = 0; rc.h = 0; ++i;
        }
        if (windowed && win.n > 0 && win.n < win.wsz) EMIT(win_min(&win).el);
    } else {
        for (;;) {
            for (; nf < K && i < l; ++i) {
                const int v1 = dna_code(s[i]);
                if (v1 < 0) { i += K; nf = 0; fw.h = 0; }
                else { cyc_eat(&fw, (uint8_t)v1); ++nf; }
            }
            if (nf < K) { free(win.ring); return nout; }      /* "All failed": returns without the tail flush (encoder.h:776) */
            if (windowed) { if (win_push(&win, fw.h, d2o_frev64(fw.h), &mn)) EMIT(mn); } else EMIT(fw.h);
            int hitn = 0;
            for (; i < l; ++i) {
                if (dna_code(s[i]) < 0) { hitn = 1; break; }
                cyc_update(&fw, (uint8_t)dna_code(s[i - K]), (uint8_t)dna_code(s[i]));
                if (windowed) { if (win_push(&win, fw.h, d2o_frev64(fw.h), &mn)) EMIT(mn); } else EMIT(fw.h);
            }
            if (!hitn) break;
            i += K; nf = 0; fw.h = 0; ++i;
        }
        if (windowed && win.n > 0 && win.n < win.wsz) EMIT(win_min(&win).el);
    }
    free(win.ring);
    return nout;
}
#undef EMIT

/* ------------------------------------------------------------------------------------------ */
/* One-permutation MinHash (LazyOnePermSetSketch<uint64_t>)                                     */
/* ------------------------------------------------------------------------------------------ */

uint32_t d2o_opmh_m(uint32_t sketchsize) { return sketchsize + (sketchsize & 1u); } /* oph.h:145 */

void d2o_opmh_reset(uint64_t *regs, double *counts, uint32_t m) {
    for (uint32_t i = 0; i < m; ++i) { regs[i] = ~0ULL; if (counts) counts[i] = 0.; }
}

/* DHasher(x) = Wang(x ^ seed_ ^ 0x533f8c2151b20f97). oph.h:44-53,55-71 */
static inline uint64_t dhash(uint64_t x) { return d2o_wang64(x ^ D2O_OPH_SEED ^ 0x533f8c2151b20f97ULL); }
static inline uint64_t dhash_inv(uint64_t h) { return d2o_wang64_inv(h) ^ 0x533f8c2151b20f97ULL ^ D2O_OPH_SEED; }

void d2o_opmh_update(uint64_t *regs, double *counts, uint32_t m, const uint64_t *hv, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t id = dhash(hv[i]);
        const uint32_t idx = (uint32_t)id % m; /* Schismatic<uint32_t>::mod truncates; div.h:256-262 */
        if (regs[idx] > id) { regs[idx] = id; if (counts) counts[idx] = 1.; }
        else if (counts) counts[idx] += (regs[idx] == id);
    }
}

/* oph.h:188-205: a candidate is promoted once seen mincount times; per-bucket candidate multiset. */
typedef struct { uint64_t id; uint32_t cnt; } pot_t;
typedef struct { pot_t *v; uint32_t n, cap; } potvec_t;
void d2o_opmh_update_mincount(uint64_t *regs, double *counts, uint32_t m, const uint64_t *hv,
                              uint64_t n, double mincount) {
    potvec_t *pots = (potvec_t *)calloc(m, sizeof(potvec_t));
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t id = dhash(hv[i]);
        const uint32_t idx = (uint32_t)id % m;
        if (regs[idx] > id) {
            potvec_t *p = &pots[idx];
            uint32_t j = 0;
            for (; j < p->n; ++j) if (p->v[j].id == id) break;
            if (j == p->n) {
                if (p->n == p->cap) { p->cap = p->cap ? p->cap * 2 : 4; p->v = (pot_t *)realloc(p->v, p->cap * sizeof(pot_t)); }
                p->v[p->n++] = (pot_t){id, 1};
            } else ++p->v[j].cnt;
            if (p->v[j].cnt >= mincount) {
                regs[idx] = id; if (counts) counts[idx] = p->v[j].cnt;
                uint32_t o = 0;
                for (uint32_t t = 0; t < p->n; ++t) if (p->v[t].id < id) p->v[o++] = p->v[t];
                p->n = o;
            }
        } else if (counts) counts[idx] += (regs[idx] == id);
    }
    for (uint32_t i = 0; i < m; ++i) free(pots[i].v);
    free(pots);
}

double d2o_opmh_card(const uint64_t *regs, uint32_t m) {
    long double sum = 0.L;
    for (uint32_t i = 0; i < m; ++i) sum = sum + (long double)regs[i] * 0x1p-64L;
    if (!sum) return INFINITY;
    return (double)((long double)m * ((long double)m / sum));
}

void d2o_opmh_sigs(const uint64_t *regs, uint32_t m, double *out) {
    uint64_t nempty = 0;
    for (uint32_t i = 0; i < m; ++i) nempty += regs[i] == ~0ULL;
    const double muld = -1.0 / (double)((uint64_t)m - nempty);
    const long double mul = muld;
    for (uint32_t i = 0; i < m; ++i) {
        const uint64_t x = regs[i];
        if (x == ~0ULL || x == 0) { out[i] = 0.; continue; }
        const uint64_t rem = ~0ULL - x + 1;
        out[i] = (double)(mul * logl(0x1p-64L * (long double)rem));
    }
}

void d2o_opmh_ids(const uint64_t *regs, uint32_t m, uint64_t *out) {
    for (uint32_t i = 0; i < m; ++i) out[i] = dhash_inv(regs[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* Full continuous SetSketch  (CSetSketch<double>)                                             */
/* ------------------------------------------------------------------------------------------ */

void d2o_css_reset(double *regs, uint32_t m) {
    for (uint64_t i = 0; i < 2ULL * m - 1; ++i) regs[i] = DBL_MAX; /* setsketch.h:139-142 */
}

/* max-tree update, setsketch.h:151-166 */
static int mvt_update(double *d, uint32_t m, uint64_t index, double x) {
    const uint64_t sz = 2ULL * m - 1;
    if (!(x < d[index])) return 0;
    for (;;) {
        d[index] = x;
        if ((index = m + (index >> 1)) >= sz) break;
        const uint64_t lhi = (index - m) << 1, rhi = lhi + 1;
        x = d[lhi] > d[rhi] ? d[lhi] : d[rhi];
        if (x >= d[index]) break;
    }
    return 1;
}

/* flog.h:14-20 */
static inline double flog_d(double x) {
    uint64_t yi; memcpy(&yi, &x, 8);
    return fma((double)yi, 1.539095918623324e-16, -709.0895657128241);
}

/* Lazy Fisher-Yates (fy.h:14-66) with the WyRand<uint32_t,2> stream (wy.h:96-145). */
typedef struct {
    uint32_t *g, *v; uint32_t n, i, c;
    uint64_t state; uint64_t buf[2]; unsigned off;
} lazyshuf_t;
static uint32_t ls_rng32(lazyshuf_t *s) {
    if (s->off + 4 > 16) { s->buf[0] = d2o_wyhash64(&s->state); s->buf[1] = d2o_wyhash64(&s->state); s->off = 0; }
    uint32_t r; memcpy(&r, (const unsigned char *)s->buf + s->off, 4); s->off += 4;
    return r;
}
static uint32_t ls_step(lazyshuf_t *s) {
    const uint32_t samp = ls_rng32(s) % (s->n - s->i);
    const uint32_t j = s->i + samp;
    const uint32_t k = s->v[j] == s->c ? s->g[j] : j;
    s->g[j] = s->v[s->i] == s->c ? s->g[s->i] : s->i;
    s->v[j] = s->c;
    if (++s->i == s->n) s->i = 0;
    return k;
}

void d2o_css_update(double *regs, uint32_t m, const uint64_t *hv, uint64_t n, uint64_t *ids) {
    lazyshuf_t ls; ls.n = m; ls.c = 0;
    ls.g = (uint32_t *)calloc(m, 4); ls.v = (uint32_t *)calloc(m, 4);
    const double INVMUL64 = 0x1p-64;
    for (uint64_t e = 0; e < n; ++e) {
        const uint64_t id = hv[e];
        double carry = 0.;
        uint64_t hid = id;
        uint64_t rv = d2o_cehash(id ^ 0xb2069fc679a8da0bULL);
        double mv = regs[2ULL * m - 2];
        double tv = (double)rv * INVMUL64;
        const double bv0 = -1. / m;
        if (bv0 * flog_d(tv) * .7 > mv) continue;
        double ev = bv0 * log(tv);
        if (ev > mv) continue;
        ls.i = 0; ++ls.c; ls.state = rv; ls.off = 16; /* reset(); seed(rv) */
        uint64_t bi = 1;
        for (;;) {
            const uint32_t idx = ls_step(&ls);
            if (mvt_update(regs, m, idx, ev)) { if (ids) ids[idx] = id; mv = regs[2ULL * m - 2]; }
            if (bi == m) break;
            rv = d2o_wyhash64(&hid);
            const double bv = -(1. / (double)(m - bi)); ++bi; /* getbeta, setsketch.h:300-302 */
            const double nv = (double)rv * INVMUL64;
            if (bv * flog_d(nv) * .7 + ev > mv) break;
            /* kahan.h:8-13 */
            double inc = fma(bv, log(nv), -carry); /* GCC contracts `bv*log(nv) - carry` at -O3 -mfma */
            const double tmp = ev + inc;
            carry = (tmp - ev) - inc;
            ev = tmp;
            if (ev > mv) break;
        }
    }
    free(ls.g); free(ls.v);
}

double d2o_css_card(const double *regs, uint32_t m) {
    double s = 0.;
    for (uint32_t i = 0; i < m; ++i) s += regs[i];
    return m / s;
}

/* ------------------------------------------------------------------------------------------ */
/* densify                                                                                      */
/* ------------------------------------------------------------------------------------------ */

uint64_t d2o_densify(double *sig, uint64_t *kmers, uint64_t S) {
    uint64_t nz = 0;
    for (uint64_t i = 0; i < S; ++i) nz += sig[i] == 0.;
    if (nz == S) return S;
    double *tmp = (double *)malloc(S * sizeof(double));
    memcpy(tmp, sig, S * sizeof(double));
    uint64_t ne = 0;
    for (uint64_t i = 0; i < S; ++i) {
        if (sig[i] != 0.) continue;
        ++ne;
        uint64_t rng = i + 0x5bf2b8bdf07c06cULL, j;
        do { j = d2o_wyhash64(&rng) % S; } while (sig[j] == 0.);
        tmp[i] = sig[j];
        if (kmers) kmers[i] = kmers[j];
    }
    memcpy(sig, tmp, S * sizeof(double));
    free(tmp);
    return ne;
}

/* ------------------------------------------------------------------------------------------ */
/* compare                                                                                      */
/* ------------------------------------------------------------------------------------------ */

void d2o_count_gtlt(const double *a, const double *b, uint64_t n, uint64_t *gt, uint64_t *lt) {
    uint64_t g = 0, l = 0;
    for (uint64_t i = 0; i < n; ++i) { g += a[i] > b[i]; l += b[i] > a[i]; }
    *gt = g; *lt = l;
}

/* sketch::eq::count_eq<double> (count_eq.h:40-45): the IEEE `==` on RegT = double, also when the k-mer ids are viewed as doubles
 * (cmp_core.cpp:501-506) -- NaN bit patterns never match, -0 equals +0 */
uint64_t d2o_count_eq(const uint64_t *a, const uint64_t *b, uint64_t n) {
    const double *x = (const double *)a, *y = (const double *)b;
    uint64_t e = 0;
    for (uint64_t i = 0; i < n; ++i) e += x[i] == y[i];
    return e;
}

static inline long double ldmax(long double a, long double b) { return a < b ? b : a; } /* std::max */
static inline long double ldmin(long double a, long double b) { return b < a ? b : a; } /* std::min */

float d2o_finalize(uint64_t c0, uint64_t c1, uint64_t S, double lhc, double rhc, int measure, int k,
                   int cmp_kind) {
    long double ret;
    const long double lhcard = lhc, rhcard = rhc;
    const long double invdenom = 1.L / S;
    const double poisson_mult = -1. / (k > 1 ? k : 1);
    if (cmp_kind == 0) { /* cmp_core.cpp:458-494 */
        const long double alpha = c0 * invdenom, beta = c1 * invdenom;
        long double eq = (1. - alpha - beta);
        const long double ucard = ldmax((lhcard + rhcard) / (2.L - alpha - beta), 0.L);
        if (eq <= 0.) return measure != D2O_POISSON_LLR ? 0.f : (float)DBL_MAX;
        if (eq <= 1e-15L) eq = 0;
        const float isz = (float)(ucard * eq), sim = (float)eq;
        switch (measure) {
            case D2O_SIMILARITY: ret = sim; break;
            case D2O_INTERSECTION: ret = isz; break;
            case D2O_CONTAINMENT: ret = isz / rhcard; break;
            case D2O_SYMMETRIC_CONTAINMENT: ret = isz / ldmin(lhcard, rhcard); break;
            case D2O_POISSON_LLR:
                ret = sim ? (double)(log(2. * sim / (1. + sim)) * poisson_mult) : (double)INFINITY; break;
            case D2O_UNION_SIZE: ret = lhcard + rhcard - isz; break;
            default: ret = -1.f;
        }
    } else { /* cmp_core.cpp:495-517 */
        ret = invdenom * c0;
        if (measure == D2O_INTERSECTION) ret *= ldmax((lhcard + rhcard) / (1.L + ret), 0.L);
        else if (measure == D2O_SYMMETRIC_CONTAINMENT) ret *= ldmax((lhcard + rhcard) / (1.L + ret), 0.L) / ldmin(lhcard, rhcard);
        else if (measure == D2O_CONTAINMENT) ret *= ldmax((lhcard + rhcard) / (1.L + ret), 0.L) / lhcard;
        else if (measure == D2O_POISSON_LLR) ret = ret ? (double)(logl(2. * ret / (1. + ret)) * poisson_mult) : (double)INFINITY;
        else if (measure == D2O_UNION_SIZE) {
            const long double isz = ret * ldmax((lhcard + rhcard) / (1.L + ret), 0.L);
            ret = lhcard + rhcard - isz;
        }
    }
    if (isnan(ret) || isinf(ret)) ret = LDBL_MAX; /* cmp_core.cpp:573 */
    return (float)ret;
}

float d2o_compare(const double *a, const double *b, uint64_t S, double lhc, double rhc, int measure,
                  int k, int cmp_kind) {
    uint64_t c0 = 0, c1 = 0;
    if (cmp_kind == 0) d2o_count_gtlt(a, b, S, &c0, &c1);
    else c0 = d2o_count_eq((const uint64_t *)a, (const uint64_t *)b, S);
    return d2o_finalize(c0, c1, S, lhc, rhc, measure, k, cmp_kind);
}

void d2o_allpairs_symmetric(const double *regs, const double *cards, uint64_t n, uint64_t S,
                            int measure, int k, int cmp_kind, float *out) {
    for (uint64_t i = 0; i < n; ++i)
        for (uint64_t j = i + 1; j < n; ++j)
            *out++ = d2o_compare(regs + i * S, regs + j * S, S, cards[i], cards[j], measure, k, cmp_kind);
}
void d2o_allpairs_asymmetric(const double *regs, const double *cards, uint64_t n, uint64_t S,
                             int measure, int k, int cmp_kind, float *out) {
    for (uint64_t i = 0; i < n; ++i)
        for (uint64_t j = 0; j < n; ++j)
            *out++ = d2o_compare(regs + i * S, regs + j * S, S, cards[i], cards[j], measure, k, cmp_kind);
}
void d2o_panel(const double *regs, const double *cards, uint64_t nf, uint64_t nq, uint64_t S,
               int measure, int k, int cmp_kind, float *out) {
    for (uint64_t i = 0; i < nf; ++i)
        for (uint64_t j = 0; j < nq; ++j)
            *out++ = d2o_compare(regs + i * S, regs + (nf + j) * S, S, cards[i], cards[nf + j], measure, k, cmp_kind);
}

/* ------------------------------------------------------------------------------------------ */
/* compressed registers: make_compressed (src/cmp_core.cpp:209-322) and the compressed branch   */
/* of compare() (src/cmp_core.cpp:362-449).  Quantised registers are returned as doubles (exact */
/* small integers) so that every counting routine above applies unchanged.                      */
/* ------------------------------------------------------------------------------------------ */
static int64_t ld_to_i64_x86(long double v) { /* static_cast<int64_t>(long double) as x86 executes it: out of range / NaN -> INT64_MIN */
    if (!(v > -9223372036854775809.0L && v < 9223372036854775808.0L)) return INT64_MIN;
    return (int64_t)v;
}
static uint64_t reg2sig(double x) { /* cmp_core.cpp:19-37, sizeof(T) == 8 */
    uint64_t v; memcpy(&v, &x, 8);
    return d2o_wang64(v ^ 0xa3407fb23cd20efULL);
}
/* fd in {1, 2, 4}; truncation <= 0: setsketch quantisation with (a, b) (fitted from the data when a or b <= 0, cmp_core.cpp:250-266;
 * setsketch.cpp:7-10), falling back to b-bit when the fit degenerates (:267-270); truncation > 0: b-bit signatures (:293-321).
 * Returns the truncation method actually used; *a, *b hold the parameters used. */
int d2o_make_compressed(const double *sigs, const uint64_t *kmers, uint64_t nsigs, double fd, int truncation,
                        long double *a_io, long double *b_io, double *out) {
    long double a = *a_io, b = *b_io;
    if (truncation <= 0) {
        const long double q = fd == 1. ? 254.3 : fd == 2. ? 65534 : fd == 4. ? 4294967294 : 15.4; /* double literals, as in the reference */
        if (a <= 0. || b <= 0.) {
            double minreg = DBL_MAX, maxreg = -DBL_MAX;
            for (uint64_t i = 0; i < nsigs; ++i) {
                const double v = sigs[i];
                if (v <= 0 || v == DBL_MAX) continue;
                if (v < minreg) minreg = v;
                if (v > maxreg) maxreg = v;
            }
            long double mx = minreg, mn = maxreg;          /* optimal_parameters(minreg, maxreg, q): named (maxreg, minreg), swapped if needed */
            if (mx < mn) { const long double t = mx; mx = mn; mn = t; }
            b = expl(logl(mx / mn) / q);
            a = mx / b;
        }
        if (a == 0. || isinf(b)) truncation = 1;
        else {
            *a_io = a; *b_io = b;
            const long double logbinv = 1.L / log1pl(b - 1.L);
            const int64_t top = (int64_t)(q + 1);
            for (uint64_t i = 0; i < nsigs; ++i) {
                const long double sub = 1.L - logl((long double)sigs[i] / a) * logbinv;
                int64_t isub = ld_to_i64_x86(sub);
                if (isub > top) isub = top;
                if (isub < 0) isub = 0;
                out[i] = (double)isub;
            }
            return 0;
        }
    }
    const int shift = fd == 1. ? 58 : fd == 2. ? 48 : fd == 4. ? 32 : 0;
    for (uint64_t i = 0; i < nsigs; ++i) {
        const uint64_t sig = (kmers ? d2o_wang64(kmers[i]) : reg2sig(sigs[i])) >> shift;
        out[i] = (double)sig;
    }
    return 1;
}

static long double g_b(long double b, long double arg) { return (1.L - powl(b, -arg)) / (1.L - 1.L / b); } /* cmp_core.cpp:323-325 */

float d2o_finalize_compressed(uint64_t c0, uint64_t c1, uint64_t S, double lhc, double rhc, int measure, int k,
                              int bbit, double fd, long double b) {
    long double ret;
    const long double lhcard = lhc, rhcard = rhc;
    const long double invdenom = 1.L / S;
    const double poisson_mult = -1. / (k > 1 ? k : 1);
    if (bbit) { /* cmp_core.cpp:406-424 */
        const long double b2pow = -ldexpl(1.L, -(int)(fd * 8.));
        ret = ldmax(0.L, fmal((long double)c0, invdenom, b2pow) / (1.L + b2pow));
        if (measure == D2O_INTERSECTION || measure == D2O_UNION_SIZE) {
            const long double isz = ldmax((lhcard + rhcard) / (2.L - (1.L - ret)), 0.L);
            ret = measure == D2O_INTERSECTION ? isz : lhcard + rhcard - isz;
        } else if (measure == D2O_CONTAINMENT) ret = ldmax((lhcard + rhcard) / (2.L - (1.L - ret)), 0.L) * ret / lhcard;
        else if (measure == D2O_POISSON_LLR) ret = ret ? (double)(logl(2. * ret / (1. + ret)) * poisson_mult) : (double)INFINITY;
        else if (measure == D2O_SYMMETRIC_CONTAINMENT) ret = ldmax((lhcard + rhcard) / (2.L - (1.L - ret)), 0.L) * ret / ldmin(lhcard, rhcard);
    } else { /* cmp_core.cpp:425-448 */
        long double alpha = c0 * invdenom, beta = c1 * invdenom, mu;
        alpha = g_b(b, alpha); beta = g_b(b, beta);            /* fd < sizeof(RegT) always holds here */
        if (alpha + beta >= 1.) mu = lhcard + rhcard;
        else mu = ldmax((lhcard + rhcard) / (2.L - alpha - beta), 0.L);
        ret = ldmax(1.L - (alpha + beta), 0.L);
        switch (measure) {
            case D2O_INTERSECTION: ret *= mu; break;
            case D2O_UNION_SIZE: ret = lhcard + rhcard - (ret * mu); break;
            case D2O_CONTAINMENT: ret = ret * mu / lhcard; break;
            case D2O_SYMMETRIC_CONTAINMENT: ret = (ret * mu) / ldmin(lhcard, rhcard); break;
            case D2O_POISSON_LLR: ret = ret ? (double)(logl(2. * ret / (1. + ret)) * poisson_mult) : (double)INFINITY; break;
            default: ;
        }
    }
    if (isnan(ret) || isinf(ret)) ret = LDBL_MAX; /* cmp_core.cpp:573 */
    return (float)ret;
}

/* shape: 0 symmetric, 1 asymmetric, 2 panel (rows = first n - nq, columns = last nq) */
void d2o_allpairs_compressed(const double *cregs, const double *cards, uint64_t n, uint64_t nq, uint64_t S, int shape,
                             int measure, int k, int bbit, double fd, long double b, float *out) {
    const uint64_t nr = shape == 2 ? n - nq : n, c0 = shape == 2 ? n - nq : 0;
    for (uint64_t i = 0; i < nr; ++i)
        for (uint64_t j = (shape == 0 ? i + 1 : c0); j < n; ++j) {
            uint64_t x = 0, y = 0;
            if (bbit) x = d2o_count_eq((const uint64_t *)(cregs + i * S), (const uint64_t *)(cregs + j * S), S);
            else d2o_count_gtlt(cregs + i * S, cregs + j * S, S, &x, &y);
            *out++ = d2o_finalize_compressed(x, y, S, cards[i], cards[j], measure, k, bbit, fd, b);
        }
}

/* ------------------------------------------------------------------------------------------ */
/* exact counting + ProbMinHash3 + BagMinHash2                                                 */
/* ------------------------------------------------------------------------------------------ */
static int cmp_u64(const void *a, const void *b) { uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return x < y ? -1 : x > y; }
uint64_t d2o_count_exact(uint64_t *hv, uint64_t n, uint64_t *keys, double *counts) {
    qsort(hv, n, 8, cmp_u64);
    uint64_t nd = 0;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i; while (j < n && hv[j] == hv[i]) ++j;
        keys[nd] = hv[i]; counts[nd] = (double)(int32_t)(j - i); ++nd; i = j;
    }
    return nd;
}

/* Counter::add with a count sketch, src/counter.h:68-77: bucket Wang(x) % cssize gets +1, or -1 when the top bit of Wang(x) is clear
 * (Counter::ct() reports COUNTSKETCH_COUNTING whenever the table exists, :18); float accumulation (exact: integers below 2^24).
 * Counter::finalize(Sketch&), :131-137: every bucket i with |count| >= threshold feeds dst.update(i, |count|) -- the element id is the bucket
 * index.  Buckets with weight 0 (threshold 0) are no-ops in both weighted sketches and are dropped here.  keys/counts need cssize slots;
 * returns the number of elements written (ascending bucket index). */
uint64_t d2o_count_sketch(const uint64_t *hv, uint64_t n, uint64_t cssize, double threshold, uint64_t *keys, double *counts) {
    float *cs = (float *)calloc(cssize, sizeof(float));
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t h = d2o_wang64(hv[i]);
        cs[h % cssize] += (h & 0x8000000000000000ULL) == 0 ? -1.f : 1.f;
    }
    uint64_t nd = 0;
    for (uint64_t i = 0; i < cssize; ++i) {
        const float v = fabsf(cs[i]);
        if ((double)v >= threshold && v > 0.f) { keys[nd] = i; counts[nd] = (double)v; ++nd; }
    }
    free(cs);
    return nd;
}

/* --parse-by-seq cardinality of one record's set sketch, src/fastxsketchbyseq.cpp:393-430: a NaN estimate becomes 0 (:398-402);
 * an estimate below 10 * sketchsize is replaced by the exact number of distinct maskfn'd k-mers (minimizers) of the record,
 * which the reference collects in a flat_hash_set by walking the record a second time (:405-422).  hv is sorted in place. */
double d2o_byseq_cardinality(double estimate, uint64_t sketchsize, uint64_t *hv, uint64_t n) {
    if (estimate != estimate) estimate = 0.;
    if (!(estimate < 10. * (double)sketchsize)) return estimate;
    qsort(hv, n, 8, cmp_u64);
    uint64_t nd = 0;
    for (uint64_t i = 0; i < n; ++i) nd += (i == 0 || hv[i] != hv[i - 1]);
    return (double)nd;
}

/* value tree of bmh.h:53-91 (update returns -1 on equality, 1 when lowered, 0 otherwise) */
static int bmh_mvt_update(double *d, uint32_t m, uint64_t index, double x) {
    const uint64_t sz = 2ULL * m - 1;
    if (x == d[index]) return -1;
    if (x < d[index]) {
        do {
            d[index] = x;
            index = m + (index >> 1);
            if (index >= sz) break;
            const uint64_t lhi = (index - m) << 1, rhi = lhi + 1;
            x = d[lhi] > d[rhi] ? d[lhi] : d[rhi];
        } while (x < d[index]);
        return 1;
    }
    return 0;
}

void d2o_pmh_reset(double *regs, uint64_t *ids, uint32_t m) {
    for (uint64_t i = 0; i < 2ULL * m - 1; ++i) regs[i] = DBL_MAX;
    if (ids) memset(ids, 0, 8ULL * m);
}

typedef struct { double lambda, c1, c2, c3, c4; } texp_t;
static texp_t texp_constants(uint32_t m) { /* bmh.h:490-502, long double then narrowed */
    const long double lambda = log1pl(1.L / (m - 1));
    const long double c1 = (expl(lambda) - 1.L) / lambda;
    const long double c2 = logl(2.L / (1.L + expl(-lambda))) / lambda;
    const long double c3 = (1.L - expl(-lambda)) / lambda;
    const long double c4 = c1 * lambda;
    texp_t t = {(double)lambda, (double)c1, (double)c2, (double)c3, (double)c4};
    return t;
}
static double texp_sample(uint64_t rngstate, const texp_t *c) { /* truncexpsamplestepped, bmh.h:507-525 */
    double x = (0x1p-64 * (double)rngstate) * c->c1;
    if (x >= 1.) {
        for (;;) {
            if ((x = 0x1p-64 * (double)d2o_wyhash64(&rngstate)) < c->c2) break;
            double yhat = 0.5 * (0x1p-64 * (double)d2o_wyhash64(&rngstate));
            double omx = 1. - x;
            if (yhat > omx) { x = omx; yhat = 1. - yhat; }
            omx = 1. - x;
            if (x <= c->c3 * (1. - yhat) || (yhat * c->c1 <= omx)) break;
            if (fma(yhat, c->c4, 1.) <= exp(c->lambda * omx)) break;
        }
    }
    return x;
}

double d2o_pmh_update(double *regs, uint64_t *ids, uint32_t m, const uint64_t *keys, const double *wts, uint64_t n, double threshold) {
    lazyshuf_t ls; ls.n = m; ls.c = 0;
    ls.g = (uint32_t *)calloc(m, 4); ls.v = (uint32_t *)calloc(m, 4);
    const texp_t tc = texp_constants(m);
    double tw = 0., twc = 0.;
    for (uint64_t e = 0; e < n; ++e) {
        const uint64_t id = keys[e]; const double w = wts[e];
        if (!(w > threshold) || w <= 0.) continue;
        { double inc = w - twc; const double tmp = tw + inc; twc = (tmp - tw) - inc; tw = tmp; }
        uint64_t hi = id;
        const double wi = 1. / w;
        uint64_t i = 0;
        uint64_t rv = d2o_wyhash64(&hi);
        double maxv = regs[2ULL * m - 2];
        double hv = wi * texp_sample(rv, &tc);
        if (hv >= maxv) continue;
        ls.i = 0; ++ls.c; ls.state = rv; ls.off = 16;
        while (hv < maxv) {
            const uint32_t idx = ls_step(&ls);
            if (bmh_mvt_update(regs, m, idx, hv)) {
                if (ids) ids[idx] = id;
                maxv = regs[2ULL * m - 2];
                if (hv >= maxv) break;
            }
            if ((hv = wi * (double)(++i)) > maxv) break;
            hv = fma(wi, texp_sample(d2o_wyhash64(&rv), &tc), hv);
        }
    }
    free(ls.g); free(ls.v);
    return tw;
}

/* ---- BagMinHash2 ---- */
typedef struct { double x, weight, minp, maxq, carry; uint64_t idx, wyv, id; } pproc_t;
static inline uint64_t d2bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }
static inline double bits2d(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static void pp_step(pproc_t *p, uint32_t m) { /* bmh.h:170-176 */
    const uint64_t xi = d2o_wyhash64(&p->wyv);
    double inc = -log((double)(xi >> 12) * 2.220446049250313e-16) / (p->maxq - p->minp);
    inc -= p->carry;
    const double tmp = p->x + inc;
    p->carry = (tmp - p->x) - inc;
    p->x = tmp;
    p->idx = xi % m;
}
static inline int pp_partially(const pproc_t *p) { return bits2d(d2bits(p->minp) + 1) <= p->weight; }
static inline int pp_fully(const pproc_t *p) { return p->maxq <= p->weight; }
static inline int pp_can_split(const pproc_t *p) { return d2bits(p->maxq) > d2bits(p->minp) + 1; }
static pproc_t pp_split(pproc_t *p) { /* bmh.h:182-206 */
    uint64_t midpoint = (d2bits(p->minp) + d2bits(p->maxq)) / 2;
    const double midval = bits2d(midpoint);
    const uint64_t rval = d2o_wyhash64(&midpoint);
    uint64_t xval = d2bits(p->x) ^ rval;
    const double pr = (midval - p->minp) / (p->maxq - p->minp);
    const double rv = (double)d2o_wyhash64(&xval) * 5.421010862427522e-20;
    const int goleft = rv < pr;
    pproc_t r = *p;
    r.minp = goleft ? midval : p->minp; r.maxq = goleft ? p->maxq : midval; r.wyv = xval;
    r.idx = ~0ULL; /* a fresh process has no index until it steps; it is never used before */
    if (goleft) p->maxq = midval; else p->minp = midval;
    return r;
}
static void bmh_hv_update(double *regs, uint64_t *ids, uint32_t m, const pproc_t *p) {
    if (bmh_mvt_update(regs, m, p->idx, p->x) && ids) ids[p->idx] = p->id;
}
/* binary min-heap on x (the reference's priority queue pops the smallest x first: operator< is reversed, bmh.h:165-166) */
typedef struct { pproc_t *v; size_t n, cap; } pheap_t;
static void ph_push(pheap_t *h, pproc_t p) {
    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 64; h->v = (pproc_t *)realloc(h->v, h->cap * sizeof(pproc_t)); }
    size_t i = h->n++; h->v[i] = p;
    while (i && h->v[(i - 1) / 2].x > h->v[i].x) { pproc_t t = h->v[i]; h->v[i] = h->v[(i - 1) / 2]; h->v[(i - 1) / 2] = t; i = (i - 1) / 2; }
}
static pproc_t ph_pop(pheap_t *h) {
    pproc_t top = h->v[0]; h->v[0] = h->v[--h->n];
    size_t i = 0;
    for (;;) { size_t l = 2 * i + 1, r = l + 1, s = i;
        if (l < h->n && h->v[l].x < h->v[s].x) s = l;
        if (r < h->n && h->v[r].x < h->v[s].x) s = r;
        if (s == i) break;
        pproc_t t = h->v[i]; h->v[i] = h->v[s]; h->v[s] = t; i = s; }
    return top;
}

double d2o_bmh_update(double *regs, uint64_t *ids, uint32_t m, const uint64_t *keys, const double *wts, uint64_t n, double threshold) {
    pheap_t heap = {0, 0, 0};
    double tw = 0., twc = 0.;
    const uint64_t top = 2ULL * m - 2;
    for (uint64_t e = 0; e < n; ++e) {
        const double w = wts[e];
        if (!(w > threshold) || w <= 0.) continue;
        { double inc = w - twc; const double tmp = tw + inc; twc = (tmp - tw) - inc; tw = tmp; }
        pproc_t p = {0., w, 0., DBL_MAX, 0., ~0ULL, keys[e], keys[e]};
        pp_step(&p, m);
        if (pp_fully(&p)) bmh_hv_update(regs, ids, m, &p);
        heap.n = 0;
        while (p.x < regs[top]) {
            while (pp_can_split(&p) && pp_partially(&p)) {
                pproc_t q = pp_split(&p);
                if (pp_fully(&p)) bmh_hv_update(regs, ids, m, &p);
                if (pp_partially(&q)) {
                    pp_step(&q, m);
                    if (pp_fully(&q)) bmh_hv_update(regs, ids, m, &q);
                    if (pp_partially(&q)) ph_push(&heap, q);
                }
            }
            if (pp_fully(&p)) {
                pp_step(&p, m);
                bmh_hv_update(regs, ids, m, &p);
                if (p.x <= regs[top]) ph_push(&heap, p);
            }
            if (heap.n == 0) break;
            p = ph_pop(&heap);
        }
    }
    free(heap.v);
    return tw;
}

/* ------------------------------------------------------------------------------------------ */
/* LSH top-k                                                                                    */
/* ------------------------------------------------------------------------------------------ */
uint32_t d2o_lsh_key(const double *sig, uint32_t table_type, uint64_t j) { /* ssi.h:320-331,355-378 */
    if (table_type == 0) return (uint32_t)d2o_wang64(d2bits(sig[j]));
    uint64_t v0 = d2o_wang64(d2bits(sig[2 * j]));
    uint64_t v1 = d2o_wang64(d2bits(sig[2 * j + 1]) ^ v0);
    return (uint32_t)(v0 ^ v1);
}

typedef struct { uint32_t key, id; } kid_t;
static int cmp_kid(const void *a, const void *b) {
    const kid_t *x = (const kid_t *)a, *y = (const kid_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id);
}
typedef struct { kid_t *tab; uint64_t n, S, ntab; int nlsh; } lshidx_t; /* table t occupies tab[t*n, (t+1)*n), sorted by (key,id) */
/* --nLSH: table types 0 .. nlsh-1 with 1, 2, 4 registers per key and S, S/2, 8S/4 tables (src/cmp_core.cpp:757-770); 1 to 3 are
 * restated directly; types 3 .. 8 hash 6, 8, ... 16 registers (nperhashes = 2 * type, 8S / nperhashes tables) through XXH3_64bits (below) */
static int g_nlsh = 2;
void d2o_set_nlsh(int nlsh) { g_nlsh = nlsh < 1 ? 2 : nlsh > 9 ? 9 : nlsh; }
/* XXH64 (the published algorithm; the reference vendors xxHash) over len bytes, len a multiple of 8 */
static inline uint64_t xxh_round(uint64_t acc, uint64_t in) { acc += in * 0xC2B2AE3D27D4EB4FULL; acc = (acc << 31) | (acc >> 33); return acc * 0x9E3779B185EBCA87ULL; }
static inline uint64_t xxh_merge(uint64_t h, uint64_t v) { h ^= xxh_round(0, v); return h * 0x9E3779B185EBCA87ULL + 0x85EBCA77C2B2AE63ULL; }
static uint64_t xxh64_words(const uint64_t *w, uint64_t nwords, uint64_t seed) {
    const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL, P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
    uint64_t h, i = 0;
    if (nwords >= 4) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        for (; i + 4 <= nwords; i += 4) { v1 = xxh_round(v1, w[i]); v2 = xxh_round(v2, w[i + 1]); v3 = xxh_round(v3, w[i + 2]); v4 = xxh_round(v4, w[i + 3]); }
        h = ((v1 << 1) | (v1 >> 63)) + ((v2 << 7) | (v2 >> 57)) + ((v3 << 12) | (v3 >> 52)) + ((v4 << 18) | (v4 >> 46));
        h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
    } else h = seed + P5;
    h += nwords * 8;
    for (; i < nwords; ++i) { h ^= xxh_round(0, w[i]); h = ((h << 27) | (h >> 37)) * P1 + P4; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}
/* key of table (type, j): hash_index, ssi.h:355-392.  type 2 = four registers: hashmem256 (:313-318) while 4(j+1) <= S, else XXH64
 * seeded with ((type << 32) ^ (type >> 32)) | j over four registers picked by wyhash64(seed) -- truncated to 32 bits -- mod S. */
/* XXH3_64bits of 17 .. 128 bytes, seed 0, default secret (xxHash 0.8.0, the published algorithm: XXH3_len_17to128_64b).  Inputs here are
 * whole 64-bit words, an even number of them, so every 16-byte lane is two aligned words.  Secret: the first 128 bytes of XXH3_kSecret as
 * little-endian words. */
static const uint64_t XXH3_SECRET64[16] = {
    0xbe4ba423396cfeb8ULL, 0x1cad21f72c81017cULL, 0xdb979083e96dd4deULL, 0x1f67b3b7a4a44072ULL, 0x78e5c0cc4ee679cbULL, 0x2172ffcc7dd05a82ULL,
    0x8e2443f7744608b8ULL, 0x4c263a81e69035e0ULL, 0xcb00c391bb52283cULL, 0xa32e531b8b65d088ULL, 0x4ef90da297486471ULL, 0xd8acdea946ef1938ULL,
    0x3f349ce33f76faa8ULL, 0x1d4f0bc7c7bbdcf9ULL, 0x3159b4cd4be0518aULL, 0x647378d9c97e9fc8ULL};
static inline uint64_t xxh3_mix16(const uint64_t *in, const uint64_t *sec) { return wymum(in[0] ^ sec[0], in[1] ^ sec[1]); }   /* mul128_fold64 */
static uint64_t xxh3_64_words(const uint64_t *in, uint64_t W) {       /* W = 4 .. 16 words (32 .. 128 bytes; 32 bytes never gets here) */
    const uint64_t len = W * 8, *sec = XXH3_SECRET64;
    uint64_t acc = len * 0x9E3779B185EBCA87ULL;
    if (len > 32) {
        if (len > 64) {
            if (len > 96) { acc += xxh3_mix16(in + 6, sec + 12); acc += xxh3_mix16(in + W - 8, sec + 14); }
            acc += xxh3_mix16(in + 4, sec + 8); acc += xxh3_mix16(in + W - 6, sec + 10);
        }
        acc += xxh3_mix16(in + 2, sec + 4); acc += xxh3_mix16(in + W - 4, sec + 6);
    }
    acc += xxh3_mix16(in, sec); acc += xxh3_mix16(in + W - 2, sec + 2);
    acc ^= acc >> 37; acc *= 0x165667919E3779F9ULL; acc ^= acc >> 32;
    return acc;
}
static uint64_t lsh_nper(int type) { return type < 3 ? (1ULL << type) : 2ULL * (uint64_t)type; }    /* registers per key, cmp_core.cpp:757-760 */
static uint32_t lsh_key_any(const double *sig, uint64_t S, uint32_t type, uint64_t j) {
    if (type < 2) return d2o_lsh_key(sig, type, j);
    const uint64_t nreg = lsh_nper((int)type);
    uint64_t v[32];
    if ((j + 1) * nreg <= S) {
        memcpy(v, sig + nreg * j, nreg * 8);
        if (nreg == 4) return (uint32_t)d2o_wang64(d2o_cehash(v[0]) ^ (d2o_cehash(v[1]) * d2o_cehash(v[2]) - v[3]));   /* hashmem256 */
        return (uint32_t)xxh3_64_words(v, nreg);                                                                    /* hashmem default: XXH3_64bits, ssi.h:352 */
    }
    /* ssi.h:375-391: XXH64 seeded with ((type << 32) ^ (type >> 32)) | j over registers picked by wyhash64(seed) -- truncated to 32 bits -- mod S:
     * eight picks per whole eight of nreg, then nreg more */
    uint64_t seed = (((uint64_t)type << 32) ^ ((uint64_t)type >> 32)) | j;
    const uint64_t seed0 = seed, nw = nreg + 8 * (nreg / 8);
    for (uint64_t r = 0; r < nw; ++r) { const uint32_t pick = (uint32_t)d2o_wyhash64(&seed) % (uint32_t)S; memcpy(&v[r], sig + pick, 8); }
    return (uint32_t)xxh64_words(v, nw, seed0);
}
static uint64_t lsh_nsubs(uint64_t S, int type) { const uint64_t nh = lsh_nper(type); return nh <= 2 ? S / nh : S * 8 / nh; }
static uint64_t lsh_tab0(uint64_t S, int type) { uint64_t t = 0; for (int q = 0; q < type; ++q) t += lsh_nsubs(S, q); return t; }
static lshidx_t lsh_build(const double *regs, uint64_t n, uint64_t S) {
    lshidx_t ix; ix.n = n; ix.S = S; ix.nlsh = g_nlsh; ix.ntab = lsh_tab0(S, ix.nlsh);
    ix.tab = (kid_t *)malloc(sizeof(kid_t) * ix.ntab * n);
    for (uint64_t t = 0; t < ix.ntab; ++t) {
        uint32_t type = 0; while (t >= lsh_tab0(S, (int)type + 1)) ++type;
        const uint64_t j = t - lsh_tab0(S, (int)type);
        kid_t *T = ix.tab + t * n;
        for (uint64_t i = 0; i < n; ++i) { T[i].key = lsh_key_any(regs + i * S, S, type, j); T[i].id = (uint32_t)i; }
        qsort(T, n, sizeof(kid_t), cmp_kid); /* bucket order = insertion order = ascending id under -p1 */
    }
    return ix;
}
static uint64_t lsh_query(const lshidx_t *ix, const double *sig, uint64_t maxcand, uint32_t *ids, uint32_t *counts) {
    uint64_t nc = 0;
    const uint64_t n = ix->n, S = ix->S;
    for (int type = ix->nlsh - 1; type >= 0 && nc < maxcand; --type) { /* most specific table type first, ssi.h:425 */
        const uint64_t nsubs = lsh_nsubs(S, type);
        for (uint64_t j = 0; j < nsubs; ++j) {
            const uint32_t key = lsh_key_any(sig, S, (uint32_t)type, j);
            const kid_t *T = ix->tab + (lsh_tab0(S, type) + j) * n;
            uint64_t lo = 0, hi = n;
            while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (T[mid].key < key) lo = mid + 1; else hi = mid; }
            for (uint64_t q = lo; q < n && T[q].key == key; ++q) {
                const uint32_t id = T[q].id;
                uint64_t f = 0; for (; f < nc; ++f) if (ids[f] == id) break;
                if (f < nc) { ++counts[f]; continue; }
                ids[nc] = id; counts[nc] = 1; ++nc;
                if (nc == maxcand) return nc; /* early stop the moment maxcand distinct ids are seen, ssi.h:437-440 */
            }
        }
    }
    return nc;
}
uint64_t d2o_lsh_query(const double *regs, uint64_t n, uint64_t S, uint64_t query, uint64_t maxcand, uint32_t *ids, uint32_t *counts) {
    lshidx_t ix = lsh_build(regs, n, S);
    uint64_t r = lsh_query(&ix, regs + query * S, maxcand, ids, counts);
    free(ix.tab);
    return r;
}

/* neighbour list = std::priority_queue<pair<float,uint32>> (max-heap) + dedup set; kept as a sorted array */
typedef struct { float d; uint32_t id; } nb_t;
typedef struct { nb_t *v; uint32_t n, cap; uint32_t *set; uint32_t nset, capset; } nlist_t;
static int nb_less(nb_t a, nb_t b) { return a.d < b.d || (a.d == b.d && a.id < b.id); }
static void nl_push(nlist_t *l, nb_t x) {
    if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 16; l->v = (nb_t *)realloc(l->v, l->cap * sizeof(nb_t)); }
    uint32_t i = l->n++;
    while (i && nb_less(x, l->v[i - 1])) { l->v[i] = l->v[i - 1]; --i; }
    l->v[i] = x;
}
static int nl_inset(const nlist_t *l, uint32_t id) { for (uint32_t i = 0; i < l->nset; ++i) if (l->set[i] == id) return 1; return 0; }
static void nl_setadd(nlist_t *l, uint32_t id) {
    if (l->nset == l->capset) { l->capset = l->capset ? l->capset * 2 : 16; l->set = (uint32_t *)realloc(l->set, l->capset * 4); }
    l->set[l->nset++] = id;
}
static void nl_setdel(nlist_t *l, uint32_t id) { for (uint32_t i = 0; i < l->nset; ++i) if (l->set[i] == id) { l->set[i] = l->set[--l->nset]; return; } }
static void nl_update(nlist_t *l, nb_t item, uint64_t k) { /* index_build.cpp:20-44 with topk > 0 */
    if (nl_inset(l, item.id)) return;
    if (l->n < k) { nl_setadd(l, item.id); nl_push(l, item); return; }
    const nb_t top = l->v[l->n - 1];
    if (item.d <= top.d) {
        if (top.d != item.d) { nl_setdel(l, top.id); --l->n; }
        nl_push(l, item); /* note: the id is NOT added to the dedup set on this branch */
    }
}
static int cmp_nb(const void *a, const void *b) { nb_t x = *(const nb_t *)a, y = *(const nb_t *)b; return nb_less(x, y) ? -1 : nb_less(y, x); }
/* --topk with --fastcmp N: the LSH index is built over the f64 signatures (index_build.cpp:93, sketch_compressed_set is false), while
 * compare() in refine_results takes the compressed branch (cmp_core.cpp:362-449) over make_compressed's registers.  Set by
 * d2o_topk_compressed around d2o_topk. */
static const double *g_cregs = NULL; static int g_c_bbit = 0; static double g_c_fd = 0.; static long double g_c_b = 0.L;
static float refine_compare(const double *regs, uint64_t S, uint64_t i, uint64_t j, const double *cards, int measure, int k, int cmp_kind) {
    if (!g_cregs) return d2o_compare(regs + i * S, regs + j * S, S, cards[i], cards[j], measure, k, cmp_kind);
    uint64_t x = 0, y = 0;
    if (g_c_bbit) x = d2o_count_eq((const uint64_t *)(g_cregs + i * S), (const uint64_t *)(g_cregs + j * S), S);
    else d2o_count_gtlt(g_cregs + i * S, g_cregs + j * S, S, &x, &y);
    return d2o_finalize_compressed(x, y, S, cards[i], cards[j], measure, k, g_c_bbit, g_c_fd, g_c_b);
}

uint64_t d2o_topk(const double *regs, const double *cards, uint64_t n, uint64_t S, int topk, int measure, int k, int cmp_kind,
                  uint64_t *indptr, uint32_t **idx, float **val) {
    lshidx_t ix = lsh_build(regs, n, S);
    uint64_t ntoquery = (uint64_t)((float)topk * 3.5f); /* index_build.cpp:57-60 (LSHDistType = float) */
    if (ntoquery > n - 1) ntoquery = n - 1;
    nlist_t *L = (nlist_t *)calloc(n, sizeof(nlist_t));
    uint32_t *ids = (uint32_t *)malloc(4 * (ntoquery + 1)), *cnt = (uint32_t *)malloc(4 * (ntoquery + 1));
    for (uint64_t q = 0; q < n; ++q) {
        const uint64_t nc = ntoquery ? lsh_query(&ix, regs + q * S, ntoquery, ids, cnt) : 0;
        for (uint64_t j = 0; j < nc; ++j) {
            const uint32_t oid = ids[j];
            if (oid == q) continue;
            const float cd = -(float)cnt[j];
            nl_update(&L[oid], (nb_t){cd, (uint32_t)q}, ntoquery);
            nl_update(&L[q], (nb_t){cd, oid}, ntoquery);
        }
    }
    /* cmp_main.h:44-49: everything except union/intersection/similarity/containment counts as a distance
     * (including symmetric containment -- a quirk of the reference that the output order depends on) */
    const int is_dist = !(measure == D2O_UNION_SIZE || measure == D2O_INTERSECTION || measure == D2O_SIMILARITY || measure == D2O_CONTAINMENT);
    const float mult = is_dist ? 1.f : -1.f;
    uint64_t nnz = 0;
    for (uint64_t i = 0; i < n; ++i) { /* refine.cpp:20-76, num_neighbors_ > 0 branch */
        nlist_t *l = &L[i];
        for (uint32_t j = 0; j < l->n; ++j)
            l->v[j].d = mult * refine_compare(regs, S, i, l->v[j].id, cards, measure, k, cmp_kind);
        qsort(l->v, l->n, sizeof(nb_t), cmp_nb);
        if (!is_dist) { uint32_t j = 0; while (j < l->n && l->v[j].d != 0.f) ++j; l->n = j; }
        if ((uint32_t)topk < l->n) { const float bs = l->v[topk - 1].d; uint32_t j = (uint32_t)topk; while (j < l->n && !(l->v[j].d > bs)) ++j; l->n = j; }
        if (!is_dist) for (uint32_t j = 0; j < l->n; ++j) l->v[j].d = -l->v[j].d;
        indptr[i] = nnz; nnz += l->n;
    }
    indptr[n] = nnz;
    *idx = (uint32_t *)malloc(4 * (nnz + 1)); *val = (float *)malloc(4 * (nnz + 1));
    for (uint64_t i = 0, o = 0; i < n; ++i) for (uint32_t j = 0; j < L[i].n; ++j, ++o) { (*idx)[o] = L[i].v[j].id; (*val)[o] = L[i].v[j].d; }
    for (uint64_t i = 0; i < n; ++i) { free(L[i].v); free(L[i].set); }
    free(L); free(ids); free(cnt); free(ix.tab);
    return nnz;
}
uint64_t d2o_topk_compressed(const double *regs, const double *cregs, const double *cards, uint64_t n, uint64_t S, int topk, int measure, int k,
                             int bbit, double fd, long double b, uint64_t *indptr, uint32_t **idx, float **val) {
    g_cregs = cregs; g_c_bbit = bbit; g_c_fd = fd; g_c_b = b;
    const uint64_t r = d2o_topk(regs, cards, n, S, topk, measure, k, 0, indptr, idx, val);
    g_cregs = NULL;
    return r;
}
/* --similarity-threshold x (NN_GRAPH_THRESHOLD), sequential (-p1) semantics: build_index with topk = -1 (src/index_build.cpp:56,26-31: every
 * candidate of every query is appended to both endpoints' lists unless already there; candidate scan up to n-1 distinct ids, :59), lists
 * sorted by (-hit count, id) (:139-141), then refine_results' threshold branch (src/refine.cpp:43-68): walk the list in that order, keep
 * entries whose measure passes the threshold, stop after 20 consecutive failures, sort by (-similarity, id).  Distances pass with v < x. */
uint64_t d2o_nn_threshold(const double *regs, const double *cards, uint64_t n, uint64_t S, double min_sim, int measure, int k, int cmp_kind,
                          uint64_t *indptr, uint32_t **idx, float **val) {
    lshidx_t ix = lsh_build(regs, n, S);
    const uint64_t ntoquery = n - 1;
    nlist_t *L = (nlist_t *)calloc(n, sizeof(nlist_t));
    uint32_t *ids = (uint32_t *)malloc(4 * (ntoquery + 1)), *cnt = (uint32_t *)malloc(4 * (ntoquery + 1));
    unsigned char *seen = (unsigned char *)calloc(n, 1);     /* per-list dedup sets, kept as one n x n bit table would be too big: rebuilt per query */
    for (uint64_t q = 0; q < n; ++q) {
        const uint64_t nc = ntoquery ? lsh_query(&ix, regs + q * S, ntoquery, ids, cnt) : 0;
        for (uint64_t j = 0; j < nc; ++j) {
            const uint32_t oid = ids[j];
            if (oid == q) continue;
            const float cd = -(float)cnt[j];
            if (!nl_inset(&L[oid], (uint32_t)q)) { nl_setadd(&L[oid], (uint32_t)q); nl_push(&L[oid], (nb_t){cd, (uint32_t)q}); }
            if (!nl_inset(&L[q], oid)) { nl_setadd(&L[q], oid); nl_push(&L[q], (nb_t){cd, oid}); }
        }
    }
    free(seen);
    const int is_dist = !(measure == D2O_UNION_SIZE || measure == D2O_INTERSECTION || measure == D2O_SIMILARITY || measure == D2O_CONTAINMENT);
    const float mult = is_dist ? 1.f : -1.f, MDIST = 3.402823466e+38f, ms = (float)min_sim;   /* LSHDistType = float; min_similarity_ is a double compared against float values */
    uint64_t nnz = 0;
    for (uint64_t i = 0; i < n; ++i) {
        nlist_t *l = &L[i];                                   /* nl_push keeps the list sorted by (d, id) = (-count, id) */
        uint32_t failures = 0, lsz = l->n;
        for (uint32_t j = 0; j < lsz; ++j) {
            const float v = d2o_compare(regs + i * S, regs + (uint64_t)l->v[j].id * S, S, cards[i], cards[l->v[j].id], measure, k, cmp_kind);
            const int pass = is_dist ? (double)v < min_sim : (double)v >= min_sim;
            if (!pass) { l->v[j].d = MDIST; if (++failures == 20) { l->n = j; break; } }
            else { l->v[j].d = v * mult; failures = 0; }
        }
        uint32_t o = 0;
        for (uint32_t j = 0; j < l->n; ++j) {
            const nb_t x = l->v[j];
            const int drop = x.d == MDIST || (is_dist ? (double)x.d > min_sim : (double)(-x.d) < min_sim);
            if (!drop) l->v[o++] = x;
        }
        l->n = o;
        qsort(l->v, l->n, sizeof(nb_t), cmp_nb);
        if (!is_dist) for (uint32_t j = 0; j < l->n; ++j) l->v[j].d = -l->v[j].d;
        indptr[i] = nnz; nnz += l->n;
    }
    (void)ms;
    indptr[n] = nnz;
    *idx = (uint32_t *)malloc(4 * (nnz + 1)); *val = (float *)malloc(4 * (nnz + 1));
    for (uint64_t i = 0, o = 0; i < n; ++i) for (uint32_t j = 0; j < L[i].n; ++j, ++o) { (*idx)[o] = L[i].v[j].id; (*val)[o] = L[i].v[j].d; }
    for (uint64_t i = 0; i < n; ++i) { free(L[i].v); free(L[i].set); }
    free(L); free(ids); free(cnt); free(ix.tab);
    return nnz;
}
void d2o_free(void *p) { free(p); }
