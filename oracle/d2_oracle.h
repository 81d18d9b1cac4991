/*
 * d2_oracle.h -- CPU restatement of the dashing2 sketch / cmp hot paths.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker.  The product (dashing2_b200/csrc) never links or calls it.
 *
 * Parity pinning: the reference ships no numeric golden vectors for this path;
 * this restatement is pinned against outputs of the reference binary itself, built unmodified from
 * /root/reference by oracle/Makefile.ref into oracle/_ref/ (see tests/golden/make_golden.py and
 * tests/test_oracle_vs_golden.py), plus the two known answers in bonsai/test/encoding.cpp:84,122.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 */
#ifndef D2_ORACLE_H
#define D2_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- integer hashes ---------------------------------------------------------------------- */
uint64_t d2o_wang64(uint64_t key);          /* bonsai/hll/include/sketch/hash.h:42-62 */
uint64_t d2o_wang64_inv(uint64_t h);        /* inverse of the above (hash.h WangHash::inverse) */
uint64_t d2o_frev64(uint64_t x);            /* bonsai/include/bonsai/encoder.h:47,59 lex_score */
uint64_t d2o_cehash(uint64_t x);            /* hash.h:858 CEHasher */
uint64_t d2o_wyhash64(uint64_t *state);     /* bonsai/hll/include/aesctr/wy.h:56-59 */
uint64_t d2o_revcomp(uint64_t kmer, int k); /* bonsai/include/bonsai/kmerutil.h:83-90 */
uint64_t d2o_xormask_for_seed(uint64_t seed); /* src/enums.cpp:133-140 */

/* ---- k-mer / minimizer stream -------------------------------------------------------------- */
/* Number of k-mer positions in one record: max(0, len-k+1) (metric unit). */
uint64_t d2o_kmer_positions(uint64_t len, int k);
/* Emit maskfn'd k-mers (or minimizers when w > k) of ONE record, in reference order.
 * Returns the number emitted; writes at most cap values (pass cap >= len+1).
 * encoder.h:212-272,282-317,444-451; qmap.h:79-87; src/fastxsketch.cpp:385-389; src/enums.h:136-140 */
uint64_t d2o_hash_stream(const char *seq, uint64_t len, int k, int w, int canon, uint64_t xormask,
                         uint64_t *out, uint64_t cap);

/* Same for the protein alphabets (alphabet = 20, 14, 6, or 8 for the 3-bit encoding; never canonical): alphabet.h:107-120,
 * rhtraits.h:52-62, encoder.h:241-306. */
uint64_t d2o_hash_stream_protein(const char *seq, uint64_t len, int k, int w, int alphabet, uint64_t xormask, uint64_t *out, uint64_t cap);
/* Same for k > 32: RollingHasher<uint64_t> over CyclicHash (bonsai encoder.h:644-865, rollinghash/cyclichash.h). */
uint64_t d2o_hash_stream_rolling(const char *seq, uint64_t len, int k, int w, int canon, uint64_t xormask,
                                 uint64_t *out, uint64_t cap);

/* ---- One-permutation MinHash (src/oph.h) --------------------------------------------------- */
#define D2O_OPH_SEED 0x8f1896f3f85ef4a3ULL  /* std::mt19937_64(0x321b919a61cb41f7)(), oph.h:59,142 */
uint32_t d2o_opmh_m(uint32_t sketchsize);   /* oph.h:145 (rounded up to even) */
void d2o_opmh_reset(uint64_t *regs, double *counts, uint32_t m);                    /* oph.h:232-239 */
void d2o_opmh_update(uint64_t *regs, double *counts, uint32_t m, const uint64_t *hv, uint64_t n); /* oph.h:176-211 (mincount<=1) */
double d2o_opmh_card(const uint64_t *regs, uint32_t m);                             /* oph.h:240-247 */
void d2o_opmh_sigs(const uint64_t *regs, uint32_t m, double *out);                  /* oph.h:248-263 */
void d2o_opmh_ids(const uint64_t *regs, uint32_t m, uint64_t *out);                 /* oph.h:264-271 */
/* --count-threshold > 1 variant (oph.h:188-205): sequential, order dependent. */
void d2o_opmh_update_mincount(uint64_t *regs, double *counts, uint32_t m, const uint64_t *hv,
                              uint64_t n, double mincount);

/* ---- Full (continuous) SetSketch, src/setsketch.h:369-423 ---------------------------------- */
/* regs: f64[2m-1] register array + max-tree (mvt_t, setsketch.h:123-167). */
void d2o_css_reset(double *regs, uint32_t m);
void d2o_css_update(double *regs, uint32_t m, const uint64_t *hv, uint64_t n, uint64_t *ids /*nullable*/);
double d2o_css_card(const double *regs, uint32_t m);       /* setsketch.h:553-561 */

/* ---- densify (src/cmp_core.cpp:577-613) ---------------------------------------------------- */
uint64_t d2o_densify(double *sig, uint64_t *kmers /*nullable*/, uint64_t sketchsize);

/* ---- compare (src/cmp_core.cpp:349-575) ---------------------------------------------------- */
enum { D2O_SIMILARITY = 0, D2O_CONTAINMENT = 1, D2O_SYMMETRIC_CONTAINMENT = 2, D2O_POISSON_LLR = 3,
       D2O_INTERSECTION = 4, D2O_UNION_SIZE = 5 };
void d2o_count_gtlt(const double *a, const double *b, uint64_t n, uint64_t *gt, uint64_t *lt); /* count_eq.h:412-445 */
uint64_t d2o_count_eq(const uint64_t *a, const uint64_t *b, uint64_t n);                      /* count_eq.h:40-56 */
/* finalisation of one pair from integer counts; cmp_kind 0 = gt/lt branch (:458-494), 1 = equality branch (:495-517) */
float d2o_finalize(uint64_t gt_or_eq, uint64_t lt, uint64_t sketchsize, double lhcard, double rhcard,
                   int measure, int k, int cmp_kind);
float d2o_compare(const double *a, const double *b, uint64_t sketchsize, double lhcard, double rhcard,
                  int measure, int k, int cmp_kind);
/* all-pairs drivers restating src/emitrect.cpp:229-326 orderings. out sizes: n(n-1)/2, n*n, nf*nq */
/* compressed registers: make_compressed (src/cmp_core.cpp:209-322), compressed compare branch (:362-449) */
int d2o_make_compressed(const double *sigs, const uint64_t *kmers /*nullable*/, uint64_t nsigs, double fd, int truncation,
                        long double *a_io, long double *b_io, double *out);
float d2o_finalize_compressed(uint64_t c0, uint64_t c1, uint64_t sketchsize, double lhcard, double rhcard, int measure, int k,
                              int bbit, double fd, long double b);
void d2o_allpairs_compressed(const double *cregs, const double *cards, uint64_t n, uint64_t nq, uint64_t S, int shape,
                             int measure, int k, int bbit, double fd, long double b, float *out);
void d2o_allpairs_symmetric(const double *regs, const double *cards, uint64_t n, uint64_t S,
                            int measure, int k, int cmp_kind, float *out);
void d2o_allpairs_asymmetric(const double *regs, const double *cards, uint64_t n, uint64_t S,
                             int measure, int k, int cmp_kind, float *out);
void d2o_panel(const double *regs, const double *cards, uint64_t nf, uint64_t nq, uint64_t S,
               int measure, int k, int cmp_kind, float *out);

#ifdef __cplusplus
}
#endif
#endif

/* ---- weighted sketches over (hashed k-mer, count) pairs: src/counter.h:68-77,118-138 ------------------ */
#ifdef __cplusplus
extern "C" {
#endif
/* Sorts hv in place and run-length encodes it: keys[i] with counts[i]; returns the number of distinct keys.
 * (The reference counts in a hash map; sketches below are order independent, so sorted order is as good.) */
uint64_t d2o_count_exact(uint64_t *hv, uint64_t n, uint64_t *keys, double *counts);
/* --countsketch-size n (src/counter.h:68-77,131-137): signed count-sketch table of n floats over the hashed k-mer stream; the elements fed to
 * the weighted sketch are (bucket index, |count|) for |count| >= threshold (zero-weight buckets dropped: no-ops). Returns their number. */
uint64_t d2o_count_sketch(const uint64_t *hv, uint64_t n, uint64_t cssize, double threshold, uint64_t *keys, double *counts);
/* --parse-by-seq (one sketch per record) cardinality rule for set sketches, src/fastxsketchbyseq.cpp:393-430: NaN -> 0, and an
 * estimate < 10 * sketchsize is replaced by the exact number of distinct hashed k-mers of the record (hv is sorted in place). */
double d2o_byseq_cardinality(double estimate, uint64_t sketchsize, uint64_t *hv, uint64_t n);
/* ProbMinHash3 (bonsai/hll/include/sketch/bmh.h:662-700; base :545-661; truncated exponential :490-525).
 * regs f64[2m-1] (value tree), ids u64[m]; returns total weight. Elements with count <= threshold are skipped
 * (src/counter.h:123). */
void d2o_pmh_reset(double *regs, uint64_t *ids, uint32_t m);
double d2o_pmh_update(double *regs, uint64_t *ids, uint32_t m, const uint64_t *keys, const double *w, uint64_t n, double threshold);
/* BagMinHash2 (bmh.h:269-316 update_2, poisson_process_t :129-207), restated per element without the carried heap
 * (the carried heap only ever holds points that cannot lower a register). */
double d2o_bmh_update(double *regs, uint64_t *ids, uint32_t m, const uint64_t *keys, const double *w, uint64_t n, double threshold);
#ifdef __cplusplus
}
#endif

/* ---- LSH top-k nearest-neighbour graph: src/ssi.h:290-453, src/index_build.cpp:20-165, src/refine.cpp:6-81,
 * src/emitnn.cpp:12-52, table geometry src/cmp_core.cpp:757-772.  Sequential (-p1) semantics. */
#ifdef __cplusplus
extern "C" {
#endif
/* --nLSH for the calls below: 2 (default: table types 0 and 1) or 1 (type 0 only); src/cmp_core.cpp:757-770. */
void d2o_set_nlsh(int nlsh);
/* LSH keys: table type 0 = one register, type 1 = two registers (default --nLSH 2); truncated to 32 bits. */
uint32_t d2o_lsh_key(const double *sig, uint32_t table_type, uint64_t j);
/* candidate scan for one query; ids/counts must hold maxcand entries; returns the number of candidates. */
uint64_t d2o_lsh_query(const double *regs, uint64_t n, uint64_t S, uint64_t query, uint64_t maxcand, uint32_t *ids, uint32_t *counts);
/* whole pipeline -> CSR (indptr[n+1], idx/val malloc'ed; free with d2o_free). Returns nnz. */
uint64_t d2o_topk(const double *regs, const double *cards, uint64_t n, uint64_t S, int topk, int measure, int k, int cmp_kind,
                  uint64_t *indptr, uint32_t **idx, float **val);
/* --topk with --fastcmp N [--bbit-sigs]: index over the f64 signatures, refinement through the compressed compare branch over cregs
 * (from d2o_make_compressed). */
uint64_t d2o_topk_compressed(const double *regs, const double *cregs, const double *cards, uint64_t n, uint64_t S, int topk, int measure, int k,
                             int bbit, double fd, long double b, uint64_t *indptr, uint32_t **idx, float **val);
/* --similarity-threshold x: every id sharing an LSH bucket, refined with the 20-consecutive-failures rule (src/refine.cpp:43-68) -> CSR. */
uint64_t d2o_nn_threshold(const double *regs, const double *cards, uint64_t n, uint64_t S, double min_sim, int measure, int k, int cmp_kind,
                          uint64_t *indptr, uint32_t **idx, float **val);
void d2o_free(void *p);
#ifdef __cplusplus
}
#endif
