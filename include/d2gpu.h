/*
 * d2gpu.h -- C ABI of libd2gpu.so: the B200 (sm_100a) implementation of dashing2's two hot paths.
 *
 * The reference (dnbaker/dashing2 @ 3906ebde, paths relative to /root/reference) has no FFI; its
 * seam for these paths is a handful of C++ functions.  Each entry point below names the reference
 * function(s) it replaces.  Conventions: plain pointers + sizes, no C++/torch types; every call
 * returns 0 on success or a negative D2G_E* code, with a message in d2g_last_error() (thread local);
 * no exception crosses this boundary.  A d2g_ctx owns one CUDA device + stream + scratch memory and
 * is thread-compatible (one thread at a time per ctx).  There is NO CPU fallback: without a CUDA
 * device d2g_init fails with D2G_ENODEVICE.
 *
 * Pointer naming: plain = host memory; *_d = device memory on the ctx's device.
 */
#ifndef D2GPU_H
#define D2GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D2G_OK 0
#define D2G_EINVAL (-1)      /* bad argument / unsupported parameter combination */
#define D2G_ENODEVICE (-2)   /* no usable CUDA device */
#define D2G_ECUDA (-3)       /* CUDA runtime error (see d2g_last_error) */
#define D2G_ENOMEM (-4)
#define D2G_EIO (-5)
#define D2G_EUNSUPPORTED (-6)/* a reference mode this library does not implement on the GPU */

typedef struct d2g_ctx d2g_ctx;

int d2g_init(d2g_ctx **ctx, int device);
void d2g_destroy(d2g_ctx *ctx);
const char *d2g_last_error(void);
const char *d2g_version(void);
/* The CUDA stream (cudaStream_t) every call on this ctx is enqueued on; for event timing. */
void *d2g_stream(d2g_ctx *ctx);
int d2g_sync(d2g_ctx *ctx);
/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches). */
uint64_t d2g_launch_count(const d2g_ctx *ctx);
/* Counters of the last call, for the roofline report.  which: D2G_STAT_REFINED = neighbour-list entries the last d2g_lsh_* call compared
 * exactly (before trimming; each reads one neighbour row of S registers, plus the owner's row once per list). */
enum { D2G_STAT_REFINED = 0, D2G_STAT_NSTATS = 1 };
uint64_t d2g_stat(const d2g_ctx *ctx, int which);
/* Per-kernel device timing for the roofline report: when enabled, CUDA events bracket the dominant
 * kernels on the ctx stream. d2g_get_timing synchronises, returns accumulated milliseconds and launch
 * count for one kernel class and resets that class. */
enum { D2G_T_SKETCH_MAIN = 0, D2G_T_SKETCH_BOOT = 1 /* Full SetSketch boot / ids passes; the BagMinHash / ProbMinHash element kernels */, D2G_T_CMP = 2 /* the pair-comparison tile kernel */,
       D2G_T_CMP_PREP = 3 /* order-code construction: keys, per-register sort, ranks */,
       D2G_T_PACK = 4 /* ASCII -> packed sequence (d2g_pack_dev) */,
       D2G_T_SORT = 5 /* counting sketches: radix sorts by (entity, value) + run-length encode; LSH: per-table key sort */,
       D2G_T_LSH_REFINE = 6 /* exact compare of the surviving neighbour-list entries */, D2G_T_NCLASSES = 7 };
int d2g_set_timing(d2g_ctx *ctx, int enabled);
int d2g_get_timing(d2g_ctx *ctx, int kernel_class, double *ms_total, uint64_t *n_launches);

/* ------------------------------------------------------------------------------------------------
 * Sketch path.  Replaces the per-file body of fastx2sketch (src/fastxsketch.cpp:303-624): k-mer
 * encode (bonsai encoder.h:241-272), canonicalise (kmerutil.h:137), windowed minimizer
 * (encoder.h:212-217, qmap.h:79-87), maskfn (src/enums.h:136-140) and the sketch update
 * (src/oph.h:176-211 / src/setsketch.h:369-423 / bmh.h:269-316 / bmh.h:662-700).
 * ---------------------------------------------------------------------------------------------- */
enum { D2G_MODE_OPMH = 0, D2G_MODE_FULL_SETSKETCH = 1, D2G_MODE_BAGMINHASH = 2, D2G_MODE_PROBMINHASH = 3 };

typedef struct {
    int32_t k;                 /* 1..32: exact 2-bit encoding (bonsai encoder.h:241-272); 33..: 64-bit cyclic rolling hash
                                  (RollingHasher, encoder.h:644-865; dispatch src/fastxsketch.cpp:399-421) */
    int32_t w;                 /* window; w <= k means unwindowed */
    int32_t canon;             /* reference default 1 (src/sketch_main.cpp:28) */
    int32_t mode;              /* D2G_MODE_* */
    uint64_t xormask;          /* maskfn XOR mask: 0 for --seed 0, else Wang(seed) (src/enums.cpp:133) */
    uint32_t sketchsize;       /* S */
    uint32_t count_threshold;  /* -m / --count-threshold; 0/1 = off.  OPMH: register = min id seen >= c times (oph.h:188-205); counting
                                  sketches: elements with count <= c are skipped (counter.h:123); Full SetSketch: D2G_EUNSUPPORTED */
    uint64_t countsketch_size; /* -c / --countsketch-size n, counting sketches only; 0 = exact counting.  n > 0: signed count sketch of n buckets,
                                  elements (bucket index, |count|) for |count| >= count_threshold (counter.h:68-77,131-137) */
    int32_t alphabet;          /* 0 (or 4) = DNA; 20 / 14 / 6 / 8 = --protein(20) / --protein14 / --protein6 / --protein8 (src/options.h:328-331;
                                  bonsai alphabet.h:107-120, rhtraits.h:52-62), never canonical, k up to 14 / 16 / 24 / 22, ASCII input only */
    int32_t reserved;
} d2g_sketch_params;

/* Number of registers per entity the OPMH sketch keeps (S rounded up to even, src/oph.h:145). */
uint32_t d2g_opmh_m(uint32_t sketchsize);
/* The metric unit: k-mer positions fed to the sketch = sum over records of max(0, len-k+1). */
uint64_t d2g_count_kmers(const uint64_t *rec_off, uint64_t n_rec, int32_t k);

/*
 * One batch of records belonging to n_entities sketches (entity = input file, or record under
 * --parse-by-seq).  seq = concatenated record bytes with line terminators already removed
 * (kseq semantics, bonsai/klib/kseq.h:178); rec_off[n_rec+1] byte offsets; rec_entity[n_rec] the
 * sketch each record feeds (non-decreasing).  Outputs are caller-allocated and may be NULL when
 * not wanted:
 *   regs_u64_out [n_entities][m]  OPMH raw 64-bit bucket minima (m = d2g_opmh_m(S)), ~0 = empty
 *   sig_out      [n_entities][S]  f64 registers exactly as the reference stores them in
 *                                 SketchingResult::signatures_ (src/fastxsketch.h:47)
 *   card_out     [n_entities]     cardinality estimate (oph.h:240-247 / setsketch.h:553-561 / total weight)
 *   ids_out      [n_entities][S]  --save-kmers ids: the hashed k-mer behind each register (oph.h:264-271; bmh.h ids_ of BagMinHash /
 *                                 ProbMinHash, the bucket index under countsketch_size; setsketch.h:400-404 for the Full SetSketch)
 * All pointers are HOST memory; the call copies in, runs the kernels, copies out and synchronises.
 */
int d2g_sketch_batch(d2g_ctx *ctx, const d2g_sketch_params *p,
                     const char *seq, const uint64_t *rec_off, const uint32_t *rec_entity,
                     uint64_t n_rec, uint32_t n_entities,
                     uint64_t *regs_u64_out, double *sig_out, double *card_out, uint64_t *ids_out,
                     uint64_t *n_kmers_hashed);

/* Exact number of distinct k-mers (minimizers when w > k) per entity.  Replaces the small-cardinality fallback of
 * --parse-by-seq (src/fastxsketchbyseq.cpp:405-430): when a set sketch's estimate is below 10 * sketchsize the reference walks
 * the record again into a hash set of maskfn'd k-mers and stores its size as the cardinality.  Same record tables as
 * d2g_sketch_batch (rec_off[0] == 0, at most 2^32 bases per call); only k, w, canon and xormask of the parameters matter
 * (maskfn is a bijection, so the count does not depend on the seed).  distinct_out [n_entities], host memory. */
int d2g_distinct_kmers(d2g_ctx *ctx, const d2g_sketch_params *p,
                       const char *seq, const uint64_t *rec_off, const uint32_t *rec_entity,
                       uint64_t n_rec, uint32_t n_entities, uint64_t *distinct_out);

/* --filterset PATH (src/d2.cpp:45-98, src/filterset.h FilterSet as a sorted hash set; the test in front of every sketch update,
 * src/fastxsketch.cpp:385-388): hashed k-mers (minimizers when w > k) found in the set never reach the sketches of later d2g_sketch_* /
 * d2g_distinct_kmers / d2g_kmer_counts calls on this ctx.  d2g_set_filterset builds the set on the device from the records of the filter
 * file (host ASCII, same record table as d2g_sketch_batch with rec_off[0] == 0; k, w, canon, xormask, alphabet of p as for the sketch
 * itself -- the reference hashes the filter file with the options of the run); *n_out (optional) = hashed values kept, duplicates
 * included.  d2g_set_filterset_values takes hashed values directly (the reference's "PATH:x" raw 64-bit file).  The kernels with the
 * membership test are separate instantiations: nothing changes for a ctx without a filter set. */
int d2g_set_filterset(d2g_ctx *ctx, const d2g_sketch_params *p, const char *seq, const uint64_t *rec_off, uint64_t n_rec, uint64_t *n_out);
int d2g_set_filterset_values(d2g_ctx *ctx, const uint64_t *values, uint64_t n);
int d2g_clear_filterset(d2g_ctx *ctx);

/* --save-kmercounts (-N): counts_out f32 [n_entities][S] = how often the element that owns each register occurs in the stream of hashed
 * k-mers (one per window when w > k) of its entity -- what the reference keeps beside the registers (src/oph.h:206-209 counts_,
 * src/setsketch.h:405-406, the weights of BagMinHash / ProbMinHash) and writes as float32 to FILE.kmercounts.f64
 * (src/sketch_core.cpp:162-171).  ids [n_entities][S]: the ids_out of the sketch call over the same batch (packed sequence, host
 * memory; mask may be NULL).  At most 2^32 bases per call; exact counting only. */
int d2g_kmer_counts(d2g_ctx *ctx, const d2g_sketch_params *p, const uint64_t *codes, const uint32_t *mask,
                    const uint64_t *rec_off, const uint32_t *rec_entity, uint64_t n_rec, uint32_t n_entities,
                    const uint64_t *ids, float *counts_out);

/* Host-side transform of OPMH bucket minima (host memory) into the reference's f64 signatures and
 * cardinality: sig = -1/(m-nempty) * logl(2^-64 * (2^64 - reg)), card = m*m / sum(reg * 2^-64), both in
 * x87 long double exactly as src/oph.h:240-263 does on the host. regs_u64 [n][d2g_opmh_m(S)]. */
int d2g_opmh_finalize(const uint64_t *regs_u64, uint32_t n_entities, uint32_t sketchsize, double *sig_out, double *card_out);

/* Same computation with inputs and outputs resident in device memory (asynchronous on
 * d2g_stream(ctx); no host copies).  seq_d holds total_len ASCII bytes; it is packed on the device first
 * (d2g_pack_dev into ctx scratch) and the sketch kernels read the packed form. */
int d2g_sketch_batch_dev(d2g_ctx *ctx, const d2g_sketch_params *p,
                         const char *seq_d, const uint64_t *rec_off_d, const uint32_t *rec_entity_d,
                         uint64_t n_rec, uint32_t n_entities, uint64_t total_len,
                         uint64_t *regs_u64_out_d, double *sig_out_d, double *card_out_d, uint64_t *ids_out_d);

/* ---- packed sequence: what the sketch kernels read, and what crosses PCIe ------------------------------------------
 * A batch of n_bases record bytes is held as
 *   codes : uint64[d2g_packed_words(n_bases)], word i = bases [32i, 32i+32), two bits per base, first base in bits 63:62
 *           (A0 C1 G2 T3, case-insensitive: bonsai/include/bonsai/alphabet.h:128);
 *   mask  : uint32[d2g_packed_words(n_bases)], word i = the same bases, one bit per base, first base in bit 31,
 *           1 = not A/C/G/T (encoder.h:254 resets the k-mer run there; in windowed mode the k-mer becomes 0, :568-571).
 * d2g_packed_words includes the padding the kernels rely on.  This replaces the byte-per-base hand-over at
 * Encoder::for_each (bonsai/include/bonsai/encoder.h:511-530): a quarter of the bytes per base over PCIe. */
uint64_t d2g_packed_words(uint64_t n_bases);
/* Host packer (AVX-512 / AVX2 / scalar, all host threads the process may use; D2G_HOST_THREADS overrides): packs the
 * concatenation of n_pieces ASCII pieces (records with line terminators already removed).  *n_invalid_words (optional)
 * receives the number of words holding a non-ACGT base: when 0 the mask need not be passed on. */
int d2g_pack_sequences(const char *const *pieces, const uint64_t *piece_len, uint64_t n_pieces,
                       uint64_t *codes, uint32_t *mask, uint64_t *n_invalid_words);
/* Device packer: total_len ASCII bytes at seq_d -> codes_d / mask_d (d2g_packed_words(total_len) words each); asynchronous. */
int d2g_pack_dev(d2g_ctx *ctx, const char *seq_d, uint64_t total_len, uint64_t *codes_d, uint32_t *mask_d);
/* d2g_sketch_batch with the sequence already packed by the caller (host memory; mask may be NULL when no base is invalid).
 * d2g_sketch_batch itself packs its ASCII input chunk by chunk on the host threads into a pinned staging ring, so both
 * calls move the same bytes; this one saves the pass over the ASCII when the caller packs while it parses. */
int d2g_sketch_batch_packed(d2g_ctx *ctx, const d2g_sketch_params *p,
                            const uint64_t *codes, const uint32_t *mask, const uint64_t *rec_off, const uint32_t *rec_entity,
                            uint64_t n_rec, uint32_t n_entities,
                            uint64_t *regs_u64_out, double *sig_out, double *card_out, uint64_t *ids_out,
                            uint64_t *n_kmers_hashed);
int d2g_sketch_batch_packed_dev(d2g_ctx *ctx, const d2g_sketch_params *p,
                                const uint64_t *codes_d, const uint32_t *mask_d, const uint64_t *rec_off_d, const uint32_t *rec_entity_d,
                                uint64_t n_rec, uint32_t n_entities, uint64_t total_len,
                                uint64_t *regs_u64_out_d, double *sig_out_d, double *card_out_d, uint64_t *ids_out_d);

/* ------------------------------------------------------------------------------------------------
 * Compare path.  Replaces compare() (src/cmp_core.cpp:349-575), densify (:577-613) and the
 * row/column orderings of emit_rectangular (src/emitrect.cpp:198-326).
 * ---------------------------------------------------------------------------------------------- */
enum { D2G_SIMILARITY = 0, D2G_CONTAINMENT = 1, D2G_SYMMETRIC_CONTAINMENT = 2, D2G_POISSON_LLR = 3,
       D2G_INTERSECTION = 4, D2G_UNION_SIZE = 5 };
enum { D2G_CMP_GTLT = 0,   /* SetSketch / OPMH registers: count a>b and a<b (cmp_core.cpp:458-494) */
       D2G_CMP_EQ = 1,     /* BagMinHash / ProbMinHash: count bitwise-equal (cmp_core.cpp:495-517) */
       D2G_CMP_SS_COMPRESSED = 2, /* --fastcmp N: log-quantised registers from d2g_make_compressed, gt/lt counts through
                                     g_b with base compressed_b (cmp_core.cpp:425-448) */
       D2G_CMP_BBIT = 3 };        /* --fastcmp N --bbit-sigs: truncated register hashes, equal count with the b-bit collision
                                     correction for regbytes*8 bits (cmp_core.cpp:406-424) */
enum { D2G_SYMMETRIC = 0,  /* condensed upper triangle, rows i<j (emitrect.cpp:290-323) */
       D2G_ASYMMETRIC = 1, /* full n x n (emitrect.cpp:249-268) */
       D2G_PANEL = 2 };    /* rows = first n-nq sketches (-F), cols = last nq (-Q) (emitrect.cpp:229-246) */

typedef struct {
    uint32_t sketchsize;   /* S registers of 8 bytes per sketch */
    int32_t cmp_kind;      /* D2G_CMP_* */
    int32_t measure;       /* D2G_* measure */
    int32_t k;             /* k-mer length, only used by D2G_POISSON_LLR */
    int32_t shape;         /* D2G_SYMMETRIC / D2G_ASYMMETRIC / D2G_PANEL */
    uint64_t n;            /* total sketches */
    uint64_t nq;           /* PANEL: number of query (column) sketches */
    double regbytes;       /* D2G_CMP_BBIT / D2G_CMP_SS_COMPRESSED: --fastcmp register size in bytes (1, 2 or 4); else ignored */
    long double compressed_b; /* D2G_CMP_SS_COMPRESSED: base b of the quantisation (from d2g_make_compressed); else ignored */
    int32_t nlsh;          /* d2g_lsh_topk*: --nLSH, number of table types of the index (src/cmp_core.cpp:757-770); 0 = the reference default 2; 1..3 implemented */
} d2g_cmp_params;

/* In-place densification of OPMH signatures (empty == 0.0), src/cmp_core.cpp:577-613. */
int d2g_densify(d2g_ctx *ctx, double *sig, uint64_t *kmers /*nullable*/, uint64_t n, uint32_t sketchsize);
int d2g_densify_dev(d2g_ctx *ctx, double *sig_d, uint64_t *kmers_d, uint64_t n, uint32_t sketchsize);

/* Register compression, make_compressed (src/cmp_core.cpp:209-322) for --fastcmp N (regbytes 1, 2 or 4) applied to f64
 * registers after sketching.  bbit == 0: SetSketch log-quantisation 1 - log(reg/a)/log(b) clamped to [0, q+1]; *a_io / *b_io
 * <= 0 asks for the data-fitted parameters (optimal_parameters, src/setsketch.cpp:7-10) and returns them.  A degenerate fit
 * falls back to b-bit exactly as the reference does.  bbit != 0: top bits of Wang(register bits ^ 0xa3407fb23cd20ef), or of
 * Wang(kmers[i]) when kmers is given.  The arithmetic is x87 long double on the host in the reference and here.  out[n][S]
 * receives the quantised registers as doubles (exact small integers): every compare entry point takes them unchanged with
 * cmp_kind = D2G_CMP_SS_COMPRESSED (and compressed_b = *b_io) or D2G_CMP_BBIT, according to *bbit_used.  Host pointers. */
int d2g_make_compressed(const double *regs, const uint64_t *kmers /*nullable*/, uint64_t n, uint32_t sketchsize, double regbytes,
                        int32_t bbit, long double *a_io, long double *b_io, double *out, int32_t *bbit_used);

/* Number of float32 values the full output holds for these parameters. */
uint64_t d2g_cmp_output_size(const d2g_cmp_params *p);
/* Rows [row_begin,row_end) of the output, as emit_rectangular would produce them; returns the number
 * of float32 values in that row range through *n_vals. */
int d2g_cmp_rows_size(const d2g_cmp_params *p, uint64_t row_begin, uint64_t row_end, uint64_t *n_vals);

/* Whole matrix, host in / host out (regs f64[n][S], cards f64[n], out f32[d2g_cmp_output_size]). */
int d2g_cmp_matrix(d2g_ctx *ctx, const d2g_cmp_params *p, const double *regs, const double *cards, float *out);
/* Rows [row_begin,row_end) of the matrix, host in / host out (out holds d2g_cmp_rows_size values).  Device->host
 * copies overlap the kernels and land directly in `out` (page-locked `out` moves at full PCIe speed). */
int d2g_cmp_rows(d2g_ctx *ctx, const d2g_cmp_params *p, const double *regs, const double *cards,
                 uint64_t row_begin, uint64_t row_end, float *out);
/* Row range with a sink: results are delivered in row order in blocks (host memory valid only during
 * the callback), so a front-end can stream them to the reference's output format. sink returns 0 to continue. */
typedef int (*d2g_sink_fn)(void *user, const float *block, uint64_t first_row, uint64_t n_rows, uint64_t n_vals);
int d2g_cmp_stream(d2g_ctx *ctx, const d2g_cmp_params *p, const double *regs, const double *cards,
                   uint64_t row_begin, uint64_t row_end, d2g_sink_fn sink, void *user);
/* Device-resident variant: regs_d/cards_d/out_d on the device; computes rows [row_begin,row_end) into
 * out_d (packed, starting at offset 0); asynchronous on the ctx stream. */
int d2g_cmp_rows_dev(d2g_ctx *ctx, const d2g_cmp_params *p, const double *regs_d, const double *cards_d,
                     uint64_t row_begin, uint64_t row_end, float *out_d);
/* Raw integer counts for one tile (rows x cols), for callers that finalise themselves:
 * c0 = #(row > col) (or #equal for D2G_CMP_EQ), c1 = #(row < col). Host pointers. */
int d2g_cmp_counts(d2g_ctx *ctx, uint32_t sketchsize, int32_t cmp_kind,
                   const double *rows, uint64_t n_rows, const double *cols, uint64_t n_cols,
                   uint32_t *c0_out, uint32_t *c1_out);

/* ------------------------------------------------------------------------------------------------
 * Several GPUs.  The reference shards its all-pairs phase over output rows inside one process (src/emitrect.cpp:198-326); here the
 * rows shard over GPUs, each context owning an NCCL communicator (libnccl.so.2 is resolved at run time when one is first asked for).
 *   d2g_init_devices      one process, several devices: contexts + their communicator (d2g_init(ctx**, devices, ndev));
 *   d2g_comm_unique_id /  one process per GPU (torchrun, MPI): rank 0 makes an id, hands it to the others (any transport), every
 *   d2g_comm_init_rank    rank joins with its own context.
 * Sketching needs no communication (files are sharded).  d2g_cmp_rows_sharded_dev is the path's one exchange step: rank r holds the
 * registers of sketches [r * n_per, min(n, (r+1) * n_per)), n_per = ceil(n / nranks), as it produced them; the ranks exchange register
 * POSITIONS (all-to-all), each ranks its share of the positions over all sketches, and the 32-bit ranks are all-gathered -- half the
 * bytes of an f64 all-gather and 1/nranks of the order-code preparation per GPU.  Every rank then computes rows [row_begin, row_end)
 * of the matrix over ALL n sketches into out_d (packed from row_begin).  Collective: every rank of the communicator must call it.
 * ---------------------------------------------------------------------------------------------- */
#define D2G_COMM_ID_BYTES 128
int d2g_init_devices(d2g_ctx **ctxs, const int *devices, int ndev);
int d2g_comm_unique_id(void *id_out /* D2G_COMM_ID_BYTES */);
int d2g_comm_init_rank(d2g_ctx *ctx, int nranks, int rank, const void *id /* D2G_COMM_ID_BYTES */);
int d2g_comm_init_all(d2g_ctx **ctxs, int n);   /* contexts of one process on distinct devices */
int d2g_comm_size(const d2g_ctx *ctx);
int d2g_comm_rank(const d2g_ctx *ctx);
int d2g_comm_destroy(d2g_ctx *ctx);
int d2g_cmp_rows_sharded_dev(d2g_ctx *ctx, const d2g_cmp_params *p, const double *local_regs_d, const double *local_cards_d,
                             uint64_t local_begin, uint64_t local_n, uint64_t row_begin, uint64_t row_end, float *out_d);
/* The same with host memory on both sides (the front-end's --gpus N): this rank's block of registers is uploaded, the exchange runs,
 * and rows [row_begin, row_end) are delivered to the sink in row order like d2g_cmp_stream.  Collective. */
int d2g_cmp_stream_sharded(d2g_ctx *ctx, const d2g_cmp_params *p, const double *local_regs, const double *local_cards,
                           uint64_t local_begin, uint64_t local_n, uint64_t row_begin, uint64_t row_end, d2g_sink_fn sink, void *user);
/* ... and into a caller buffer holding d2g_cmp_rows_size values (device->host copies overlap the kernels, as in d2g_cmp_rows).  Collective. */
int d2g_cmp_rows_sharded(d2g_ctx *ctx, const d2g_cmp_params *p, const double *local_regs, const double *local_cards,
                         uint64_t local_begin, uint64_t local_n, uint64_t row_begin, uint64_t row_end, float *out);

/* ------------------------------------------------------------------------------------------------
 * LSH-assisted top-k neighbour graph (--topk K).  Replaces build_index (src/index_build.cpp:53-165) over
 * SetSketchIndex (src/ssi.h:290-453, default --nLSH 2), refine_results (src/refine.cpp:6-81) and the CSR
 * assembly of emit_neighbors (src/emitnn.cpp:12-52).  Output is the reference's sequential (-p1) result
 * (its multi-threaded output is timing dependent).  regs f64[n][S] (already densified for OPMH), cards f64[n]:
 * host memory.  indptr_out: caller-allocated u64[n+1]; *idx_out / *val_out are malloc'ed (release with
 * d2g_free) and hold indptr_out[n] entries: neighbour ids and similarities (or distances), best first.
 * ---------------------------------------------------------------------------------------------- */
int d2g_lsh_topk(d2g_ctx *ctx, const d2g_cmp_params *p, const double *regs, const double *cards,
                 int32_t topk, uint64_t *indptr_out, uint32_t **idx_out, float **val_out);

/* Lists [row_begin,row_end) of the same graph (indptr_out: u64[row_end-row_begin+1], starting at 0).  The neighbour lists are
 * independent once the candidates of ALL queries are known, so several GPUs each build the (replicated) index, scan all
 * queries and replay / refine / trim only their own range of lists; concatenated in rank order the pieces are the graph. */
int d2g_lsh_topk_rows(d2g_ctx *ctx, const d2g_cmp_params *p, const double *regs, const double *cards,
                      int32_t topk, uint64_t row_begin, uint64_t row_end, uint64_t *indptr_out, uint32_t **idx_out, float **val_out);

/* General form.  index_regs: the registers the LSH index is built over (NULL = regs); regs: the registers refinement compares -- with
 * cmp_kind D2G_CMP_SS_COMPRESSED / D2G_CMP_BBIT the output of d2g_make_compressed, while index_regs are the f64 signatures (--topk with
 * --fastcmp: the reference builds the index before it compresses, src/cmp_core.cpp:741-799, and refine_results goes through the compressed
 * branch of compare(), :362-449).  topk > 0: top-k lists as above.  topk <= 0: similarity-threshold graph (--similarity-threshold x,
 * NN_GRAPH_THRESHOLD; src/options.h:309, src/index_build.cpp:26-31,56-60, src/refine.cpp:43-68): candidate lists without a cap, first
 * arrival keeps its hit count, lists walked in (-hits, id) order keeping measure >= min_similarity (distances: < min_similarity) until
 * 20 consecutive failures, sorted best first.  Threshold graphs hold one uncapped list per query in shared memory: n <= 25601. */
int d2g_lsh_graph(d2g_ctx *ctx, const d2g_cmp_params *p, const double *index_regs, const double *regs, const double *cards,
                  int32_t topk, double min_similarity, uint64_t row_begin, uint64_t row_end,
                  uint64_t *indptr_out, uint32_t **idx_out, float **val_out);

void d2g_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
