// api_cmp.cu -- compare path of the C ABI (include/d2gpu.h): compare(), densify, make_compressed and the orderings of emit_rectangular.
#include "api_internal.h"
#include "cmp_kernels.cuh"
#include "cmp16_kernels.cuh"
#include <cub/device/device_segmented_radix_sort.cuh>

// -------------------------------------------------------------------------------------------------
// compare path
// -------------------------------------------------------------------------------------------------
int check_cmp_params(const d2g_cmp_params *p) {
    if (!p) return fail(D2G_EINVAL, "null params");
    if (p->sketchsize == 0) return fail(D2G_EINVAL, "sketchsize must be > 0");
    if (p->cmp_kind < D2G_CMP_GTLT || p->cmp_kind > D2G_CMP_BBIT) return fail(D2G_EINVAL, "bad cmp_kind %d", p->cmp_kind);
    if (p->cmp_kind >= D2G_CMP_SS_COMPRESSED) {
        if (p->regbytes != 1. && p->regbytes != 2. && p->regbytes != 4.) return fail(D2G_EINVAL, "compressed registers: regbytes must be 1, 2 or 4 (got %g)", p->regbytes);
        if (p->cmp_kind == D2G_CMP_SS_COMPRESSED && !(p->compressed_b > 1.L)) return fail(D2G_EINVAL, "compressed registers: base b must be > 1");
        if (p->sketchsize > 65535) return fail(D2G_EUNSUPPORTED, "compressed registers: sketchsize > 65535 not supported");
    }
    if (p->measure < 0 || p->measure > D2G_UNION_SIZE) return fail(D2G_EINVAL, "bad measure %d", p->measure);
    if (p->shape < 0 || p->shape > D2G_PANEL) return fail(D2G_EINVAL, "bad shape %d", p->shape);
    if (p->shape == D2G_PANEL && p->nq > p->n) return fail(D2G_EINVAL, "nq > n");
    return D2G_OK;
}
namespace {
uint64_t n_rows(const d2g_cmp_params *p) { return p->shape == D2G_PANEL ? p->n - p->nq : p->n; }
uint64_t n_cols(const d2g_cmp_params *p) { return p->shape == D2G_PANEL ? p->nq : p->n; }
uint64_t rows_size(const d2g_cmp_params *p, uint64_t r0, uint64_t r1) {
    if (p->shape == D2G_SYMMETRIC) {
        auto tri = [&](uint64_t i) { return i * p->n - i * (i + 1) / 2; };
        return tri(r1) - tri(r0);
    }
    return (r1 - r0) * n_cols(p);
}
} // namespace

int make_consts(d2g_ctx *c, const d2g_cmp_params *p, d2g::CmpConsts *k) {
    const uint32_t S = p->sketchsize;
    k->invdenom = xf::from_long_double(1.L / S);
    k->eps = xf::from_long_double(1e-15L);
    k->poisson_mult = -1. / std::max(1, p->k);
    k->S = S; k->measure = p->measure; k->cmp_kind = p->cmp_kind;
    k->fast_sim = (p->measure == D2G_SIMILARITY && p->cmp_kind == D2G_CMP_GTLT && (S & (S - 1)) == 0) ? 1 : 0;
    k->eq_llr_lut = nullptr; k->lut80 = nullptr;
    if (p->cmp_kind >= D2G_CMP_SS_COMPRESSED) {
        // Everything that depends on the integer counts alone is x87 long-double arithmetic on the host in the reference
        // (powl in g_b, fmal, logl); tabulate it over the S + 1 possible counts (cmp_core.cpp:323-325,406-432).
        const long double invdenom = 1.L / S;
        std::vector<xf::f80> l80(S + 1);
        std::vector<long double> lv(S + 1);
        if (p->cmp_kind == D2G_CMP_BBIT) {
            const long double b2pow = -ldexpl(1.L, -(int)(p->regbytes * 8.));
            for (uint32_t e = 0; e <= S; ++e) lv[e] = std::max(0.L, fmal((long double)(uint64_t)e, invdenom, b2pow) / (1.L + b2pow));
        } else {
            const long double b = p->compressed_b;
            for (uint32_t e = 0; e <= S; ++e) lv[e] = (1.L - powl(b, -((uint64_t)e * invdenom))) / (1.L - 1.L / b);
        }
        for (uint32_t e = 0; e <= S; ++e) l80[e] = xf::from_long_double(lv[e]);
        if (int rc = c->clut80.reserve((S + 1) * sizeof(xf::f80))) return rc;
        CU(cudaMemcpyAsync(c->clut80.p, l80.data(), (S + 1) * sizeof(xf::f80), cudaMemcpyHostToDevice, c->stream));
        k->lut80 = c->clut80.as<xf::f80>();
        if (p->measure == D2G_POISSON_LLR) {
            auto llr = [&](long double ret) -> float {   // sim2dist on a long double argument (cmp_core.cpp:361) + :573
                ret = ret ? (long double)(double)(logl(2. * ret / (1. + ret)) * k->poisson_mult) : (long double)INFINITY;
                if (isnan(ret) || isinf(ret)) ret = __LDBL_MAX__;
                return (float)ret;
            };
            std::vector<float> lut;
            if (p->cmp_kind == D2G_CMP_BBIT) {
                lut.resize(S + 1);
                for (uint32_t e = 0; e <= S; ++e) lut[e] = llr(lv[e]);
            } else {
                lut.resize((size_t)(S + 1) * (S + 2) / 2);
                auto work = [&](uint32_t g0, uint32_t g1) {
                    for (uint32_t g = g0; g < g1; ++g)
                        for (uint32_t l = 0; l + g <= S; ++l)
                            lut[d2g::tri_index(g, l, S)] = llr(std::max(1.L - (lv[g] + lv[l]), 0.L));
                };
                const unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 32u));
                std::vector<std::thread> th;
                for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t]() { for (uint32_t g = t; g <= S; g += nt) work(g, g + 1); });
                for (auto &x : th) x.join();
            }
            if (int rc = c->clut.reserve(lut.size() * 4)) return rc;
            CU(cudaMemcpyAsync(c->clut.p, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice, c->stream));
            c->lut_S = 0; c->lut_k = -1;              // the equality-branch cache below no longer describes clut
            k->eq_llr_lut = c->clut.as<float>();
        }
        CU(cudaStreamSynchronize(c->stream));           // the host vectors go out of scope
        return D2G_OK;
    }
    if (p->cmp_kind == D2G_CMP_EQ && p->measure == D2G_POISSON_LLR) {
        if (c->lut_S != S || c->lut_k != p->k) {
            // equality branch of the Mash distance is long-double logl on the host in the reference
            // (cmp_core.cpp:361,509); it only depends on the integer count, so tabulate it here.
            std::vector<float> lut(S + 1);
            const long double invdenom = 1.L / S;
            for (uint32_t e = 0; e <= S; ++e) {
                long double ret = invdenom * e;
                ret = ret ? (long double)(double)(logl(2. * ret / (1. + ret)) * k->poisson_mult) : (long double)INFINITY;
                if (isnan(ret) || isinf(ret)) ret = __LDBL_MAX__;
                lut[e] = (float)ret;
            }
            if (int rc = c->clut.reserve((S + 1) * 4)) return rc;
            CU(cudaMemcpyAsync(c->clut.p, lut.data(), (S + 1) * 4, cudaMemcpyHostToDevice, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            c->lut_S = S; c->lut_k = p->k;
        }
        k->eq_llr_lut = c->clut.as<float>();
    }
    return D2G_OK;
}

namespace {
// kinds 0 and 2 count (a > b, a < b); kinds 1 and 3 count bitwise-equal registers
inline bool counts_gtlt(int cmp_kind) { return cmp_kind == D2G_CMP_GTLT || cmp_kind == D2G_CMP_SS_COMPRESSED; }

// f64 tile kernel over rows [r0,r1) x columns [c0,c1) (global sketch ids); `base` carries the output mapping.
int launch_cmp_f64(d2g_ctx *c, const d2g_cmp_params *p, d2g::CmpArgs a, uint64_t r0, uint64_t r1, uint64_t c0, uint64_t c1,
                   const int *use_flag, int want) {
    if (r1 <= r0 || c1 <= c0) return D2G_OK;
    a.row0 = r0; a.row1 = r1; a.col0 = c0; a.col1 = c1; a.use_flag = use_flag; a.want = want;
    const uint64_t tiles_i = (r1 - r0 + d2g::CMP_T - 1) / d2g::CMP_T;
    a.tiles_j = (c1 - c0 + d2g::CMP_T - 1) / d2g::CMP_T;
    const uint64_t grid = tiles_i * a.tiles_j;
    if (grid > 0x7fffffffULL) return fail(D2G_EINVAL, "row block too large for one launch");
    KernelTimer kt(c, use_flag ? D2G_T_CMP_PREP : D2G_T_CMP);
    if (counts_gtlt(p->cmp_kind)) d2g::cmp_tile_kernel<0><<<(unsigned)grid, d2g::CMP_THREADS, 0, c->stream>>>(a);
    else d2g::cmp_tile_kernel<1><<<(unsigned)grid, d2g::CMP_THREADS, 0, c->stream>>>(a);
    c->launches++;
    CU(cudaGetLastError());
    return D2G_OK;
}

__global__ void fill_offsets_kernel(int64_t *offs, uint32_t nseg, uint64_t stride) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nseg) offs[i] = (int64_t)((uint64_t)i * stride);
}

// Per-register-position sort of the job's sketches + dense ranks, in groups of register positions so that one segmented
// sort holds < 2^31 items.  Writes half codes into codes16 (blocked layout) or, when grank != nullptr, the u32 ranks.
int c16_sort_rank(d2g_ctx *c, const d2g_cmp_params *p, d2g::C16Job j, uint16_t *codes16, uint32_t *grank, int *flag) {
    using namespace d2g;
    const uint32_t S = j.S;
    const uint64_t U = (uint64_t)j.nA + j.nB;
    const uint32_t group = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(S, 0x7fffffffULL / std::max<uint64_t>(1, U)));
    const uint64_t items_max = U * group;
    auto al = [](uint64_t b) { return (b + 255) / 256 * 256; };
    uint64_t off = 0;
    const uint64_t o_kA = off; off += al(items_max * 8); const uint64_t o_kB = off; off += al(items_max * 8);
    const uint64_t o_iA = off; off += al(items_max * 4); const uint64_t o_iB = off; off += al(items_max * 4);
    const uint64_t o_offs = off; off += al(((uint64_t)group + 1) * 8);
    if (int rc = c->c16buf.reserve(off)) return rc;
    unsigned char *B = c->c16buf.as<unsigned char>();
    uint64_t *kA = (uint64_t *)(B + o_kA), *kB = (uint64_t *)(B + o_kB);
    uint32_t *iA = (uint32_t *)(B + o_iA), *iB = (uint32_t *)(B + o_iB);
    int64_t *offs = (int64_t *)(B + o_offs);
    cudaStream_t st = c->stream;
    fill_offsets_kernel<<<(group + 1 + 255) / 256, 256, 0, st>>>(offs, group, U);
    c->launches++;
    for (uint32_t s0 = 0; s0 < S; s0 += group) {
        j.s_begin = s0; j.s_count = std::min(group, S - s0);
        const uint64_t items = U * j.s_count;
        const dim3 gk((unsigned)((U + 31) / 32), (j.s_count + 31) / 32);
        if (counts_gtlt(p->cmp_kind)) c16_keys_kernel<0><<<gk, 256, 0, st>>>(j, kA, iA, flag);
        else c16_keys_kernel<1><<<gk, 256, 0, st>>>(j, kA, iA, flag);
        size_t need = 0;
        cub::DeviceSegmentedRadixSort::SortPairs(nullptr, need, kA, kB, iA, iB, (int)items, (int)j.s_count, offs, offs + 1, 0, 64, st);
        if (int rc = c->wtmp.reserve(need + 256)) return rc;
        size_t tbytes = c->wtmp.cap;
        CU(cub::DeviceSegmentedRadixSort::SortPairs(c->wtmp.p, tbytes, kA, kB, iA, iB, (int)items, (int)j.s_count, offs, offs + 1, 0, 64, st));
        c16_rank_kernel<<<j.s_count, 256, 0, st>>>(j, kB, iB, codes16, grank, flag);
        c->launches += 2 + 11;
        CU(cudaGetLastError());
    }
    return D2G_OK;
}

// Multi-job comparisons: rank ALL sketches [g0, g0 + N) once per register position (u32 ranks in HBM); each job then
// derives its own dense codes from them with c16_local_codes_kernel instead of sorting its sketches again.
int c16_build_global(d2g_ctx *c, const d2g_cmp_params *p, const double *regs_d, uint64_t g0, uint64_t N) {
    using namespace d2g;
    auto &g = c->c16g;
    const uint32_t S = p->sketchsize;
    if (g.valid && g.regs == regs_d && g.g0 == g0 && g.N == N && g.S == S && g.kind == p->cmp_kind) return D2G_OK;
    g.valid = false;
    if (int rc = c->c16grank.reserve(N * S * 4)) return rc;
    if (int rc = c->c16flag.reserve(256)) return rc;
    KernelTimer kt(c, D2G_T_CMP_PREP);
    CU(cudaMemsetAsync(c->c16flag.p, 0, 4, c->stream));
    C16Job j{};
    j.regs = regs_d; j.S = S; j.gA0 = g0; j.nA = (uint32_t)N; j.nB = 0; j.posB0 = 0; j.KP = 0;
    if (int rc = c16_sort_rank(c, p, j, nullptr, c->c16grank.as<uint32_t>(), c->c16flag.as<int>())) return rc;
    g.valid = true; g.regs = regs_d; g.g0 = g0; g.N = N; g.S = S; g.kind = p->cmp_kind;
    return D2G_OK;
}

// One comparison job on 16-bit order codes (cmp16_kernels.cuh): the sketches [lo1,hi1) (and [lo2,hi2) when
// hi2 > lo2) are ranked per register position, coded, and rows [r0,r1) x columns [c0,c1) are compared.
// [r0,r1) must lie inside range 1; [c0,c1) inside range 2 when it exists, else inside range 1.
int run_cmp16_job(d2g_ctx *c, const d2g_cmp_params *p, const d2g::CmpArgs &base, uint64_t lo1, uint64_t hi1, uint64_t lo2, uint64_t hi2,
                  uint64_t r0, uint64_t r1, uint64_t c0, uint64_t c1) {
    using namespace d2g;
    if (r1 <= r0 || c1 <= c0) return D2G_OK;
    const uint32_t S = p->sketchsize;
    const bool two = hi2 > lo2;
    C16Job j{};
    j.regs = base.regs; j.S = S; j.gA0 = lo1; j.nA = (uint32_t)(hi1 - lo1); j.gB0 = two ? lo2 : 0; j.nB = two ? (uint32_t)(hi2 - lo2) : 0;
    j.posB0 = (j.nA + C16_BLK - 1) / C16_BLK * C16_BLK;
    j.KP = ((S + 1) / 2 + C16_KC - 1) / C16_KC * C16_KC;
    j.s_begin = 0; j.s_count = S;
    const uint64_t U = (uint64_t)j.nA + j.nB;
    const uint64_t nblocks = (uint64_t)j.posB0 / C16_BLK + (j.nB + C16_BLK - 1) / C16_BLK + 2;   // +2: a row tile reads two blocks
    const uint64_t code_bytes = nblocks * j.KP * C16_BLK * 4;
    if (int rc = c->c16codes.reserve(code_bytes)) return rc;
    if (int rc = c->c16flag.reserve(256)) return rc;
    int *flag = c->c16flag.as<int>();
    cudaStream_t st = c->stream;
    auto &cc = c->c16cache;
    auto &gl = c->c16g;
    // gt/lt registers with power-of-two S and no raw counts wanted: the != count alone determines the result
    const bool pow2 = (S & (S - 1)) == 0;
    const int mode = (!counts_gtlt(p->cmp_kind) || (p->cmp_kind == D2G_CMP_GTLT && pow2 && !base.c0_out && !getenv("D2G_C16_NO_NE"))) ? 1 : 0;
    const bool cached = cc.valid && cc.regs == base.regs && cc.lo1 == lo1 && cc.hi1 == hi1 && cc.lo2 == lo2 && cc.hi2 == hi2 && cc.S == S && cc.kind == p->cmp_kind && cc.mode == mode;
    if (!cached) {
        KernelTimer kt(c, D2G_T_CMP_PREP);
        cc.valid = true; cc.regs = base.regs; cc.lo1 = lo1; cc.hi1 = hi1; cc.lo2 = lo2; cc.hi2 = hi2; cc.S = S; cc.kind = p->cmp_kind; cc.mode = mode;
        CU(cudaMemsetAsync(c->c16codes.p, 0, code_bytes, st));
        const bool use_hash = mode == 1 && U <= C16_HASH_MAX_SKETCHES && !getenv("D2G_C16_NO_HASH") && !c->c16_sharded;
        if (use_hash) {
            // != only: injective codes suffice -> open-addressing table per register position, no sort
            const uint32_t TS = (uint32_t)std::max<uint64_t>(64, U + U / 2);
            const size_t smem = (size_t)TS * 8;
            // the flag also carries "a register of the global ranking was NaN" for the jobs that derive their codes from it: while those
            // ranks are live it is only ever raised, never cleared (a set flag costs a redundant f64 pass, a lost one wrong counts)
            if (!gl.valid) CU(cudaMemsetAsync(flag, 0, 4, st));
            if (counts_gtlt(p->cmp_kind)) {
                CU(cudaFuncSetAttribute(c16_hash_codes_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                c16_hash_codes_kernel<0><<<(S + 3) / 4, 512, smem, st>>>(j, TS, c->c16codes.as<uint16_t>(), flag);
            } else {
                CU(cudaFuncSetAttribute(c16_hash_codes_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                c16_hash_codes_kernel<1><<<(S + 3) / 4, 512, smem, st>>>(j, TS, c->c16codes.as<uint16_t>(), flag);
            }
            c->launches++;
            CU(cudaGetLastError());
        }
        const bool use_global = !use_hash && gl.valid && gl.regs == base.regs && gl.S == S && gl.kind == p->cmp_kind && lo1 >= gl.g0 && hi1 <= gl.g0 + gl.N &&
                                (!two || (lo2 >= gl.g0 && hi2 <= gl.g0 + gl.N));
        if (use_global) {
            const size_t smem = 2 * ((gl.N + 31) / 32) * 4;
            CU(cudaFuncSetAttribute(c16_local_codes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            c16_local_codes_kernel<<<S, 256, smem, st>>>(j, c->c16grank.as<uint32_t>(), gl.g0, (uint32_t)gl.N, c->c16codes.as<uint16_t>());
            c->launches++;
            CU(cudaGetLastError());
        } else if (!use_hash) {
            if (!gl.valid) CU(cudaMemsetAsync(flag, 0, 4, st));
            if (int rc = c16_sort_rank(c, p, j, c->c16codes.as<uint16_t>(), nullptr, flag)) return rc;
        }
    }
    (void)U;
    C16Args a;
    a.codes = c->c16codes.as<uint32_t>(); a.KP = j.KP;
    auto view = [&](uint64_t lo, uint64_t hi, uint64_t rlo, uint32_t rpos0, uint32_t &blk0, uint32_t &n, uint64_t &g0) {
        const uint64_t blk = (lo - rlo) / C16_BLK;
        g0 = rlo + blk * C16_BLK; blk0 = rpos0 / C16_BLK + (uint32_t)blk; n = (uint32_t)(hi - g0);
    };
    view(r0, r1, lo1, 0, a.a_blk0, a.n_a, a.gi0);
    if (two) view(c0, c1, lo2, j.posB0, a.b_blk0, a.n_b, a.gj0);
    else view(c0, c1, lo1, 0, a.b_blk0, a.n_b, a.gj0);
    a.o = base; a.o.row0 = r0; a.o.row1 = r1; a.o.col0 = c0; a.o.col1 = c1; a.o.use_flag = nullptr; a.o.want = 0;
    a.use_flag = flag; a.want = 0;
    const uint64_t tiles_i = (a.n_a + C16_TM - 1) / C16_TM;
    a.tiles_j = (a.n_b + C16_TN - 1) / C16_TN;
    const uint64_t grid = tiles_i * a.tiles_j;
    if (grid > 0x7fffffffULL) return fail(D2G_EINVAL, "comparison job too large for one launch");
    {
        a.ne_is_gt = (mode == 1 && p->cmp_kind == D2G_CMP_GTLT) ? 1 : 0;
        a.one = 1;
        int acc = 1;
        if (const char *ev = getenv("D2G_C16_ACC")) acc = atoi(ev);
        KernelTimer kt(c, D2G_T_CMP);
        auto go = [&](auto kern) -> int {
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C16_SMEM));
            kern<<<(unsigned)grid, C16_THREADS, C16_SMEM, st>>>(a);
            return D2G_OK;
        };
        int rc;
        if (mode == 0) rc = acc == 0 ? go(cmp16_tile_kernel<0, 0>) : go(cmp16_tile_kernel<0, 1>);
        else rc = acc == 0 ? go(cmp16_tile_kernel<1, 0>) : go(cmp16_tile_kernel<1, 1>);
        if (rc) return rc;
        c->launches++;
        CU(cudaGetLastError());
    }
    // registers the codes cannot express (NaN): the f64 kernel recomputes the job, gated on the device flag
    return launch_cmp_f64(c, p, base, r0, r1, c0, c1, flag, 1);
}

// rows [r0,r1) of the output into out_d / counts (packed from row r0)
int launch_cmp(d2g_ctx *c, const d2g_cmp_params *p, const d2g::CmpConsts &k, const double *regs_d, const double *cards_d,
               uint64_t r0, uint64_t r1, float *out_d, uint32_t *c0_d, uint32_t *c1_d, uint64_t sym_lo = ~0ULL) {
    if (r1 <= r0) return D2G_OK;
    d2g::CmpArgs a{};
    a.regs = regs_d; a.cards = cards_d; a.n = p->n; a.out_row0 = r0;
    a.col_base = p->shape == D2G_PANEL ? p->n - p->nq : 0;
    a.ncols = n_cols(p); a.shape = p->shape; a.out = out_d; a.c0_out = c0_d; a.c1_out = c1_d; a.c = k;
    if (a.ncols == 0) return D2G_OK;
    const uint32_t S = p->sketchsize;
    // columns this row range needs; sym_lo lets successive row blocks of one call share one code space
    const uint64_t cb = p->shape == D2G_SYMMETRIC ? std::min(r0, sym_lo) : a.col_base, ce = a.col_base + a.ncols;
    const uint64_t nR = r1 - r0, nC = ce - cb;
    // path choice: codes pay a per-job sort of the registers, worth it from ~4e9 register comparisons on; jobs that only
    // count != and fit the shared-memory table (run_cmp16_job) build their codes in one cheap pass: from ~2e8 on
    const bool ne_only = !counts_gtlt(p->cmp_kind) || (p->cmp_kind == D2G_CMP_GTLT && (S & (S - 1)) == 0 && !c0_d);
    const bool hashable = ne_only && ((cb <= r0 && r1 <= ce) ? nC : nR + nC) <= d2g::C16_HASH_MAX_SKETCHES;
    int path = (double)nR * (double)nC * (double)S >= (hashable ? 2.0e8 : 4.0e9) ? 1 : 0;
    if (const char *ev = getenv("D2G_CMP_PATH")) path = !strcmp(ev, "codes") ? 1 : (!strcmp(ev, "f64") ? 0 : path);
    uint64_t M = 63232;                                                 // sketches per job: <= 63487 ranks, multiple of 128
    if (const char *ev = getenv("D2G_C16_MAXJOB")) M = std::max<uint64_t>(256, std::min<uint64_t>(M, strtoull(ev, nullptr, 10) / 128 * 128));  // test knob
    M = std::min<uint64_t>(M, (0x7fffffffULL / S) / 128 * 128);          // one segmented sort holds < 2^31 items
    if (S > 65535 || M < 256) path = 0;                                  // 16-bit counters / degenerate blocks
    if (c->c16_sharded) {                                                 // the registers of the other ranks are not here: order codes only
        if (!path && (S > 65535 || M < 256)) return fail(D2G_EUNSUPPORTED, "sharded comparison needs sketchsize <= 65535");
        path = 1;
    }
    if (!path) return launch_cmp_f64(c, p, a, r0, r1, cb, ce, nullptr, 0);
    auto up64 = [](uint64_t x) { return (x + 63) / 64 * 64; };
    // Decomposition into jobs of at most Mj sketches: one range when the rows lie inside the columns, two ranges
    // otherwise, block pairs when that is too many sketches (diagonal blocks of a symmetric comparison share one range).
    struct Job { uint64_t lo1, hi1, lo2, hi2, r0, r1, c0, c1; };
    auto plan = [&](uint64_t Mj, std::vector<Job> &jobs) {
        jobs.clear();
        if (cb <= r0 && r1 <= ce && nC <= Mj) { jobs.push_back({cb, ce, 0, 0, r0, r1, cb, ce}); return; }
        if (up64(nR) + nC <= Mj) { jobs.push_back({r0, r1, cb, ce, r0, r1, cb, ce}); return; }
        uint64_t BR, BC;
        if (p->shape == D2G_SYMMETRIC) BR = BC = Mj / 2;
        else if (up64(nR) <= Mj / 2) { BR = nR; BC = (Mj - up64(nR)) / 64 * 64; }
        else if (nC <= Mj / 2) { BC = nC; BR = (Mj - nC) / 128 * 128; }
        else BR = BC = Mj / 2;
        for (uint64_t rb = r0; rb < r1; rb += BR) {
            const uint64_t re = std::min(r1, rb + BR);
            for (uint64_t cc = cb; cc < ce; cc += BC) {
                const uint64_t cf = std::min(ce, cc + BC);
                if (p->shape == D2G_SYMMETRIC && cf <= rb + 1) continue;     // block entirely on/below the diagonal
                if (cc == rb && cf == re) jobs.push_back({rb, re, 0, 0, rb, re, cc, cf});
                else jobs.push_back({rb, re, cc, cf, rb, re, cc, cf});
            }
        }
    };
    auto job_sketches = [](const Job &j) { return (j.hi1 - j.lo1) + (j.hi2 - j.lo2); };
    std::vector<Job> jobs, hjobs;
    plan(M, jobs);
    const uint64_t g0 = std::min(r0, cb), g1 = std::max(r1, ce), N = g1 - g0;
    bool use_global = jobs.size() > 1 && 2 * ((N + 31) / 32) * 4 <= 200 * 1024 && N < 0xFFFFFFF0ULL && !getenv("D2G_C16_NO_GLOBAL");   // bitmap + prefix in shared memory
    if (c->c16_sharded) use_global = true;                                // ranks of all sketches were gathered: every job derives its codes from them
    if (ne_only && !getenv("D2G_C16_NO_HASH") && M > d2g::C16_HASH_MAX_SKETCHES && !c->c16_sharded) {
        // != only: jobs of <= 16 384 sketches take their codes from a hash table (no sort).  Pick the cheaper plan with measured
        // per-element costs (B200, S=4096): segmented sort + ranks 1.2e-10 s, local codes from global ranks 0.8e-11 s, hash codes 2.8e-11 s.
        plan(d2g::C16_HASH_MAX_SKETCHES, hjobs);
        double cost_sorted = 0, cost_hashed = 0;
        for (const Job &j : jobs) cost_sorted += (double)job_sketches(j) * S * (use_global ? 0.8e-11 : 1.2e-10);
        if (use_global) cost_sorted += (double)N * S * 1.2e-10;
        if (jobs.size() == 1 && job_sketches(jobs[0]) <= d2g::C16_HASH_MAX_SKETCHES) cost_sorted = 1e30;   // hashed inside run_cmp16_job anyway
        for (const Job &j : hjobs) cost_hashed += (double)job_sketches(j) * S * 2.8e-11 + 15e-6;
        if (cost_hashed < cost_sorted) { jobs.swap(hjobs); use_global = false; }
    }
    if (use_global && !c->c16_sharded) { if (int rc = c16_build_global(c, p, regs_d, g0, N)) return rc; }
    for (const Job &j : jobs)
        if (int rc = run_cmp16_job(c, p, a, j.lo1, j.hi1, j.lo2, j.hi2, j.r0, j.r1, j.c0, j.c1)) return rc;
    return D2G_OK;
}

} // namespace

extern "C" int d2g_make_compressed(const double *regs, const uint64_t *kmers, uint64_t n, uint32_t S, double fd, int32_t bbit,
                                   long double *a_io, long double *b_io, double *out, int32_t *bbit_used) {
    if (!regs || !out || !a_io || !b_io) return fail(D2G_EINVAL, "null argument");
    if (fd != 1. && fd != 2. && fd != 4.) return fail(D2G_EINVAL, "regbytes must be 1, 2 or 4 (got %g)", fd);
    const uint64_t nsigs = n * S;
    long double a = *a_io, b = *b_io;
    auto parallel = [&](auto fn) {
        unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 64u));
        if (nsigs < (1u << 16)) nt = 1;
        std::vector<std::thread> th;
        const uint64_t per = (nsigs + nt - 1) / nt;
        for (unsigned t = 0; t < nt; ++t) { const uint64_t lo = t * per, hi = std::min(nsigs, lo + per); if (lo < hi) th.emplace_back(fn, lo, hi); }
        for (auto &x : th) x.join();
    };
    if (!bbit) {
        const long double q = fd == 1. ? 254.3 : fd == 2. ? 65534 : 4294967294;   // double literals, as in the reference (cmp_core.cpp:249)
        if (a <= 0. || b <= 0.) {                       // cmp_core.cpp:250-266
            double minreg = __DBL_MAX__, maxreg = -__DBL_MAX__;
            for (uint64_t i = 0; i < nsigs; ++i) {
                const double v = regs[i];
                if (v <= 0 || v == __DBL_MAX__) continue;
                minreg = std::min(minreg, v); maxreg = std::max(maxreg, v);
            }
            long double mx = minreg, mn = maxreg;       // optimal_parameters(minreg, maxreg, q), src/setsketch.h:563-566
            if (mx < mn) std::swap(mx, mn);
            b = expl(logl(mx / mn) / q);                // src/setsketch.cpp:7-10
            a = mx / b;
        }
        if (a == 0. || isinf(b)) bbit = 1;              // cmp_core.cpp:267-270
        else {
            *a_io = a; *b_io = b;
            const long double logbinv = 1.L / log1pl(b - 1.L);
            const int64_t top = (int64_t)(q + 1);
            parallel([&](uint64_t lo, uint64_t hi) {
                for (uint64_t i = lo; i < hi; ++i) {
                    const long double sub = 1.L - logl((long double)regs[i] / a) * logbinv;
                    // static_cast<int64_t>(long double) as x86 executes it: out of range / NaN -> INT64_MIN
                    int64_t isub = (sub > -9223372036854775809.0L && sub < 9223372036854775808.0L) ? (int64_t)sub : INT64_MIN;
                    out[i] = (double)std::max<int64_t>(0, std::min(top, isub));
                }
            });
            if (bbit_used) *bbit_used = 0;
            return D2G_OK;
        }
    }
    const int shift = fd == 1. ? 58 : fd == 2. ? 48 : 32;   // cmp_core.cpp:306-320
    parallel([&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            uint64_t v;
            if (kmers) v = d2g::wang64(kmers[i]);
            else { memcpy(&v, regs + i, 8); v = d2g::wang64(v ^ 0xa3407fb23cd20efULL); }   // reg2sig, cmp_core.cpp:19-37
            out[i] = (double)(v >> shift);
        }
    });
    if (bbit_used) *bbit_used = 1;
    return D2G_OK;
}

extern "C" {

uint64_t d2g_cmp_output_size(const d2g_cmp_params *p) { return p ? rows_size(p, 0, n_rows(p)) : 0; }

int d2g_cmp_rows_size(const d2g_cmp_params *p, uint64_t r0, uint64_t r1, uint64_t *n_vals) {
    if (int rc = check_cmp_params(p)) return rc;
    if (r0 > r1 || r1 > n_rows(p)) return fail(D2G_EINVAL, "bad row range");
    if (n_vals) *n_vals = rows_size(p, r0, r1);
    return D2G_OK;
}

int d2g_densify_dev(d2g_ctx *c, double *sig_d, uint64_t *kmers_d, uint64_t n, uint32_t S) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (!n || !S) return D2G_OK;
    CU(cudaSetDevice(c->device));
    const uint64_t tot = n * S;
    if (int rc = c->ctmp.reserve(tot * 8)) return rc;
    if (kmers_d) if (int rc = c->cktmp.reserve(tot * 8)) return rc;
    d2g::densify_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(sig_d, kmers_d, n, S, c->ctmp.as<double>(), c->cktmp.as<uint64_t>());
    c->launches++;
    CU(cudaMemcpyAsync(sig_d, c->ctmp.p, tot * 8, cudaMemcpyDeviceToDevice, c->stream));
    if (kmers_d) CU(cudaMemcpyAsync(kmers_d, c->cktmp.p, tot * 8, cudaMemcpyDeviceToDevice, c->stream));
    return D2G_OK;
}

int d2g_densify(d2g_ctx *c, double *sig, uint64_t *kmers, uint64_t n, uint32_t S) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (!n || !S) return D2G_OK;
    CU(cudaSetDevice(c->device));
    const uint64_t tot = n * S;
    if (int rc = c->cregs.reserve(tot * 8)) return rc;
    CU(cudaMemcpyAsync(c->cregs.p, sig, tot * 8, cudaMemcpyHostToDevice, c->stream));
    if (kmers) { if (int rc = c->ids.reserve(tot * 8)) return rc; CU(cudaMemcpyAsync(c->ids.p, kmers, tot * 8, cudaMemcpyHostToDevice, c->stream)); }
    if (int rc = d2g_densify_dev(c, c->cregs.as<double>(), kmers ? c->ids.as<uint64_t>() : nullptr, n, S)) return rc;
    CU(cudaMemcpyAsync(sig, c->cregs.p, tot * 8, cudaMemcpyDeviceToHost, c->stream));
    if (kmers) CU(cudaMemcpyAsync(kmers, c->ids.p, tot * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return D2G_OK;
}

int d2g_cmp_rows_dev(d2g_ctx *c, const d2g_cmp_params *p, const double *regs_d, const double *cards_d,
                     uint64_t r0, uint64_t r1, float *out_d) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_cmp_params(p)) return rc;
    if (r0 > r1 || r1 > n_rows(p)) return fail(D2G_EINVAL, "bad row range");
    CU(cudaSetDevice(c->device));
    c->c16cache.valid = false; c->c16g.valid = false;
    d2g::CmpConsts k;
    if (int rc = make_consts(c, p, &k)) return rc;
    return launch_cmp(c, p, k, regs_d, cards_d, r0, r1, out_d, nullptr, nullptr);
}

// Rows [r0,r1) in row blocks of <= ~64M values.  Kernels run on the ctx stream, the device->host copies on the copy
// stream (block b+1 computes while block b drains).  direct_out != nullptr: results go straight into the caller's
// buffer (full speed when it is pinned); otherwise through two pinned staging buffers to the sink, in row order.
// shard != nullptr: regs / cards are this rank's block of sketches only (sharded_prepare runs the exchange; collective)
struct ShardArg { uint64_t local_begin, local_n; };
static int sharded_prepare_fwd(d2g_ctx *c, const d2g_cmp_params *p, const double *local_regs_d, const double *local_cards_d, uint64_t local_begin, uint64_t local_n);
static int cmp_blocks(d2g_ctx *c, const d2g_cmp_params *p, const double *regs, const double *cards,
                      uint64_t r0, uint64_t r1, d2g_sink_fn sink, void *user, float *direct_out, const ShardArg *shard = nullptr) {
    CU(cudaSetDevice(c->device));
    c->c16cache.valid = false; c->c16g.valid = false;
    const uint32_t S = p->sketchsize;
    const uint64_t n_up = shard ? shard->local_n : p->n;
    if (int rc = c->cregs.reserve(n_up * S * 8 + 8)) return rc;
    if (int rc = c->ccards.reserve(n_up * 8 + 8)) return rc;
    if (n_up) {
        CU(cudaMemcpyAsync(c->cregs.p, regs, n_up * S * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->ccards.p, cards, n_up * 8, cudaMemcpyHostToDevice, c->stream));
    }
    d2g::CmpConsts k;
    if (int rc = make_consts(c, p, &k)) return rc;
    const double *regs_d = c->cregs.as<double>(), *cards_d = c->ccards.as<double>();
    struct ShardedScope { d2g_ctx *c; ~ShardedScope() { c->c16_sharded = false; } } scope{c};
    if (shard) {
        if (int rc = sharded_prepare_fwd(c, p, regs_d, cards_d, shard->local_begin, shard->local_n)) return rc;
        c->c16_sharded = true;
        regs_d = c->c16g.regs; cards_d = c->xcards.as<double>();
    }
    const uint64_t max_vals = 64ULL << 20;
    const uint64_t ncol = n_cols(p);
    uint64_t rows_per = std::max<uint64_t>(d2g::CMP_T, max_vals / std::max<uint64_t>(1, ncol) / d2g::CMP_T * d2g::CMP_T);
    uint64_t cap_vals = 0;
    for (uint64_t b = r0; b < r1; b += rows_per) cap_vals = std::max(cap_vals, rows_size(p, b, std::min(r1, b + rows_per)));
    if (int rc = c->cout.reserve(2 * cap_vals * 4)) return rc;
    if (!direct_out) {
        if (int rc = c->pin[0].reserve(cap_vals * 4)) return rc;
        if (int rc = c->pin[1].reserve(cap_vals * 4)) return rc;
    }
    struct Pending { uint64_t b0, b1, nv; int slot; bool live; } pend{0, 0, 0, 0, false};
    uint64_t done_vals = 0, iblk = 0;
    for (uint64_t b = r0; b < r1; b += rows_per, ++iblk) {
        const int slot = (int)(iblk & 1);
        const uint64_t e = std::min(r1, b + rows_per), nv = rows_size(p, b, e);
        float *out_d = c->cout.as<float>() + (uint64_t)slot * cap_vals;
        if (iblk >= 2) CU(cudaStreamWaitEvent(c->stream, c->evd[slot], 0));       // the copy of block b-2 has left this slot
        if (int rc = launch_cmp(c, p, k, regs_d, cards_d, b, e, out_d, nullptr, nullptr, r0)) return rc;
        CU(cudaEventRecord(c->ev[slot], c->stream));
        CU(cudaStreamWaitEvent(c->copy_stream, c->ev[slot], 0));
        float *dst = direct_out ? direct_out + done_vals : (float *)c->pin[slot].p;
        CU(cudaMemcpyAsync(dst, out_d, nv * 4, cudaMemcpyDeviceToHost, c->copy_stream));
        CU(cudaEventRecord(c->evd[slot], c->copy_stream));
        done_vals += nv;
        if (!direct_out) {
            if (pend.live) {
                CU(cudaEventSynchronize(c->evd[pend.slot]));
                if (sink(user, (const float *)c->pin[pend.slot].p, pend.b0, pend.b1 - pend.b0, pend.nv)) return fail(D2G_EIO, "sink aborted");
            }
            pend = {b, e, nv, slot, true};
        }
    }
    if (pend.live) {
        CU(cudaEventSynchronize(c->evd[pend.slot]));
        if (sink(user, (const float *)c->pin[pend.slot].p, pend.b0, pend.b1 - pend.b0, pend.nv)) return fail(D2G_EIO, "sink aborted");
    }
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->stream));
    return D2G_OK;
}

int d2g_cmp_stream(d2g_ctx *c, const d2g_cmp_params *p, const double *regs, const double *cards,
                   uint64_t r0, uint64_t r1, d2g_sink_fn sink, void *user) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_cmp_params(p)) return rc;
    if (r0 > r1 || r1 > n_rows(p)) return fail(D2G_EINVAL, "bad row range");
    if (!sink) return fail(D2G_EINVAL, "null sink");
    if (r0 == r1 || p->n == 0) return D2G_OK;
    return cmp_blocks(c, p, regs, cards, r0, r1, sink, user, nullptr);
}

int d2g_cmp_rows(d2g_ctx *c, const d2g_cmp_params *p, const double *regs, const double *cards,
                 uint64_t r0, uint64_t r1, float *out) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_cmp_params(p)) return rc;
    if (r0 > r1 || r1 > n_rows(p)) return fail(D2G_EINVAL, "bad row range");
    if (r0 == r1 || p->n == 0) return D2G_OK;
    if (!out) return fail(D2G_EINVAL, "null output");
    return cmp_blocks(c, p, regs, cards, r0, r1, nullptr, nullptr, out);
}

int d2g_cmp_matrix(d2g_ctx *c, const d2g_cmp_params *p, const double *regs, const double *cards, float *out) {
    if (int rc = check_cmp_params(p)) return rc;
    return d2g_cmp_rows(c, p, regs, cards, 0, n_rows(p), out);
}

int d2g_cmp_counts(d2g_ctx *c, uint32_t S, int32_t cmp_kind, const double *rows, uint64_t nr, const double *cols, uint64_t nc,
                   uint32_t *c0_out, uint32_t *c1_out) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (!nr || !nc) return D2G_OK;
    CU(cudaSetDevice(c->device));
    d2g_cmp_params p{};
    p.sketchsize = S; p.cmp_kind = cmp_kind; p.measure = D2G_SIMILARITY; p.k = 31; p.shape = D2G_PANEL; p.n = nr + nc; p.nq = nc;
    if (int rc = check_cmp_params(&p)) return rc;
    c->c16cache.valid = false; c->c16g.valid = false;
    if (int rc = c->cregs.reserve(p.n * S * 8)) return rc;
    CU(cudaMemcpyAsync(c->cregs.p, rows, nr * S * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->cregs.as<double>() + nr * S, cols, nc * S * 8, cudaMemcpyHostToDevice, c->stream));
    if (int rc = c->cout.reserve(2 * nr * nc * 4)) return rc;
    uint32_t *c0_d = c->cout.as<uint32_t>(), *c1_d = c0_d + nr * nc;
    d2g::CmpConsts k;
    if (int rc = make_consts(c, &p, &k)) return rc;
    if (int rc = launch_cmp(c, &p, k, c->cregs.as<double>(), nullptr, 0, nr, nullptr, c0_d, cmp_kind == D2G_CMP_GTLT ? c1_d : nullptr)) return rc;
    if (c0_out) CU(cudaMemcpyAsync(c0_out, c0_d, nr * nc * 4, cudaMemcpyDeviceToHost, c->stream));
    if (c1_out && cmp_kind == D2G_CMP_GTLT) CU(cudaMemcpyAsync(c1_out, c1_d, nr * nc * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return D2G_OK;
}

} // extern "C"

// -------------------------------------------------------------------------------------------------
// sharded comparison: sketches live on the rank that made them, one exchange step over NCCL
// -------------------------------------------------------------------------------------------------
#include "nccl_dl.h"
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(D2G_ECUDA, "%s failed: %s", #call, d2g_nccl::g_api.GetErrorString(r_)); } while (0)

namespace {
// Collective.  Rank r holds sketches [r * n_per, min(n, (r+1) * n_per)), n_per = ceil(n / nranks).  What the tile kernel needs of a
// sketch is not its registers but, per register position, the rank of the value among all n sketches (cmp16_kernels.cuh).  So:
//   1. all-to-all (grouped ncclSend / ncclRecv): rank r receives register positions [r * Sp, (r+1) * Sp) of every sketch, Sp = ceil(S / nranks);
//   2. rank r sorts its Sp positions over all n sketches (1 / nranks of the single-GPU preparation) -> u32 ranks [Sp][n_pad];
//   3. ncclAllGather of the rank slices (4 bytes per register instead of the 8 of an f64 all-gather) and of the cardinalities.
// Afterwards ctx->c16g describes the gathered ranks and launch_cmp derives every job's 16-bit codes from them.
int sharded_prepare(d2g_ctx *c, const d2g_cmp_params *p, const double *local_regs_d, const double *local_cards_d, uint64_t local_begin, uint64_t local_n) {
    using namespace d2g;
    if (!c->nccl_comm) return fail(D2G_EINVAL, "this context owns no communicator (d2g_comm_init_rank / d2g_init_devices)");
    const uint64_t n = p->n, R = (uint64_t)c->nranks, r = (uint64_t)c->rank;
    const uint32_t S = p->sketchsize;
    if (S > 65535) return fail(D2G_EUNSUPPORTED, "sharded comparison needs sketchsize <= 65535");
    if (p->cmp_kind >= D2G_CMP_SS_COMPRESSED && false) return fail(D2G_EUNSUPPORTED, "unreachable");
    const uint64_t n_per = (n + R - 1) / R, n_pad = n_per * R;
    const uint64_t exp_begin = std::min(n, r * n_per), exp_n = std::min(n, (r + 1) * n_per) - exp_begin;
    if (local_begin != exp_begin || local_n != exp_n)
        return fail(D2G_EINVAL, "rank %d of %d must hold sketches [%llu, %llu) of %llu (got [%llu, %llu))", c->rank, c->nranks, (unsigned long long)exp_begin,
                    (unsigned long long)(exp_begin + exp_n), (unsigned long long)n, (unsigned long long)local_begin, (unsigned long long)(local_begin + local_n));
    if (2 * ((n_pad + 31) / 32) * 4 > 200 * 1024) return fail(D2G_EUNSUPPORTED, "sharded comparison: too many sketches (%llu) for the shared-memory rank bitmap", (unsigned long long)n);
    const uint32_t Sp = (uint32_t)((S + R - 1) / R);
    const uint64_t blk = n_per * Sp;                                      // doubles per (source rank, destination rank) block
    cudaStream_t st = c->stream;
    ncclComm_t comm = static_cast<ncclComm_t>(c->nccl_comm);
    const auto &nc = d2g_nccl::g_api;
    if (int rc = c->xsend.reserve(blk * R * 8)) return rc;
    if (int rc = c->xrecv.reserve(blk * R * 8)) return rc;
    if (int rc = c->xrank.reserve((uint64_t)Sp * n_pad * 4)) return rc;
    if (int rc = c->c16grank.reserve((uint64_t)Sp * R * n_pad * 4)) return rc;
    if (int rc = c->xcards.reserve(n_pad * 8 + n_per * 8)) return rc;
    if (int rc = c->c16flag.reserve(256)) return rc;
    KernelTimer kt(c, D2G_T_CMP_PREP);
    double *send = c->xsend.as<double>(), *recv = c->xrecv.as<double>();
    // 1. my registers, split by destination: block d = positions [d * Sp, (d+1) * Sp) of my sketches as an [n_per][Sp] matrix (zero padded)
    CU(cudaMemsetAsync(send, 0, blk * R * 8, st));
    for (uint64_t d = 0; d < R; ++d) {
        const uint64_t c0 = d * Sp;
        if (c0 >= S || !local_n) continue;
        const uint64_t wcols = std::min<uint64_t>(Sp, S - c0);
        CU(cudaMemcpy2DAsync(send + d * blk, (size_t)Sp * 8, local_regs_d + c0, (size_t)S * 8, (size_t)wcols * 8, (size_t)local_n, cudaMemcpyDeviceToDevice, st));
    }
    NC(nc.GroupStart());
    for (uint64_t d = 0; d < R; ++d) {
        NC(nc.Send(send + d * blk, blk, ncclFloat64, (int)d, comm, st));
        NC(nc.Recv(recv + d * blk, blk, ncclFloat64, (int)d, comm, st));
    }
    NC(nc.GroupEnd());
    // 2. recv is the [n_pad][Sp] matrix of my positions over all sketches: rank every position
    CU(cudaMemsetAsync(c->c16flag.p, 0, 4, st));
    C16Job j{};
    j.regs = recv; j.S = Sp; j.gA0 = 0; j.nA = (uint32_t)n_pad; j.nB = 0; j.posB0 = 0; j.KP = 0;
    if (int rc = c16_sort_rank(c, p, j, nullptr, c->xrank.as<uint32_t>(), c->c16flag.as<int>())) return rc;
    // 3. gather the rank slices, the cardinalities and the "a register was NaN" flag
    NC(nc.AllGather(c->xrank.p, c->c16grank.p, (size_t)Sp * n_pad, ncclUint32, comm, st));
    double *cards_all = c->xcards.as<double>(), *cards_mine = cards_all + n_pad;
    CU(cudaMemsetAsync(cards_mine, 0, n_per * 8, st));
    if (local_n) CU(cudaMemcpyAsync(cards_mine, local_cards_d, local_n * 8, cudaMemcpyDeviceToDevice, st));
    NC(nc.AllGather(cards_mine, cards_all, n_per, ncclFloat64, comm, st));
    NC(nc.AllReduce(c->c16flag.p, c->c16flag.p, 1, ncclInt32, ncclMax, comm, st));
    int h_flag = 0;
    CU(cudaMemcpyAsync(&h_flag, c->c16flag.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (h_flag) return fail(D2G_EUNSUPPORTED, "sharded comparison: a register is NaN; gather the registers and use d2g_cmp_rows_dev (its f64 kernel handles NaN as the reference does)");
    auto &g = c->c16g;
    g.valid = true; g.regs = reinterpret_cast<const double *>(c->c16grank.p); g.g0 = 0; g.N = n_pad; g.S = S; g.kind = p->cmp_kind;
    return D2G_OK;
}
} // namespace

static int sharded_prepare_fwd(d2g_ctx *c, const d2g_cmp_params *p, const double *local_regs_d, const double *local_cards_d, uint64_t local_begin, uint64_t local_n) {
    return sharded_prepare(c, p, local_regs_d, local_cards_d, local_begin, local_n);
}

// Host in / host out variant for a front-end whose devices each hold the sketches they made: the local block is uploaded, the exchange
// runs, and rows [row_begin, row_end) are streamed to the sink exactly as d2g_cmp_stream does.  Collective.
extern "C" int d2g_cmp_stream_sharded(d2g_ctx *c, const d2g_cmp_params *p, const double *local_regs, const double *local_cards,
                                      uint64_t local_begin, uint64_t local_n, uint64_t r0, uint64_t r1, d2g_sink_fn sink, void *user) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_cmp_params(p)) return rc;
    if (r0 > r1 || r1 > n_rows(p)) return fail(D2G_EINVAL, "bad row range");
    if (!sink) return fail(D2G_EINVAL, "null sink");
    const ShardArg sh{local_begin, local_n};
    return cmp_blocks(c, p, local_regs, local_cards, r0, r1, sink, user, nullptr, &sh);
}

// The same into a caller buffer (page-locked memory moves at full PCIe speed): rows [row_begin, row_end) packed from row_begin.  Collective.
extern "C" int d2g_cmp_rows_sharded(d2g_ctx *c, const d2g_cmp_params *p, const double *local_regs, const double *local_cards,
                                    uint64_t local_begin, uint64_t local_n, uint64_t r0, uint64_t r1, float *out) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_cmp_params(p)) return rc;
    if (r0 > r1 || r1 > n_rows(p)) return fail(D2G_EINVAL, "bad row range");
    if (r1 > r0 && !out) return fail(D2G_EINVAL, "null output");
    const ShardArg sh{local_begin, local_n};
    return cmp_blocks(c, p, local_regs, local_cards, r0, r1, nullptr, nullptr, out, &sh);
}

extern "C" int d2g_cmp_rows_sharded_dev(d2g_ctx *c, const d2g_cmp_params *p, const double *local_regs_d, const double *local_cards_d,
                                        uint64_t local_begin, uint64_t local_n, uint64_t r0, uint64_t r1, float *out_d) {
    if (!c) return fail(D2G_EINVAL, "null ctx");
    if (int rc = check_cmp_params(p)) return rc;
    if (r0 > r1 || r1 > n_rows(p)) return fail(D2G_EINVAL, "bad row range");
    CU(cudaSetDevice(c->device));
    c->c16cache.valid = false; c->c16g.valid = false;
    d2g::CmpConsts k;
    if (int rc = make_consts(c, p, &k)) return rc;
    if (int rc = sharded_prepare(c, p, local_regs_d, local_cards_d, local_begin, local_n)) return rc;
    c->c16_sharded = true;
    const int rc = launch_cmp(c, p, k, c->c16g.regs, c->xcards.as<double>(), r0, r1, out_d, nullptr, nullptr);
    c->c16_sharded = false;
    return rc;
}
