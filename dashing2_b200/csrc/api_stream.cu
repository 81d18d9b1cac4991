// api_stream.cu -- element streams (stream_kernels.cuh): which parameter sets take them, and the producers' launchers.
#define D2G_STREAM_PRODUCERS
#include "api_sketch_launch.h"
#include <cub/device/device_scan.cuh>

bool is_stream_mode(const d2g_sketch_params *p) {
    if (p->alphabet != 0 && p->alphabet != 4) return true;           // protein
    if (p->k > 32) return true;                                      // rolling hash
    return p->w > p->k && !p->canon;                                 // -C with a window
}

namespace {
__global__ void scale_off_kernel(const uint64_t *rec_off, uint64_t n, uint64_t mult, uint64_t *out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = rec_off[i] * mult;
}

// CyclicHash character tables: 256 draws of WyRand<uint64_t> seeded with the 32-bit truncation of seed1 ^ seed2
// (rollinghash/cyclichash.h CyclicHash::seed -> characterhash.h CharacterHash::seed(uint32_t); aesctr/wy.h:112 maps seed 0 to 1337)
void cyclic_table(uint64_t s1, uint64_t s2, uint64_t out4[4]) {
    uint64_t state = (uint32_t)(s1 ^ s2);
    if (!state) state = 1337;
    for (int i = 0; i < 4; ++i) out4[i] = d2g::wyhash64(state);      // only the DNA codes 0..3 are ever looked up in a valid window
}
inline uint64_t rotl_h(uint64_t x, int r) { r &= 63; return r ? (x << r) | (x >> (64 - r)) : x; }

// alphabet.h:107-120 through make_lut (:30-58): comma-separated groups get codes 0, 1, ...; both cases
void alpha_lut(const char *groups, int8_t lut[256]) {
    memset(lut, -1, 256);
    int id = 0;
    for (const char *q = groups; *q; ++q) {
        if (*q == ',') { ++id; continue; }
        lut[(unsigned char)(*q | 32)] = (int8_t)id; lut[(unsigned char)(*q & 0xdf)] = (int8_t)id;
    }
}
} // namespace

int prepare_stream(d2g_ctx *c, const d2g_sketch_params *p, const d2g::PackedSeq &seq_d, const uint8_t *ascii_d, const uint64_t *rec_off_d,
                   uint64_t n_rec, uint64_t total_len, StreamView *out) {
    cudaStream_t st = c->stream;
    const bool protein = p->alphabet != 0 && p->alphabet != 4;
    const bool windowed = p->w > p->k;
    const bool rolling = !protein && p->k > 32;
    const uint64_t mult = (rolling && p->canon && windowed) ? 2 : 1;
    out->p = *p;
    out->p.k = 1; out->p.w = windowed ? p->w - p->k + 1 : 0; out->p.alphabet = 0; out->p.canon = 1;   // not a stream mode any more
    out->total_len = total_len * mult;
    if (int rc = c->items.reserve((total_len * mult + 64) * 8)) return rc;
    if (int rc = c->itemcnt.reserve((n_rec + 1) * 4)) return rc;
    if (int rc = c->itemoff.reserve((n_rec + 1) * 8)) return rc;
    uint64_t *items = c->items.as<uint64_t>();
    uint32_t *cnt = c->itemcnt.as<uint32_t>();
    scale_off_kernel<<<(unsigned)((n_rec + 1 + 255) / 256), 256, 0, st>>>(rec_off_d, n_rec + 1, mult, c->itemoff.as<uint64_t>());
    c->launches++;
    out->rec_off_d = c->itemoff.as<uint64_t>();
    out->seq = seq_d;
    out->seq.items = items; out->seq.item_cnt = cnt;
    out->seq.item_flags = rolling ? 0u : d2g::STREAM_SKIP_ONES;
    if (n_rec == 0 || total_len == 0) { if (n_rec) CU(cudaMemsetAsync(cnt, 0, n_rec * 4, st)); return D2G_OK; }

    if (protein) {
        if (!ascii_d) return fail(D2G_EINVAL, "protein alphabets need the record bytes (d2g_sketch_batch / d2g_sketch_batch_dev), not packed DNA");
        const int A = p->alphabet;
        int8_t lut[256];
        alpha_lut(A == 20 ? "A,C,D,E,F,G,H,I,K,L,M,N,P,Q,R,S,T,V,W,Y" : A == 14 ? "A,C,D,EQ,FY,G,H,IV,KR,LM,N,P,ST,W"
                  : A == 6 ? "AST,CP,DHNEKQR,FWY,G,ILMV" : "AST,C,DHN,EKQR,FWY,G,ILMV,P", lut);
        if (int rc = c->slut.reserve(256)) return rc;
        CU(cudaMemcpyAsync(c->slut.p, lut, 256, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));                               // lut is a stack array
        d2g::ProteinConsts pc;
        pc.mul = (uint64_t)A; pc.k = p->k; pc.windowed = windowed; pc.bitmask = A == 8;
        pc.mask = A == 8 ? (~0ULL >> (64 - p->k)) : (uint64_t)std::pow((double)A, (double)p->k);   // rhtraits.h:57-61
        d2g::protein_kernel<<<(unsigned)((n_rec + 127) / 128), 128, 0, st>>>(ascii_d, rec_off_d, n_rec, c->slut.as<int8_t>(), pc, items, cnt);
        c->launches++;
        CU(cudaGetLastError());
        return D2G_OK;
    }
    if (rolling) {
        d2g::RollConsts rc;
        rc.K = p->k; rc.canon = p->canon; rc.windowed = windowed;
        cyclic_table(1337, 137, rc.tf);                              // encoder.h:673 defaults, src/d2.h:136
        cyclic_table(137ULL * 1337ULL, 137ULL ^ 1337ULL, rc.tr);    // encoder.h:680: the reverse-complement hasher
        for (int ch = 0; ch < 4; ++ch) {                             // K copies of one complemented base (see stream_kernels.cuh)
            uint64_t q = 0;
            for (int j = 0; j < p->k; ++j) q ^= rotl_h(rc.tr[3 - ch], j);
            rc.q[ch] = q;
        }
        const uint64_t nsegs = total_len / (uint64_t)(p->k + 1) + n_rec + 1;
        if (int r2 = c->segs.reserve(nsegs * sizeof(d2g::RollSeg))) return r2;
        if (int r2 = c->nseg.reserve((n_rec + 1) * 4)) return r2;
        d2g::roll_walk_kernel<<<(unsigned)((n_rec * 32 + 127) / 128), 128, 0, st>>>(seq_d, rec_off_d, n_rec, rc, c->segs.as<d2g::RollSeg>(), c->nseg.as<uint32_t>(), cnt);
        const uint64_t nchunks = (total_len + d2g::ROLL_CHUNK - 1) / d2g::ROLL_CHUNK;
        d2g::roll_fill_kernel<<<(unsigned)((nchunks + 127) / 128), 128, 0, st>>>(seq_d, rec_off_d, n_rec, total_len, rc, c->segs.as<d2g::RollSeg>(), c->nseg.as<uint32_t>(), items);
        c->launches += 2;
        CU(cudaGetLastError());
        return D2G_OK;
    }
    // -C with a window, k <= 32
    const uint64_t nw = d2g::packed_words(total_len);
    const uint32_t *emask = seq_d.mask;
    if (p->k >= 31) {
        if (int rc = c->svmask.reserve(nw * 4)) return rc;
        CU(cudaMemsetAsync(c->svmask.p, 0xFF, nw * 4, st));          // padding reads as invalid
        const uint64_t nthreads = (total_len + 31) / 32 * 32;
        d2g::ncw_virtual_invalid_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(seq_d, rec_off_d, n_rec, total_len, c->svmask.as<uint32_t>(), nw);
        c->launches++;
        emask = c->svmask.as<uint32_t>();
    }
    if (int rc = c->sflag.reserve((total_len + 1) * 4)) return rc;
    if (int rc = c->sexcl.reserve((total_len + 1) * 4)) return rc;
    if (total_len + 1 >= 0xFFFFFFFFULL) return fail(D2G_EINVAL, "-C with a window: at most 2^32 bases per batch");
    uint32_t *flag = c->sflag.as<uint32_t>(), *excl = c->sexcl.as<uint32_t>();
    d2g::ncw_flag_kernel<<<(unsigned)((total_len + 1 + 255) / 256), 256, 0, st>>>(emask, rec_off_d, n_rec, total_len, p->k, flag);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, flag, excl, total_len + 1, st);
    if (int rc = c->stmp.reserve(tb + 256)) return rc;
    CU(cub::DeviceScan::ExclusiveSum(c->stmp.p, tb, flag, excl, total_len + 1, st));
    d2g::ncw_fill_kernel<<<(unsigned)((std::max(total_len, n_rec) + 255) / 256), 256, 0, st>>>(seq_d, rec_off_d, n_rec, total_len, p->k, flag, excl, items, cnt);
    c->launches += 4;
    CU(cudaGetLastError());
    return D2G_OK;
}
