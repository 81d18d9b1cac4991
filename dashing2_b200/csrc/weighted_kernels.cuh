// weighted_kernels.cuh -- K4/K5/K6: exact k-mer counting and the weighted sketches built on it.
//
// Reference: Counter::add / finalize (/root/reference/src/counter.h:68-77,118-138) fills a hash map
// hashed-k-mer -> count and feeds every (id, count) to ProbMinHash3 (bonsai/hll/include/sketch/bmh.h:662-700,
// truncated exponential :490-525) or BagMinHash2 (bmh.h:269-316, poisson_process_t :129-207).
// Device design (no hash map):
//   emit   : the sketch kernel with EmitConsumer writes one hashed value per k-mer / per window (every
//            window counts, no minimizer de-duplication) into the slot range of its own span;
//   count  : radix sort by (entity, value) + run-length encode -> distinct (entity, id, count) triples;
//   sketch : one thread per distinct element replays the element's point sequence against the per-entity
//            registers in HBM (atomicMin on order-preserving keys), pruned by a per-entity bound T;
//   verify : T is a guess (~ m ln m / total weight); if afterwards max(register) <= T the registers are exact
//            (every point below T was generated and nothing above can matter), otherwise T is quadrupled and
//            the entity is redone.  Both sketches are element-order independent (see oracle/d2_oracle.c).
#pragma once
#include "common.cuh"
#include "devlog.cuh"
#include "fss_kernels.cuh"

namespace d2g {

// ---- emit ------------------------------------------------------------------------------------------
struct EmitConsumer {
    struct Params { uint64_t *out_hv; uint32_t *out_ent; uint64_t span; };   // both arrays have one slot per base, pre-filled with ~0
    static constexpr bool kEveryWindow = true;
    static constexpr int kMinBlocks = 4;
    static __host__ __device__ size_t smem_bytes(uint32_t, bool) { return 16; }
    Params p; unsigned int *cur; uint32_t ent; uint64_t base;
    __device__ __forceinline__ void init(unsigned char *smem, const Params &pp, bool) {
        p = pp; cur = reinterpret_cast<unsigned int *>(smem); base = (uint64_t)blockIdx.x * p.span; ent = 0xFFFFFFFFu;
        if (threadIdx.x == 0) *cur = 0;
    }
    // slots [base + off, ...) belong to the region of the new entity: never write its values below `off`
    __device__ __forceinline__ void begin_entity(uint32_t e, uint64_t off) {
        __syncthreads();
        if (threadIdx.x == 0 && *cur < (unsigned int)off) *cur = (unsigned int)off;
        ent = e;
        __syncthreads();
    }
    __device__ __forceinline__ void consume(uint64_t hv) {
        const unsigned int slot = atomicAdd(cur, 1u);
        p.out_hv[base + slot] = hv; p.out_ent[base + slot] = ent;
    }
    __device__ __forceinline__ void end_tile(uint32_t) {}
    __device__ __forceinline__ void flush(uint32_t) {}
};

// ---- run-length encode of the sorted (entity, value) stream ----------------------------------------
// `shift` drops low bits that do not take part in the identity of an element (the sign bit of a count-sketch key, below)
__global__ void rle_flag_kernel(const uint64_t *hv, const uint32_t *ent, uint64_t n, uint32_t *flag, int shift = 0) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t e = ent[i];
    flag[i] = (e != 0xFFFFFFFFu && (i == 0 || ent[i - 1] != e || (hv[i - 1] >> shift) != (hv[i] >> shift))) ? 1u : 0u;
}
// ---- --countsketch-size n: Counter::add / finalize with a count sketch (src/counter.h:68-77,131-137) ----------------
// The reference adds +1 / -1 (top bit of Wang(x) set / clear) to bucket Wang(x) % n of a float table and afterwards feeds every bucket i
// with |count| >= threshold to the weighted sketch as element (i, |count|).  Device: every emitted value becomes the key
// (bucket << 1) | positive; after the sort by (entity, key) a bucket is one run with its negative entries first, so its signed sum is
// (#positive - #negative) from one binary search for the boundary -- integers, as the float sums are below 2^24.
__global__ void cs_key_kernel(uint64_t *hv, const uint32_t *ent, uint64_t n, uint64_t cssize) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n || ent[i] == 0xFFFFFFFFu) return;
    const uint64_t h = wang64(hv[i]);
    hv[i] = ((h % cssize) << 1) | (h >> 63);
}
__global__ void cs_run_weight_kernel(const uint64_t *hv, const uint32_t *pos, uint64_t nu, const unsigned long long *n_valid, uint32_t *wts) {
    const uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (u >= nu) return;
    const uint64_t i = pos[u], end = (u + 1 < nu) ? pos[u + 1] : *n_valid;
    uint64_t a = i, b = end;                              // first entry of the run with the positive bit
    while (a < b) { const uint64_t mid = (a + b) >> 1; if (hv[mid] & 1ULL) b = mid; else a = mid + 1; }
    const long long sum = (long long)(end - a) - (long long)(a - i);
    wts[u] = (uint32_t)(sum < 0 ? -sum : sum);
}
// pos[u] = index of the u-th run head; the element count per run is pos[u+1]-pos[u] (sentinel slots sort last)
__global__ void rle_scatter_kernel(const uint32_t *flag, const uint32_t *excl, const uint32_t *ent, uint64_t n, uint32_t *pos, unsigned long long *n_valid) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) pos[excl[i]] = (uint32_t)i;
    if (ent[i] != 0xFFFFFFFFu && (i + 1 == n || ent[i + 1] == 0xFFFFFFFFu)) *n_valid = i + 1;   // end of the real data
}
// total weight per entity (sum of counts above the threshold; integers, so exact in double)
__global__ void weight_sum_kernel(const uint32_t *ent, const uint32_t *pos, uint64_t nu, const unsigned long long *n_valid, double threshold,
                                  unsigned long long *wsum, const uint32_t *wts = nullptr) {
    const uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t e = 0xFFFFFFFFu; unsigned long long c = 0;
    if (u < nu) {
        const uint64_t i = pos[u], end = (u + 1 < nu) ? pos[u + 1] : *n_valid;
        e = ent[i];
        c = wts ? wts[u] : end - i;
        if (!((double)c > threshold)) c = 0;
    }
    // the runs are sorted by entity: a warp almost always holds one entity -> one atomic per warp and entity, not per element
    unsigned todo = __ballot_sync(0xffffffffu, c != 0);
    while (todo) {
        const int leader = __ffs((int)todo) - 1;
        const uint32_t le = __shfl_sync(0xffffffffu, e, leader);
        const bool mine = c != 0 && e == le;
        unsigned long long v = mine ? c : 0;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((int)(threadIdx.x & 31) == leader) atomicAdd(wsum + le, v);
        todo &= ~__ballot_sync(0xffffffffu, mine);
    }
}

// per-entity first guess of the bound: registers receive points at total rate W/m, so their maximum is about
// m (ln m + gamma) / W; start four times above that
__global__ void weighted_guess_kernel(const unsigned long long *wsum, uint32_t n_ent, uint32_t m, double *T, uint32_t *state) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ent) return;
    const double W = (double)wsum[e];
    T[e] = W > 0. ? 4. * (double)m * (log((double)m) + 1.) / W : 0.;
    state[e] = W > 0. ? 0u : 2u;   // 0 = to do, 1 = redo with larger T, 2 = exact
}
// one CTA per entity: exact iff every register is filled and the largest is <= T
__global__ void weighted_verify_kernel(uint64_t *keys, uint32_t m, double *T, uint32_t *state, unsigned int *n_redo) {
    __shared__ uint64_t red[256];
    const uint32_t e = blockIdx.x;
    if (state[e] == 2u) return;
    uint64_t mx = 0;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) mx = max(mx, keys[(uint64_t)e * m + i]);
    red[threadIdx.x] = mx; __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) { if ((int)threadIdx.x < s) red[threadIdx.x] = max(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
    const bool ok = red[0] != FSS_KEY_EMPTY && dunkey(red[0]) <= T[e];
    __syncthreads();
    if (ok) { if (threadIdx.x == 0) state[e] = 2u; return; }
    // redo: registers are reset (they only hold valid upper bounds, but a clean slate keeps the run identical)
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) keys[(uint64_t)e * m + i] = FSS_KEY_EMPTY;
    if (threadIdx.x == 0) { T[e] = T[e] < 1e290 ? T[e] * 4. : 1.7976931348623157e308; state[e] = 1u; atomicAdd(n_redo, 1u); }
}

// ---- ProbMinHash3 -----------------------------------------------------------------------------------
struct TexpConsts { double lambda, c1, c2, c3, c4; };   // bmh.h:490-502, computed in long double on the host
__device__ __forceinline__ double texp_sample(uint64_t rngstate, const TexpConsts &c) {   // bmh.h:507-525
    double x = __dmul_rn(__dmul_rn(0x1p-64, __ull2double_rn(rngstate)), c.c1);
    if (x >= 1.) {
        for (;;) {
            if ((x = __dmul_rn(0x1p-64, __ull2double_rn(wyhash64(rngstate)))) < c.c2) break;
            double yhat = __dmul_rn(0.5, __dmul_rn(0x1p-64, __ull2double_rn(wyhash64(rngstate))));
            double omx = __dadd_rn(1., -x);
            if (yhat > omx) { x = omx; yhat = __dadd_rn(1., -yhat); }
            omx = __dadd_rn(1., -x);
            if (x <= __dmul_rn(c.c3, __dadd_rn(1., -yhat)) || __dmul_rn(yhat, c.c1) <= omx) break;
            if (__fma_rn(yhat, c.c4, 1.) <= exp(__dmul_rn(c.lambda, omx))) break;   // exp only decides a rejection; an ulp cannot flip it in practice
        }
    }
    return x;
}

template <class PermState>
__device__ __forceinline__ bool pmh_walk(uint64_t id, double w, uint32_t m, double T, const TexpConsts &tc, uint64_t *keys, PermState &ps, uint64_t *ids = nullptr) {
    uint64_t hi = id;
    const double wi = 1. / w;
    uint64_t rv = wyhash64(hi);
    double hv = __dmul_rn(wi, texp_sample(rv, tc));
    if (hv > T) return true;
    WalkRng rng; rng.seed(rv);
    for (uint32_t i = 0;;) {
        const uint32_t samp = rng.next() % (m - i);
        uint32_t idx;
        if (!ps.step(i, samp, idx)) return false;
        const uint64_t kk = dkey(hv);
        if (ids) { if (kk == keys[idx]) ids[idx] = id; }             // ids pass: the registers are final, the element that equals one owns it
        else if (kk < keys[idx]) atomicMin(reinterpret_cast<unsigned long long *>(keys + idx), (unsigned long long)kk);
        if (++i >= m) return true;                                   // one full permutation touches every register
        hv = __dmul_rn(wi, (double)i);
        if (hv > T) return true;
        hv = __fma_rn(wi, texp_sample(wyhash64(rv), tc), hv);       // bmh.h:697
        if (hv > T) return true;
    }
}

struct WeightedArgs {
    const uint64_t *hv; const uint32_t *ent; const uint32_t *pos; uint64_t nu; const unsigned long long *n_valid;
    double threshold; uint32_t m;
    const double *T; const uint32_t *state; uint64_t *keys;
    uint64_t *ovf; unsigned long long *ovf_count; uint64_t ovf_cap;   // (unique index) of elements needing the dense walk
    unsigned int *error;
    const uint32_t *wts; int id_shift;   // count sketch: weight of run u (else its length), element id = key >> id_shift (else the key)
    uint64_t *ids;                       // non-null: --save-kmers pass over final registers (every entity, no register update), ids [n_ent][m]
    unsigned long long *next;            // bmh_kernel: element counter of its persistent grid (zeroed before every launch)
};

__global__ void pmh_kernel(const WeightedArgs a, const TexpConsts tc) {
    const uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (u >= a.nu) return;
    const uint64_t i = a.pos[u], end = (u + 1 < a.nu) ? a.pos[u + 1] : *a.n_valid;
    const uint32_t e = a.ent[i];
    if (!a.ids && a.state[e] == 2u) return;
    const double w = a.wts ? (double)a.wts[u] : (double)(end - i);
    if (!(w > a.threshold)) return;
    SparsePerm sp;
    if (!pmh_walk(a.hv[i] >> a.id_shift, w, a.m, a.T[e], tc, a.keys + (uint64_t)e * a.m, sp, a.ids ? a.ids + (uint64_t)e * a.m : nullptr)) {
        const unsigned long long g = atomicAdd(a.ovf_count, 1ULL);
        if (g < a.ovf_cap) a.ovf[g] = u; else atomicExch(a.error, 1u);
    }
}
__global__ void pmh_longwalk_kernel(const WeightedArgs a, const TexpConsts tc, uint32_t *scratch) {
    const uint64_t slot = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, nslots = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n = min((uint64_t)*a.ovf_count, a.ovf_cap);
    DensePerm dp; dp.g = scratch + slot * 2ULL * a.m; dp.v = dp.g + a.m; dp.c = 0;
    for (uint64_t q = slot; q < n; q += nslots) {
        const uint64_t u = a.ovf[q];
        const uint64_t i = a.pos[u], end = (u + 1 < a.nu) ? a.pos[u + 1] : *a.n_valid;
        const uint32_t e = a.ent[i];
        ++dp.c;
        pmh_walk(a.hv[i] >> a.id_shift, a.wts ? (double)a.wts[u] : (double)(end - i), a.m, a.T[e], tc, a.keys + (uint64_t)e * a.m, dp,
                 a.ids ? a.ids + (uint64_t)e * a.m : nullptr);
    }
}

// ---- BagMinHash2 ------------------------------------------------------------------------------------
struct PProc { double x, minp, maxq, carry; uint64_t wyv; };
constexpr int BMH_STACK = 48;

__device__ __forceinline__ uint32_t bmh_step(PProc &p, uint32_t m, const FastMod32 &fm64hint) {   // bmh.h:170-176
    const uint64_t xi = wyhash64(p.wyv);
    const double u = __dmul_rn(__ull2double_rn(xi >> 12), 2.220446049250313e-16);
    double inc = __ddiv_rn(-ref_log(u), __dadd_rn(p.maxq, -p.minp));
    inc = __dadd_rn(inc, -p.carry);
    const double tmp = __dadd_rn(p.x, inc);
    p.carry = __dadd_rn(__dadd_rn(tmp, -p.x), -inc);
    p.x = tmp;
    (void)fm64hint;
    return (uint32_t)(xi % m);
}
__device__ __forceinline__ bool bmh_partially(const PProc &p, double w) {
    return __longlong_as_double(__double_as_longlong(p.minp) + 1) <= w;
}
__device__ __forceinline__ bool bmh_can_split(const PProc &p) {
    return (uint64_t)__double_as_longlong(p.maxq) > (uint64_t)__double_as_longlong(p.minp) + 1;
}
__device__ __forceinline__ PProc bmh_split(PProc &p) {   // bmh.h:182-206
    uint64_t midpoint = ((uint64_t)__double_as_longlong(p.minp) + (uint64_t)__double_as_longlong(p.maxq)) / 2;
    const double midval = __longlong_as_double((long long)midpoint);
    const uint64_t rval = wyhash64(midpoint);
    uint64_t xval = (uint64_t)__double_as_longlong(p.x) ^ rval;
    const double pr = __ddiv_rn(__dadd_rn(midval, -p.minp), __dadd_rn(p.maxq, -p.minp));
    const double rv = __dmul_rn(__ull2double_rn(wyhash64(xval)), 5.421010862427522e-20);
    const bool goleft = rv < pr;
    PProc r = p;
    r.minp = goleft ? midval : p.minp; r.maxq = goleft ? p.maxq : midval; r.wyv = xval;
    if (goleft) p.maxq = midval; else p.minp = midval;
    return r;
}

// One lane per element, but elements differ wildly in how far their split tree has to be explored (ncu, one thread per element and
// its whole tree: 3.8 of 32 lanes busy on average).  So the tree walk is a flat state machine -- one loop iteration is one split, one
// step of a fully covered process, or one pop -- and a lane whose element is finished draws the next one from a global counter in the
// same iteration.  Same operations in the same order per element as update_2 (bmh.h:269-316).  Persistent grid; a.next counts elements.
__global__ void __launch_bounds__(128) bmh_kernel(const WeightedArgs a) {
    const FastMod32 fm{};
    PProc stack[BMH_STACK]; uint32_t sidx[BMH_STACK]; int sp = 0;
    PProc p{0., 0., 0., 0., 0}; uint32_t pidx = 0;
    bool active = false, more = true;
    double w = 0., T = 0.; uint64_t id = 0; uint64_t *keys = nullptr, *ids = nullptr;
    auto apply = [&](uint32_t idx, double x) {
        const uint64_t kk = dkey(x);
        if (ids) { if (kk == keys[idx]) ids[idx] = id; }
        else if (kk < keys[idx]) atomicMin(reinterpret_cast<unsigned long long *>(keys + idx), (unsigned long long)kk);
    };
    while (__any_sync(0xffffffffu, active || more)) {
        int step_who = 0;                      // 0 none, 1 = p (first step of a new element), 2 = the split-off half q, 3 = p (fully covered: next point)
        bool pop = false;
        PProc q{0., 0., 0., 0., 0};
        if (!active) {
            if (more) {
                const unsigned long long u = atomicAdd(a.next, 1ULL);
                if (u >= a.nu) more = false;
                else {
                    const uint64_t i0 = a.pos[u], end = (u + 1 < a.nu) ? a.pos[u + 1] : *a.n_valid;
                    const uint32_t e = a.ent[i0];
                    w = a.wts ? (double)a.wts[u] : (double)(end - i0);
                    if ((a.ids || a.state[e] != 2u) && w > a.threshold) {
                        T = a.T[e]; keys = a.keys + (uint64_t)e * a.m; id = a.hv[i0] >> a.id_shift;
                        ids = a.ids ? a.ids + (uint64_t)e * a.m : nullptr;
                        p = PProc{0., 0., 1.7976931348623157e308, 0., id};
                        sp = 0; active = true; step_who = 1;
                    }
                }
            }
        } else if (!(p.x < T)) pop = true;                            // bmh.h:282: only processes below the current maximum are expanded
        else if (bmh_can_split(p) && bmh_partially(p, w)) {
            q = bmh_split(p);
            if (p.maxq <= w) apply(pidx, p.x);
            if (bmh_partially(q, w)) step_who = 2;
        } else if (p.maxq <= w) step_who = 3;
        else pop = true;
        if (step_who) {
            PProc t = step_who == 2 ? q : p;
            const uint32_t idx = bmh_step(t, a.m, fm);
            if (step_who == 2) {
                if (t.maxq <= w) apply(idx, t.x);
                if (bmh_partially(t, w) && t.x < T) {                 // anything at or above T can never be expanded again
                    if (sp == BMH_STACK) { atomicExch(a.error, 2u); active = false; more = false; }
                    else { stack[sp] = t; sidx[sp] = idx; ++sp; }
                }
            } else {
                p = t; pidx = idx;
                if (step_who == 3 || p.maxq <= w) apply(pidx, p.x);
                if (step_who == 3 && !(p.x < T)) pop = true;          // else: re-expand the same process (equivalent to push + pop)
            }
        }
        if (pop) {
            if (sp == 0) active = false;
            else { --sp; p = stack[sp]; pidx = sidx[sp]; }
        }
    }
}

} // namespace d2g
