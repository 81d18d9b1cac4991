// fss_kernels.cuh -- K3: Full (continuous) SetSketch register construction on the device.
//
// Reference: CSetSketch<double>::update, /root/reference/src/setsketch.h:369-423 (max-tree mvt_t
// :123-167, lazy Fisher-Yates bonsai/hll/include/sketch/fy.h:14-66, WyRand<uint32_t,2> aesctr/wy.h,
// flog.h, kahan.h).  Per element id the reference generates an increasing sequence
//     ev_0 = -log(CEHasher(id ^ C) * 2^-64) / m,   ev_t = ev_{t-1} (+Kahan) -log(wy_t(id) * 2^-64) / (m - t)
// and applies reg[pi_t] = min(reg[pi_t], ev_t) along a per-element random permutation pi until ev_t
// exceeds the current maximum register.  The early exit (and its 0.7*flog pre-test) only skips updates
// that could not lower any register, so the final registers are the element-wise minimum over all
// elements of all (pi_t, ev_t) -- independent of element order.  That is what is computed here:
//
//   boot   : over a sample of the stream, the t=0 contribution alone (register pi_0 receives ev_0; the
//            largest CEHasher value per register gives the smallest ev_0) -> a per-entity upper bound T
//            of the final maximum register;
//   main   : every element whose ev_0 <= T walks its sequence while ev_t <= T, atomicMin into registers
//            held in shared memory (order-preserving u64 keys of the doubles).  Elements are first
//            compacted into a shared-memory queue so the walk runs on dense warps;
//   long   : walks longer than FSS_SPARSE steps (only when T is loose: tiny inputs) are replayed by a
//            second kernel with a dense permutation state in HBM scratch.
// log() is ref_log (devlog.cuh), bit-identical to the host libm; the Kahan increment is the fused
// fma(b, log, -carry) that the reference binary executes (GCC contracts `b*log(v) - carry`).
#pragma once
#include "common.cuh"
#include "devlog.cuh"
#include "sketch_kernels.cuh"

namespace d2g {

constexpr uint64_t FSS_XOR = 0xb2069fc679a8da0bULL;   // setsketch.h:376
constexpr int FSS_SPARSE = 24;                         // walk steps kept in the per-thread sparse permutation

constexpr uint64_t FSS_KEY_EMPTY = 0x7fefffffffffffffULL | 0x8000000000000000ULL; // dkey(DBL_MAX), setsketch.h:130

// flog.h:14-20
__device__ __forceinline__ double flog_d(double x) {
    return __fma_rn((double)(unsigned long long)__double_as_longlong(x), 1.539095918623324e-16, -709.0895657128241);
}

// ---- boot: t = 0 contributions only --------------------------------------------------------------
struct FssBootConsumer {
    struct Params { uint64_t *maxrv; FastMod32 fm; uint32_t m; };   // maxrv [n_entities][m], zero-initialised
    static __host__ __device__ size_t smem_bytes(uint32_t m, bool) { return (size_t)m * 8; }
    uint64_t *s; Params p;
    __device__ __forceinline__ void init(unsigned char *smem, const Params &pp, bool) {
        p = pp; s = reinterpret_cast<uint64_t *>(smem);
        for (uint32_t i = threadIdx.x; i < p.m; i += SK_THREADS) s[i] = 0;
    }
    static constexpr bool kEveryWindow = false;
    static constexpr int kMinBlocks = 4;
    __device__ __forceinline__ void begin_entity(uint32_t, uint64_t) {}
    __device__ __forceinline__ void consume(uint64_t hv) {
        const uint64_t rv = cehash(hv ^ FSS_XOR);
        uint64_t st = rv;
        const uint32_t idx = fastmod32((uint32_t)wyhash64(st), p.fm);   // first draw of the permutation, fy.h:36-37
        if (rv > s[idx]) atomicMax(reinterpret_cast<unsigned long long *>(s + idx), (unsigned long long)rv);
    }
    __device__ __forceinline__ void end_tile(uint32_t) {}
    __device__ __forceinline__ void flush(uint32_t ent) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < p.m; i += SK_THREADS) {
            const uint64_t v = s[i];
            if (v) { atomicMax(reinterpret_cast<unsigned long long *>(p.maxrv + (uint64_t)ent * p.m + i), (unsigned long long)v); s[i] = 0; }
        }
        __syncthreads();
    }
};

// ---- guess and verify: the bound T without a boot pass ---------------------------------------------
// Every element reaches register i at a time that is marginally Exp(1) (the m arrival times of one element are the
// order statistics of m iid exponentials), so a register of n elements is Exp(n) and the largest of m registers is about
// (ln m + gamma) / n.  For inputs with many elements per register the bound is GUESSED as three times that, with n
// estimated from the sequence length (P(max > guess) ~ m^-2 when the estimate is right), and VERIFIED afterwards: if every
// final register is <= the guess, nothing the guess pruned could have lowered a register and the result is exact; an
// entity that fails (repetitive sequence: far fewer distinct elements than estimated) is redone through the boot pass.
static __global__ void fss_entity_positions_kernel(const uint64_t *rec_off, const uint32_t *rec_entity, uint64_t n_rec, uint32_t ent_base, int need,
                                            unsigned long long *npos) {
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint64_t l = rec_off[r + 1] - rec_off[r];
    if (l >= (uint64_t)need) atomicAdd(npos + (rec_entity[r] - ent_base), (unsigned long long)(l - need + 1));
}
// state: 0 = guessed bound (main pass A), 1 = boot pass + main pass B
static __global__ void fss_guess_kernel(const unsigned long long *npos, uint32_t n_ent, uint32_t m, int wsz, int allow_guess, double *T, double *Tguess, uint32_t *state) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ent) return;
    const double n_est = (double)npos[e] * (wsz > 1 ? 2. / (wsz + 1.) : 1.);
    const double lm = log((double)m) + 0.5772156649;
    const bool guess = allow_guess && n_est >= 2. * (double)m * lm;
    const double g = guess ? 3. * lm / n_est : 1.7976931348623157e308;
    T[e] = g; Tguess[e] = g; state[e] = guess ? 0u : 1u;
}
// one CTA per entity in state 0: exact iff every register is filled and the largest is <= the guess; else -> state 1, registers cleared
static __global__ void fss_verify_kernel(uint64_t *keys, uint32_t m, const double *Tguess, uint32_t *state, unsigned int *n_redo) {
    __shared__ uint64_t red[256];
    const uint32_t ent = blockIdx.x;
    if (state[ent] != 0) { if (threadIdx.x == 0) atomicAdd(n_redo, 1u); return; }
    uint64_t mx = 0;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) mx = max(mx, keys[(uint64_t)ent * m + i]);
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) { if ((int)threadIdx.x < s) red[threadIdx.x] = max(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
    const bool ok = red[0] < FSS_KEY_EMPTY && dunkey(red[0]) <= Tguess[ent];
    __syncthreads();
    if (ok) return;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) keys[(uint64_t)ent * m + i] = FSS_KEY_EMPTY;
    if (threadIdx.x == 0) { state[ent] = 1; atomicAdd(n_redo, 1u); }
}

// one CTA per entity: T = max_i ev_0(maxrv_i) (DBL_MAX when some register was never hit by the sample)
static __global__ void fss_threshold_kernel(const uint64_t *maxrv, uint32_t m, double *T, const uint32_t *state) {
    __shared__ double red[256];
    const uint32_t ent = blockIdx.x;
    if (state && state[ent] != 1) return;
    const double bv0 = -1. / m;
    double mx = 0.;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        const uint64_t rv = maxrv[(uint64_t)ent * m + i];
        const double ev = rv ? __dmul_rn(bv0, ref_log(__dmul_rn(__ull2double_rn(rv), 0x1p-64))) : 1.7976931348623157e308;
        mx = fmax(mx, ev);
    }
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) { if ((int)threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
    if (threadIdx.x == 0) T[blockIdx.x] = red[0];
}

// ---- the per-element walk ------------------------------------------------------------------------
struct WalkRng {   // WyRand<uint32_t, 2>: two 64-bit draws per refill, served as four little-endian u32
    uint64_t state, b0, b1; int off;
    __device__ __forceinline__ void seed(uint64_t s) { state = s; off = 4; }
    __device__ __forceinline__ uint32_t next() {
        if (off == 4) { b0 = wyhash64(state); b1 = wyhash64(state); off = 0; }
        const uint64_t w = (off & 2) ? b1 : b0;
        const uint32_t r = (off & 1) ? (uint32_t)(w >> 32) : (uint32_t)w;
        ++off;
        return r;
    }
};

// Where a walk delivers (register, key): plain u64 keys (HBM), or the CTA-local 32-bit filter in front of the HBM keys.
struct Keys64Sink {
    uint64_t *keys;
    __device__ __forceinline__ void put(uint32_t idx, uint64_t kk) const {
        if (kk < keys[idx]) atomicMin(reinterpret_cast<unsigned long long *>(keys + idx), (unsigned long long)kk);
    }
};
// approx[i] = high word of the smallest key THIS CTA has delivered to register i (0xFFFFFFFF = none yet): a key whose high
// word is larger cannot lower the register, everything else goes to the entity's keys in HBM with atomicMin.  16 KiB
// instead of 32 KiB of shared memory per CTA at S = 4096 (four CTAs per SM), and the CTA-local maximum that tightens the
// bound T is still available (rounded up to the next high word).
struct ApproxSink {
    uint32_t *approx; uint64_t *gkeys;
    __device__ __forceinline__ void put(uint32_t idx, uint64_t kk) const {
        const uint32_t h = (uint32_t)(kk >> 32), cur = approx[idx];
        if (h > cur) return;
        if (h < cur) atomicMin(approx + idx, h);
        if (kk < gkeys[idx]) atomicMin(reinterpret_cast<unsigned long long *>(gkeys + idx), (unsigned long long)kk);
    }
};

// Replays the sequence of one element against a register sink with threshold T.
// PermState provides step(i, samp) -> register index (lazy Fisher-Yates) and may report overflow.
template <class PermState, class Sink>
__device__ __forceinline__ bool fss_walk(uint64_t x, uint32_t m, double T, const Sink &sink, PermState &ps) {
    uint64_t hid = x;
    uint64_t rv = cehash(x ^ FSS_XOR);
    const double tv = __dmul_rn(__ull2double_rn(rv), 0x1p-64);
    const double bv0 = -1. / m;
    double ev = __dmul_rn(bv0, ref_log(tv));
    if (ev > T) return true;
    WalkRng rng; rng.seed(rv);
    double carry = 0.;
    uint32_t bi = 1;
    for (uint32_t i = 0;; ++i) {
        const uint32_t samp = rng.next() % (m - i);
        uint32_t idx;
        if (!ps.step(i, samp, idx)) return false;                       // sparse state exhausted
        sink.put(idx, dkey(ev));
        if (bi == m) return true;
        rv = wyhash64(hid);
        const double bv = -(1. / (double)(m - bi)); ++bi;               // getbeta, setsketch.h:300-302
        const double nv = __dmul_rn(__ull2double_rn(rv), 0x1p-64);
        if (__fma_rn(__dmul_rn(bv, flog_d(nv)), .7, ev) > T) return true; // conservative pre-test, setsketch.h:418
        const double inc = __fma_rn(bv, ref_log(nv), -carry);           // kahan.h:8-13 (contracted by the reference build)
        const double tmp = __dadd_rn(ev, inc);
        carry = __dadd_rn(__dadd_rn(tmp, -ev), -inc);
        ev = tmp;
        if (ev > T) return true;
    }
}

struct SparsePerm {   // lazy Fisher-Yates over a handful of touched slots (fy.h:35-47)
    uint32_t key[FSS_SPARSE], val[FSS_SPARSE]; int n = 0;
    __device__ __forceinline__ uint32_t get(uint32_t j) const { for (int q = 0; q < n; ++q) if (key[q] == j) return val[q]; return j; }
    __device__ __forceinline__ bool step(uint32_t i, uint32_t samp, uint32_t &out) {
        const uint32_t j = i + samp;
        out = get(j);
        const uint32_t gi = get(i);
        for (int q = 0; q < n; ++q) if (key[q] == j) { val[q] = gi; return true; }
        if (n == FSS_SPARSE) return false;
        key[n] = j; val[n] = gi; ++n;
        return true;
    }
};

struct DensePerm {    // the reference's layout: g/v arrays with a generation counter
    uint32_t *g, *v; uint32_t c;
    __device__ __forceinline__ bool step(uint32_t i, uint32_t samp, uint32_t &out) {
        const uint32_t j = i + samp;
        out = v[j] == c ? g[j] : j;
        g[j] = v[i] == c ? g[i] : i;
        v[j] = c;
        return true;
    }
};

struct FssMainConsumer {
    struct Params {
        uint64_t *keys;            // [n_entities][m], initialised to FSS_KEY_EMPTY
        double *T;                 // [n_entities] upper bound of the final maximum register; tightened by every CTA (atomicMin)
        uint64_t *ovf; unsigned long long *ovf_count; uint64_t ovf_cap;   // (x, entity) pairs for the long-walk kernel
        uint32_t m;
    };
    // survivors are batched over tiles so the walk runs on full warps.  Between two end_tile calls a tile pushes at most
    // SK_TILE survivors (unwindowed) or, in windowed mode, the staged minimizers (SK_SCAP) -- more only for adversarial
    // windows that change their minimizer at every position, which then overflow to the long-walk list.
    static __host__ __device__ constexpr int qcap(bool windowed) { return windowed ? SK_SCAP + 960 : SK_TILE + 64; }
    static constexpr int DCAP = 192;          // walks that outran the sparse state wait here for a tighter threshold
    static constexpr uint32_t APPROX_EMPTY = 0xFFFFFFFFu, APPROX_DBLMAX = (uint32_t)(FSS_KEY_EMPTY >> 32);
    static __host__ __device__ size_t smem_bytes(uint32_t m, bool windowed) { return (size_t)((m + 1) / 2) * 8 + (size_t)(qcap(windowed) + DCAP) * 8 + 64; }
    uint32_t *approx; uint64_t *queue, *defer, *bcast; int *qn, *dn, *nx; Params p; double T; uint64_t rvmin; int drain_at, QCAP; uint32_t cur;
    __device__ __forceinline__ void init(unsigned char *smem, const Params &pp, bool windowed) {
        QCAP = qcap(windowed);
        drain_at = QCAP - (windowed ? SK_SCAP : SK_TILE); cur = 0;
        p = pp; approx = reinterpret_cast<uint32_t *>(smem); queue = reinterpret_cast<uint64_t *>(smem) + (p.m + 1) / 2; defer = queue + QCAP; bcast = defer + DCAP;
        qn = reinterpret_cast<int *>(bcast + 4); dn = qn + 1; nx = dn + 1;
        for (uint32_t i = threadIdx.x; i < p.m; i += SK_THREADS) approx[i] = APPROX_EMPTY;
        if (threadIdx.x == 0) { *qn = 0; *dn = 0; *nx = 0; }
        T = 1.7976931348623157e308; rvmin = 0;
    }
    static __device__ __forceinline__ uint64_t rvmin_for(double t, uint32_t m) {   // every rv below this certainly has ev_0 > t
        if (!(t < 1e300)) return 0;
        const double scaled = exp(-t * (double)m) * (1. - 1e-6) * 18446744073709551616.0;
        return scaled >= 18446744073709549568.0 ? 0xFFFFFFFFFFFFF800ULL : (scaled <= 0. ? 0 : (uint64_t)scaled);
    }
    static constexpr bool kEveryWindow = false;
    static constexpr int kMinBlocks = 4;
    __device__ __forceinline__ void begin_entity(uint32_t ent, uint64_t) {
        T = *reinterpret_cast<volatile double *>(p.T + ent); rvmin = rvmin_for(T, p.m); cur = ent;
    }
    __device__ __forceinline__ void consume(uint64_t hv) {
        if (cehash(hv ^ FSS_XOR) < rvmin) return;
        const int slot = atomicAdd(qn, 1);
        if (slot < QCAP) queue[slot] = hv;
        else {
            const unsigned long long g = atomicAdd(p.ovf_count, 1ULL);
            if (g < p.ovf_cap) { p.ovf[2 * g] = hv; p.ovf[2 * g + 1] = cur; }
        }
    }
    // All threads.  Replays the queued elements (register updates go to the entity's keys in HBM through the CTA-local
    // filter), then tightens the threshold: the final registers are element-wise <= what this CTA delivered, so the largest
    // local entry bounds the final maximum; the bound is shared with the other CTAs working on the same entity through p.T.
    // Walks that outrun the sparse permutation state (threshold still loose) are retried after the tightening;
    // only if that does not help do they go to the long-walk kernel.
    __device__ __forceinline__ void drain(uint32_t ent, bool final) {
        for (int round = 0;; ++round) {
            const int n = min(*qn, QCAP);
            const ApproxSink sink{approx, p.keys + (uint64_t)ent * p.m};
            // Walks differ in length (most end after one or two steps, some take ten): every lane draws its next element from a shared
            // counter the moment its walk ends, and the loop body holds a single log evaluation that serves the first point of a new walk
            // and the next point of a running one alike -- the lanes of a warp stay busy instead of waiting for the longest walk.
            // Same arithmetic, point by point, as fss_walk (setsketch.h:369-423).
            {
                const uint32_t m = p.m;
                const double bv0 = -1. / m;
                bool active = false, more = true;
                uint64_t x = 0, hid = 0; double ev = 0., carry = 0.; uint32_t i = 0;
                WalkRng rng; rng.seed(0);
                SparsePerm sp;
                while (__any_sync(0xffffffffu, active || more)) {
                    double nv, bv; bool fresh = false;
                    if (!active && more) {
                        const int q = atomicAdd(nx, 1);
                        if (q < n) {
                            x = queue[q]; hid = x;
                            const uint64_t rv = cehash(x ^ FSS_XOR);
                            nv = __dmul_rn(__ull2double_rn(rv), 0x1p-64); bv = bv0;
                            rng.seed(rv); sp.n = 0; carry = 0.; i = 0;
                            active = true; fresh = true;
                        } else more = false;
                    } else if (active) {
                        const uint64_t rv = wyhash64(hid);
                        bv = -(1. / (double)(m - (i + 1)));                               // getbeta, setsketch.h:300-302 (i + 1 points delivered so far)
                        nv = __dmul_rn(__ull2double_rn(rv), 0x1p-64);
                        ++i;
                        if (__fma_rn(__dmul_rn(bv, flog_d(nv)), .7, ev) > T) active = false;    // conservative pre-test, setsketch.h:418
                    }
                    if (active) {
                        const double lg = ref_log(nv);
                        if (fresh) ev = __dmul_rn(bv, lg);
                        else {
                            const double inc = __fma_rn(bv, lg, -carry);                 // kahan.h:8-13 (contracted by the reference build)
                            const double tmp = __dadd_rn(ev, inc);
                            carry = __dadd_rn(__dadd_rn(tmp, -ev), -inc);
                            ev = tmp;
                        }
                        if (ev > T) active = false;
                        else {
                            const uint32_t samp = rng.next() % (m - i);
                            uint32_t idx;
                            if (!sp.step(i, samp, idx)) {                                // sparse state exhausted: retry after the threshold has tightened
                                const int slot = atomicAdd(dn, 1);
                                if (slot < DCAP) defer[slot] = x;
                                else {
                                    const unsigned long long g = atomicAdd(p.ovf_count, 1ULL);
                                    if (g < p.ovf_cap) { p.ovf[2 * g] = x; p.ovf[2 * g + 1] = ent; }
                                }
                                active = false;
                            } else {
                                sink.put(idx, dkey(ev));
                                if (i + 1 == m) active = false;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) *nx = 0;
            uint32_t mx = 0;
            for (uint32_t i = threadIdx.x; i < p.m; i += SK_THREADS) mx = max(mx, approx[i]);
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if ((threadIdx.x & 31) == 0) queue[threadIdx.x >> 5] = mx;      // queue is free now
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t m2 = 0;
                for (int w = 0; w < SK_THREADS / 32; ++w) m2 = max(m2, (uint32_t)queue[w]);
                double t = T;
                if (m2 < APPROX_DBLMAX) {                                   // every register has been delivered a finite value
                    const double tl = dunkey(((uint64_t)m2 << 32) | 0xFFFFFFFFULL);   // largest key with that high word
                    if (tl < t && tl >= 0.) { t = tl; atomicMin(reinterpret_cast<unsigned long long *>(p.T + ent), (unsigned long long)__double_as_longlong(tl)); }
                }
                const double tg = *reinterpret_cast<volatile double *>(p.T + ent);
                if (tg < t) t = tg;
                bcast[0] = (uint64_t)__double_as_longlong(t); bcast[1] = rvmin_for(t, p.m);
                bcast[2] = (t < T) ? 1 : 0;
            }
            __syncthreads();
            const bool improved = bcast[2] != 0;
            T = __longlong_as_double((long long)bcast[0]); rvmin = bcast[1];
            const int nd = min(*dn, DCAP);
            __syncthreads();
            if (nd == 0) { if (threadIdx.x == 0) { *qn = 0; *dn = 0; } __syncthreads(); return; }
            if (improved && (final || round == 0)) {
                // retry the deferred walks now with the tighter threshold
                for (int q = threadIdx.x; q < nd; q += SK_THREADS) queue[q] = defer[q];
                if (threadIdx.x == 0) { *qn = nd; *dn = 0; }
                __syncthreads();
                continue;
            }
            if (!final && improved) {
                // keep them for the next drain
                for (int q = threadIdx.x; q < nd; q += SK_THREADS) queue[q] = defer[q];
                if (threadIdx.x == 0) { *qn = nd; *dn = 0; }
                __syncthreads();
                return;
            }
            // no progress possible here: hand them to the long-walk kernel
            for (int q = threadIdx.x; q < nd; q += SK_THREADS) {
                const unsigned long long g = atomicAdd(p.ovf_count, 1ULL);
                if (g < p.ovf_cap) { p.ovf[2 * g] = defer[q]; p.ovf[2 * g + 1] = ent; }
            }
            if (threadIdx.x == 0) { *qn = 0; *dn = 0; }
            __syncthreads();
            return;
        }
    }
    __device__ __forceinline__ void end_tile(uint32_t ent) {
        __syncthreads();
        if (*qn > drain_at) drain(ent, false);           // uniform: every thread reads the same counter after the barrier
    }
    __device__ __forceinline__ void flush(uint32_t ent) {
        __syncthreads();
        drain(ent, true);
        for (uint32_t i = threadIdx.x; i < p.m; i += SK_THREADS) approx[i] = APPROX_EMPTY;   // the keys themselves are already in HBM
        __syncthreads();
    }
};

// one CTA per entity: T[e] = largest final register (keys are order preserving; an entity without elements keeps its bound)
static __global__ void fss_final_bound_kernel(const uint64_t *keys, uint32_t m, double *T) {
    __shared__ uint64_t red[256];
    const uint32_t e = blockIdx.x;
    uint64_t mx = 0;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) mx = max(mx, keys[(uint64_t)e * m + i]);
    red[threadIdx.x] = mx; __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) { if ((int)threadIdx.x < s) red[threadIdx.x] = max(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
    if (threadIdx.x == 0 && red[0] != FSS_KEY_EMPTY) T[e] = dunkey(red[0]);
}

// ---- --save-kmers: ids_[idx] = id wherever CSetSketch::update lowers a register (setsketch.h:400-404) -----------------
// A second, read-only pass once the registers are final: every element whose first point is not above the entity's verified bound
// T replays its walk, and a point that EQUALS its register names the element that set it.  No 128-bit atomics in the hot pass,
// and no cost at all without --save-kmers.  Walks that outrun the sparse permutation state go to the long-walk kernel below.
struct IdsSink {
    const uint64_t *keys; uint64_t *ids; uint64_t x;
    __device__ __forceinline__ void put(uint32_t idx, uint64_t kk) const { if (kk == keys[idx]) ids[idx] = x; }
};
struct FssIdsConsumer {
    struct Params { const uint64_t *keys; const double *T; uint64_t *ids; uint64_t *ovf; unsigned long long *ovf_count; uint64_t ovf_cap; uint32_t m; };
    static __host__ __device__ size_t smem_bytes(uint32_t, bool) { return 16; }
    static constexpr bool kEveryWindow = false;
    static constexpr int kMinBlocks = 2;
    Params p; double T; uint64_t rvmin; uint32_t cur;
    __device__ __forceinline__ void init(unsigned char *, const Params &pp, bool) { p = pp; T = 1.7976931348623157e308; rvmin = 0; cur = 0; }
    __device__ __forceinline__ void begin_entity(uint32_t ent, uint64_t) { T = p.T[ent]; rvmin = FssMainConsumer::rvmin_for(T, p.m); cur = ent; }
    __device__ __forceinline__ void consume(uint64_t hv) {
        if (cehash(hv ^ FSS_XOR) < rvmin) return;
        SparsePerm sp;
        const IdsSink sink{p.keys + (uint64_t)cur * p.m, p.ids + (uint64_t)cur * p.m, hv};
        if (!fss_walk(hv, p.m, T, sink, sp)) {
            const unsigned long long g = atomicAdd(p.ovf_count, 1ULL);
            if (g < p.ovf_cap) { p.ovf[2 * g] = hv; p.ovf[2 * g + 1] = cur; }
        }
    }
    __device__ __forceinline__ void end_tile(uint32_t) {}
    __device__ __forceinline__ void flush(uint32_t) {}
};
static __global__ void fss_longwalk_ids_kernel(const uint64_t *ovf, const unsigned long long *ovf_count, uint64_t ovf_cap, uint32_t m,
                                        const double *T, const uint64_t *keys, uint64_t *ids, uint32_t *scratch) {
    const uint64_t slot = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t nslots = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n = min((uint64_t)*ovf_count, ovf_cap);
    DensePerm dp; dp.g = scratch + slot * 2ULL * m; dp.v = dp.g + m; dp.c = 0;
    for (uint64_t e = slot; e < n; e += nslots) {
        const uint64_t x = ovf[2 * e]; const uint32_t ent = (uint32_t)ovf[2 * e + 1];
        ++dp.c;
        fss_walk(x, m, T[ent], IdsSink{keys + (uint64_t)ent * m, ids + (uint64_t)ent * m, x}, dp);
    }
}

// long walks: one thread per queued element, dense permutation state in HBM scratch (slot-private)
static __global__ void fss_longwalk_kernel(const uint64_t *ovf, const unsigned long long *ovf_count, uint64_t ovf_cap, uint32_t m,
                                    const double *T, uint64_t *keys, uint32_t *scratch) {
    const uint64_t slot = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t nslots = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n = min((uint64_t)*ovf_count, ovf_cap);
    DensePerm dp; dp.g = scratch + slot * 2ULL * m; dp.v = dp.g + m; dp.c = 0;   // scratch zero-initialised; generations start at 1
    for (uint64_t e = slot; e < n; e += nslots) {
        const uint64_t x = ovf[2 * e]; const uint32_t ent = (uint32_t)ovf[2 * e + 1];
        ++dp.c;
        fss_walk(x, m, T[ent], Keys64Sink{keys + (uint64_t)ent * m}, dp);
    }
}

// keys -> doubles (first S registers) and cardinality m / sum(reg) (setsketch.h:553-561; the reference
// sums under `omp simd`, i.e. in compiler-chosen order: sequential here, agreement ~1e-15 relative)
static __global__ void fss_finalize_kernel(const uint64_t *keys, uint32_t n_ent, uint32_t m, double *sig, double *card) {
    const uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (sig && e < (uint64_t)n_ent * m) sig[e] = dunkey(keys[e]);
    if (card && e < n_ent) {
        double s = 0.;
        for (uint32_t i = 0; i < m; ++i) s = __dadd_rn(s, dunkey(keys[e * m + i]));
        card[e] = (double)m / s;
    }
}

} // namespace d2g
