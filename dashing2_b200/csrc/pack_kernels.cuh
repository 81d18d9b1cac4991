// pack_kernels.cuh -- ASCII bases -> the packed sequence layout every sketch kernel reads.
//
// Packed layout of a batch (base positions are offsets into the concatenated record bytes):
//   codes : uint64 words, word i = bases [32i, 32i+32), two bits per base, first base in bits 63:62
//           (A0 C1 G2 T3, case-insensitive -- /root/reference/bonsai/include/bonsai/alphabet.h:128 DNA4);
//   mask  : uint32 words, word i = the same 32 bases, one bit per base, first base in bit 31; 1 = not A/C/G/T
//           (the reference resets the k-mer run on such a base, encoder.h:254, or feeds k-mer 0 in windowed mode, :568-571).
// Both arrays are padded with PACK_PAD_WORDS words past the last base so that tile loads need no bounds check;
// bases beyond the end read as invalid.  The host packer (pack_host.cpp) writes the identical layout.
#pragma once
#include "common.cuh"

namespace d2g {

constexpr uint64_t PACK_PAD_WORDS = 256;   // 8192 bases: more than one tile halo
__host__ __device__ __forceinline__ uint64_t packed_words(uint64_t n_bases) { return ((n_bases + 127) / 128) * 4 + PACK_PAD_WORDS; }

struct PackedSeq {
    const uint64_t *codes; const uint32_t *mask;
    // set only for the element streams of stream_kernels.cuh (k > 32, -C with a window, protein): the records' elements in push order,
    // item_cnt[r] of them from items[rec_off[r]] on (rec_off then counts item slots), STREAM_* flags
    const uint64_t *items = nullptr; const uint32_t *item_cnt = nullptr; uint32_t item_flags = 0;
};

// 4 ASCII bytes (first base in the low byte) -> 8 bits of codes (first base in the two MSBs) and a
// 4-bit invalid mask (first base in bit 3).
__device__ __forceinline__ void decode4(uint32_t v, uint32_t &codes, uint32_t &inv) {
    uint32_t x = (v >> 1) & 0x03030303u;               // A0 C1 T2 G3
    x ^= (x >> 1) & 0x01010101u;                        // A0 C1 G2 T3
    codes = (x * 0x40100401u) >> 24;
    // the letter each code stands for, looked up per byte with PRMT; any difference (case folded) marks the base invalid
    uint32_t s = (x | (x >> 4)) & 0x00FF00FFu;
    s = (s | (s >> 8)) & 0xFFFFu;
    const uint32_t z = (v & 0xDFDFDFDFu) ^ __byte_perm(0x54474341u, 0u, s);
    const uint32_t nz = (((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z) & 0x80808080u;
    inv = ((nz >> 7) * 0x08040201u) >> 24;
}

// one thread per word of 32 bases
static __global__ void pack_ascii_kernel(const uint8_t *seq, uint64_t n_bases, uint64_t n_words, uint64_t *codes, uint32_t *mask) {
    const uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const uint64_t b0 = w * 32;
    if (b0 >= n_bases) { codes[w] = 0; mask[w] = 0xFFFFFFFFu; return; }
    uint4 v[2];
    if (b0 + 32 <= n_bases && (reinterpret_cast<uintptr_t>(seq) & 15) == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(seq) + 2 * w;
        v[0] = __ldg(src); v[1] = __ldg(src + 1);
    } else {                                            // ragged end / unaligned buffer: byte loads, zero fill (zero is invalid)
        uint32_t t[8];
        #pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t x = 0;
            #pragma unroll
            for (int j = 0; j < 4; ++j) { const uint64_t b = b0 + 4 * i + j; if (b < n_bases) x |= (uint32_t)seq[b] << (8 * j); }
            t[i] = x;
        }
        v[0] = make_uint4(t[0], t[1], t[2], t[3]); v[1] = make_uint4(t[4], t[5], t[6], t[7]);
    }
    uint64_t cw = 0; uint32_t mw = 0;
    #pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t c0, c1, c2, c3, i0, i1, i2, i3;
        decode4(v[h].x, c0, i0); decode4(v[h].y, c1, i1); decode4(v[h].z, c2, i2); decode4(v[h].w, c3, i3);
        cw = (cw << 32) | ((c0 << 24) | (c1 << 16) | (c2 << 8) | c3);
        mw = (mw << 16) | ((i0 << 12) | (i1 << 8) | (i2 << 4) | i3);
    }
    codes[w] = cw; mask[w] = mw;
}

} // namespace d2g
