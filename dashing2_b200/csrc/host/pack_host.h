// pack_host.h -- host-side packer of the sequence layout in csrc/pack_kernels.cuh (implementation: pack_host.cpp, plain C++
// with AVX-512 / AVX2 / scalar paths chosen at run time).  Internal to libd2gpu; the C ABI wraps it as d2g_pack_sequences.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <functional>

namespace d2g_host {
unsigned host_threads();                                            // D2G_HOST_THREADS, else the CPUs this process may run on (<= 64)
void parallel_for(size_t n, const std::function<void(size_t)> &fn); // persistent pool shared by all contexts
// Words [w0, w1) of the packed form of the n bases at seq (codes / mask indexed from w0).  Returns the number of words that hold
// an invalid base inside the data.  Multi-threaded.
uint64_t pack_contiguous(const char *seq, uint64_t n, uint64_t w0, uint64_t w1, uint64_t *codes, uint32_t *mask);
// The same for the concatenation of pieces, words [0, n_words).
uint64_t pack_pieces(const char *const *pieces, const uint64_t *piece_len, uint64_t n_pieces, uint64_t n_words, uint64_t *codes, uint32_t *mask);
const char *pack_isa_name();
}  // namespace d2g_host
