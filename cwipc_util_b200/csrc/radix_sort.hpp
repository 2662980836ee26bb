// radix_sort.hpp -- stable LSD radix sort of 64-bit words on a bit range (sm_100a).
#pragma once

#include "runtime.hpp"

namespace cwcu {

// Sort the n words in `a` ascending by bits [begin_bit, end_bit); words equal on that range keep
// their relative order.  `b` is scratch of the same size.  Returns the buffer (a or b) that holds
// the result.  All work is queued on `s`; nothing is synchronised.
uint64_t *radix_sort_u64(uint64_t *a, uint64_t *b, size_t n, int begin_bit, int end_bit, int dev, cudaStream_t s);

} // namespace cwcu
