// Internal declarations shared by the translation units of libcu2b.so (not part of the ABI).
#ifndef CU2B_INTERNAL_H_
#define CU2B_INTERNAL_H_

#include <stddef.h>
#include <stdint.h>

#include "cu2b.h"

// Records a printf-style message for cu2b_last_error() (thread local) and returns `code`.
cu2b_status cu2b_fail(cu2b_status code, const char *fmt, ...)
#if defined(__GNUC__)
    __attribute__((format(printf, 2, 3)))
#endif
    ;

// Row padding rule of the device layout: factor rows are stored with a pitch of kp floats,
// kp = n_factors rounded up to a multiple of 4, so every row starts on a 16-byte boundary and
// can be moved with 128-bit accesses. Padding elements are zero and stay zero under the update.
static inline int cu2b_padded_factors(int k) { return (k + 3) & ~3; }

// Threshold below which the chunked file readers stay single-threaded (host_io.cpp).
size_t cu2b_io_parallel_min_bytes();

#endif  // CU2B_INTERNAL_H_
