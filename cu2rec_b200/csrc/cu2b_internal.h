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

// Row placement rule shared by the session (engine.cu) and the DSGD partition (host_io.cpp): slot[r] = row
// offset inside a range of n rows starting at matrix row `first_row` for the item of popularity rank r.
void cu2b_paired_slots(int n, int first_row, int rows_per_block, int *slot);
int cu2b_rows_per_l2_block(int n_factors);  // factor rows per 1 KB
// cu2b_dsgd_partition with the placement rule of a given row size
cu2b_status cu2b_dsgd_partition_rows(const cu2b_rating *train, int64_t n, int rows, int cols, int world, int rows_per_block,
                                     int *user_block, int *user_local, int *users_per_block, int *item_new,
                                     int *item_block_ptr, int64_t *block_nnz);

// Threshold below which the chunked file readers stay single-threaded (host_io.cpp).
size_t cu2b_io_parallel_min_bytes();

#endif  // CU2B_INTERNAL_H_
