/* The timed loops of bench.py's strong-scaling section as a compiled caller would run them: N back-to-back calls of one C ABI
 * entry point issued from C (the crate's callers are Rust; a Python `for` adds 1-2 us and its jitter to every call, which on
 * 8 ranks shows up as start skew between the GPUs). Loaded by bench.py with ctypes next to the library itself.
 *   gcc -O2 -std=c99 -shared -fPIC -Iinclude tools/call_loop.c -o tools/bin/libec_call_loop.so -Lerased_cells_b200/lib -lerased_cells_b200 */
#include "erased_cells_b200.h"

/* n x ec_buf_min_max_sharded (comm != NULL) or ec_buf_min_max; returns the first failing status */
int ec_loop_min_max(ec_comm* comm, const ec_buf* strip, int n, ec_value* mn, ec_value* mx) {
    for (int i = 0; i < n; ++i) {
        const ec_status s = comm ? ec_buf_min_max_sharded(comm, strip, 0, mn, mx) : ec_buf_min_max(strip, 0, mn, mx);
        if (s != EC_OK) return (int)s;
    }
    return 0;
}
