/* The C ABI from plain C (C99): the header must be valid C, and the README example must run through it.
 * tests/test_gpu_cpp_port.py builds this with gcc and runs it on the GPU box. */
#include <stdio.h>
#include <string.h>

#include "erased_cells_b200.h"

#define CHECK(call) do { ec_status s_ = (call); if (s_ != EC_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, (int)s_, ec_last_error()); return 1; } } while (0)

int main(void) {
    const uint8_t a[3] = {1, 2, 3};
    const uint16_t b[3] = {2, 4, 6};
    ec_buf *ba, *bb, *q, *r;
    ec_value half, mn, mx;
    double out[3];
    int ord;
    if (ec_abi_version() != EC_ABI_VERSION) return 2;
    if (ec_ctype_union(EC_INT32, EC_FLOAT32) != EC_FLOAT64 || !ec_ctype_can_fit_into(EC_UINT8, EC_INT16)) return 3;
    CHECK(ec_buf_from_host(EC_UINT8, a, 3, &ba));
    CHECK(ec_buf_from_host(EC_UINT16, b, 3, &bb));
    CHECK(ec_buf_binary(EC_DIV, ba, bb, &q));           /* buf1 / buf2 */
    memset(&half, 0, sizeof half);
    half.ct = EC_FLOAT64;
    { const double h = 0.5; memcpy(&half.bits, &h, 8); }
    CHECK(ec_buf_scalar(EC_MUL, q, &half, &r));          /* ... * 0.5 */
    if (ec_buf_ctype(r) != EC_FLOAT64 || ec_buf_len(r) != 3) return 4;
    CHECK(ec_buf_to_host(r, out, sizeof out));
    if (out[0] != 0.25 || out[1] != 0.25 || out[2] != 0.25) return 5;
    CHECK(ec_buf_min_max(r, NULL, &mn, &mx));
    if (mn.bits != mx.bits) return 6;
    CHECK(ec_buf_cmp(r, r, &ord));
    if (ord != 0) return 7;
    if (ec_buf_convert(r, EC_INT32, &q) != EC_NARROWING) return 8;   /* Error::NarrowingError{Float64, Int32} */
    { uint8_t s, d; ec_last_narrowing(&s, &d); if (s != EC_FLOAT64 || d != EC_INT32) return 9; }
    ec_buf_free(ba); ec_buf_free(bb); ec_buf_free(r);
    printf("c abi ok, %llu kernel launches\n", (unsigned long long)ec_kernel_launches());
    return 0;
}
