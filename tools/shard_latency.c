/* Latency of reductions on a raster that ONE process keeps sharded over G GPUs (ec_init_devices), through the plain C
 * ABI a Rust / C caller would use: f32 side x side `min_max` under the three cross-GPU finishes, masked counts, and an
 * element-wise op for scale. Wall clock around the host-visible call (that is what a caller waits for), best and mean.
 *   gcc -O2 -std=c99 -Iinclude tools/shard_latency.c -o tools/bin/shard_latency -Lerased_cells_b200/lib -lerased_cells_b200 -Wl,-rpath,$PWD/erased_cells_b200/lib
 *   tools/bin/shard_latency 32768 0,1,2,3,4,5,6,7 */
#define _POSIX_C_SOURCE 199309L
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "erased_cells_b200.h"

static double now_us(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e6 + t.tv_nsec * 1e-3;
}
static char g_out[8192];
static size_t g_len = 0;
static void emit(const char* fmt, ...) {  /* the JSON line is printed in one piece at the end (libraries write to stdout too) */
    va_list ap;
    va_start(ap, fmt);
    g_len += (size_t)vsnprintf(g_out + g_len, sizeof g_out - g_len, fmt, ap);
    va_end(ap);
}
#define CK(x) do { ec_status s_ = (x); if (s_ != EC_OK) { fprintf(stderr, "%s failed: %d %s\n", #x, (int)s_, ec_last_error()); exit(1); } } while (0)

int main(int argc, char** argv) {
    size_t side = argc > 1 ? strtoull(argv[1], NULL, 0) : 16384;
    int devs[16], n = 0;
    const char* list = argc > 2 ? argv[2] : "0";
    for (const char* p = list; *p && n < 16;) { char* e; devs[n++] = (int)strtol(p, &e, 10); p = *e == ',' ? e + 1 : e; }
    const int iters = argc > 3 ? atoi(argv[3]) : 50;
    CK(ec_init_devices(devs, n));
    ec_set_min_max_cache(0);  /* every timed call reads the raster */
    ec_set_shard_min_cells((size_t)1 << 20);
    const size_t cells = side * side;
    ec_buf* a;
    CK(ec_buf_synth(EC_FLOAT32, cells, 0xEC40, 0, 2, -1e4, 1e4, 0, NULL, &a));
    CK(ec_synchronize());
    emit("{\"side\": %zu, \"cells\": %zu, \"devices\": \"%s\", \"strips\": %d", side, cells, list, ec_buf_shard_count(a));
    static const char* names[3] = {"host_fold", "peer_mailbox", "nccl"};
    ec_value mn0, mx0;
    for (int mode = 0; mode < 3; ++mode) {
        if (mode > 0 && n == 1) break;
        ec_set_shard_finish(mode);
        ec_value mn, mx;
        for (int i = 0; i < 5; ++i) CK(ec_buf_min_max(a, NULL, &mn, &mx));
        double best = 1e30, sum = 0;
        for (int i = 0; i < iters; ++i) {
            const double t0 = now_us();
            CK(ec_buf_min_max(a, NULL, &mn, &mx));
            const double dt = now_us() - t0;
            if (dt < best) best = dt;
            sum += dt;
        }
        if (mode == 0) { mn0 = mn; mx0 = mx; }
        emit(", \"min_max_%s_us\": [%.1f, %.1f], \"bits_%s\": [\"%llx\", \"%llx\"], \"same_as_host_fold_%s\": %s", names[mode], best, sum / iters, names[mode],
               (unsigned long long)mn.bits, (unsigned long long)mx.bits, names[mode], (mn.bits == mn0.bits && mx.bits == mx0.bits) ? "true" : "false");
    }
    ec_set_shard_finish(0);
    {   /* masked: NoData mask (counts come with it), masked min_max, counts */
        ec_value nd;
        CK(ec_buf_get(a, 12345, &nd));
        ec_mask* m;
        CK(ec_mask_from_nodata(a, EC_NODATA_VALUE, &nd, &m));  /* first call: allocations */
        ec_mask_free(m);
        double t0 = now_us();
        CK(ec_mask_from_nodata(a, EC_NODATA_VALUE, &nd, &m));
        size_t d, nodata;
        CK(ec_mask_counts(m, &d, &nodata));
        const double t_build = now_us() - t0;
        t0 = now_us();
        CK(ec_mask_counts(m, &d, &nodata));
        const double t_counts = now_us() - t0;
        ec_value mn, mx;
        CK(ec_buf_min_max(a, m, &mn, &mx));
        double best = 1e30;
        for (int i = 0; i < iters; ++i) { t0 = now_us(); CK(ec_buf_min_max(a, m, &mn, &mx)); const double dt = now_us() - t0; if (dt < best) best = dt; }
        emit(", \"from_nodata_plus_counts_us\": %.1f, \"counts_again_us\": %.2f, \"valid\": %zu, \"masked_min_max_us\": %.1f", t_build, t_counts, d, best);
        ec_mask_free(m);
    }
    {   /* element-wise: f32 -> f64 convert, strips side by side */
        ec_buf* o;
        CK(ec_buf_convert(a, EC_FLOAT64, &o));
        CK(ec_synchronize());
        ec_buf_free(o);
        double best = 1e30;
        for (int i = 0; i < 10; ++i) {
            const double t0 = now_us();
            CK(ec_buf_convert(a, EC_FLOAT64, &o));
            CK(ec_synchronize());
            const double dt = now_us() - t0;
            if (dt < best) best = dt;
            ec_buf_free(o);
        }
        emit(", \"convert_f32_f64_us\": %.1f, \"convert_GBps\": %.1f", best, 12.0 * cells / best / 1e3);
    }
    {   /* statistics (extension): min_max + moments pass per strip, exact sums folded on the host */
        ec_statistics st;
        CK(ec_buf_statistics(a, NULL, &st));
        double best = 1e30;
        for (int i = 0; i < 10; ++i) { const double t0 = now_us(); CK(ec_buf_statistics(a, NULL, &st)); const double dt = now_us() - t0; if (dt < best) best = dt; }
        emit(", \"statistics_us\": %.1f, \"mean\": %.17g, \"stddev\": %.17g", best, st.mean, st.stddev);
    }
    emit("}");
    printf("RESULT %s\n", g_out);
    ec_buf_free(a);
    return 0;
}
