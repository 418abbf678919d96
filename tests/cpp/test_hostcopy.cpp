// Host-side copy pool of the staged pageable-memory path (erased_cells_b200/csrc/ec_hostcopy.hpp): no CUDA.
// Copies ragged sizes with every thread count, from several caller threads at once, then sleeps long enough for the
// workers to park on their condition variable and exits: process exit must not wait for them.
#include <chrono>
#include <cstdio>
#include <numeric>

#include "../../erased_cells_b200/csrc/ec_hostcopy.hpp"

static int fails = 0;
static void check(bool ok, const char* what, size_t n, int t) {
    if (!ok) { std::printf("FAIL %s n=%zu threads=%d\n", what, n, t); ++fails; }
}
int main() {
    std::vector<uint8_t> src((size_t(9) << 20) + 4099);
    uint32_t x = 0x9E3779B9u;
    for (auto& b : src) { x = x * 1664525u + 1013904223u; b = uint8_t(x >> 24); }
    for (int nt = 0; nt < 2; ++nt) {
        ec::g_copy_nt = nt;  // memcpy, then the streaming-store flavour (falls back where there is no AVX2)
#if defined(__x86_64__)
        if (nt && !__builtin_cpu_supports("avx2")) break;
#else
        if (nt) break;
#endif
        for (size_t n : {size_t(0), size_t(1), size_t(4097), (size_t(1) << 20) - 1, size_t(1) << 20, (size_t(3) << 20) + 31, src.size()})
            for (int t : {1, 2, 3, 8, 13})
                for (size_t skew : {size_t(0), size_t(5)}) {  // unaligned source and destination
                    if (n + skew > src.size()) continue;
                    std::vector<uint8_t> dst(n + skew + 64, 0xAB);
                    ec::copy_pool().copy(dst.data() + skew, src.data() + skew, n, t);
                    check(std::memcmp(dst.data() + skew, src.data() + skew, n) == 0, "bytes", n, t);
                    bool clean = true;
                    for (size_t i = 0; i < skew; ++i) clean &= dst[i] == 0xAB;
                    for (size_t i = n + skew; i < dst.size(); ++i) clean &= dst[i] == 0xAB;
                    check(clean, "wrote outside the range", n, t);
                }
    }
    // several callers at once take turns
    std::vector<std::thread> callers;
    std::atomic<int> bad{0};
    for (int c = 0; c < 4; ++c)
        callers.emplace_back([&, c] {
            std::vector<uint8_t> dst(src.size());
            for (int r = 0; r < 6; ++r) {
                std::fill(dst.begin(), dst.end(), uint8_t(c));
                ec::copy_pool().copy(dst.data(), src.data(), src.size() - c, 4 + c);
                if (std::memcmp(dst.data(), src.data(), src.size() - c) != 0) ++bad;
            }
        });
    for (auto& t : callers) t.join();
    check(bad == 0, "concurrent callers", src.size(), 4);
    std::this_thread::sleep_for(std::chrono::milliseconds(300));  // the workers are asleep by now
    std::printf(fails ? "HOSTCOPY_FAILED\n" : "HOSTCOPY_OK\n");
    return fails ? 1 : 0;
}
