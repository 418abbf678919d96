// One process, two row strips: the sharded handles of the C ABI driven from C++ through include/erased_cells.hpp.
// Runs on a one-GPU box by listing CUDA device 0 twice (two logical devices with their own streams and block caches —
// the same code path as two GPUs); `test_sharded <dev0,dev1,...>` takes real devices instead.
// Every result is checked against plain host arithmetic (C++ doubles = the reference's f64 ops on this platform).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "erased_cells.hpp"

using namespace erased_cells;
static int failed = 0, passed = 0;
#define CHECK(c) do { if (c) ++passed; else { ++failed; std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); } } while (0)

int main(int argc, char** argv) {
    std::vector<int> devs = {0, 0};
    if (argc > 1) {
        devs.clear();
        for (char* p = std::strtok(argv[1], ","); p; p = std::strtok(nullptr, ",")) devs.push_back(std::atoi(p));
    }
    const int G = init_devices(devs);
    CHECK(G == int(devs.size()));
    set_shard_min_cells(1024);
    const size_t n = 100000 + 77;  // ragged: the last strip takes the remainder
    std::vector<uint8_t> a(n);
    std::vector<uint16_t> b(n);
    for (size_t i = 0; i < n; ++i) { a[i] = uint8_t(i * 7 + 3); b[i] = uint16_t((i * 2654435761u) >> 7) % 5; }  // zeros in b: x/0 and 0/0
    CellBuffer A(a), B(b);
    CHECK(A.shard_count() == G && B.shard_count() == G);
    size_t covered = 0;
    for (int g = 0; g < G; ++g) { const ec_shard_info s = A.shard(g); CHECK(s.offset == covered && s.offset % 128 == 0 && s.logical_device == g); covered += s.len; }
    CHECK(covered == n);
    // README: buf1 / buf2 * 0.5, strip by strip on the strips' devices
    const CellBuffer r = A / B * CellValue(0.5);
    CHECK(r.shard_count() == G && r.cell_type() == CellType::Float64 && r.len() == n);
    const std::vector<double> got = r.to_vec<double>();
    bool same = got.size() == n;
    for (size_t i = 0; same && i < n; ++i) {
        const double want = double(a[i]) / double(b[i]) * 0.5;
        same = std::memcmp(&want, &got[i], 8) == 0 || (std::isnan(want) && std::isnan(got[i]));
    }
    CHECK(same);
    // reductions finish across the strips; x/0 = inf must come out as the maximum, the 0/0 NaNs sort above it
    std::vector<int16_t> c(n);
    for (size_t i = 0; i < n; ++i) c[i] = int16_t((i * 40503u) & 0xFFFF);
    int16_t lo = c[0], hi = c[0];
    for (int16_t v : c) { lo = v < lo ? v : lo; hi = v > hi ? v : hi; }
    const CellBuffer Cb(c);
    const auto mm = Cb.min_max();
    CHECK(mm.first == CellValue(lo) && mm.second == CellValue(hi));
    // NoData mask + masked chain + counts on sharded handles
    std::vector<int16_t> d(c);
    for (size_t i = 0; i < n; i += 97) d[i] = -32768;
    MaskedCellBuffer M = MaskedCellBuffer::from_vec_with_nodata(d, NoData<int16_t>(NoData<int16_t>::Default));
    size_t nodata = (n + 96) / 97;
    for (size_t i = 0; i < n; ++i) if (i % 97 != 0 && d[i] == -32768) ++nodata;
    CHECK(M.counts().second == nodata && M.counts().first == n - nodata);
    const MaskedCellBuffer D = (M - M) * CellValue(2.0);
    const auto dm = D.min_max();
    CHECK(dm.first == CellValue(0.0) && dm.second == CellValue(0.0) && D.counts().first == n - nodata);
    // single cells land in the right strip; equality is decided on the devices
    CellBuffer A2 = A;
    CHECK(A2 == A);
    const size_t idx = A.shard(G - 1).offset + 5;
    A2.put(idx, CellValue(uint8_t(a[idx] ^ 1)));
    CHECK(!(A2 == A) && A2.get(idx) == CellValue(uint8_t(a[idx] ^ 1)) && A.get(idx) == CellValue(a[idx]));
    // operands of different lengths: the result is re-partitioned (zip truncation, src/buffer.rs:327)
    std::vector<uint16_t> shorter(b.begin(), b.begin() + 60001);
    const CellBuffer S(shorter);
    const std::vector<double> z = (A + S).to_vec<double>();
    bool zs = z.size() == shorter.size();
    for (size_t i = 0; zs && i < z.size(); ++i) zs = z[i] == double(a[i]) + double(shorter[i]);
    CHECK(zs);
    std::printf("sharded: %d passed, %d failed\n", passed, failed);
    return failed ? 1 : 0;
}
