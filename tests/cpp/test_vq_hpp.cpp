// Driver for include/vq.hpp, the C++ host mirror of the vq crate API.
//   test_vq_hpp host          validation order / error kinds and messages (no GPU needed), prints "host ok"
//   test_vq_hpp stream S N K  prints choose_multiple(N, K) then 5 x choose(N) of StdRng::seed_from_u64(S)
//   test_vq_hpp gpu           trains PQ / TSVQ / BQ / SQ on seeded data and prints results as hex words
// Built and run by tests/test_cpp_host.py.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "vq.hpp"

#define EXPECT_THROW(kind_, substr, stmt)                                                              \
    do {                                                                                               \
        bool ok = false;                                                                               \
        try { stmt; } catch (const vq::VqError& e) {                                                   \
            ok = e.kind == vq::ErrorKind::kind_ && std::string(e.what()).find(substr) != std::string::npos; \
            if (!ok) std::printf("wrong error for `%s`: %s\n", #stmt, e.what());                        \
        }                                                                                              \
        if (!ok) { std::printf("FAILED: %s\n", #stmt); return 1; }                                      \
    } while (0)

static std::vector<std::vector<float>> make_rows(size_t n, size_t dim, uint32_t seed) {
    std::vector<std::vector<float>> rows(n, std::vector<float>(dim));
    uint32_t s = seed;
    for (auto& r : rows)
        for (auto& v : r) { s = s * 1664525u + 1013904223u; v = (float)(int32_t)(s >> 8) / 8388608.0f - 1.0f + (float)((s >> 3) & 7); }
    return rows;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "host";
    if (mode == "host") {
        using namespace vq;
        auto rows = make_rows(10, 8, 1);
        // src/pq.rs:91-117 then src/core/vector.rs:396-410, in that order, all before the GPU is touched
        EXPECT_THROW(EmptyInput, "Empty input", ProductQuantizer({}, 2, 2, 1, Distance::Euclidean, 42));
        auto ragged = rows; ragged[3].resize(7);
        EXPECT_THROW(DimensionMismatch, "expected 8, found 7", ProductQuantizer(ragged, 2, 2, 1, Distance::Euclidean, 42));
        EXPECT_THROW(InvalidParameter, "'m': must be at most the data dimension (8)", ProductQuantizer(rows, 16, 2, 1, Distance::Euclidean, 42));
        EXPECT_THROW(InvalidParameter, "'m': dimension (8) must be divisible by m", ProductQuantizer(rows, 3, 2, 1, Distance::Euclidean, 42));
        EXPECT_THROW(InvalidParameter, "'k': must be greater than 0", ProductQuantizer(rows, 2, 0, 1, Distance::Euclidean, 42));
        EXPECT_THROW(InvalidParameter, "'k': not enough data points (10) for 11 clusters", ProductQuantizer(rows, 2, 11, 1, Distance::Euclidean, 42));
        EXPECT_THROW(EmptyInput, "Empty input", TSVQ({}, 3, Distance::Euclidean));
        EXPECT_THROW(DimensionMismatch, "expected 8, found 7", TSVQ(ragged, 3, Distance::Euclidean));
        EXPECT_THROW(InvalidParameter, "'threshold': must be finite", BinaryQuantizer(NAN, 0, 1));
        EXPECT_THROW(InvalidParameter, "'low/high': low must be less than high", BinaryQuantizer(0.f, 5, 5));
        EXPECT_THROW(InvalidParameter, "'min': must be finite", ScalarQuantizer(INFINITY, 1.f, 4));
        EXPECT_THROW(InvalidParameter, "'max': must be greater than min", ScalarQuantizer(1.f, 1.f, 4));
        EXPECT_THROW(InvalidParameter, "'levels': must be at least 2", ScalarQuantizer(0.f, 1.f, 1));
        EXPECT_THROW(InvalidParameter, "'levels': must be no more than 256 to fit in u8", ScalarQuantizer(0.f, 1.f, 257));
        ScalarQuantizer sq(0.f, 1.f, 11);
        if (sq.step() != (1.f - 0.f) / 10.f || sq.levels() != 11) { std::printf("FAILED: step\n"); return 1; }
        if (std::string(distance_name(Distance::CosineDistance)) != "cosine") return 1;
        EXPECT_THROW(DimensionMismatch, "expected 3, found 2", distance_compute(Distance::Euclidean, {1, 2, 3}, {1, 2}));
        std::printf("host ok\n");
        return 0;
    }
    if (mode == "stream") {
        uint64_t seed = std::strtoull(argv[2], nullptr, 10), n = std::strtoull(argv[3], nullptr, 10), k = std::strtoull(argv[4], nullptr, 10);
        auto rng = vq::rand09::StdRng::seed_from_u64(seed);
        for (auto v : rng.sample_indices(n, k)) std::printf("%" PRIu64 " ", v);
        std::printf("\n");
        for (int i = 0; i < 5; ++i) std::printf("%" PRIu64 " ", rng.random_range(n));
        std::printf("\n");
        return 0;
    }
    if (mode == "nogpu") {  // must fail loudly: there is no CPU fallback
        try { vq::Engine e(0); } catch (const vq::VqError& e) {
            std::printf("%s\n", e.what());
            return e.kind == vq::ErrorKind::FfiError ? 0 : 1;
        }
        std::printf("an engine was created\n");
        return 2;
    }
    if (mode == "gpu") {
        using namespace vq;
        auto rows = make_rows(600, 16, 7);
        ProductQuantizer pq(rows, 2, 16, 4, Distance::CosineDistance, 42);
        std::printf("cb");
        for (float v : pq.codebooks()) { uint32_t b; std::memcpy(&b, &v, 4); std::printf(" %08x", b); }
        std::printf("\nq");
        for (uint16_t h : pq.quantize(rows[5])) std::printf(" %04x", h);
        std::printf("\n");
        TSVQ t(rows, 3, Distance::Euclidean);
        std::printf("t");
        for (uint16_t h : t.quantize(rows[9])) std::printf(" %04x", h);
        std::printf("\n");
        BinaryQuantizer bq(0.5f, 0, 1);
        ScalarQuantizer sq(-1.f, 8.f, 256);
        std::printf("b");
        for (uint8_t c : bq.quantize(rows[0])) std::printf(" %u", c);
        std::printf("\ns");
        for (uint8_t c : sq.quantize(rows[0])) std::printf(" %u", c);
        std::printf("\nd %.9g\n", distance_compute(Distance::Manhattan, rows[0], rows[1]));
        // library-owned communicator with one rank: the row-sharded constructor must give the same codebooks
        {
            auto eng = Engine::shared();
            eng->comm_init(Engine::comm_unique_id(), 0, 1);
            auto info = eng->comm_info();
            std::vector<float> flat;
            for (const auto& r : rows) flat.insert(flat.end(), r.begin(), r.end());
            RowShard sh{0, rows.size()};
            ProductQuantizer pqs(flat.data(), rows.size(), 16, 2, 16, 4, Distance::CosineDistance, 42, sh, eng);
            std::printf("comm %d %d %d\n", info.first, info.second, pqs.codebooks() == pq.codebooks() ? 1 : 0);
            eng->comm_destroy();
        }
        return 0;
    }
    return 64;
}
