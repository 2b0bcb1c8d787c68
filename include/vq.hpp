// vq.hpp -- C++ host-side mirror of the vq crate's public API over the C ABI (include/vqb200.h).
//
// north_star asks for a Rust host; no Rust toolchain exists in this environment, so -- as the
// reference is compiled code -- the host side above the C ABI is written in C++ with the crate's
// own names, argument meaning, validation order and error kinds:
//
//   vq::Distance                      src/core/distance.rs:8-65
//   vq::VqError (+ kinds)             src/core/error.rs:5-28
//   vq::BinaryQuantizer               src/bq.rs:55-118
//   vq::ScalarQuantizer               src/sq.rs:63-151
//   vq::ProductQuantizer              src/pq.rs:83-209     (construction == training, like the crate)
//   vq::TSVQ                          src/tsvq.rs:195-265
//   vq::rand09::StdRng                the `rand 0.9` calls of src/core/vector.rs:412-413,450
//                                     (PARITY UNPINNED: the rand crate is not vendored in the reference,
//                                      see DESIGN.md section 2; the stream can be replaced through IndexSource)
//
// `quantize` takes ONE vector and returns the reconstructed f16 centroid values (PQ/TSVQ) or one
// byte per element (BQ/SQ), exactly like the Quantizer trait (src/core/quantizer.rs:29-63); the
// `*_batch` methods are the batch forms.  f16 values are carried as IEEE binary16 bit patterns.
// Header-only; link with -lvqb200.  There is no CPU fallback: constructing an Engine without an
// sm_100 GPU throws.
#pragma once

#include <cmath>
#include <cstdint>
#include <array>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

#include "vqb200.h"

namespace vq {

// ------------------------------------------------------------------------------------ errors
enum class ErrorKind { DimensionMismatch, EmptyInput, InvalidParameter, InvalidData, FfiError };

class VqError : public std::runtime_error {  // src/core/error.rs:5-28 (messages follow the #[error(..)] strings)
public:
    VqError(ErrorKind k, const std::string& msg) : std::runtime_error(msg), kind(k) {}
    static VqError dimension_mismatch(size_t expected, size_t found) {
        return VqError(ErrorKind::DimensionMismatch,
                       "Dimension mismatch: expected " + std::to_string(expected) + ", found " + std::to_string(found));
    }
    static VqError empty_input() { return VqError(ErrorKind::EmptyInput, "Empty input: at least one vector is required"); }
    static VqError invalid_parameter(const std::string& parameter, const std::string& reason) {
        return VqError(ErrorKind::InvalidParameter, "Invalid parameter '" + parameter + "': " + reason);
    }
    static VqError ffi(const std::string& msg) { return VqError(ErrorKind::FfiError, "FFI error: " + msg); }
    ErrorKind kind;
};

// ------------------------------------------------------------------------------------ engine
class Engine {
public:
    explicit Engine(int device = 0) {
        int rc = vqb_ctx_create(device, &ctx_);
        if (rc == VQB_ERR_UNSUPPORTED_DEVICE)
            throw VqError::ffi("vq needs an sm_100 (B200) GPU: no CUDA device or wrong architecture; there is no CPU fallback");
        if (rc != VQB_SUCCESS) throw VqError::ffi("vqb_ctx_create failed (" + std::to_string(rc) + ")");
    }
    ~Engine() { if (ctx_) vqb_ctx_destroy(ctx_); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    vqb_ctx* ctx() const { return ctx_; }
    void check(int rc) const {
        if (rc == VQB_SUCCESS) return;
        const char* m = vqb_last_error(ctx_);
        std::string msg = m ? m : "";
        if (rc == VQB_ERR_EMPTY_INPUT) throw VqError::empty_input();
        if (rc == VQB_ERR_INVALID_INPUT || rc == VQB_ERR_DIM_MISMATCH || rc == VQB_ERR_NULL_PTR)
            throw VqError(ErrorKind::InvalidParameter, msg.empty() ? "invalid input" : msg);
        throw VqError::ffi(msg.empty() ? "status " + std::to_string(rc) : msg);
    }
    static std::shared_ptr<Engine> shared(int device = 0) {  // one lazily created engine per process
        static std::shared_ptr<Engine> e;
        if (!e) e = std::make_shared<Engine>(device);
        return e;
    }
    // ---- multi-GPU: one process and one Engine per GPU; the library owns the NCCL communicator.  Rank 0
    // draws an id and ships its VQB_COMM_ID_BYTES to the other ranks by any means (MPI, a file, a socket); every rank
    // then calls comm_init (collective).  Training with a RowShard issues one ncclAllReduce per iteration.
    static std::array<uint8_t, VQB_COMM_ID_BYTES> comm_unique_id() {
        std::array<uint8_t, VQB_COMM_ID_BYTES> id{};
        if (vqb_comm_unique_id(id.data()) != VQB_SUCCESS) throw VqError::ffi("vqb_comm_unique_id failed (is libnccl.so.2 loadable?)");
        return id;
    }
    void comm_init(const std::array<uint8_t, VQB_COMM_ID_BYTES>& id, int rank, int world) { check(vqb_comm_init_rank(ctx_, id.data(), rank, world)); }
    void comm_destroy() { check(vqb_comm_destroy(ctx_)); }
    std::pair<int, int> comm_info() const {  // {rank, world}; {0, 1} without a communicator
        int r = 0, w = 1;
        check(vqb_comm_info(ctx_, &r, &w));
        return {r, w};
    }
private:
    vqb_ctx* ctx_ = nullptr;
};

// The rows this rank holds of a training set that is sharded over the ranks of the engine's communicator: rows
// [row_offset, row_offset + n_local) of n_global.  Initial and re-seed indices are GLOBAL row numbers (the same stream on
// every rank); the owner of a row contributes it through the per-iteration sum.
struct RowShard { uint64_t row_offset = 0, n_global = 0; };

inline std::string get_simd_backend() { return vqb_backend_name(); }  // src/core/hsdlib_ffi.rs:144-155

// ------------------------------------------------------------------------------------ Distance
enum class Distance : int {  // enum order of src/core/distance.rs:8-17 == VQB_* ids
    SquaredEuclidean = VQB_SQUARED_EUCLIDEAN, Euclidean = VQB_EUCLIDEAN, Manhattan = VQB_MANHATTAN,
    CosineDistance = VQB_COSINE,
    Chebyshev = VQB_CHEBYSHEV   // EXTENSION: not in the reference; distance_compute and ProductQuantizer encoding only
};
inline const char* distance_name(Distance d) {  // distance.rs:21-29
    switch (d) {
        case Distance::SquaredEuclidean: return "squared_euclidean";
        case Distance::Euclidean: return "euclidean";
        case Distance::Manhattan: return "manhattan";
        case Distance::Chebyshev: return "chebyshev";
        default: return "cosine";
    }
}
inline float distance_compute(Distance d, const std::vector<float>& a, const std::vector<float>& b,
                              std::shared_ptr<Engine> eng = nullptr) {  // distance.rs:48-65
    if (a.size() != b.size()) throw VqError::dimension_mismatch(a.size(), b.size());
    if (!eng) eng = Engine::shared();
    float out = 0.f;
    eng->check(vqb_distance_batch(eng->ctx(), (int)d, a.data(), b.data(), 1, a.size(), &out));
    return out;
}

// ------------------------------------------------------------------------------------ rand 0.9
namespace rand09 {
inline uint32_t rotl32(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
inline void chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]) {
    uint32_t st[16] = {0x61707865u, 0x3320646Eu, 0x79622D32u, 0x6B206574u, key[0], key[1], key[2], key[3], key[4], key[5],
                       key[6], key[7], (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t x[16];
    std::memcpy(x, st, sizeof(x));
    auto qr = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
    };
    for (int r = 0; r < rounds / 2; ++r) {
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + st[i];
}

// StdRng = ChaCha12, 64-bit block counter, stream 0, 4-block (64-word) output buffer.
class StdRng {
public:
    static StdRng seed_from_u64(uint64_t state) {  // rand_core: PCG32 expansion to 8 little-endian words
        StdRng r;
        for (int i = 0; i < 8; ++i) {
            state = state * 6364136223846793005ull + 11634580027462260723ull;
            uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
            uint32_t rot = (uint32_t)(state >> 59);
            r.key_[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
        }
        return r;
    }
    uint32_t next_u32() {
        if (index_ >= 64) generate(0);
        return buf_[index_++];
    }
    uint64_t next_u64() {  // rand_core BlockRng::next_u64
        if (index_ < 63) {
            uint64_t lo = buf_[index_], hi = buf_[index_ + 1];
            index_ += 2;
            return (hi << 32) | lo;
        }
        if (index_ >= 64) {
            generate(2);
            return ((uint64_t)buf_[1] << 32) | buf_[0];
        }
        uint64_t lo = buf_[63];
        generate(1);
        return ((uint64_t)buf_[0] << 32) | lo;
    }
    // rng.random_range(..n) for usize: sampled as u32 when n fits (0.9 portability rule); Canon's method
    uint64_t random_range(uint64_t n) {
        if (n == 0) throw std::invalid_argument("empty range");
        if (n > 0xFFFFFFFFull) {
            unsigned __int128 prod = (unsigned __int128)next_u64() * n;
            uint64_t result = (uint64_t)(prod >> 64), lo = (uint64_t)prod;
            if (lo > (uint64_t)(0 - n)) {
                uint64_t new_hi = (uint64_t)(((unsigned __int128)next_u64() * n) >> 64);
                if (lo + new_hi < lo) ++result;
            }
            return result;
        }
        uint32_t r = (uint32_t)n;
        uint64_t prod = (uint64_t)next_u32() * r;
        uint32_t result = (uint32_t)(prod >> 32), lo = (uint32_t)prod;
        if (lo > (uint32_t)(0u - r)) {
            uint32_t new_hi = (uint32_t)(((uint64_t)next_u32() * r) >> 32);
            if ((uint64_t)lo + new_hi > 0xFFFFFFFFull) ++result;
        }
        return result;
    }
    // rand::seq::index::sample(rng, length, amount): `amount` distinct indices in algorithm order
    std::vector<uint64_t> sample_indices(uint64_t length, uint64_t amount) {
        if (amount > length) throw std::invalid_argument("`amount` of samples must be less than or equal to `length`");
        if (length > 0xFFFFFFFFull) return sample_rejection(length, amount, true);
        const int j = length >= 500000 ? 1 : 0;
        if (amount < 163) {
            const float c0[2] = {1.6f, 8.0f / 45.0f}, c1[2] = {10.0f, 70.0f / 9.0f};
            float amount_fp = (float)amount, m4 = c0[j] * amount_fp;
            if (amount > 11 && (float)length < (c1[j] + m4) * amount_fp) return sample_inplace(length, amount);
            return sample_floyd(length, amount);
        }
        const float c[2] = {270.0f, 330.0f / 9.0f};
        if ((float)length < c[j] * (float)amount) return sample_inplace(length, amount);
        return sample_rejection(length, amount, false);
    }
private:
    void generate(int index) {
        for (int b = 0; b < 4; ++b) chacha_block(key_, counter_ + b, 0, 12, buf_ + 16 * b);
        counter_ += 4;
        index_ = index;
    }
    uint32_t uniform_u32(uint32_t length) {  // Uniform::<u32>::new(0, length).sample
        uint32_t thresh = (0u - length) % length;
        for (;;) {
            uint64_t prod = (uint64_t)next_u32() * length;
            if ((uint32_t)prod >= thresh) return (uint32_t)(prod >> 32);
        }
    }
    uint64_t uniform_u64(uint64_t length) {
        uint64_t thresh = (0 - length) % length;
        for (;;) {
            unsigned __int128 prod = (unsigned __int128)next_u64() * length;
            if ((uint64_t)prod >= thresh) return (uint64_t)(prod >> 64);
        }
    }
    std::vector<uint64_t> sample_floyd(uint64_t length, uint64_t amount) {
        std::vector<uint64_t> ind;
        for (uint64_t j = length - amount; j < length; ++j) {
            uint64_t t = random_range(j + 1);
            for (auto& v : ind)
                if (v == t) { v = j; break; }
            ind.push_back(t);
        }
        return ind;
    }
    std::vector<uint64_t> sample_inplace(uint64_t length, uint64_t amount) {
        std::vector<uint64_t> ind(length);
        for (uint64_t i = 0; i < length; ++i) ind[i] = i;
        for (uint64_t i = 0; i < amount; ++i) {
            uint64_t j = i + random_range(length - i);
            std::swap(ind[i], ind[j]);
        }
        ind.resize(amount);
        return ind;
    }
    std::vector<uint64_t> sample_rejection(uint64_t length, uint64_t amount, bool wide) {
        std::unordered_set<uint64_t> seen;
        std::vector<uint64_t> out;
        for (uint64_t i = 0; i < amount; ++i)
            for (;;) {
                uint64_t pos = wide ? uniform_u64(length) : uniform_u32((uint32_t)length);
                if (seen.insert(pos).second) { out.push_back(pos); break; }
            }
        return out;
    }
    uint32_t key_[8] = {0};
    uint64_t counter_ = 0;
    uint32_t buf_[64];
    int index_ = 64;
};
}  // namespace rand09

// Where ProductQuantizer takes its row indices from.  Default = rand09::StdRng::seed_from_u64(seed + subspace)
// (src/pq.rs:130): choose_multiple for the initial centroids, choose for every re-seed.
struct IndexSource {
    virtual ~IndexSource() = default;
    virtual std::vector<uint64_t> choose_multiple(uint32_t subspace, uint64_t n, uint64_t k) = 0;
    virtual uint64_t choose(uint32_t subspace, uint64_t n) = 0;
};
class Rand09Source : public IndexSource {
public:
    Rand09Source(uint64_t seed, size_t m) {
        for (size_t i = 0; i < m; ++i) rng_.push_back(rand09::StdRng::seed_from_u64(seed + i));  // wrapping add, pq.rs:130
    }
    std::vector<uint64_t> choose_multiple(uint32_t s, uint64_t n, uint64_t k) override { return rng_[s].sample_indices(n, std::min(k, n)); }
    uint64_t choose(uint32_t s, uint64_t n) override { return rng_[s].random_range(n); }
private:
    std::vector<rand09::StdRng> rng_;
};

// ------------------------------------------------------------------------------------ BQ
class BinaryQuantizer {  // src/bq.rs
public:
    BinaryQuantizer(float threshold, uint8_t low, uint8_t high, std::shared_ptr<Engine> eng = nullptr)
        : threshold_(threshold), low_(low), high_(high), eng_(std::move(eng)) {
        if (!std::isfinite(threshold))
            throw VqError::invalid_parameter("threshold", "must be finite (not NaN or infinite)");     // bq.rs:56-61
        if (low >= high) throw VqError::invalid_parameter("low/high", "low must be less than high");  // bq.rs:62-67
    }
    float threshold() const { return threshold_; }
    uint8_t low() const { return low_; }
    uint8_t high() const { return high_; }
    std::vector<uint8_t> quantize(const std::vector<float>& v) {  // bq.rs:94-105
        std::vector<uint8_t> out(v.size());
        if (!v.empty()) engine().check(vqb_bq_quantize(engine().ctx(), v.data(), v.size(), threshold_, low_, high_, out.data()));
        return out;
    }
    std::vector<float> dequantize(const std::vector<uint8_t>& c) {  // bq.rs:107-118
        std::vector<float> out(c.size());
        if (!c.empty()) engine().check(vqb_bq_dequantize(engine().ctx(), c.data(), c.size(), low_, high_, out.data()));
        return out;
    }
private:
    Engine& engine() { if (!eng_) eng_ = Engine::shared(); return *eng_; }
    float threshold_; uint8_t low_, high_;
    std::shared_ptr<Engine> eng_;
};

// ------------------------------------------------------------------------------------ SQ
class ScalarQuantizer {  // src/sq.rs
public:
    ScalarQuantizer(float min, float max, size_t levels, std::shared_ptr<Engine> eng = nullptr)
        : min_(min), max_(max), levels_(levels), eng_(std::move(eng)) {
        if (!std::isfinite(min)) throw VqError::invalid_parameter("min", "must be finite (not NaN or infinite)");         // sq.rs:64-69
        if (!std::isfinite(max)) throw VqError::invalid_parameter("max", "must be finite (not NaN or infinite)");         // sq.rs:70-75
        if (max <= min) throw VqError::invalid_parameter("max", "must be greater than min");                              // sq.rs:76-81
        if (levels < 2) throw VqError::invalid_parameter("levels", "must be at least 2");                                 // sq.rs:82-87
        if (levels > 256) throw VqError::invalid_parameter("levels", "must be no more than 256 to fit in u8");            // sq.rs:88-93
        step_ = (max - min) / (float)(levels - 1);                                                                        // sq.rs:94
    }
    float min() const { return min_; }
    float max() const { return max_; }
    size_t levels() const { return levels_; }
    float step() const { return step_; }
    std::vector<uint8_t> quantize(const std::vector<float>& v) {  // sq.rs:130-144
        std::vector<uint8_t> out(v.size());
        if (!v.empty()) engine().check(vqb_sq_quantize(engine().ctx(), v.data(), v.size(), min_, max_, step_, (uint32_t)levels_, out.data()));
        return out;
    }
    std::vector<float> dequantize(const std::vector<uint8_t>& c) {  // sq.rs:146-151
        std::vector<float> out(c.size());
        if (!c.empty()) engine().check(vqb_sq_dequantize(engine().ctx(), c.data(), c.size(), min_, step_, out.data()));
        return out;
    }
private:
    Engine& engine() { if (!eng_) eng_ = Engine::shared(); return *eng_; }
    float min_, max_, step_ = 0.f; size_t levels_;
    std::shared_ptr<Engine> eng_;
};

// ------------------------------------------------------------------------------------ PQ
class ProductQuantizer {  // src/pq.rs
public:
    // ProductQuantizer::new(training_data: &[&[f32]], m, k, max_iters, distance, seed)  (pq.rs:83-90).
    // `rows` are the row slices; every row must have rows[0].size() elements.
    ProductQuantizer(const std::vector<std::vector<float>>& rows, size_t m, size_t k, size_t max_iters, Distance distance,
                     uint64_t seed, std::shared_ptr<Engine> eng = nullptr, IndexSource* indices = nullptr) {
        if (rows.empty()) throw VqError::empty_input();                                                   // pq.rs:91-93
        const size_t dim = rows[0].size();
        for (const auto& r : rows)
            if (r.size() != dim) throw VqError::dimension_mismatch(dim, r.size());                          // pq.rs:95-104
        std::vector<float> flat(rows.size() * dim);
        for (size_t i = 0; i < rows.size(); ++i) std::memcpy(flat.data() + i * dim, rows[i].data(), dim * sizeof(float));
        init(flat.data(), rows.size(), dim, m, k, max_iters, distance, seed, std::move(eng), indices);
    }
    // Zero-copy form: `data` is n x dim row-major (host or device pointer).
    ProductQuantizer(const float* data, size_t n, size_t dim, size_t m, size_t k, size_t max_iters, Distance distance,
                     uint64_t seed, std::shared_ptr<Engine> eng = nullptr, IndexSource* indices = nullptr) {
        if (n == 0) throw VqError::empty_input();
        init(data, n, dim, m, k, max_iters, distance, seed, std::move(eng), indices);
    }
    // Row-sharded form (no counterpart in the single-process crate): `data` holds this rank's n_local rows of
    // shard.n_global; the engine must have a communicator (Engine::comm_init).  Every rank gets the same codebooks.
    ProductQuantizer(const float* data, size_t n_local, size_t dim, size_t m, size_t k, size_t max_iters, Distance distance,
                     uint64_t seed, const RowShard& shard, std::shared_ptr<Engine> eng, IndexSource* indices = nullptr) {
        if (shard.n_global == 0) throw VqError::empty_input();
        init(data, n_local, dim, m, k, max_iters, distance, seed, std::move(eng), indices, &shard);
    }
    ~ProductQuantizer() { if (pq_) vqb_pq_destroy(pq_); }
    ProductQuantizer(const ProductQuantizer&) = delete;
    ProductQuantizer& operator=(const ProductQuantizer&) = delete;

    size_t num_subspaces() const { return m_; }        // pq.rs:144-161
    size_t sub_dim() const { return dim_ / m_; }
    size_t dim() const { return dim_; }
    Distance distance_metric() const { return distance_; }
    const std::vector<float>& codebooks() const { return cb_; }   // [m][k][sub_dim]
    const std::vector<uint32_t>& iters_run() const { return iters_; }

    std::vector<uint16_t> quantize(const std::vector<float>& v) {  // pq.rs:167-199 -> Vec<f16> of length dim
        if (v.size() != dim_) throw VqError::dimension_mismatch(dim_, v.size());
        std::vector<uint16_t> out(dim_);
        eng_->check(vqb_pq_encode(pq_, v.data(), 1, VQB_ASSIGN_AUTO, nullptr, 1, out.data()));
        return out;
    }
    std::vector<float> dequantize(const std::vector<uint16_t>& q) {  // pq.rs:201-209
        if (q.size() != dim_) throw VqError::dimension_mismatch(dim_, q.size());
        std::vector<float> out(dim_);
        eng_->check(vqb_f16_dequantize(eng_->ctx(), q.data(), q.size(), out.data()));
        return out;
    }
    // batch forms (ROADMAP.md:30-31): either output may be null; pointers may be host or device
    void quantize_batch(const float* x, size_t n, uint8_t* codes /* n*m, k <= 256 */, uint16_t* recon_f16 /* n*dim */) {
        eng_->check(vqb_pq_encode(pq_, x, n, VQB_ASSIGN_AUTO, codes, 1, recon_f16));
    }
    void decode(const uint8_t* codes, size_t n, float* out) { eng_->check(vqb_pq_decode(pq_, codes, 1, n, out)); }

private:
    struct ReseedState { IndexSource* src; uint64_t n; };
    static uint64_t reseed_tramp(void* user, uint32_t s) {
        auto* st = static_cast<ReseedState*>(user);
        return st->src->choose(s, st->n);
    }
    void init(const float* data, size_t n_local, size_t dim, size_t m, size_t k, size_t max_iters, Distance distance, uint64_t seed,
              std::shared_ptr<Engine> eng, IndexSource* indices, const RowShard* shard = nullptr) {
        const size_t n = shard ? (size_t)shard->n_global : n_local;   // the crate's checks and the index stream see the whole set
        if (m == 0) throw VqError::invalid_parameter("m", "must be greater than 0");  // the crate panics on dim % 0 (pq.rs:112)
        if (dim < m) throw VqError::invalid_parameter("m", "must be at most the data dimension (" + std::to_string(dim) + ")");  // pq.rs:106-111
        if (dim % m) throw VqError::invalid_parameter("m", "dimension (" + std::to_string(dim) + ") must be divisible by m");   // pq.rs:112-117
        if (k == 0) throw VqError::invalid_parameter("k", "must be greater than 0");                                            // vector.rs:399-404
        if (n < k) throw VqError::invalid_parameter("k", "not enough data points (" + std::to_string(n) + ") for " + std::to_string(k) + " clusters");  // vector.rs:405-410
        m_ = m; k_ = k; dim_ = dim; distance_ = distance;
        eng_ = eng ? std::move(eng) : Engine::shared();
        Rand09Source own(seed, m);
        IndexSource* src = indices ? indices : &own;
        std::vector<uint64_t> init_idx;
        init_idx.reserve(m * k);
        for (size_t s = 0; s < m; ++s) {
            auto v = src->choose_multiple((uint32_t)s, n, k);   // vector.rs:413
            init_idx.insert(init_idx.end(), v.begin(), v.end());
        }
        ReseedState st{src, n};
        vqb_train_opts o;
        std::memset(&o, 0, sizeof(o));
        o.struct_size = sizeof(o);
        o.update_mode = VQB_UPDATE_ORDERED;   // the reference's summation order
        o.assign_mode = VQB_ASSIGN_AUTO;
        o.reseed = &ProductQuantizer::reseed_tramp;
        o.reseed_user = &st;
        if (shard) { o.flags = VQB_TRAIN_USE_COMM; o.row_offset = shard->row_offset; o.n_global = shard->n_global; }
        cb_.assign(m * k * (dim / m), 0.f);
        iters_.assign(m, 0);
        eng_->check(vqb_pq_train(eng_->ctx(), data, n_local, dim, m, k, max_iters, init_idx.data(), &o, cb_.data(), iters_.data()));
        eng_->check(vqb_pq_create(eng_->ctx(), cb_.data(), m, k, dim / m, (int)distance, &pq_));
    }
    size_t m_ = 0, k_ = 0, dim_ = 0;
    Distance distance_ = Distance::Euclidean;
    std::shared_ptr<Engine> eng_;
    std::vector<float> cb_;
    std::vector<uint32_t> iters_;
    vqb_pq* pq_ = nullptr;
};

// ------------------------------------------------------------------------------------ TSVQ
class TSVQ {  // src/tsvq.rs
public:
    TSVQ(const std::vector<std::vector<float>>& rows, size_t max_depth, Distance distance, std::shared_ptr<Engine> eng = nullptr) {
        if (rows.empty()) throw VqError::empty_input();                                                   // tsvq.rs:200-202
        dim_ = rows[0].size();
        for (const auto& r : rows)
            if (r.size() != dim_) throw VqError::dimension_mismatch(dim_, r.size());                        // tsvq.rs:204-213
        std::vector<float> flat(rows.size() * dim_);
        for (size_t i = 0; i < rows.size(); ++i) std::memcpy(flat.data() + i * dim_, rows[i].data(), dim_ * sizeof(float));
        distance_ = distance;
        eng_ = eng ? std::move(eng) : Engine::shared();
        eng_->check(vqb_tsvq_train(eng_->ctx(), flat.data(), rows.size(), dim_, max_depth, (int)distance, &t_));
    }
    ~TSVQ() { if (t_) vqb_tsvq_destroy(t_); }
    TSVQ(const TSVQ&) = delete;
    TSVQ& operator=(const TSVQ&) = delete;
    size_t dim() const { return dim_; }                     // tsvq.rs:226-234
    Distance distance_metric() const { return distance_; }
    std::vector<uint16_t> quantize(const std::vector<float>& v) {  // tsvq.rs:239-255
        if (v.size() != dim_) throw VqError::dimension_mismatch(dim_, v.size());
        std::vector<uint16_t> out(dim_);
        eng_->check(vqb_tsvq_encode(t_, v.data(), 1, nullptr, out.data()));
        return out;
    }
    std::vector<float> dequantize(const std::vector<uint16_t>& q) {  // tsvq.rs:257-265
        if (q.size() != dim_) throw VqError::dimension_mismatch(dim_, q.size());
        std::vector<float> out(dim_);
        eng_->check(vqb_f16_dequantize(eng_->ctx(), q.data(), q.size(), out.data()));
        return out;
    }
    void quantize_batch(const float* x, size_t n, uint32_t* leaf, uint16_t* recon_f16) {
        eng_->check(vqb_tsvq_encode(t_, x, n, leaf, recon_f16));
    }
private:
    size_t dim_ = 0;
    Distance distance_ = Distance::Euclidean;
    std::shared_ptr<Engine> eng_;
    vqb_tsvq* t_ = nullptr;
};

}  // namespace vq
