// Host side of the drop-in `PolarCode` class: code construction, encoder and the batched
// BPSK/AWGN BLER harness, on top of the GPU decoder's C ABI (include/polar_b200.h).
// Reference behaviour followed: PolarC/PolarCode.cpp (cited per function).
#include "PolarCode.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>

#include "polar_b200.h"

namespace {
void check(int rc, const char* what) {
    if (rc != POLAR_B200_OK)
        throw std::runtime_error(std::string(what) + ": " + polar_b200_strerror(rc));
}
}  // namespace

PolarCode::PolarCode(uint8_t num_layers, uint16_t info_length, double epsilon, uint16_t crc_size)
    : _n(num_layers), _info_length(info_length), _crc_size(crc_size), _design_epsilon(epsilon) {
    const char* ex = getenv("POLAR_B200_EXACT");
    if (ex && *ex && *ex != '0') arithmetic_mode = POLAR_B200_MODE_F64;
    if (const char* m = getenv("POLAR_B200_MODE")) {
        const std::string v(m);
        if (v == "fp32") arithmetic_mode = POLAR_B200_MODE_FP32;
        else if (v == "strict") arithmetic_mode = POLAR_B200_MODE_STRICT;
        else if (v == "f64") arithmetic_mode = POLAR_B200_MODE_F64;
    }
    _block_length = (uint16_t)(1 << _n);
    _frozen_bits.resize(_block_length);
    _bit_rev_order.resize(_block_length);
    create_bit_rev_order();
    initialize_frozen_bits();
}

PolarCode::~PolarCode() {
    if (_ctx) polar_b200_destroy(_ctx);
}

// PolarCode.cpp:647-656.
void PolarCode::create_bit_rev_order() {
    for (unsigned i = 0; i < _block_length; ++i) {
        unsigned r = 0;
        for (unsigned b = 0; b < _n; ++b)
            if (i & (1u << b)) r |= 1u << (_n - 1 - b);
        _bit_rev_order[i] = (uint16_t)r;
    }
}

// PolarCode.cpp:17-58. The arithmetic order of the Bhattacharyya recursion and the shape of
// the std::sort call (uint16_t sequence, int-taking comparator, ties left to libstdc++'s
// introsort) are kept, because the *permutation* of exactly tied channels decides which info
// index maps to which position.
void PolarCode::initialize_frozen_bits() {
    const int N = _block_length;
    std::vector<double> z(N, _design_epsilon);
    for (int stage = 0; stage < _n; ++stage) {
        const int inc = 1 << stage;
        for (int j = 0; j < inc; ++j)
            for (int i = 0; i < N; i += 2 * inc) {
                const double c1 = z[i + j], c2 = z[i + j + inc];
                z[i + j] = c1 + c2 - c1 * c2;
                z[i + j + inc] = c1 * c2;
            }
    }
    _channel_order_descending.resize(N);
    for (int i = 0; i < N; ++i) _channel_order_descending[i] = (uint16_t)i;
    std::sort(_channel_order_descending.begin(), _channel_order_descending.end(),
              [&](int i1, int i2) { return z[_bit_rev_order.at(i1)] < z[_bit_rev_order.at(i2)]; });

    const int eff = _info_length + _crc_size;
    for (int i = 0; i < N; ++i) _frozen_bits[_channel_order_descending[i]] = (i < eff) ? 0 : 1;

    _crc_matrix.resize(_crc_size);
    for (int r = 0; r < _crc_size; ++r) {
        _crc_matrix[r].resize(_info_length);
        for (int j = 0; j < _info_length; ++j) _crc_matrix[r][j] = (uint8_t)(rand() % 2);
    }
}

// PolarCode.cpp:60-91.
std::vector<uint8_t> PolarCode::encode(std::vector<uint8_t> info_bits) {
    const int N = _block_length;
    std::vector<uint8_t> u(N, 0), coded(N);
    for (int i = 0; i < _info_length; ++i) u[_channel_order_descending[i]] = info_bits.at(i);
    for (int r = 0; r < _crc_size; ++r) {
        unsigned acc = 0;
        for (int j = 0; j < _info_length; ++j) acc += (unsigned)_crc_matrix[r][j] * info_bits.at(j);
        u[_channel_order_descending[_info_length + r]] = (uint8_t)(acc % 2);
    }
    for (int stage = 0; stage < _n; ++stage) {
        const int inc = 1 << stage;
        for (int base = 0; base < N; base += 2 * inc)
            for (int j = 0; j < inc; ++j) u[base + j] = (uint8_t)((u[base + j] + u[base + j + inc]) % 2);
    }
    for (int i = 0; i < N; ++i) coded[i] = u[_bit_rev_order[i]];
    return coded;
}

polar_b200_ctx* PolarCode::device_ctx(int min_batch) {
    if (_ctx && _ctx_batch >= min_batch) return _ctx;
    if (_ctx) {     // the ctx (tables, scratch, counters, handles given out) stays; only its host staging grows
        check(polar_b200_reserve(_ctx, min_batch), "polar_b200_reserve");
        _ctx_batch = min_batch;
        return _ctx;
    }
    std::vector<uint8_t> flat((size_t)_crc_size * _info_length);
    for (int r = 0; r < _crc_size; ++r)
        std::copy(_crc_matrix[r].begin(), _crc_matrix[r].end(), flat.begin() + (size_t)r * _info_length);
    int batch = std::max(min_batch, 1);
    check(polar_b200_create(&_ctx, device, _n, _info_length, _crc_size, _frozen_bits.data(),
                            _channel_order_descending.data(), _crc_size ? flat.data() : nullptr, 127, batch),
          "polar_b200_create");
    _ctx_batch = batch;
    return _ctx;
}

void PolarCode::decode_scl_llr_batch_packed(const float* llr, int B, uint16_t list_size, uint32_t* info_packed) {
    if (list_size >= 128)   // the reference's uint8_t loop counters never terminate here (PolarCode.cpp:525)
        throw std::invalid_argument("PolarCode: list_size must be < 128");
    check(polar_b200_decode_scl_llr_host_ex(device_ctx(B), llr, B, list_size, info_packed, arithmetic_mode, nullptr),
          "polar_b200_decode_scl_llr_host_ex");
}

void PolarCode::decode_scl_llr_batch_packed_double(const double* llr, int B, uint16_t list_size, uint32_t* info_packed) {
    if (list_size >= 128) throw std::invalid_argument("PolarCode: list_size must be < 128");
    if (arithmetic_mode == POLAR_B200_MODE_F64) return decode_scl_llr_batch_packed_f64(llr, B, list_size, info_packed);
    if (arithmetic_mode == POLAR_B200_MODE_STRICT) {
        check(polar_b200_decode_scl_llr_f64_strict_host(device_ctx(B), llr, B, list_size, info_packed, nullptr),
              "polar_b200_decode_scl_llr_f64_strict_host");
        return;
    }
    std::vector<float> f((size_t)B * _block_length);
    for (size_t i = 0; i < f.size(); ++i) f[i] = (float)llr[i];
    decode_scl_llr_batch_packed(f.data(), B, list_size, info_packed);
}

void PolarCode::decode_scl_llr_batch_packed_f64(const double* llr, int B, uint16_t list_size, uint32_t* info_packed) {
    if (list_size >= 128) throw std::invalid_argument("PolarCode: list_size must be < 128");
    check(polar_b200_decode_scl_llr_f64_host(device_ctx(B), llr, B, list_size, info_packed, nullptr),
          "polar_b200_decode_scl_llr_f64_host");
}

std::vector<uint8_t> PolarCode::decode_scl_llr_batch(const float* llr, int B, uint16_t list_size) {
    const int KW = info_words(), K = _info_length;
    std::vector<uint32_t> packed((size_t)B * KW);
    decode_scl_llr_batch_packed(llr, B, list_size, packed.data());
    std::vector<uint8_t> out((size_t)B * K);
    for (int b = 0; b < B; ++b)
        for (int j = 0; j < K; ++j) out[(size_t)b * K + j] = (packed[(size_t)b * KW + (j >> 5)] >> (j & 31)) & 1u;
    return out;
}

void PolarCode::decode_scl_llr_device(const float* llr_dev, int B, uint16_t list_size, uint32_t* info_packed_dev,
                                      void* cuda_stream, float* margin_dev) {
    check(polar_b200_decode_scl_llr_ex(device_ctx(1), llr_dev, B, list_size, info_packed_dev, arithmetic_mode, margin_dev,
                                       cuda_stream),
          "polar_b200_decode_scl_llr_ex");
}

// PolarCode.cpp:130-148: one codeword by value in, K bytes out.
std::vector<uint8_t> PolarCode::decode_scl_llr(std::vector<double> llr, uint16_t list_size) {
    if (arithmetic_mode != POLAR_B200_MODE_FP32) {
        // one codeword is latency-bound either way: straight to double, the reference's own arithmetic
        llr.at(_block_length - 1);
        std::vector<uint32_t> packed(info_words());
        decode_scl_llr_batch_packed_f64(llr.data(), 1, list_size, packed.data());
        std::vector<uint8_t> out(_info_length);
        for (int j = 0; j < _info_length; ++j) out[j] = (packed[j >> 5] >> (j & 31)) & 1u;
        return out;
    }
    std::vector<float> f(_block_length);
    for (int i = 0; i < _block_length; ++i) f[i] = (float)llr.at(i);
    return decode_scl_llr_batch(f.data(), 1, list_size);
}

void PolarCode::decode_scl_p1_batch_packed(const double* p1, const double* p0, int B, uint16_t list_size, uint32_t* info_packed) {
    if (list_size >= 128) throw std::invalid_argument("PolarCode: list_size must be < 128");
    check(polar_b200_decode_scl_p1_host(device_ctx(B), p1, p0, B, list_size, info_packed, nullptr),
          "polar_b200_decode_scl_p1_host");
}

// PolarCode.cpp:110-128: one codeword by value in (p1 first, then p0), K bytes out.
std::vector<uint8_t> PolarCode::decode_scl_p1(std::vector<double> p1, std::vector<double> p0, uint16_t list_size) {
    p1.at(_block_length - 1);          // the reference reads with .at(): short input throws std::out_of_range (:122-123)
    p0.at(_block_length - 1);
    std::vector<uint32_t> packed(info_words());
    decode_scl_p1_batch_packed(p1.data(), p0.data(), 1, list_size, packed.data());
    std::vector<uint8_t> out(_info_length);
    for (int j = 0; j < _info_length; ++j) out[j] = (packed[j >> 5] >> (j & 31)) & 1u;
    return out;
}

// PolarCode.cpp:658-785. Three phases instead of one nested loop:
//   1. draw info bits / noise for every run with the reference's RNG objects in its call order
//      (rand() every 100th run :703-707, default_random_engine + normal_distribution :688-689,:708-710)
//      and encode (:712);
//   2. for every Eb/N0 build the LLR batch of all runs (:744-753) and decode it with every list size
//      on the GPU -- including the cells the reference would skip;
//   3. replay the reference's counting rules (:725-742, :758-769) over the per-cell success flags.
// Skipped cells never influence counted ones, so the table equals the sequential loop's.
std::vector<std::vector<double>> PolarCode::get_bler_quick(std::vector<double> ebno_vec,
                                                           std::vector<uint8_t> list_size_vec) {
    const int max_err = bler_max_err, max_runs = bler_max_runs;
    const int N = _block_length, K = _info_length, KW = info_words();
    const size_t nl = list_size_vec.size(), ne = ebno_vec.size();
    std::vector<std::vector<double>> bler(nl, std::vector<double>(ne, 0)), num_err = bler, num_run = bler;

    const double N_0 = 1.0;
    std::normal_distribution<double> gauss_dist(0.0f, N_0);
    std::default_random_engine generator;
    auto t1 = std::chrono::high_resolution_clock::now();

    // phase 1
    std::vector<double> noise((size_t)max_runs * N);
    std::vector<float> bpsk((size_t)max_runs * N);
    std::vector<uint32_t> truth((size_t)max_runs * KW, 0);
    std::vector<uint8_t> info_bits(K, 0);
    for (int run = 0; run < max_runs; ++run) {
        if ((run % 100) == 0)
            for (int i = 0; i < K; ++i) info_bits[i] = (uint8_t)(rand() % 2);
        for (int i = 0; i < N; ++i) noise[(size_t)run * N + i] = gauss_dist(generator);
        std::vector<uint8_t> coded = encode(info_bits);
        for (int i = 0; i < N; ++i) bpsk[(size_t)run * N + i] = 2.0f * coded[i] - 1.0f;
        for (int j = 0; j < K; ++j) truth[(size_t)run * KW + (j >> 5)] |= (uint32_t)(info_bits[j] & 1) << (j & 31);
    }

    // phase 2: ok[l][e][run]
    std::vector<uint8_t> ok(nl * ne * (size_t)max_runs, 0);
    std::vector<double> llr64((size_t)max_runs * N);
    std::vector<uint32_t> dec((size_t)max_runs * KW);
    for (size_t ie = 0; ie < ne; ++ie) {
        const double a = std::pow(10.0f, ebno_vec[ie] / 20) * std::sqrt(((double)K) / ((double)N));
        for (size_t i = 0; i < noise.size(); ++i) {
            const double r = a * (double)bpsk[i] + std::sqrt(N_0 / 2) * noise[i];
            llr64[i] = -4 * r * a / N_0;
        }
        for (size_t il = 0; il < nl; ++il) {
            decode_scl_llr_batch_packed_double(llr64.data(), max_runs, list_size_vec[il], dec.data());
            for (int run = 0; run < max_runs; ++run) {
                bool same = true;
                for (int w = 0; w < KW; ++w) same &= dec[(size_t)run * KW + w] == truth[(size_t)run * KW + w];
                ok[(il * ne + ie) * max_runs + run] = same ? 1 : 0;
            }
        }
    }

    // phase 3
    for (int run = 0; run < max_runs; ++run) {
        if (bler_verbose && max_runs >= 100 && (run % (max_runs / 100)) == 0) {
            auto t2 = std::chrono::high_resolution_clock::now();
            auto us = std::chrono::duration_cast<std::chrono::microseconds>(t2 - t1).count();
            std::cout << "Running iteration " << run << "; time elapsed = " << us / 1000 / 1000 << " seconds"
                      << "; percent complete = " << (100 * run) / max_runs << "." << std::endl;
        }
        for (size_t il = 0; il < nl; ++il) {
            bool decoded_lower = false;
            for (size_t ie = 0; ie < ne; ++ie) {
                if (num_err[il][ie] > max_err) continue;
                num_run[il][ie]++;
                if (decoded_lower) continue;
                if (ok[(il * ne + ie) * max_runs + run]) decoded_lower = true;
                else num_err[il][ie]++;
            }
        }
    }
    for (size_t il = 0; il < nl; ++il)
        for (size_t ie = 0; ie < ne; ++ie) bler[il][ie] = num_err[il][ie] / num_run[il][ie];
    return bler;
}

// The multi-GPU form of the BLER loop (PolarCode.cpp:696-775): shards of the codeword index space on every device,
// the counters reduced with ncclAllReduce.
std::vector<std::vector<double>> PolarCode::bler_sweep(const std::vector<double>& ebno_vec, const std::vector<uint8_t>& list_size,
                                                       long long total, unsigned long long seed, std::vector<int> devices,
                                                       std::vector<long long>* counts_out) {
    const int ne = (int)ebno_vec.size(), nl = (int)list_size.size();
    if (ne < 1 || nl < 1 || total < 0) throw std::invalid_argument("PolarCode::bler_sweep: empty sweep");
    if (devices.empty()) {
        const int nd = polar_b200_device_count();
        if (nd < 1) throw std::runtime_error("PolarCode::bler_sweep: no CUDA device (there is no CPU fallback)");
        for (int d = 0; d < nd; ++d) devices.push_back(d);
    }
    const int nd = (int)devices.size();
    std::vector<int> lists(list_size.begin(), list_size.end());
    for (int L : lists) if (L < 1 || L >= 128) throw std::invalid_argument("PolarCode: list_size must be in 1..127");
    std::vector<uint8_t> flat((size_t)_crc_size * _info_length);
    for (int r = 0; r < _crc_size; ++r)
        std::copy(_crc_matrix[r].begin(), _crc_matrix[r].end(), flat.begin() + (size_t)r * _info_length);
    const size_t cells = (size_t)nl * ne * 2;
    std::vector<std::vector<long long>> local(nd, std::vector<long long>(cells, 0));
    std::vector<polar_b200_ctx*> ctxs(nd, nullptr);
    std::vector<int> rcs(nd, 0);
    std::vector<std::thread> workers;
    for (int d = 0; d < nd; ++d)
        workers.emplace_back([&, d] {
            // this device's shard: a contiguous block of the global index space
            const long long lo = total * d / nd, hi = total * (d + 1) / nd;
            int rc = polar_b200_create(&ctxs[d], devices[d], _n, _info_length, _crc_size, _frozen_bits.data(),
                                       _channel_order_descending.data(), _crc_size ? flat.data() : nullptr, 127, 1);
            if (!rc) rc = polar_b200_bler_sweep(ctxs[d], seed, lo, hi - lo, ebno_vec.data(), ne, lists.data(), nl, arithmetic_mode,
                                                local[d].data(), nullptr);
            rcs[d] = rc;
        });
    for (auto& w : workers) w.join();
    int first_rc = 0;
    for (int d = 0; d < nd; ++d) { if (ctxs[d]) polar_b200_destroy(ctxs[d]); if (rcs[d] && !first_rc) first_rc = rcs[d]; }
    check(first_rc, "polar_b200_bler_sweep");
    // the one collective: sum of the counters over the devices
    std::vector<polar_b200_comm*> comms(nd, nullptr);
    int rc = polar_b200_comm_init_all(comms.data(), nd, devices.data());
    if (rc == POLAR_B200_E_NONCCL && nd == 1) rc = POLAR_B200_OK;        // a single device has nothing to reduce
    else if (!rc) {
        std::vector<long long*> ptrs(nd);
        for (int d = 0; d < nd; ++d) ptrs[d] = local[d].data();
        rc = polar_b200_comm_allreduce_i64_group(comms.data(), nd, ptrs.data(), (int)cells);
    }
    for (auto* m : comms) if (m) polar_b200_comm_destroy(m);
    check(rc, "polar_b200_comm_allreduce_i64_group");
    std::vector<std::vector<double>> bler(nl, std::vector<double>(ne, 0.0));
    for (int il = 0; il < nl; ++il)
        for (int ie = 0; ie < ne; ++ie) {
            const long long e = local[0][((size_t)il * ne + ie) * 2], r = local[0][((size_t)il * ne + ie) * 2 + 1];
            bler[il][ie] = r ? (double)e / (double)r : 0.0;
        }
    if (counts_out) *counts_out = local[0];
    return bler;
}

// ---------------------------------------------------------------------------------------------
// C wrappers so the Python package (ctypes) can drive the host class. Exceptions are turned into
// a thread-local message + nonzero return.
// ---------------------------------------------------------------------------------------------
namespace {
thread_local std::string g_last_error;
template <class F>
int guarded(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { g_last_error = e.what(); return 1; }
    catch (...) { g_last_error = "unknown C++ exception"; return 1; }
}
}  // namespace

extern "C" {

const char* polar_host_last_error(void) { return g_last_error.c_str(); }

// reseed != 0: srand(1) first, i.e. the rand() state of a fresh process (what main.cpp sees).
void* polar_host_create(int n, int K, double epsilon, int crc, int reseed, int device) {
    PolarCode* p = nullptr;
    if (guarded([&] { if (reseed) srand(1); p = new PolarCode((uint8_t)n, (uint16_t)K, epsilon, (uint16_t)crc); p->device = device; }))
        return nullptr;
    return p;
}
void polar_host_destroy(void* h) { delete static_cast<PolarCode*>(h); }

void polar_host_get_construction(void* h, uint8_t* frozen, uint16_t* order, uint8_t* crc_matrix, uint16_t* bitrev) {
    PolarCode* p = static_cast<PolarCode*>(h);
    const int N = p->block_length(), K = p->info_length();
    if (frozen) std::copy(p->frozen_bits().begin(), p->frozen_bits().end(), frozen);
    if (order) std::copy(p->channel_order().begin(), p->channel_order().end(), order);
    if (bitrev) std::copy(p->bit_rev_order().begin(), p->bit_rev_order().end(), bitrev);
    if (crc_matrix)
        for (int r = 0; r < p->crc_size(); ++r) std::copy(p->crc_matrix()[r].begin(), p->crc_matrix()[r].end(), crc_matrix + (size_t)r * K);
    (void)N;
}

int polar_host_encode(void* h, const uint8_t* info, int B, uint8_t* coded) {
    PolarCode* p = static_cast<PolarCode*>(h);
    return guarded([&] {
        const int N = p->block_length(), K = p->info_length();
        for (int b = 0; b < B; ++b) {
            std::vector<uint8_t> out = p->encode(std::vector<uint8_t>(info + (size_t)b * K, info + (size_t)(b + 1) * K));
            std::copy(out.begin(), out.end(), coded + (size_t)b * N);
        }
    });
}

// single-codeword reference-shaped call (double LLRs in, K bytes out)
int polar_host_decode_scl_llr(void* h, const double* llr, int L, uint8_t* info_out) {
    PolarCode* p = static_cast<PolarCode*>(h);
    return guarded([&] {
        std::vector<uint8_t> out = p->decode_scl_llr(std::vector<double>(llr, llr + p->block_length()), (uint16_t)L);
        std::copy(out.begin(), out.end(), info_out);
    });
}

int polar_host_decode_batch_packed(void* h, const float* llr, int B, int L, uint32_t* info_packed) {
    PolarCode* p = static_cast<PolarCode*>(h);
    return guarded([&] { p->decode_scl_llr_batch_packed(llr, B, (uint16_t)L, info_packed); });
}

int polar_host_decode_device(void* h, const float* llr_dev, int B, int L, uint32_t* info_packed_dev, void* stream,
                             float* margin_dev) {
    PolarCode* p = static_cast<PolarCode*>(h);
    return guarded([&] { p->decode_scl_llr_device(llr_dev, B, (uint16_t)L, info_packed_dev, stream, margin_dev); });
}

int polar_host_decode_batch_packed_f64(void* h, const double* llr, int B, int L, uint32_t* info_packed) {
    PolarCode* p = static_cast<PolarCode*>(h);
    return guarded([&] { p->decode_scl_llr_batch_packed_f64(llr, B, (uint16_t)L, info_packed); });
}

int polar_host_decode_p1_batch_packed(void* h, const double* p1, const double* p0, int B, int L, uint32_t* info_packed) {
    PolarCode* p = static_cast<PolarCode*>(h);
    return guarded([&] { p->decode_scl_p1_batch_packed(p1, p0, B, (uint16_t)L, info_packed); });
}

// single-codeword reference-shaped call (PolarCode.h:31)
int polar_host_decode_scl_p1(void* h, const double* p1, const double* p0, int L, uint8_t* info_out) {
    PolarCode* p = static_cast<PolarCode*>(h);
    return guarded([&] {
        const int N = p->block_length();
        std::vector<uint8_t> out = p->decode_scl_p1(std::vector<double>(p1, p1 + N), std::vector<double>(p0, p0 + N), (uint16_t)L);
        std::copy(out.begin(), out.end(), info_out);
    });
}

void polar_host_set_exact(void* h, int exact) {
    static_cast<PolarCode*>(h)->arithmetic_mode = exact ? POLAR_B200_MODE_F64 : POLAR_B200_MODE_STRICT;
}
void polar_host_set_mode(void* h, int mode) { static_cast<PolarCode*>(h)->arithmetic_mode = mode; }
int polar_host_get_mode(void* h) { return static_cast<PolarCode*>(h)->arithmetic_mode; }

int polar_host_decode_batch_packed_double(void* h, const double* llr, int B, int L, uint32_t* info_packed) {
    PolarCode* p = static_cast<PolarCode*>(h);
    return guarded([&] { p->decode_scl_llr_batch_packed_double(llr, B, (uint16_t)L, info_packed); });
}

void* polar_host_ctx(void* h, int min_batch) {
    PolarCode* p = static_cast<PolarCode*>(h);
    void* c = nullptr;
    if (guarded([&] { c = p->device_ctx(min_batch); })) return nullptr;
    return c;
}

// counts: [n_list][n_ebno][2] int64 out (reduced over the devices); devices may be NULL (= all visible)
int polar_host_bler_sweep(void* h, const double* ebno, int n_ebno, const uint8_t* lists, int n_list, long long total,
                          unsigned long long seed, const int* devices, int n_devices, long long* counts) {
    PolarCode* p = static_cast<PolarCode*>(h);
    return guarded([&] {
        std::vector<long long> c;
        p->bler_sweep(std::vector<double>(ebno, ebno + n_ebno), std::vector<uint8_t>(lists, lists + n_list), total, seed,
                      devices ? std::vector<int>(devices, devices + n_devices) : std::vector<int>(), &c);
        std::copy(c.begin(), c.end(), counts);
    });
}

int polar_host_get_bler_quick(void* h, const double* ebno, int n_ebno, const uint8_t* lists, int n_list,
                              int max_err, int max_runs, int verbose, double* bler_out) {
    PolarCode* p = static_cast<PolarCode*>(h);
    return guarded([&] {
        p->bler_max_err = max_err; p->bler_max_runs = max_runs; p->bler_verbose = verbose != 0;
        std::vector<std::vector<double>> r = p->get_bler_quick(std::vector<double>(ebno, ebno + n_ebno),
                                                               std::vector<uint8_t>(lists, lists + n_list));
        for (int i = 0; i < n_list; ++i)
            for (int j = 0; j < n_ebno; ++j) bler_out[i * n_ebno + j] = r[i][j];
    });
}

}  // extern "C"
