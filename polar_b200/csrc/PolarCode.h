// Host-side drop-in for the reference's `class PolarCode` (PolarC/PolarCode.h:16-90).
//
// Same public surface -- constructor (num_layers, info_length, epsilon, crc_size), encode,
// decode_scl_p1, decode_scl_llr, get_bler_quick -- so that the reference's own driver
// (PolarC/main.cpp:15,32) compiles against this header unchanged. The LLR-domain decoder,
// the only thing the reference's BLER loop calls (PolarCode.cpp:756), runs on the GPU
// through the C ABI of include/polar_b200.h; there is no CPU decoder behind it and every
// decode throws std::runtime_error if the CUDA library cannot be used.
//
// Additions over the reference surface are the batched entry points a throughput user
// needs (decode_scl_llr_batch*, accessors for the construction tables).
#ifndef POLAR_B200_POLARCODE_H
#define POLAR_B200_POLARCODE_H

#include <cstdint>
#include <vector>

struct polar_b200_ctx;

class PolarCode {
public:
    // PolarCode.h:19-28. num_layers = log2(block length) as in the C++ reference (the MATLAB twin
    // takes N). Side effect kept: consumes crc_size * info_length values of rand() for the
    // random parity ("CRC") matrix, PolarCode.cpp:51-56.
    PolarCode(uint8_t num_layers, uint16_t info_length, double epsilon, uint16_t crc_size);
    ~PolarCode();
    PolarCode(const PolarCode&) = delete;
    PolarCode& operator=(const PolarCode&) = delete;

    // PolarCode.h:30 / PolarCode.cpp:60-91. Host code (negligible next to decoding).
    std::vector<uint8_t> encode(std::vector<uint8_t> info_bits);
    // PolarCode.h:31 / PolarCode.cpp:110-128, 375-420. Probability-domain list decoding, one codeword, on the
    // GPU in double (the reference's harness never calls it, PolarCode.cpp:755 is commented out).
    std::vector<uint8_t> decode_scl_p1(std::vector<double> p1, std::vector<double> p0, uint16_t list_size);
    // PolarCode.h:32 / PolarCode.cpp:130-190. One codeword, GPU, latency-bound; kept for source
    // compatibility. Evaluated in double like the reference unless arithmetic_mode is FP32.
    std::vector<uint8_t> decode_scl_llr(std::vector<double> llr, uint16_t list_size);
    // PolarCode.h:34 / PolarCode.cpp:658-785. Same RNG objects in the same call order, same
    // counting rules (early stop, decoded-at-lower-Eb/N0 shortcut) replayed on the host over
    // per-cell success flags; all decodes of a sweep are batched on the GPU.
    std::vector<std::vector<double>> get_bler_quick(std::vector<double> ebno_vec, std::vector<uint8_t> list_size);

    // ---- batched extensions ----
    // llr: host, [B][N] floats. Returns [B][K] bytes (0/1).
    std::vector<uint8_t> decode_scl_llr_batch(const float* llr, int B, uint16_t list_size);
    // Same with packed output: [B][info_words()] little-endian words.
    void decode_scl_llr_batch_packed(const float* llr, int B, uint16_t list_size, uint32_t* info_packed);
    // Reference-precision mode (double LLRs, the reference's literal formulas in double on the GPU).
    void decode_scl_llr_batch_packed_f64(const double* llr, int B, uint16_t list_size, uint32_t* info_packed);
    // Probability domain, batched: p1, p0 host [B][N] doubles; packed output.
    void decode_scl_p1_batch_packed(const double* p1, const double* p0, int B, uint16_t list_size, uint32_t* info_packed);
    // double LLRs in arithmetic_mode (STRICT: float kernels + double re-decode of the flagged codewords on these doubles)
    void decode_scl_llr_batch_packed_double(const double* llr, int B, uint16_t list_size, uint32_t* info_packed);
    // Device pointers, asynchronous on `cuda_stream` (a cudaStream_t). margin_dev: optional [B] floats (decision margins).
    void decode_scl_llr_device(const float* llr_dev, int B, uint16_t list_size, uint32_t* info_packed_dev, void* cuda_stream,
                               float* margin_dev = nullptr);

    // Monte-Carlo BLER over `total` codewords (codeword g at Eb/N0 point g % ebno_vec.size()), generated, decoded and
    // compared on the GPUs: the index range is split over `devices` (default: every visible device), one host thread and
    // one decoder context per device, and the (num_err, num_run) counters are summed with ncclAllReduce. No early stop,
    // unlike get_bler_quick (PolarCode.cpp:725-742 is sequential in run order); result[list][ebno] = num_err / num_run.
    // counts (optional): [list][ebno][2] receives the reduced counters.
    std::vector<std::vector<double>> bler_sweep(const std::vector<double>& ebno_vec, const std::vector<uint8_t>& list_size,
                                                long long total, unsigned long long seed = 1,
                                                std::vector<int> devices = std::vector<int>(),
                                                std::vector<long long>* counts = nullptr);

    int block_length() const { return _block_length; }
    int info_length() const { return _info_length; }
    int crc_size() const { return _crc_size; }
    int num_layers() const { return _n; }
    int info_words() const { return (_info_length + 31) / 32; }
    const std::vector<uint8_t>& frozen_bits() const { return _frozen_bits; }
    const std::vector<uint16_t>& channel_order() const { return _channel_order_descending; }
    const std::vector<std::vector<uint8_t>>& crc_matrix() const { return _crc_matrix; }
    const std::vector<uint16_t>& bit_rev_order() const { return _bit_rev_order; }
    polar_b200_ctx* device_ctx(int min_batch);   // creates / grows the GPU context on demand

    // knobs of get_bler_quick, defaults = the reference's constants (PolarCode.cpp:661-662)
    int bler_max_err = 100;
    int bler_max_runs = 1000;
    bool bler_verbose = true;      // the reference's "Running iteration ..." lines
    int device = 0;                // CUDA device ordinal
    // Arithmetic of every decode (include/polar_b200.h, POLAR_B200_MODE_*): 0 = FP32 kernels alone, 1 = STRICT (default:
    // FP32 kernels, codewords with a decision closer than tau decoded again in double -- reproduces the reference's
    // decisions), 2 = F64 (everything in double, an order of magnitude slower). Environment: POLAR_B200_MODE=fp32|strict|f64
    // (POLAR_B200_EXACT=1 is f64).
    int arithmetic_mode = 1;

private:
    uint8_t _n;
    uint16_t _info_length;
    uint16_t _block_length;
    uint16_t _crc_size;
    double _design_epsilon;

    std::vector<uint8_t> _frozen_bits;
    std::vector<uint16_t> _channel_order_descending;
    std::vector<std::vector<uint8_t>> _crc_matrix;
    std::vector<uint16_t> _bit_rev_order;

    polar_b200_ctx* _ctx = nullptr;
    int _ctx_batch = 0;

    void initialize_frozen_bits();
    void create_bit_rev_order();
};

#endif  // POLAR_B200_POLARCODE_H
