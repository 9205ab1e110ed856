/*
 * polar_b200 -- C ABI of the B200 (sm_100a) LLR-domain SC / SCL polar decoder.
 *
 * This is the drop-in boundary for the reference's hot path. The reference
 * (tavildar/Polar, PolarC/) has no FFI layer of its own: its boundary is the
 * C++ class surface PolarC/PolarCode.h:19-34. The host class shipped with this
 * library (polar_b200/csrc/PolarCode.h, same public signatures) is implemented on
 * top of the entry points below, and every entry point names the reference
 * interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this ABI;
 *   - return value 0 = OK, negative = POLAR_B200_E_* (invalid use), positive =
 *     a cudaError_t passed through; polar_b200_strerror() explains both;
 *   - no exceptions cross the ABI;
 *   - a ctx is bound to one device; calls on one ctx must be serialised by the
 *     caller (like the reference object, PolarCode.h:56-68, it is not
 *     re-entrant); different ctxs are independent;
 *   - "device" pointers are CUDA device pointers on the ctx's device; the
 *     *_host entry points take ordinary (ideally pinned) host memory and do the
 *     transfers themselves;
 *   - LLR sign convention is the reference's: LLR = ln P(y|0)/P(y|1), positive
 *     means bit 0 (PolarCode.cpp:752 with BPSK 0 -> -1 at :715);
 *   - decoded info bits are packed little-endian: bit j of codeword b is
 *     (info_packed[b * polar_b200_info_words(K) + j/32] >> (j%32)) & 1, where j is
 *     the reference's info index (decoded[j] of PolarCode.cpp:171-174).
 */
#ifndef POLAR_B200_H
#define POLAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct polar_b200_ctx polar_b200_ctx;

enum {
    POLAR_B200_OK = 0,
    POLAR_B200_E_ARG = -1,        /* null pointer / out-of-range argument            */
    POLAR_B200_E_UNSUPPORTED = -2,/* valid for the reference, not for this build     */
    POLAR_B200_E_NOGPU = -3,      /* no usable CUDA device: there is NO CPU fallback */
    POLAR_B200_E_BATCH = -4,      /* (ABI 1: B exceeded max_batch; staging now grows on demand, no longer returned) */
    POLAR_B200_E_LIST = -5,       /* list size < 1, > max_list or > 127              */
    POLAR_B200_E_NONCCL = -6,     /* libnccl.so.2 not loadable (polar_b200_comm_* only) */
    POLAR_B200_E_NCCL = -7        /* an NCCL call failed                             */
};

/* ABI version of this header; polar_b200_abi_version() must return the same value. */
#define POLAR_B200_ABI_VERSION 2

/*
 * Arithmetic modes of the LLR-domain decoder.
 *   FP32    the throughput kernels alone: float LLRs, the reference's rules of order, hardware exp2/log2. Decisions
 *           taken on a margin smaller than float rounding can differ from the double reference (measured: about 2 in
 *           10^6 codewords at 1 dB).
 *   STRICT  FP32, plus: every kernel records the smallest margin (gap between the worst kept and the best dropped
 *           fork metric; |LLR| for list 1; runner-up gap of the final pick) each codeword was decided with, and every
 *           codeword whose margin is below tau (polar_b200_set_strict_tau) is decoded again in double. Block lengths /
 *           lists without a margin-reporting kernel run entirely in double. This is what the drop-in class uses.
 *           List size 1 at N = 2^8..2^12 runs plain SC on the pruned tree (sc_ssc.cuh: subtrees without frozen leaves
 *           are decided by the signs of their root LLRs, which is exact as long as no deciding LLR is within tau of
 *           zero -- the margin this mode checks anyway; FP32 mode uses the same kernel and decodes the codewords with
 *           an exactly-zero deciding LLR leaf by leaf in fp32).
 *   F64     everything in double with the reference's literal formulas (PolarCode.cpp:438-446, 483, 505-506).
 *   MINSUM  opt-in, NOT the reference's arithmetic: min-sum check nodes everywhere (the
 *           reference's own fallback branch, PolarCode.cpp:442-446) and the hardware-friendly metric update (PM += |LLR|
 *           when a decision contradicts the LLR's sign). No transcendental functions; BLER is measurably worse (bench.py
 *           --mode minsum reports the delta). N = 2048 and N = 512, lists 1..32 (POLAR_B200_E_UNSUPPORTED elsewhere).
 */
enum { POLAR_B200_MODE_FP32 = 0, POLAR_B200_MODE_STRICT = 1, POLAR_B200_MODE_F64 = 2, POLAR_B200_MODE_MINSUM = 3 };
#define POLAR_B200_DEFAULT_STRICT_TAU 1.0e-5f
int polar_b200_abi_version(void);

/* Human-readable text for any value returned by this library. */
const char* polar_b200_strerror(int code);

/* Words of packed output per codeword: ceil(K/32). */
int polar_b200_info_words(int K);

/*
 * Page-locked host memory for the *_host entry points (plain malloc'ed memory works too, the copies are then staged by
 * the driver). write_combined != 0: not cached by the CPU -- fast to fill sequentially and to read from the device over
 * PCIe (no snooping), very slow to read back on the CPU: for LLR input buffers only. NULL on failure.
 */
void* polar_b200_host_alloc(size_t bytes, int write_combined);
int polar_b200_host_free(void* p);

/*
 * The table of compiled first-pass kernel variants (test hook): entry i serves block length 2^nlog and list sizes in
 * (2^(lanes_log2 - 1), 2^lanes_log2]; POLAR_B200_FAST_VARIANT=<i> in the environment forces it for matching calls and
 * POLAR_B200_INFO_KERNEL_KIND then reports 1 + i.
 */
int polar_b200_fast_variant_count(void);
int polar_b200_fast_variant_desc(int index, int* nlog, int* lanes_log2, int* warps_per_block);

/*
 * Test hooks of the list-size-1 first pass of the STRICT and FP32 modes (sc_ssc.cuh: plain SC on the pruned decoding tree, N = 2^8..2^12;
 * POLAR_B200_SSC=0 in the environment switches it off, POLAR_B200_INFO_KERNEL_KIND reports 500). No GPU needed.
 * _schedule: the per-code operation list built from frozen_mask ([2^n] bytes); returns its length (0: this code is not
 * served by that kernel), writes it when cap is large enough. _positions: where output bit j is gathered from, [K].
 */
int polar_b200_ssc_schedule(int n, const uint8_t* frozen_mask, uint32_t* ops_out, int cap);
int polar_b200_ssc_positions(int n, const uint16_t* info_order, int K, uint16_t* pos_out);

/* CUDA devices visible to this process (0 when there is none). */
int polar_b200_device_count(void);

/*
 * Create a decoder for one polar code on one device.
 *
 * Replaces the per-object state the reference builds in its constructor
 * (PolarCode.h:19-28 -> PolarCode.cpp:17-58, 647-656). The construction itself
 * (Bhattacharyya recursion, std::sort, rand() parity matrix) stays on the host
 * C++ side and is passed in as data, so this library is independent of
 * rand()/std::sort quirks:
 *   n            log2 of the block length N, 1 <= n <= 15 (the reference's block length is a uint16_t,
 *                PolarCode.h:40); n <= 13 runs on the warp kernels, n = 14, 15 on the block-per-codeword kernel
 *   K            info bits; crc_bits parity ("CRC") bits, K + crc_bits <= N
 *   frozen_mask  [N] bytes, 1 = frozen, index = decoding position phi (_frozen_bits)
 *   info_order   [K + crc_bits] = prefix of _channel_order_descending: position of info
 *                bit j (j < K) and of parity bit r (at K + r)
 *   crc_matrix   [crc_bits][K] row-major 0/1 (_crc_matrix); may be NULL when crc_bits == 0
 *   max_list     largest list size that will be requested (1..127, the reference's own limit:
 *                its loop counters are uint8_t, PolarCode.cpp:497-605)
 *   max_batch    expected largest B of the *_host entry points (initial size of the staging buffers, which grow on demand;
 *                the device-pointer entry points accept any B)
 */
int polar_b200_create(polar_b200_ctx** out, int device, int n, int K, int crc_bits,
                      const uint8_t* frozen_mask, const uint16_t* info_order,
                      const uint8_t* crc_matrix, int max_list, int max_batch);

int polar_b200_destroy(polar_b200_ctx* ctx);

/*
 * Decode B codewords. Replaces B calls of
 *   std::vector<uint8_t> PolarCode::decode_scl_llr(std::vector<double> llr, uint16_t list_size)
 * (PolarCode.h:32, PolarCode.cpp:130-190, 422-644).
 *   llr          device, [B][N] row-major fp32, the reference's channel order
 *   L            list size, 1 <= L <= max_list (L = 1 is plain SC); need not be a power of two.
 *                Lists up to 32 run one codeword (or several) per warp; lists 33..127 run one
 *                codeword per 64- or 128-thread block (same rules, lower throughput)
 *   info_packed  device, [B][polar_b200_info_words(K)]
 *   cuda_stream  a cudaStream_t (NULL = default stream); the call is asynchronous on it
 */
int polar_b200_decode_scl_llr(polar_b200_ctx* ctx, const float* llr, int B, int L,
                              uint32_t* info_packed, void* cuda_stream);

/*
 * Same with an explicit arithmetic mode (polar_b200_decode_scl_llr is mode FP32).
 *   margin   device, [B] floats out, or NULL: the smallest decision margin of every codeword (fast kernels only:
 *            POLAR_B200_E_UNSUPPORTED where (n, L) has none). +inf = no close decision at all.
 * llr should be 16-byte aligned; other addresses are decoded by the generic kernel (slower).
 * One call per ctx is in flight at a time: a call on another stream first waits (on the device) for the previous
 * call of this ctx, because all calls share the ctx's scratch buffers.
 */
int polar_b200_decode_scl_llr_ex(polar_b200_ctx* ctx, const float* llr, int B, int L, uint32_t* info_packed,
                                 int mode, float* margin, void* cuda_stream);

/* Margin threshold of STRICT mode (default POLAR_B200_DEFAULT_STRICT_TAU; env POLAR_B200_STRICT_TAU overrides). */
int polar_b200_set_strict_tau(polar_b200_ctx* ctx, float tau);

/* Grow the staging buffers of the *_host entry points to max_batch codewords (they also grow on demand). */
int polar_b200_reserve(polar_b200_ctx* ctx, int max_batch);

/*
 * Same, host memory in and out: H2D copy of llr, decode, D2H copy of the packed bits; the
 * result is valid on return. The batch is cut into a few chunks whose transfers and decodes
 * overlap on internal streams (pinned host memory makes the copies truly asynchronous);
 * `cuda_stream` is only synchronised on entry. This is the end-to-end path the host class and
 * bench.py's `e2e` leg use. Staging grows on demand (polar_b200_reserve).
 */
int polar_b200_decode_scl_llr_host(polar_b200_ctx* ctx, const float* llr_host, int B, int L,
                                   uint32_t* info_packed_host, void* cuda_stream);
int polar_b200_decode_scl_llr_host_ex(polar_b200_ctx* ctx, const float* llr_host, int B, int L,
                                      uint32_t* info_packed_host, int mode, void* cuda_stream);
/*
 * STRICT mode on double LLRs (what PolarCode::get_bler_quick feeds the decoder, PolarCode.cpp:752-756): the fp32
 * kernels decode the LLRs rounded to float, the flagged codewords are decoded again in double on the caller's
 * doubles. Host memory, synchronous.
 */
int polar_b200_decode_scl_llr_f64_strict_host(polar_b200_ctx* ctx, const double* llr_host, int B, int L,
                                              uint32_t* info_packed_host, void* cuda_stream);

/*
 * Reference-precision mode: the same decoder evaluated in double with the reference's literal
 * formulas (exp/log box-plus below |LLR| = 40, log(1+exp(x)) metrics; PolarCode.cpp:438-446, 483,
 * 505-506) on double LLRs, i.e. the arithmetic of PolarCode::decode_scl_llr itself. For callers that
 * need the reference's decisions on near-tied codewords (arithmetic contract); an
 * order of magnitude slower than the fp32 path (generic kernel, software exp/log).
 *   llr: [B][N] double, device (first form) or host (second form, synchronous; staging grows on demand).
 */
int polar_b200_decode_scl_llr_f64(polar_b200_ctx* ctx, const double* llr, int B, int L,
                                  uint32_t* info_packed, void* cuda_stream);
int polar_b200_decode_scl_llr_f64_host(polar_b200_ctx* ctx, const double* llr_host, int B, int L,
                                       uint32_t* info_packed_host, void* cuda_stream);

/*
 * Probability-domain list decoder. Replaces B calls of
 *   std::vector<uint8_t> PolarCode::decode_scl_p1(std::vector<double> p1, std::vector<double> p0, uint16_t list_size)
 * (PolarCode.h:31, PolarCode.cpp:110-128 -> decode_scl :150-190 with recursivelyCalcP :375-420): tree entries are
 * likelihood pairs, every refreshed layer is divided by its maximum over all live paths, forks are ranked by the
 * likelihood pair (:510-514), the final pick takes the largest likelihood of the last decided bit (:631-637).
 * Evaluated in double with individually rounded products (no FMA contraction), i.e. the reference's own
 * arithmetic. Note the reference's argument order: p1 first.
 *   p1, p0: [B][N] double, P(y_i | 1) and P(y_i | 0) in the reference's channel order; device pointers (first
 *   form, asynchronous on the stream) or host pointers (second form, synchronous; staging grows on demand).
 * The reference's BLER harness never calls this decoder (PolarCode.cpp:755 is commented out); it is provided so
 * that the class surface is complete, on the block-per-codeword kernel (not the throughput path).
 */
int polar_b200_decode_scl_p1(polar_b200_ctx* ctx, const double* p1, const double* p0, int B, int L,
                             uint32_t* info_packed, void* cuda_stream);
int polar_b200_decode_scl_p1_host(polar_b200_ctx* ctx, const double* p1_host, const double* p0_host, int B, int L,
                                  uint32_t* info_packed_host, void* cuda_stream);

/*
 * Block-error flags: block_err[b] = 1 iff any of the K info bits differ. Replaces the
 * comparison loop of the BLER harness (PolarCode.cpp:758-764). All device pointers;
 * n_err (device, may be NULL) is incremented by the number of block errors.
 */
int polar_b200_count_errors(polar_b200_ctx* ctx, const uint32_t* info_packed,
                            const uint32_t* truth_packed, int B, uint8_t* block_err,
                            unsigned long long* n_err, void* cuda_stream);

/*
 * Synthetic front end on the device: B codewords with global indices first_index .. first_index+B-1.
 * Replaces the per-run generation of the reference's BLER loop -- info bits and noise
 * (PolarCode.cpp:703-710), encoder (:60-91, :712), BPSK/AWGN channel and LLR (:715, :744-753, N0 = 1) --
 * with a counter-based generator (Philox4x32-10, key = seed, counter = (codeword index, draw)), so a
 * codeword depends only on (seed, index), not on batching or on the number of GPUs. The reference's own
 * RNG sequence (rand(), std::default_random_engine) is NOT reproduced here; the host class's
 * get_bler_quick keeps doing that on the host.
 *   ebno_db   host, [n_ebno] (n_ebno <= 64); codeword i uses ebno_db[i % n_ebno]
 *   llr       device, [B][N] fp32 out;  truth_packed  device, [B][ceil(K/32)] out (the info bits)
 * The channel is evaluated in double like the reference (r = a s + sqrt(1/2) z, llr = -4 r a; Box-Muller on 53-bit
 * uniforms, tails to |z| = 8.5) and rounded to float once. K <= 2048, N <= 8192, at most 32 parity bits
 * (POLAR_B200_E_UNSUPPORTED beyond).
 */
int polar_b200_synthesize(polar_b200_ctx* ctx, unsigned long long seed, long long first_index, int B,
                          const double* ebno_db, int n_ebno, float* llr, uint32_t* truth_packed,
                          void* cuda_stream);

/*
 * Monte-Carlo BLER sweep on this ctx's device: the codewords with global indices first_index .. first_index+count-1
 * (codeword g at Eb/N0 point g % n_ebno, as in polar_b200_synthesize) are synthesised, decoded with every list size and
 * compared with the transmitted info bits on the device, chunk by chunk; only the counters come back. Replaces the body
 * of the reference's BLER loop (PolarCode.cpp:696-775: generation :703-716, :744-753, decode :756, comparison :758-769)
 * without its sequential early-stop shortcuts (:725-742). Per chunk and list size this is one decode launch with the
 * block-error count fused into its tail (plus, in STRICT mode, the second pass over the flagged codewords).
 *   lists   host, [n_list] list sizes;  mode  POLAR_B200_MODE_*
 *   counts  host, [n_list][n_ebno][2] out: (num_err, num_run) per cell, to be summed over shards
 * Synchronous. Shards of one sweep (other devices / ranks) use the same seed and disjoint index ranges.
 */
int polar_b200_bler_sweep(polar_b200_ctx* ctx, unsigned long long seed, long long first_index, long long count,
                          const double* ebno_db, int n_ebno, const int* lists, int n_list, int mode,
                          long long* counts, void* cuda_stream);

/*
 * The one collective of a sharded sweep: sum the int64 counters over GPUs with ncclAllReduce.
 * One communicator per GPU: either one process per GPU (rank 0 calls polar_b200_comm_unique_id, the 128 bytes travel
 * by any means, every rank calls polar_b200_comm_init_rank) or one process driving several GPUs
 * (polar_b200_comm_init_all fills comms[ndev]; polar_b200_comm_allreduce_i64_group reduces all of them in one NCCL group).
 * values: host vectors, summed in place. libnccl.so.2 is loaded at run time (POLAR_B200_E_NONCCL if absent).
 */
typedef struct polar_b200_comm polar_b200_comm;
int polar_b200_comm_unique_id(unsigned char* id128);
int polar_b200_comm_init_rank(polar_b200_comm** out, int device, int nranks, int rank, const unsigned char* id128);
int polar_b200_comm_init_all(polar_b200_comm** comms, int ndev, const int* devices);
int polar_b200_comm_allreduce_i64(polar_b200_comm* comm, long long* values, int n);
int polar_b200_comm_allreduce_i64_group(polar_b200_comm** comms, int ncomm, long long** values, int n);
int polar_b200_comm_destroy(polar_b200_comm* comm);

/* Introspection (all return -1 for an unknown key). */
enum {
    POLAR_B200_INFO_KERNEL_LAUNCHES = 0, /* kernels this ctx has launched so far            */
    POLAR_B200_INFO_SM_COUNT = 1,
    POLAR_B200_INFO_WARPS_PER_BLOCK = 2, /* of the last decode launch                       */
    POLAR_B200_INFO_BLOCKS = 3,          /* grid size of the last decode launch             */
    POLAR_B200_INFO_SMEM_BYTES = 4,      /* dynamic shared memory of the last decode launch */
    POLAR_B200_INFO_SCRATCH_BYTES = 5,   /* device scratch owned by the ctx                 */
    POLAR_B200_INFO_KERNEL_KIND = 6,     /* last decode: 0 = generic kernel, 1 + i = fast variant i, -1 = f64 mode,
                                            500 = plain SC on the pruned tree, 1000 + i = min-sum build i, -2 = wide-list kernel (lists 33..127),
                                            -3 = wide-list kernel in f64, -4 = probability-domain decoder */
    POLAR_B200_INFO_HOST_CHUNKS = 7,     /* chunks the last *_host call was pipelined in            */
    POLAR_B200_INFO_LAST_FLAGGED = 8     /* codewords the last STRICT call decoded again in double (waits for it) */
};
long long polar_b200_get_info(polar_b200_ctx* ctx, int key);

#ifdef __cplusplus
}
#endif
#endif /* POLAR_B200_H */
