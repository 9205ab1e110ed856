// polar_b200 -- LLR-domain SC / SCL polar decoder for B200 (sm_100a); C ABI in include/polar_b200.h.
//
// This file: the generic warp kernel (any n <= 13, any list <= 32, float or double; also strict mode's third pass),
// the error-count and synthetic-front-end kernels, the device-side BLER sweep, the NCCL communicator wrappers, and the
// C ABI with its dispatch by arithmetic mode: scl_fast.cuh variants (fast_parts.cu) where one exists for (n, list) --
// followed in STRICT mode by scl_exact.cuh on the flagged codewords --, scl_wide.cuh for lists 33..127, n = 14, 15 and
// the probability domain, the generic kernel otherwise.
//
// What is computed is the reference's decode_scl_llr (PolarC/PolarCode.cpp:130-190,
// 422-644): Tal-Vardy successive-cancellation list decoding with LLR path metrics,
// the reference's own check-node rule (exact box-plus below |LLR| = 40, sign-min
// above, PolarCode.cpp:438-446) and its literal softplus metric updates
// (PolarCode.cpp:483,505-506). How it is computed is not the reference's:
//
//   * one WARP owns a group of 32/W codewords (W = list size rounded up to a power of
//     two); lane = one list path of one codeword, so the 2L-way fork / prune is a
//     handful of warp shuffles and ballots and there is no block-level barrier at all;
//   * every layer array is stored "path-interleaved" ([beta][lane], one 128-byte row per
//     tree node position): a lane's own column is written, any column can be read, and
//     both are bank-conflict-free in shared memory and fully coalesced in HBM/L2;
//   * the reference's lazy copy (ref-counted array pools, PolarCode.cpp:195-373) becomes a
//     5-bit column pointer per (path, layer), packed in two 64-bit registers; cloning a
//     path is a register shuffle. All paths recompute a given layer at the same time and
//     always write their own column, so a pointer can never dangle;
//   * arrays are kept in "butterfly" order (pairs (beta, beta+M) instead of the
//     reference's (2beta, 2beta+1)); the bit-reversal this implies is folded into the
//     first layer's channel reads. Partial sums are bit-packed, and combining the two
//     halves of a node is a word concatenation;
//   * decided bits are not stored per path: the last partial-sum update yields the
//     re-encoded codeword of every path, and one packed polar transform turns it back
//     into u-hat for the parity ("CRC") check and the output gather;
//   * the small layers live in shared memory, the big ones (touched rarely) in a
//     per-warp scratch in HBM that stays L2-resident.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include <dlfcn.h>

#include <new>
#include <vector>

#if __has_include(<nccl.h>)
#include <nccl.h>
#else
// minimal declarations of the NCCL C API (the library itself is loaded at run time, see polar_b200_comm_*)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt64 = 4 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
#endif

#include "polar_b200.h"

#include "polar_dev.cuh"
#include "fast_variants.cuh"
#include "sc_ssc.cuh"
#include "scl_wide.cuh"
#include "scl_exact.cuh"

namespace {

// One warp decodes groups of 32/W codewords. Dynamic shared memory per warp:
//   smem_x_rows rows of 32 floats (LLR layers lamS..n-1), smem_s_rows rows of 32 words
//   (partial-sum layers >= max(lamS,1), except layer n which is a register), 32 bytes of scatter
//   scratch.
template <class Real, class In>
__global__ void __launch_bounds__(256) scl_decode_kernel(const DecodeArgsT<Real, In> a) {
    constexpr int ROWB = 32 * (int)sizeof(Real);         // bytes of one [32 lanes] row
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int gwarp = blockIdx.x * warps_per_block + warp_in_block;
    const int total_warps = gridDim.x * warps_per_block;

    const int n = a.n, N = 1 << n, L = a.L, W = a.W, lamS = a.lamS;
    const int G = 32 / W;                      // codewords per warp
    const int slot = lane & (W - 1);           // path index within the codeword
    const int gbase = lane & ~(W - 1);         // first lane of my codeword
    const unsigned gmask_lo = (W == 32) ? FULL_MASK : ((1u << W) - 1u);
    const int NW = (N + 31) >> 5;              // words in an N-bit vector
    const int KW = (a.K + 31) >> 5;

    const size_t per_warp_smem = (size_t)a.smem_x_rows * ROWB + (size_t)a.smem_s_rows * 128 + 32;
    unsigned char* my_smem = smem_raw + per_warp_smem * warp_in_block;
    Real* sx = reinterpret_cast<Real*>(my_smem);
    uint32_t* ss = reinterpret_cast<uint32_t*>(my_smem + (size_t)a.smem_x_rows * ROWB);
    unsigned char* srcof = my_smem + (size_t)a.smem_x_rows * ROWB + (size_t)a.smem_s_rows * 128;
    Real* gx = a.gx + a.gx_stride * gwarp;
    uint32_t* gs = a.gs + a.gs_stride * gwarp;

    // Row address of LLR layer lam (1 <= lam <= n-1), position beta.
    auto xrow = [&](int lam, int beta) -> Real* {
        if (lam >= lamS) return sx + ((size_t)((1 << (n - lamS + 1)) - (1 << (n - lam + 1)) + beta) << 5);
        return gx + ((size_t)(N - (1 << (n - lam + 1)) + beta) << 5);
    };
    // Row address of partial-sum layer lam (0 <= lam <= n-1), word w.
    const int lamSS = lamS < 1 ? 1 : lamS;
    auto srow = [&](int lam, int w) -> uint32_t* {
        if (lam >= lamSS) return ss + ((size_t)(a.s_off[lam] + w) << 5);
        return gs + ((size_t)(a.s_off[lam] + w) << 5);
    };

    // list mode: decode only the codewords named by a.list[0 .. *a.count) (rows of llr / out are still indexed by codeword)
    const int nB = a.list ? *a.count : a.B;
    for (int grp = gwarp; grp * G < nB; grp += total_warps) {
        const int idx = grp * G + (lane / W);
        const bool valid = idx < nB;
        const int row = valid ? idx : nB - 1;
        const In* chan = a.llr + (size_t)(a.list ? a.list[row] : row) * N;

        // Per-path state. Reference bookkeeping reproduced: the free-path stack is filled
        // 0..L-1 (PolarCode.cpp:250-256) and the first path popped is L-1 (:259-263).
        bool active = valid && (slot == L - 1);
        Real pm = 0;
        unsigned long long px = 0, ps = 0;      // column pointers, 5 bits per layer (index lam-1)
        uint32_t s_n = 0;                       // partial sum of layer n (the last even leaf's bit)
        int stk = slot;                         // lane gbase+j holds free-path stack entry j
        int sp = L - 1;                         // stack height (uniform within a codeword)
        Real lam_n = 0;                        // LLR of layer n (decision LLR)
        uint32_t frozen_word = 0;

        for (int phi = 0; phi < N; ++phi) {
            // ---- refresh LLR layers lam_top..n (PolarCode.cpp:422-455) ----
            const int lam_top = (phi == 0) ? 1 : n - (__ffs(phi) - 1);
            for (int lam = lam_top; lam <= n; ++lam) {
                const int M = 1 << (n - lam);
                const bool is_g = (lam == lam_top) && (phi != 0);
                const Real* src = nullptr;
                if (lam > 1) src = xrow(lam - 1, 0) + get_ptr(px, lam - 2);
                const uint32_t* sw = nullptr;
                if (is_g && lam < n) sw = srow(lam, 0) + get_ptr(ps, lam - 1);
                Real* dst = (lam < n) ? xrow(lam, 0) + lane : nullptr;
                if (active) {
                    for (int i = 0; i < M; ++i) {
                        Real x0, x1;
                        int beta = i;
                        if (lam == 1) {
                            // channel layer: reference pairs are (2k, 2k+1); k runs in memory
                            // order, the result lands at the bit-reversed position.
                            x0 = (Real)chan[2 * i]; x1 = (Real)chan[2 * i + 1];
                            beta = (n > 1) ? (int)(__brev((unsigned)i) >> (33 - n)) : 0;
                        } else {
                            x0 = src[(size_t)i << 5];
                            x1 = src[(size_t)(i + M) << 5];
                        }
                        Real y;
                        if (is_g) {
                            uint32_t bit;
                            if (lam == n) bit = s_n & 1u;
                            else bit = (sw[(size_t)(beta >> 5) << 5] >> (beta & 31)) & 1u;
                            y = x1 + (bit ? -x0 : x0);                      // PolarCode.cpp:448-451
                        } else {
                            y = Arith<Real>::f(x0, x1);                              // PolarCode.cpp:438-446
                        }
                        if (lam == n) lam_n = y; else dst[(size_t)beta << 5] = y;
                    }
                }
                if (lam < n) px = set_ptr(px, lam - 1, lane);
            }

            // ---- leaf decision ----
            if ((phi & 31) == 0) frozen_word = a.frozen_words[phi >> 5];
            const bool frozen = (frozen_word >> (phi & 31)) & 1u;
            uint32_t u = 0;
            if (frozen) {
                // PolarCode.cpp:475-487
                if (active) pm += Arith<Real>::softplus(-lam_n);
            } else {
                // PolarCode.cpp:489-607. Metrics are kept positive (m = -probForks).
                const Real m0 = pm + Arith<Real>::softplus(-lam_n);
                const Real m1 = pm + Arith<Real>::softplus(lam_n);
                const unsigned act_all = __ballot_sync(FULL_MASK, active);
                const unsigned act_g = (act_all >> gbase) & gmask_lo;
                const int A = __popc(act_g);
                bool keep0 = active, keep1 = active;
                const bool need_select = active && (2 * A > L);
                bool slow = false;
                if (need_select) {
                    slow = true;
                }
                if (__any_sync(FULL_MASK, need_select)) {
                    // fast exit: every likely fork strictly beats every unlikely fork, list is full
                    const Real lo = rmin<Real>(m0, m1), hi = rmax<Real>(m0, m1);
                    const Real worst_likely = group_max<Real>(active ? lo : -Arith<Real>::inf(), W);
                    const Real best_unlikely = group_min<Real>(active ? hi : Arith<Real>::inf(), W);
                    if (need_select && A == L && best_unlikely > worst_likely) {
                        slow = false;
                        keep0 = (m0 <= m1);      // m0 == m1 cannot happen here (would not be strict)
                        keep1 = !keep0;
                    }
                    if (__any_sync(FULL_MASK, slow)) {
                        // exact rule: keep the rho best of the 2A forks under (metric asc, fork index asc)
                        // == PolarCode.cpp:528-553 (sort, threshold, '>' pass then '==' pass in index order).
                        int r0 = 0, r1 = 0;
                        for (int j = 0; j < W; ++j) {
                            const Real o0 = __shfl_sync(FULL_MASK, m0, gbase + j);
                            const Real o1 = __shfl_sync(FULL_MASK, m1, gbase + j);
                            const bool oa = (act_g >> j) & 1u;
                            if (oa) {
                                // other fork indices 2j, 2j+1; mine 2*slot, 2*slot+1
                                r0 += (o0 < m0) || (o0 == m0 && j < slot);
                                r0 += (o1 < m0) || (o1 == m0 && j < slot);
                                r1 += (o0 < m1) || (o0 == m1 && j <= slot);
                                r1 += (o1 < m1) || (o1 == m1 && j < slot);
                            }
                        }
                        if (slow) {
                            const int rho = L;   // 2A > L here
                            keep0 = r0 < rho;
                            keep1 = r1 < rho;
                        }
                    }
                }
                const bool kill = active && !keep0 && !keep1;
                const bool clone = keep0 && keep1;
                const unsigned kill_all = __ballot_sync(FULL_MASK, kill);
                const unsigned clone_all = __ballot_sync(FULL_MASK, clone);
                if ((kill_all | clone_all) == 0) {
                    if (active) { u = keep1 ? 1u : 0u; pm = keep1 ? m1 : m0; }
                } else {
                    const unsigned Kg = (kill_all >> gbase) & gmask_lo;
                    const unsigned Cg = (clone_all >> gbase) & gmask_lo;
                    const int nk = __popc(Kg), nc = __popc(Cg);
                    // killPath pushes in ascending path order (PolarCode.cpp:555-560, :292)
                    if (slot >= sp && slot < sp + nk) stk = (int)__fns(Kg, 0, slot - sp + 1);
                    const int sp2 = sp + nk;
                    // clonePath pops for ascending l (PolarCode.cpp:562-570, :275-276)
                    const int ci = __popc(Cg & ((1u << slot) - 1u));
                    const int tgt = __shfl_sync(FULL_MASK, stk, gbase + ((sp2 - 1 - ci) & (W - 1)));
                    sp = sp2 - nc;
                    srcof[lane] = (unsigned char)lane;
                    __syncwarp();
                    if (clone) srcof[gbase + tgt] = (unsigned char)lane;
                    __syncwarp();
                    const int src_lane = srcof[lane];
                    __syncwarp();
                    const bool is_new = (src_lane != lane);
                    const Real src_m1 = __shfl_sync(FULL_MASK, m1, src_lane);
                    const unsigned long long src_px = __shfl_sync(FULL_MASK, px, src_lane);
                    const unsigned long long src_ps = __shfl_sync(FULL_MASK, ps, src_lane);
                    const uint32_t src_sn = __shfl_sync(FULL_MASK, s_n, src_lane);
                    if (is_new) {
                        active = true; pm = src_m1; u = 1u; px = src_px; ps = src_ps; s_n = src_sn;
                    } else if (kill) {
                        active = false; pm = 0;
                    } else if (active) {
                        u = keep0 ? 0u : 1u;
                        pm = keep0 ? m0 : m1;
                    }
                }
            }

            // ---- partial sums (PolarCode.cpp:457-473), bit-packed, butterfly order ----
            if ((phi & 1) == 0) {
                s_n = u;
            } else {
                const int t = __ffs(~phi) - 1;           // trailing ones of phi, 1..n
                const int lam_end = n - t;               // layer whose S array receives the result
                uint32_t P = u;
                int lam = n;
                while (lam > lam_end && (n - lam) < 5) {
                    const int M = 1 << (n - lam);
                    uint32_t Sw;
                    if (lam == n) Sw = s_n;
                    else Sw = srow(lam, 0)[get_ptr(ps, lam - 1)];
                    P = ((Sw ^ P) & ((1u << M) - 1u)) | (P << M);
                    --lam;
                }
                if (lam == lam_end) {
                    srow(lam, 0)[lane] = P;
                } else {
                    const int Wd = 1 << (t - 5);          // words of the destination vector
                    uint32_t* D = srow(lam_end, 0) + lane;
                    D[(size_t)(Wd - 1) << 5] = P;
                    for (; lam > lam_end; --lam) {
                        const int mw = 1 << (n - lam - 5);
                        const int base = Wd - mw;
                        const uint32_t* S = srow(lam, 0) + get_ptr(ps, lam - 1);
                        for (int w = 0; w < mw; ++w)
                            D[(size_t)(base - mw + w) << 5] = S[(size_t)w << 5] ^ D[(size_t)(base + w) << 5];
                    }
                }
                if (lam_end >= 1) ps = set_ptr(ps, lam_end - 1, lane);
            }
            __syncwarp();
        }

        // ---- u-hat of every path: packed polar transform of the re-encoded codeword (layer 0) ----
        uint32_t* D = srow(0, 0) + lane;
        for (int sw = NW >> 1; sw >= 1; sw >>= 1)
            for (int i = 0; i < NW; ++i)
                if ((i & sw) == 0) D[(size_t)i << 5] ^= D[(size_t)(i + sw) << 5];
        bool pass = true;
        {
            // in-word stages + parity ("CRC") rows, PolarCode.cpp:93-108
            for (int i = 0; i < NW; ++i) {
                uint32_t w = D[(size_t)i << 5];
                if (N > 16) w ^= (w >> 16) & 0x0000FFFFu;
                if (N > 8) w ^= (w >> 8) & 0x00FF00FFu;
                if (N > 4) w ^= (w >> 4) & 0x0F0F0F0Fu;
                if (N > 2) w ^= (w >> 2) & 0x33333333u;
                w ^= (w >> 1) & 0x55555555u;
                D[(size_t)i << 5] = w;
            }
            for (int r = 0; r < a.crc; ++r) {
                uint32_t acc = 0;
                for (int i = 0; i < NW; ++i) acc ^= D[(size_t)i << 5] & a.crc_masks[(size_t)r * NW + i];
                if (__popc(acc) & 1) pass = false;
            }
        }
        // ---- final pick, PolarCode.cpp:609-644 ----
        const unsigned act_all = __ballot_sync(FULL_MASK, active);
        const unsigned pass_all = __ballot_sync(FULL_MASK, active && pass);
        const unsigned pass_g = (pass_all >> gbase) & gmask_lo;
        const bool use_parity = (a.crc != 0) && (pass_g != 0);
        const bool eligible = active && (use_parity ? pass : true) && (pm < Arith<Real>::inf());
        const Real best = group_min<Real>(eligible ? pm : Arith<Real>::inf(), W);
        const unsigned cand_all = __ballot_sync(FULL_MASK, eligible && pm == best);
        const unsigned cand_g = (cand_all >> gbase) & gmask_lo;
        const int win_slot = cand_g ? (__ffs(cand_g) - 1) : 0;
        const bool win_active = (act_all >> (gbase + win_slot)) & 1u;
        __syncwarp();

        // ---- output gather: decoded[j] = u-hat[order[j]], PolarCode.cpp:171-174 ----
        for (int g = 0; g < G; ++g) {
            const int wl = __shfl_sync(FULL_MASK, gbase + win_slot, g * W);
            const bool wa = __shfl_sync(FULL_MASK, (int)win_active, g * W);
            if (grp * G + g >= nB) break;
            const int cwg = a.list ? a.list[grp * G + g] : grp * G + g;
            const uint32_t* U = srow(0, 0) + wl;
            for (int t = lane; t < KW; t += 32) {
                uint32_t word = 0;
                if (wa) {
                    const int jmax = min(32, a.K - 32 * t);
                    for (int i = 0; i < jmax; ++i) {
                        const int pos = a.info_order[32 * t + i];
                        word |= ((U[(size_t)(pos >> 5) << 5] >> (pos & 31)) & 1u) << i;
                    }
                }
                a.out[(size_t)cwg * KW + t] = word;
            }
        }
        __syncwarp();
    }
}

__global__ void count_errors_kernel(const uint32_t* dec, const uint32_t* truth, int B, int KW,
                                    uint8_t* block_err, unsigned long long* n_err) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    bool err = false;
    if (b < B) {
        for (int w = 0; w < KW; ++w) err |= dec[(size_t)b * KW + w] != truth[(size_t)b * KW + w];
        if (block_err) block_err[b] = err ? 1 : 0;
    }
    const unsigned m = __ballot_sync(FULL_MASK, err);
    if (n_err && (threadIdx.x & 31) == 0 && m) atomicAdd(n_err, (unsigned long long)__popc(m));
}

// block errors per Eb/N0 point (codeword with global index g belongs to point g % n_ebno); `skip` lists codewords that
// are counted elsewhere (strict mode's second pass counts the ones it decodes again)
__global__ void count_errors_bucket_kernel(const uint32_t* dec, const uint32_t* truth, int B, int KW, long long first_index,
                                           int n_ebno, unsigned long long* err) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    bool e = false;
    for (int w = 0; w < KW; ++w) e |= dec[(size_t)b * KW + w] != truth[(size_t)b * KW + w];
    if (e) atomicAdd(err + (int)((unsigned long long)(first_index + b) % (unsigned)n_ebno), 1ull);
}

// ---- synthetic front end: info bits -> reference encoder -> BPSK/AWGN -> fp32 LLRs, one warp per codeword ----
// Replaces the per-run generation of the reference's BLER loop (PolarCode.cpp:703-716 info bits + noise,
// :60-91 encoder, :744-753 channel and LLR) with a counter-based generator: Philox4x32-10 keyed by the seed,
// counter = (global codeword index, draw index), so a codeword depends only on (seed, index).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct SynthArgs {
    float* llr;                     // [B][N]
    uint32_t* truth;                // [B][KW]
    const uint16_t* inv_order;      // [N]: info index j (< K), K + r for parity bit r, 0xFFFF for frozen positions
    const uint32_t* crc_rows;       // [crc][KW]: parity matrix rows packed over the info index
    const double* amp;              // [n_ebno]: a = 10^(EbN0/20) sqrt(K/N)   (PolarCode.cpp:744-745)
    unsigned long long seed;
    long long first_index;
    int B, n, K, crc, n_ebno;
};

__global__ void __launch_bounds__(128) synth_kernel(const SynthArgs a) {
    __shared__ uint32_t sm_info[4][64];      // K <= 2048 info bits per warp
    __shared__ uint32_t sm_u[4][256];        // N <= 8192 bits per warp
    __shared__ uint32_t sm_crc[4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int N = 1 << a.n, NW = (N + 31) >> 5, KW = (a.K + 31) >> 5;
    uint32_t* info = sm_info[wib];
    uint32_t* u = sm_u[wib];
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    for (int b = blockIdx.x * 4 + wib; b < a.B; b += gridDim.x * 4) {
        const unsigned long long gidx = (unsigned long long)(a.first_index + b);
        const uint32_t g0 = (uint32_t)gidx, g1 = (uint32_t)(gidx >> 32);
        // info bits: draw index space 0 (counter word 3 = 0)
        for (int w4 = lane; w4 * 4 < KW; w4 += 32) {
            uint32_t r[4];
            philox4x32_10(g0, g1, (uint32_t)w4, 0u, k0, k1, r);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int w = w4 * 4 + q;
                if (w < KW) {
                    uint32_t x = r[q];
                    if (32 * w + 32 > a.K) x &= (1u << (a.K - 32 * w)) - 1u;      // last partial word
                    info[w] = x;
                    a.truth[(size_t)b * KW + w] = x;
                }
            }
        }
        __syncwarp();
        // parity ("CRC") bits, PolarCode.cpp:68-74
        if (lane == 0) sm_crc[wib] = 0;
        __syncwarp();
        for (int r = 0; r < a.crc; ++r) {
            uint32_t acc = 0;
            for (int w = lane; w < KW; w += 32) acc ^= info[w] & a.crc_rows[(size_t)r * KW + w];
            acc = __reduce_xor_sync(FULL_MASK, acc);
            if (lane == 0 && (__popc(acc) & 1)) sm_crc[wib] |= 1u << r;
        }
        __syncwarp();
        const uint32_t crcbits = sm_crc[wib];
        // u vector in decoding order, PolarCode.cpp:65-74
        for (int w = lane; w < NW; w += 32) {
            uint32_t word = 0;
            const int nb = (N - 32 * w) < 32 ? (N - 32 * w) : 32;
            for (int i = 0; i < nb; ++i) {
                const unsigned j = a.inv_order[32 * w + i];
                uint32_t bit = 0;
                if (j < (unsigned)a.K) bit = (info[j >> 5] >> (j & 31)) & 1u;
                else if (j != 0xFFFFu) bit = (crcbits >> (j - a.K)) & 1u;
                word |= bit << i;
            }
            u[w] = word;
        }
        __syncwarp();
        // x = u F^(x)n: in-word stages, then word-level stages (PolarCode.cpp:76-83)
        for (int w = lane; w < NW; w += 32) {
            uint32_t x = u[w];
            if (N > 1) x ^= (x >> 1) & 0x55555555u;
            if (N > 2) x ^= (x >> 2) & 0x33333333u;
            if (N > 4) x ^= (x >> 4) & 0x0F0F0F0Fu;
            if (N > 8) x ^= (x >> 8) & 0x00FF00FFu;
            if (N > 16) x ^= (x >> 16) & 0x0000FFFFu;
            u[w] = x;
        }
        __syncwarp();
        for (int sw = 1; sw < NW; sw <<= 1) {
            for (int w = lane; w < NW; w += 32)
                if ((w & sw) == 0) u[w] ^= u[w + sw];
            __syncwarp();
        }
        // coded[i] = x[bitrev(i)] (PolarCode.cpp:85-87), BPSK 0 -> -1, r = a s + sqrt(1/2) z, llr = -4 r a (:747,:752),
        // all in double like the reference and rounded to float once at the end. One Philox
        // call = two 53-bit uniforms = one Box-Muller pair; the tails reach |z| = 8.5 (u1 >= 2^-54).
        const double amp = a.amp[(int)(gidx % (unsigned long long)a.n_ebno)];
        float* out = a.llr + (size_t)b * N;
        for (int i2 = lane; i2 * 2 < N; i2 += 32) {
            uint32_t r[4];
            philox4x32_10(g0, g1, (uint32_t)i2, 1u, k0, k1, r);       // draw index space 1: noise
            const double u1 = ((double)(((unsigned long long)r[0] << 21) | (r[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
            const double u2 = ((double)(((unsigned long long)r[2] << 21) | (r[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
            const double rad = sqrt(-2.0 * log(u1));
            double sn, cs;
            sincospi(2.0 * u2, &sn, &cs);
            const double z[2] = {rad * cs, rad * sn};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int i = i2 * 2 + q;
                if (i < N) {
                    const unsigned p = a.n ? (__brev((unsigned)i) >> (32 - a.n)) : 0u;
                    const double c = (double)((u[p >> 5] >> (p & 31)) & 1u);
                    const double rx = amp * (2.0 * c - 1.0) + 0.70710678118654752440 * z[q];
                    out[i] = (float)(-4.0 * rx * amp);
                }
            }
        }
        __syncwarp();
    }
}

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    if (!s || !*s) return dflt;
    return atoi(s);
}

}  // namespace

struct polar_b200_ctx {
    int device = 0;
    int n = 0, N = 0, K = 0, crc = 0, max_list = 0, max_batch = 0;
    int sm_count = 0;
    int first_info = 0;                    // first unfrozen decoding position
    int KW = 0, NW = 0;
    uint32_t* d_frozen = nullptr;
    uint16_t* d_order = nullptr;
    uint32_t* d_crc_masks = nullptr;
    uint16_t* d_inv_order = nullptr;       // synth: decoding position -> info / parity index
    uint32_t* d_crc_rows = nullptr;        // synth: parity matrix rows packed over the info index
    double* d_amp = nullptr;               // synth: per-Eb/N0 amplitudes (up to 64)
    void* d_gx = nullptr;                  // generic-kernel LLR scratch (float or double rows)
    double* d_llr64_stage = nullptr;       // host entry point of the f64 mode
    double* d_prob_stage = nullptr;        // host entry point of the probability-domain decoder: p0 then p1
    uint32_t* d_gs = nullptr;
    size_t gx_stride = 0, gs_stride = 0;   // per warp, elements (of the last launch)
    size_t gx_bytes = 0, gs_bytes = 0;     // capacity
    float* d_llr_stage = nullptr;
    uint32_t* d_out_stage = nullptr;
    // host entry point: H2D / decode / D2H of consecutive chunks overlap on three internal streams
    cudaStream_t st_h2d = nullptr, st_run = nullptr, st_d2h = nullptr;
    static constexpr int kMaxChunks = 16;
    cudaEvent_t ev_in[kMaxChunks] = {}, ev_done[kMaxChunks] = {};
    int last_chunks = 0;
    void* d_wgx = nullptr;                 // scratch of the wide-list kernel (lists 33..127), grow-only
    uint32_t* d_wgs = nullptr;
    size_t wgx_bytes = 0, wgs_bytes = 0;
    float* d_fgx = nullptr;                // scratch of the fast kernel
    uint32_t* d_fgs = nullptr;
    int fast_variant = -1;
    size_t fgx_bytes = 0, fgs_bytes = 0;   // capacity of the two scratch buffers (shared by all variants, grown on demand)
    size_t l2_window = 0;
    float l2_ratio = 1.0f;
    // strict mode: codewords whose smallest decision margin is below strict_tau are re-decoded in double
    int* d_flag_list = nullptr;            // [flag_cap] codeword indices, filled by the fast kernel
    int* d_flag_count = nullptr;           // [2]: counter of the current call, copy kept for polar_b200_get_info
    int flag_cap = 0;
    float strict_tau = 0.0f;
    float* d_sw_llr = nullptr;             // polar_b200_bler_sweep: one chunk of synthesised LLRs / info bits / decoded bits
    uint32_t* d_sw_truth = nullptr;
    uint32_t* d_sw_out = nullptr;
    unsigned long long* d_sw_err = nullptr;
    int sw_chunk = 0, sw_cells = 0;
    double* d_ex_gx = nullptr;             // scratch of the block-per-codeword double decoder (scl_exact.cuh), grow-only
    size_t ex_gx_bytes = 0;
    double* d_cvt = nullptr;               // float -> double conversion buffer of the kernels that take one input type
    size_t cvt_bytes = 0;
    // plain SC on the pruned tree (sc_ssc.cuh), strict mode's first pass at list size 1: per-code schedule and output map
    uint32_t* d_ssc_sched = nullptr;
    uint16_t* d_ssc_pos = nullptr;
    ssc::Layout ssc_lay = {};
    bool ssc_ok = false, ssc_prepared = false;
    long long last_flagged = -1;           // resolved lazily (device counter)
    bool flagged_pending = false;
    // one call in flight per ctx: every entry point makes its stream wait for the previous call's last launch
    cudaEvent_t ev_last = nullptr;
    cudaStream_t last_stream = nullptr;
    bool have_last = false;
    // host-side staging of the strict double entry point
    float* h_f32 = nullptr; size_t h_f32_bytes = 0;
    int* h_list = nullptr; size_t h_list_bytes = 0;
    double* h_gather = nullptr; size_t h_gather_bytes = 0;
    uint32_t* h_out2 = nullptr; size_t h_out2_bytes = 0;
    long long launches = 0;
    int last_wpb = 0, last_blocks = 0, last_smem = 0, last_kernel = 0;
    size_t scratch_bytes = 0;
};

namespace {

#define CU_TRY(expr)                                  \
    do {                                              \
        cudaError_t e__ = (expr);                     \
        if (e__ != cudaSuccess) return (int)e__;      \
    } while (0)

// fused block-error counting (fastcommon::Args / exact::Args): null truth = off
struct CountSpec {
    const uint32_t* truth = nullptr;
    unsigned long long* err = nullptr;
    long long first_index = 0;
    int n_ebno = 1;
};

struct LaunchPlan {
    int wpb, blocks, lamS, smem_x_rows, smem_s_rows, smem_bytes;
    int s_off[kMaxN + 2];
    size_t gx_rows, gs_rows;
};

// Decide which layers live in shared memory. Layers lamS..n-1 of the LLR tree
// (2^(n-lamS+1) - 2 rows) and the partial-sum layers >= max(lamS,1) go to shared memory.
LaunchPlan make_plan(const polar_b200_ctx* c, int elem = 4, int warps_per_sm_override = 0) {
    const int rowb = 32 * elem;                       // bytes of one LLR row
    LaunchPlan p;
    memset(&p, 0, sizeof(p));
    const int n = c->n;
    p.wpb = env_int("POLAR_B200_WPB", 4);
    const int warps_per_sm = warps_per_sm_override > 0 ? warps_per_sm_override : env_int("POLAR_B200_WARPS_PER_SM", 16);
    if (p.wpb < 1) p.wpb = 1;
    if (p.wpb > 8) p.wpb = 8;
    int blocks_per_sm = warps_per_sm / p.wpb;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    p.blocks = c->sm_count * blocks_per_sm;
    const int budget_per_warp = (200 * 1024) / (blocks_per_sm * p.wpb);
    int lamS = env_int("POLAR_B200_LAMS", -1);
    if (lamS < 1 || lamS > n) {
        lamS = n;   // smallest footprint, then grow while it fits
        while (lamS > 1) {
            const int cand = lamS - 1;
            int xr = (1 << (n - cand + 1)) - 2;
            int sr = 0;
            for (int lam = (cand < 1 ? 1 : cand); lam <= n - 1; ++lam) sr += ((1 << (n - lam)) + 31) / 32;
            if (xr * rowb + sr * 128 + 32 > budget_per_warp) break;
            lamS = cand;
        }
    }
    p.lamS = lamS;
    p.smem_x_rows = (1 << (n - lamS + 1)) - 2;
    const int lamSS = lamS < 1 ? 1 : lamS;
    int off = 0;
    for (int lam = 0; lam < lamSS; ++lam) { p.s_off[lam] = off; off += ((1 << (n - lam)) + 31) / 32; }
    p.gs_rows = off;
    off = 0;
    for (int lam = lamSS; lam <= n - 1; ++lam) { p.s_off[lam] = off; off += ((1 << (n - lam)) + 31) / 32; }
    p.smem_s_rows = off;
    p.gx_rows = (size_t)(1 << n) - ((size_t)1 << (n - lamS + 1));
    p.smem_bytes = (p.smem_x_rows * rowb + p.smem_s_rows * 128 + 32) * p.wpb;
    return p;
}

// Scratch of the generic kernel: grow-only by bytes (the strides travel with every launch, so plans with different
// layer splits or element types share the two buffers; a launch still running on another stream is ordered by the
// ctx's one-call-in-flight rule before anything is freed -- cudaFree synchronises the device).
int ensure_scratch(polar_b200_ctx* c, const LaunchPlan& p, int elem = 4) {
    const int warps = p.blocks * p.wpb;
    c->gx_stride = (p.gx_rows ? p.gx_rows : 1) * 32;
    c->gs_stride = (p.gs_rows ? p.gs_rows : 1) * 32;
    const size_t need_gx = c->gx_stride * warps * elem, need_gs = c->gs_stride * warps * sizeof(uint32_t);
    if (need_gx > c->gx_bytes) {
        if (c->d_gx) cudaFree(c->d_gx);
        c->d_gx = nullptr; c->gx_bytes = 0;
        CU_TRY(cudaMalloc(&c->d_gx, need_gx));
        c->gx_bytes = need_gx;
    }
    if (need_gs > c->gs_bytes) {
        if (c->d_gs) cudaFree(c->d_gs);
        c->d_gs = nullptr; c->gs_bytes = 0;
        CU_TRY(cudaMalloc(&c->d_gs, need_gs));
        c->gs_bytes = need_gs;
    }
    c->scratch_bytes = c->gx_bytes + c->gs_bytes;
    return 0;
}

// The variant table lives in fast_parts.cu (four translation units, fast_variants.cuh); concatenated here in order.
struct VariantTable {
    std::vector<FastVariant> v;
    VariantTable() {
        v.insert(v.end(), kFastPart0, kFastPart0 + kFastPartN0);
        v.insert(v.end(), kFastPart1, kFastPart1 + kFastPartN1);
        v.insert(v.end(), kFastPart2, kFastPart2 + kFastPartN2);
        v.insert(v.end(), kFastPart3, kFastPart3 + kFastPartN3);
    }
};
const VariantTable& variant_table() { static VariantTable t; return t; }
#define kFastVariants (variant_table().v.data())
#define kNumFastVariants ((int)variant_table().v.size())


int pick_fast_variant(const polar_b200_ctx* c, int L, int B) {
    const int n = c->n;
    if (env_int("POLAR_B200_FORCE_GENERIC", 0) || env_int("POLAR_B200_FORCE_WIDE", 0) || n > kMaxNWarp) return -1;
    int wlog = 0;
    while ((1 << wlog) < L) ++wlog;                 // lanes per codeword
    if (L < 1 || wlog > 5) return -1;
    const int forced = env_int("POLAR_B200_FAST_VARIANT", -1);
    if (forced >= 0 && forced < kNumFastVariants && kFastVariants[forced].nlog == n && kFastVariants[forced].wlog == wlog)
        return forced;
    // One-block-per-SM variants with the per-round barrier. Measured on B200: +25% (N=2048) / +19% (N=512) at list 32
    // for any batch of 8+ rounds per warp; for lists <= 4 +15% at 14 rounds per warp but -13..-19% at 3.5 rounds
    // (nothing to re-synchronise yet, and the warps of a sub-partition wait for the slowest of them every round).
    // POLAR_B200_SYNC: 0 = never; 1 (default) = always with one codeword per warp (lists 17..32), otherwise when the
    // batch is at least kSyncMinRounds rounds per warp; 2 = always.
    constexpr int kSyncMinRounds = 6;
    const int sync = env_int("POLAR_B200_SYNC", 1);
    if (sync >= 1)
        for (int i = 0; i < kNumFastVariants; ++i) {
            const FastVariant& v = kFastVariants[i];
            if (v.nlog != n || v.wlog != wlog || v.wpb <= 4) continue;
            const long long per_round = (long long)c->sm_count * v.bps * v.wpb * (32 >> wlog);
            if (sync >= 2 || wlog == 5 || (long long)B >= kSyncMinRounds * per_round) return i;
        }
    for (int i = 0; i < kNumFastVariants; ++i)
        if (kFastVariants[i].nlog == n && kFastVariants[i].wlog == wlog && kFastVariants[i].wpb <= 4) return i;
    return -1;
}

// list / count (device, may be null): decode only the listed codewords (strict mode's re-decode); B is then the capacity
// of the list and the grid is sized for a handful of codewords per SM (the count is not known on the host).
template <class Real, class In = Real>
int decode_generic(polar_b200_ctx* c, const In* llr, int B, int L, uint32_t* info_packed, cudaStream_t st,
                   const int* list = nullptr, const int* count = nullptr) {
    LaunchPlan p = make_plan(c, (int)sizeof(Real), list ? env_int("POLAR_B200_REDECODE_WARPS_PER_SM", 8) : 0);
    int rc = ensure_scratch(c, p, (int)sizeof(Real));
    if (rc) return rc;
    DecodeArgsT<Real, In> a;
    memset(&a, 0, sizeof(a));
    a.llr = llr; a.out = info_packed; a.list = list; a.count = count;
    a.frozen_words = c->d_frozen; a.info_order = c->d_order; a.crc_masks = c->d_crc_masks;
    a.gx = static_cast<Real*>(c->d_gx); a.gs = c->d_gs; a.gx_stride = c->gx_stride; a.gs_stride = c->gs_stride;
    a.B = B; a.n = c->n; a.K = c->K; a.crc = c->crc; a.L = L;
    int W = 1; while (W < L) W <<= 1;
    a.W = W; a.lamS = p.lamS; a.smem_x_rows = p.smem_x_rows; a.smem_s_rows = p.smem_s_rows;
    memcpy(a.s_off, p.s_off, sizeof(a.s_off));
    const int G = 32 / W;
    const int groups = (B + G - 1) / G;
    int blocks = p.blocks;
    const int need_blocks = (groups + p.wpb - 1) / p.wpb;
    if (blocks > need_blocks) blocks = need_blocks;
    CU_TRY(cudaFuncSetAttribute(scl_decode_kernel<Real, In>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes));
    scl_decode_kernel<Real, In><<<blocks, p.wpb * 32, p.smem_bytes, st>>>(a);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    if (!list) {        // the second / third pass of strict mode leaves the first pass's kernel on record
        c->last_wpb = p.wpb; c->last_blocks = blocks; c->last_smem = p.smem_bytes;
        c->last_kernel = sizeof(Real) == 8 ? -1 : 0;
    }
    return POLAR_B200_OK;
}

// ---- lists 33..127: one block of W = 64 / 128 threads per codeword (scl_wide.cuh) ----
template <class Dom, int W>
int decode_wide_w(polar_b200_ctx* c, const typename Dom::Real* in0, const typename Dom::Real* in1, int B, int L,
                  uint32_t* info_packed, cudaStream_t st) {
    using Real = typename Dom::Real;
    using Val = typename Dom::Val;
    const int n = c->n, elem = (int)sizeof(Val);
    const int blocks_per_sm = env_int("POLAR_B200_WIDE_BPS", 1024 / W >= 8 ? 8 : 4);
    const int budget = (200 * 1024) / blocks_per_sm;
    const int fixed = W * (2 * (int)sizeof(Real) + 32 + 12) + (16 + 4 * (int)sizeof(Real)) * (W / 32);
    auto s_rows_from = [&](int lam0) { int sr = 0; for (int lam = (lam0 < 1 ? 1 : lam0); lam <= n - 1; ++lam) sr += ((1 << (n - lam)) + 31) / 32; return sr; };
    int lamS = n;
    while (lamS > 1) {
        const int cand = lamS - 1;
        const int xr = (1 << (n - cand + 1)) - 2;
        if (xr * W * elem + s_rows_from(cand) * W * 4 + fixed > budget) break;
        lamS = cand;
    }
    wide::Args<Dom> a;
    memset(&a, 0, sizeof(a));
    a.lamS = lamS;
    a.smem_x_rows = (1 << (n - lamS + 1)) - 2;
    const int lamSS = lamS < 1 ? 1 : lamS;
    int off = 0;
    for (int lam = 0; lam < lamSS; ++lam) { a.s_off[lam] = off; off += ((1 << (n - lam)) + 31) / 32; }
    const size_t gs_rows = off;
    off = 0;
    for (int lam = lamSS; lam <= n - 1; ++lam) { a.s_off[lam] = off; off += ((1 << (n - lam)) + 31) / 32; }
    a.smem_s_rows = off;
    const size_t gx_rows = (size_t)(1 << n) - ((size_t)1 << (n - lamS + 1));
    const int smem = a.smem_x_rows * W * elem + a.smem_s_rows * W * 4 + fixed;
    a.gx_stride = (gx_rows ? gx_rows : 1) * W;
    a.gs_stride = (gs_rows ? gs_rows : 1) * W;
    // persistent blocks; at the largest block lengths the per-block scratch (all layers below lamS) is tens of MB,
    // so the grid is also capped by a scratch budget (POLAR_B200_WIDE_SCRATCH_GB, default 16 GB of the 180 GB)
    int max_blocks = c->sm_count * blocks_per_sm;
    {
        const double per_block = (double)a.gx_stride * elem + (double)a.gs_stride * sizeof(uint32_t);
        const double budget = (double)env_int("POLAR_B200_WIDE_SCRATCH_GB", 16) * 1073741824.0;
        const int fit = (int)(budget / per_block);
        if (max_blocks > fit) max_blocks = fit < 1 ? 1 : fit;
    }
    int blocks = max_blocks;
    if (blocks > B) blocks = B;
    const size_t need_gx = a.gx_stride * (size_t)max_blocks * elem;
    const size_t need_gs = a.gs_stride * (size_t)max_blocks * sizeof(uint32_t);
    if (need_gx > c->wgx_bytes) {
        if (c->d_wgx) cudaFree(c->d_wgx);
        c->d_wgx = nullptr; c->wgx_bytes = 0;
        CU_TRY(cudaMalloc(&c->d_wgx, need_gx));
        c->wgx_bytes = need_gx;
    }
    if (need_gs > c->wgs_bytes) {
        if (c->d_wgs) cudaFree(c->d_wgs);
        c->d_wgs = nullptr; c->wgs_bytes = 0;
        CU_TRY(cudaMalloc(&c->d_wgs, need_gs));
        c->wgs_bytes = need_gs;
    }
    a.in0 = in0; a.in1 = in1; a.out = info_packed;
    a.frozen_words = c->d_frozen; a.info_order = c->d_order; a.crc_masks = c->d_crc_masks;
    a.gx = static_cast<Val*>(c->d_wgx); a.gs = c->d_wgs;
    a.B = B; a.n = n; a.K = c->K; a.crc = c->crc; a.L = L;
    CU_TRY(cudaFuncSetAttribute(wide::scl_wide_kernel<Dom, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    wide::scl_wide_kernel<Dom, W><<<blocks, W, smem, st>>>(a);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    c->last_wpb = W / 32; c->last_blocks = blocks; c->last_smem = smem;
    c->last_kernel = Dom::kNormalise ? -4 : (sizeof(Real) == 8 ? -3 : -2);
    c->scratch_bytes = c->wgx_bytes + c->wgs_bytes;
    return POLAR_B200_OK;
}

template <class Real>
int decode_wide(polar_b200_ctx* c, const Real* llr, int B, int L, uint32_t* info_packed, cudaStream_t st) {
    if (L <= 64) return decode_wide_w<wide::LlrDom<Real>, 64>(c, llr, nullptr, B, L, info_packed, st);
    return decode_wide_w<wide::LlrDom<Real>, 128>(c, llr, nullptr, B, L, info_packed, st);
}

// probability domain (decode_scl_p1): any list size on the block-per-codeword kernel, double like the reference
int decode_prob(polar_b200_ctx* c, const double* p0, const double* p1, int B, int L, uint32_t* info_packed, cudaStream_t st) {
    if (L <= 32) return decode_wide_w<wide::ProbDom, 32>(c, p0, p1, B, L, info_packed, st);
    if (L <= 64) return decode_wide_w<wide::ProbDom, 64>(c, p0, p1, B, L, info_packed, st);
    return decode_wide_w<wide::ProbDom, 128>(c, p0, p1, B, L, info_packed, st);
}

// any list size: lists <= 32 on one warp, 33..127 on one block per codeword
template <class Real>
int decode_any(polar_b200_ctx* c, const Real* llr, int B, int L, uint32_t* info_packed, cudaStream_t st) {
    if (L > 32 || c->n > kMaxNWarp || env_int("POLAR_B200_FORCE_WIDE", 0)) return decode_wide<Real>(c, llr, B, L, info_packed, st);
    return decode_generic<Real>(c, llr, B, L, info_packed, st);
}

// plain SC on the pruned tree (sc_ssc.cuh): list size 1 with a flag list -- the rate-1 node shortcut is exact only while
// every deciding LLR has a trustworthy sign, which the margin / flag list / second pass take care of (strict mode: below
// tau, double; fp32 mode: exact zeros, leaf by leaf in fp32). One block per SM; 8 codewords per warp.
int decode_ssc(polar_b200_ctx* c, const float* llr, int B, uint32_t* out, cudaStream_t st, float* margin, int* flag_list,
               int* flag_count, float tau, int cw_base, const CountSpec* cs) {
    const ssc::Layout& lay = c->ssc_lay;
    const int smem = lay.bytes * lay.warps;
    // builds by warps per SM (the fewer warps, the more registers a thread may use, the deeper the channel prefetch)
    void (*kern)(const ssc::Args) = lay.warps > 10 ? ssc::sc_ssc_kernel<16, 4>
                                    : env_int("POLAR_B200_SSC_STAGE", 8) == 4 ? ssc::sc_ssc_kernel<10, 4> : ssc::sc_ssc_kernel<10, 8>;
    if (!c->ssc_prepared) {
        CU_TRY(cudaFuncSetAttribute(ssc::sc_ssc_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU_TRY(cudaFuncSetAttribute(ssc::sc_ssc_kernel<10, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU_TRY(cudaFuncSetAttribute(ssc::sc_ssc_kernel<10, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        c->ssc_prepared = true;
    }
    ssc::Args A;
    memset(&A, 0, sizeof(A));
    fastcommon::Args& a = A.a;
    a.llr = llr; a.out = out; a.frozen_words = c->d_frozen; a.info_order = c->d_order; a.crc_masks = c->d_crc_masks;
    a.B = B; a.K = c->K; a.crc = c->crc; a.L = 1;
    a.margin = margin; a.flag_list = flag_list; a.flag_count = flag_count; a.cw_base = cw_base;
    a.truth = cs ? cs->truth : nullptr; a.err = cs ? cs->err : nullptr;
    a.first_index = cs ? cs->first_index : 0; a.n_ebno = cs ? cs->n_ebno : 1;
    const double q = (double)tau * 16777216.0;
    a.tauq_flag = q >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)q;
    a.tauq = a.tauq_flag; a.tau = tau;
    A.sched = c->d_ssc_sched; A.pos = c->d_ssc_pos; A.lay = lay;
    A.sync_rounds = env_int("POLAR_B200_SSC_SYNC", 1);
    // Warps per SM: as many as fit (lay.warps) for the number of rounds that takes, but no more than fills those rounds
    // evenly -- the kernel is latency-bound, every warp fewer on an SM makes the others faster, and a small batch is
    // better spread thinly over all SMs than packed onto a few (N = 512, 4 096 codewords: 32 blocks x 16 warps took 83 us).
    const long long groups = ((long long)B + 7) / 8;
    const long long rounds = (groups + (long long)c->sm_count * lay.warps - 1) / ((long long)c->sm_count * lay.warps);
    int wpb = (int)((groups + (long long)c->sm_count * rounds - 1) / ((long long)c->sm_count * rounds));
    wpb = env_int("POLAR_B200_SSC_WARPS", wpb);
    if (wpb < 1 || wpb > lay.warps) wpb = lay.warps;
    int blocks = c->sm_count;
    const int need = (int)((groups + wpb - 1) / wpb);
    if (blocks > need) blocks = need;
    kern<<<blocks, wpb * 32, lay.bytes * wpb, st>>>(A);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    c->last_wpb = wpb; c->last_blocks = blocks; c->last_smem = lay.bytes * wpb;
    c->last_kernel = 500;
    return POLAR_B200_OK;
}

// margin (device, [>= cw_base + B], may be null) / flags: see fastcommon::Args. cw_base = index of llr's first row in the caller's
// batch (the host entry point decodes chunk by chunk but keeps one flag list).
// variant >= 0: entry of the reference-arithmetic table; variant <= -100: entry -100 - variant of the min-sum table.
int decode_fast(polar_b200_ctx* c, int variant, const float* llr, int B, int L, uint32_t* out, cudaStream_t st,
                float* margin = nullptr, int* flag_list = nullptr, int* flag_count = nullptr, float tau = 0.0f, int cw_base = 0,
                const CountSpec* cs = nullptr) {
    const FastVariant& v = variant <= -100 ? kFastMsPart[-100 - variant] : kFastVariants[variant];
    if (variant >= 0 && v.wlog == 0 && flag_list != nullptr && c->ssc_ok && env_int("POLAR_B200_SSC", 1) != 0 &&
        env_int("POLAR_B200_FAST_VARIANT", -1) < 0)
        return decode_ssc(c, llr, B, out, st, margin, flag_list, flag_count, tau, cw_base, cs);
    int blocks = c->sm_count * v.bps;
    const int warps = blocks * v.wpb;
    const size_t need_gx = v.gx_floats * warps * sizeof(float), need_gs = v.gs_words * warps * sizeof(uint32_t);
    if (need_gx > c->fgx_bytes) {
        if (c->d_fgx) cudaFree(c->d_fgx);
        c->d_fgx = nullptr; c->fgx_bytes = 0;
        CU_TRY(cudaMalloc(&c->d_fgx, need_gx));
        c->fgx_bytes = need_gx;
    }
    if (need_gs > c->fgs_bytes) {
        if (c->d_fgs) cudaFree(c->d_fgs);
        c->d_fgs = nullptr; c->fgs_bytes = 0;
        CU_TRY(cudaMalloc(&c->d_fgs, need_gs));
        c->fgs_bytes = need_gs;
    }
    if (c->fast_variant != variant) {
        CU_TRY(v.prepare());
        c->fast_variant = variant;
        c->scratch_bytes = (v.gx_floats * sizeof(float) + v.gs_words * sizeof(uint32_t)) * warps;
        // optional: pin the per-warp LLR scratch in L2 with an access-policy window. Measured on
        // B200 (profiles/): no gain over the default LRU behaviour (-1..-4%), so it is off by default.
        c->l2_window = 0;
        if (env_int("POLAR_B200_L2_PERSIST", 0)) {
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
            size_t want = v.gx_floats * warps * sizeof(float);
            if (max_persist > 0 && max_window > 0) {
                size_t lim = want < (size_t)max_persist ? want : (size_t)max_persist;
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, lim) == cudaSuccess) {
                    c->l2_window = want < (size_t)max_window ? want : (size_t)max_window;
                    c->l2_ratio = (float)((double)lim / (double)c->l2_window);
                    if (c->l2_ratio > 1.0f) c->l2_ratio = 1.0f;
                }
            }
            cudaGetLastError();
        }
    }
    fastcommon::Args a;
    a.llr = llr; a.out = out; a.frozen_words = c->d_frozen; a.info_order = c->d_order; a.crc_masks = c->d_crc_masks;
    a.gx = c->d_fgx; a.gs = c->d_fgs; a.B = B; a.K = c->K; a.crc = c->crc; a.L = L;
    a.margin = margin; a.flag_list = flag_list; a.flag_count = flag_count; a.cw_base = cw_base;
    a.truth = cs ? cs->truth : nullptr; a.err = cs ? cs->err : nullptr;
    a.first_index = cs ? cs->first_index : 0; a.n_ebno = cs ? cs->n_ebno : 1;
    {
        // thresholds in the kernels' fixed point (Q8.24). With a margin buffer every gap is recorded (calibration);
        // otherwise only those that will flag the codeword.
        const double q = (double)tau * 16777216.0;
        a.tauq_flag = flag_list ? (q >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)q) : 0u;
        a.tauq = margin ? 0xFFFFFFFFu : a.tauq_flag;
        a.tau = margin ? INFINITY : (flag_list ? tau : 0.0f);
    }
    {
        // leading frozen leaves handled by the cooperative phase (whole 16-leaf blocks below layer T's first node)
        const int mt = (1 << v.nlog) >> v.T;
        int pa = (c->first_info / 16) * 16;
        if (pa > mt - 16) pa = mt - 16;
        if (pa < 0 || v.wlog != 5 || env_int("POLAR_B200_PHASE_A", 1) == 0) pa = 0;
        a.PA = pa;
    }
    const int cw_per_block = v.wpb * (32 >> v.wlog);
    const int need = (B + cw_per_block - 1) / cw_per_block;
    if (blocks > need) blocks = need;
    cudaLaunchAttribute attr[1];
    int nattr = 0;
    if (c->l2_window) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = c->d_fgx;
        attr[0].val.accessPolicyWindow.num_bytes = c->l2_window;
        attr[0].val.accessPolicyWindow.hitRatio = c->l2_ratio;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        nattr = 1;
    }
    CU_TRY(v.launch(a, blocks, st, attr, nattr));
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    c->last_wpb = v.wpb; c->last_blocks = blocks; c->last_smem = v.smem_per_warp * v.wpb;
    c->last_kernel = variant <= -100 ? 1000 + (-100 - variant) : 1 + variant;
    return POLAR_B200_OK;
}


// ---- one call in flight per ctx (the scratch buffers are per ctx): a call on another stream than the previous one
// first waits for the previous call's last launch ----
int call_begin(polar_b200_ctx* c, cudaStream_t st) {
    if (!c->ev_last) CU_TRY(cudaEventCreateWithFlags(&c->ev_last, cudaEventDisableTiming));
    if (c->have_last && c->last_stream != st) CU_TRY(cudaStreamWaitEvent(st, c->ev_last, 0));
    return 0;
}
int call_end(polar_b200_ctx* c, cudaStream_t st) {
    CU_TRY(cudaEventRecord(c->ev_last, st));
    c->last_stream = st; c->have_last = true;
    return 0;
}
// wait for everything this ctx has queued (error paths, buffer growth)
void drain(polar_b200_ctx* c) {
    if (c->st_h2d) { cudaStreamSynchronize(c->st_h2d); cudaStreamSynchronize(c->st_run); cudaStreamSynchronize(c->st_d2h); }
    if (c->have_last) cudaEventSynchronize(c->ev_last);
}
int grow_host(void** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    CU_TRY(cudaMallocHost(p, need));
    *cap = need;
    return 0;
}
int ensure_flags(polar_b200_ctx* c, int B) {
    if (!c->d_flag_count) CU_TRY(cudaMalloc(&c->d_flag_count, 2 * sizeof(int)));
    if (B > c->flag_cap) {
        if (c->d_flag_list) cudaFree(c->d_flag_list);
        c->d_flag_list = nullptr; c->flag_cap = 0;
        CU_TRY(cudaMalloc(&c->d_flag_list, 2 * (size_t)B * sizeof(int)));     // second half: the third pass's list
        c->flag_cap = B;
    }
    return 0;
}

// ---- one block per codeword, double precision (scl_exact.cuh): STRICT mode's re-decode of the flagged codewords ----
constexpr int kMaxNExact = 13;
template <class In>
int decode_exact(polar_b200_ctx* c, const In* llr, int B, int L, uint32_t* out, cudaStream_t st,
                 const int* list = nullptr, const int* count = nullptr, const CountSpec* cs = nullptr,
                 int* list2 = nullptr, int* count2 = nullptr) {
    const int n = c->n, N = c->N;
    exact::Args<In> a;
    memset(&a, 0, sizeof(a));
    // Measured on B200 (N = 2048): 256 threads x 3 blocks/SM decode one codeword in 4.6 ms, 128 threads x 6 blocks/SM in
    // 6.1 ms but twice as many side by side. Long lists send about 1 % of a batch to the second pass (hundreds of codewords:
    // throughput matters), short lists a few dozen (latency matters).
    const bool many = L > 16;
    int bps = env_int("POLAR_B200_EXACT_BPS", many ? 6 : 3);
    if (bps < 1) bps = 1;
    if (bps > 8) bps = 8;
    const int budget = (220 * 1024) / bps;
    int off = 0;
    for (int lam = 0; lam <= n - 1; ++lam) { a.s_off[lam] = off; off += ((1 << (n - lam)) + 31) / 32; }
    a.smem_s_rows = off;
    const int fixed = a.smem_s_rows * 128 + 2 * 16 * 32 + 48 + (int)sizeof(exact::Tables) + 16;
    int lamS = n;                                    // smallest footprint, then grow while it fits
    while (lamS > 1) {
        const int xr = (1 << (n - (lamS - 1) + 1)) - 2;
        if (xr * 256 + fixed > budget) break;
        --lamS;
    }
    a.lamS = lamS;
    a.smem_x_rows = (1 << (n - lamS + 1)) - 2;
    const int smem = a.smem_x_rows * 256 + fixed;
    if (smem > 227 * 1024) return POLAR_B200_E_UNSUPPORTED;
    const size_t gx_rows = (size_t)N - ((size_t)1 << (n - lamS + 1));
    a.gx_stride = (gx_rows ? gx_rows : 1) * 32;
    int blocks = c->sm_count * bps;
    if (!list && blocks > B) blocks = B;
    const size_t need = a.gx_stride * (size_t)(c->sm_count * bps) * sizeof(double);
    if (need > c->ex_gx_bytes) {
        if (c->d_ex_gx) cudaFree(c->d_ex_gx);
        c->d_ex_gx = nullptr; c->ex_gx_bytes = 0;
        CU_TRY(cudaMalloc(&c->d_ex_gx, need));
        c->ex_gx_bytes = need;
    }
    a.llr = llr; a.out = out; a.list = list; a.count = count; a.list2 = list2; a.count2 = count2;
    a.frozen_words = c->d_frozen; a.info_order = c->d_order; a.crc_masks = c->d_crc_masks;
    a.gx = c->d_ex_gx;
    a.truth = cs ? cs->truth : nullptr; a.err = cs ? cs->err : nullptr;
    a.first_index = cs ? cs->first_index : 0; a.n_ebno = cs ? cs->n_ebno : 1;
    a.B = B; a.n = n; a.K = c->K; a.crc = c->crc; a.L = L;
    int W = 1; while (W < L) W <<= 1;
    a.W = W;
    a.big = env_int("POLAR_B200_EXACT_BIG", 32);
    if (a.big < 32) a.big = 32;
    CU_TRY(cudaFuncSetAttribute(exact::scl_exact_kernel<In>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int nt = env_int("POLAR_B200_EXACT_THREADS", many ? 128 : exact::NT);
    if (nt < 32) nt = 32;
    if (nt > exact::NT) nt = exact::NT;
    nt &= ~31;
    exact::scl_exact_kernel<In><<<blocks, nt, smem, st>>>(a);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    if (!list) { c->last_wpb = exact::NT / 32; c->last_blocks = blocks; c->last_smem = smem; c->last_kernel = -5; }
    return POLAR_B200_OK;
}

// STRICT mode's second pass over the flagged codewords
int redecode_flagged(polar_b200_ctx* c, const float* llr, int B, int L, uint32_t* out, cudaStream_t st,
                     const CountSpec* cs = nullptr) {
    if (c->n <= kMaxNExact && L <= 32 && (cs || env_int("POLAR_B200_REDECODE_GENERIC", 0) == 0)) {
        if (cs) return decode_exact<float>(c, llr, B, L, out, st, c->d_flag_list, c->d_flag_count, cs);
        // second pass: double with short dependency chains; third pass (normally empty): whatever the second pass found
        // to hinge on an exact cancellation, with the reference's literal formulas
        int* list2 = c->d_flag_list + c->flag_cap;
        int* count2 = c->d_flag_count + 1;
        CU_TRY(cudaMemsetAsync(count2, 0, sizeof(int), st));
        int rc = decode_exact<float>(c, llr, B, L, out, st, c->d_flag_list, c->d_flag_count, nullptr, list2, count2);
        if (rc) return rc;
        return decode_generic<double, float>(c, llr, B, L, out, st, list2, count2);
    }
    return decode_generic<double, float>(c, llr, B, L, out, st, c->d_flag_list, c->d_flag_count);
}

__global__ void to_double_kernel(const float* in, double* out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}

// everything in double on float LLRs (reference precision: the reference itself widens its input to double)
int decode_f64_from_float(polar_b200_ctx* c, const float* llr, int B, int L, uint32_t* out, cudaStream_t st) {
    if (L > 32 || c->n > kMaxNWarp || env_int("POLAR_B200_FORCE_WIDE", 0)) {
        const size_t nel = (size_t)B * c->N;
        if (nel * sizeof(double) > c->cvt_bytes) {
            if (c->d_cvt) cudaFree(c->d_cvt);
            c->d_cvt = nullptr; c->cvt_bytes = 0;
            CU_TRY(cudaMalloc(&c->d_cvt, nel * sizeof(double)));
            c->cvt_bytes = nel * sizeof(double);
        }
        to_double_kernel<<<c->sm_count * 8, 256, 0, st>>>(llr, c->d_cvt, nel);
        CU_TRY(cudaGetLastError());
        c->launches += 1;
        return decode_wide<double>(c, c->d_cvt, B, L, out, st);
    }
    if (env_int("POLAR_B200_F64_EXACT_KERNEL", 0) && c->n <= kMaxNExact) return decode_exact<float>(c, llr, B, L, out, st);
    return decode_generic<double, float>(c, llr, B, L, out, st);
}

int upload_amplitudes(polar_b200_ctx* c, const double* ebno_db, int n_ebno, cudaStream_t st) {
    double amp[64];
    for (int i = 0; i < n_ebno; ++i)      // PolarCode.cpp:744-745
        amp[i] = pow(10.0, ebno_db[i] / 20.0) * sqrt((double)c->K / (double)c->N);
    CU_TRY(cudaMemcpyAsync(c->d_amp, amp, n_ebno * sizeof(double), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));    // amp[] is a stack buffer
    return 0;
}
int synth_launch(polar_b200_ctx* c, unsigned long long seed, long long first_index, int B, int n_ebno, float* llr,
                 uint32_t* truth_packed, cudaStream_t st) {
    SynthArgs a;
    a.llr = llr; a.truth = truth_packed; a.inv_order = c->d_inv_order; a.crc_rows = c->d_crc_rows; a.amp = c->d_amp;
    a.seed = seed; a.first_index = first_index; a.B = B; a.n = c->n; a.K = c->K; a.crc = c->crc; a.n_ebno = n_ebno;
    int blocks = (B + 3) / 4;
    const int cap = c->sm_count * 16;
    if (blocks > cap) blocks = cap;
    synth_kernel<<<blocks, 128, 0, st>>>(a);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return POLAR_B200_OK;
}

float strict_tau(const polar_b200_ctx* c) {
    const char* e = getenv("POLAR_B200_STRICT_TAU");
    if (e && *e) return (float)atof(e);
    return c->strict_tau;
}

// Device-resident decode in one of the three arithmetic modes (include/polar_b200.h). llr / out: device pointers.
// cw_base / zero_count / redecode serve the chunked host entry point: chunks share one flag list and are re-decoded
// together after the last chunk.
// cs: count block errors against cs->truth (fused into the kernels' tails where they support it, a separate small kernel
// otherwise; `out` must then be a real buffer).
int decode_mode(polar_b200_ctx* c, const float* llr, int B, int L, uint32_t* out, int mode, float* margin, cudaStream_t st,
                int fv_forced = -2, int cw_base = 0, bool zero_count = true, bool redecode = true, const CountSpec* cs = nullptr) {
    auto count_after = [&](int rc) {
        if (rc || !cs) return rc;
        count_errors_bucket_kernel<<<(B + 255) / 256, 256, 0, st>>>(out, cs->truth, B, c->KW, cs->first_index, cs->n_ebno, cs->err);
        c->launches += 1;
        return (int)cudaGetLastError();
    };
    // the fast kernels read the channel rows with 16-byte loads; anything else goes to the generic kernel (scalar loads)
    const bool aligned = (reinterpret_cast<uintptr_t>(llr) & 15u) == 0;
    const int fv = !aligned ? -1 : (fv_forced >= -1 ? fv_forced : pick_fast_variant(c, L, B));
    if (mode == POLAR_B200_MODE_MINSUM) {
        // opt-in non-parity arithmetic: only where a min-sum build of a fast kernel exists
        if (!aligned || L > 32) return POLAR_B200_E_UNSUPPORTED;
        int wlog = 0;
        while ((1 << wlog) < L) ++wlog;
        for (int i = 0; i < kFastMsPartN; ++i)
            if (kFastMsPart[i].nlog == c->n && kFastMsPart[i].wlog == wlog)
                return decode_fast(c, -100 - i, llr, B, L, out, st, margin, nullptr, nullptr, 0.0f, 0, cs);
        return POLAR_B200_E_UNSUPPORTED;
    }
    if (margin && fv < 0) return POLAR_B200_E_UNSUPPORTED;       // margins are reported by the fast kernels only
    if (mode == POLAR_B200_MODE_FP32) {
        if (fv >= 0 && L == 1 && !cs && c->ssc_ok && env_int("POLAR_B200_SSC", 1) != 0 && env_int("POLAR_B200_FAST_VARIANT", -1) < 0) {
            // list size 1: the pruned-tree kernel (sc_ssc.cuh). Its rate-1 shortcut is not the leaf-by-leaf decoder on a
            // codeword with an exactly-zero deciding LLR (a tie): those few -- none on a real channel -- are listed by the
            // kernel (threshold = one step of its fixed point, 6e-8) and decoded leaf by leaf by the generic fp32 kernel.
            // The list is local to this call (row 0 = llr's first row), so chunks of the host entry point stand alone.
            int rc = ensure_flags(c, B);
            if (rc) return rc;
            CU_TRY(cudaMemsetAsync(c->d_flag_count, 0, sizeof(int), st));
            rc = decode_fast(c, fv, llr, B, L, out, st, margin, c->d_flag_list, c->d_flag_count, 1.0f / 16777216.0f, 0, nullptr);
            if (rc) return rc;
            const int kind = c->last_kernel;
            rc = decode_generic<float, float>(c, llr, B, L, out, st, c->d_flag_list, c->d_flag_count);
            c->last_kernel = kind;
            return rc;
        }
        if (fv >= 0) return decode_fast(c, fv, llr, B, L, out, st, margin, nullptr, nullptr, 0.0f, 0, cs);
        return count_after(decode_any<float>(c, llr, B, L, out, st));
    }
    if (mode == POLAR_B200_MODE_F64 || fv < 0) return count_after(decode_f64_from_float(c, llr, B, L, out, st));
    // strict: fp32 everywhere, double where a decision was closer than tau
    int rc = ensure_flags(c, cw_base + B);
    if (rc) return rc;
    if (zero_count) CU_TRY(cudaMemsetAsync(c->d_flag_count, 0, sizeof(int), st));
    rc = decode_fast(c, fv, llr, B, L, out, st, margin, c->d_flag_list, c->d_flag_count, strict_tau(c), cw_base, cs);
    if (rc) return rc;
    c->flagged_pending = true;
    if (!redecode) return POLAR_B200_OK;
    return redecode_flagged(c, llr, B, L, out, st, cs);
}

// Host memory in and out: the batch is cut into chunks whose H2D copy, decode and D2H copy overlap on three internal
// streams. redecode_on_device = false (strict mode only): flags are left in d_flag_list / d_flag_count for the caller.
int host_pipeline(polar_b200_ctx* c, const float* llr_host, int B, int L, uint32_t* info_packed_host, int mode,
                  cudaStream_t user_stream, bool redecode_on_device) {
    CU_TRY(cudaStreamSynchronize(user_stream));     // work queued by the caller comes first
    if (!c->st_h2d) {
        CU_TRY(cudaStreamCreateWithFlags(&c->st_h2d, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&c->st_run, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&c->st_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < polar_b200_ctx::kMaxChunks; ++i) {
            CU_TRY(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
        }
    }
    int rc = call_begin(c, c->st_run);
    if (rc) return rc;
    c->have_last = false;                           // this call ends with everything synchronised
    const bool strict = mode == POLAR_B200_MODE_STRICT;
    if (mode == POLAR_B200_MODE_MINSUM) {
        CU_TRY(cudaMemcpyAsync(c->d_llr_stage, llr_host, (size_t)B * c->N * sizeof(float), cudaMemcpyHostToDevice, c->st_run));
        rc = decode_mode(c, c->d_llr_stage, B, L, c->d_out_stage, mode, nullptr, c->st_run);
        if (rc) return rc;
        CU_TRY(cudaMemcpyAsync(info_packed_host, c->d_out_stage, (size_t)B * c->KW * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->st_run));
        CU_TRY(cudaStreamSynchronize(c->st_run));
        c->last_chunks = 1;
        return POLAR_B200_OK;
    }
    int fv = pick_fast_variant(c, L, B);
    if (mode == POLAR_B200_MODE_F64 || fv < 0 && (strict || L > 32 || c->n > kMaxNWarp || env_int("POLAR_B200_FORCE_WIDE", 0))) {
        // double everywhere, wide lists, N > 8192: a single chunk on the run stream
        CU_TRY(cudaMemcpyAsync(c->d_llr_stage, llr_host, (size_t)B * c->N * sizeof(float), cudaMemcpyHostToDevice, c->st_run));
        rc = (mode == POLAR_B200_MODE_FP32) ? decode_wide<float>(c, c->d_llr_stage, B, L, c->d_out_stage, c->st_run)
                                            : decode_f64_from_float(c, c->d_llr_stage, B, L, c->d_out_stage, c->st_run);
        if (rc) return rc;
        CU_TRY(cudaMemcpyAsync(info_packed_host, c->d_out_stage, (size_t)B * c->KW * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->st_run));
        CU_TRY(cudaStreamSynchronize(c->st_run));
        c->last_chunks = 1;
        if (strict) { if ((rc = ensure_flags(c, 1))) return rc; CU_TRY(cudaMemset(c->d_flag_count, 0, sizeof(int))); }
        return POLAR_B200_OK;
    }
    // Chunks are whole "rounds" of the persistent grid (every warp decodes the same number of
    // codewords per chunk), so splitting costs no extra tail; about 6 chunks hide the PCIe time.
    // The kernel variant is chosen for the chunk size (every launch is one chunk).
    auto round_size = [&](int v) {
        if (v >= 0) return c->sm_count * kFastVariants[v].bps * kFastVariants[v].wpb * (32 >> kFastVariants[v].wlog);
        LaunchPlan p = make_plan(c, 4);
        int W = 1; while (W < L) W <<= 1;
        return p.blocks * p.wpb * (32 / W);
    };
    auto chunk_size = [&](int per_round) {
        const int rounds = (B + per_round - 1) / per_round;
        long long ch = (long long)((rounds + 5) / 6) * per_round;
        if (env_int("POLAR_B200_HOST_CHUNKS", 1) == 0 || ch <= 0 || ch > B) ch = B;
        return ch;
    };
    long long chunk = chunk_size(round_size(fv));
    const int fv2 = pick_fast_variant(c, L, (int)chunk);
    if (fv2 != fv) { fv = fv2; chunk = chunk_size(round_size(fv)); }
    if (fv >= 0 && kFastVariants[fv].wlog <= 1 && env_int("POLAR_B200_HOST_CHUNKS", 1) == 1) {
        // lists <= 2 decode at least twice as fast as PCIe delivers the LLRs (8 KB per codeword at N = 2048): the copy is the
        // critical path, so cut the batch into six equal chunks even if that is less than one round of the grid --
        // only the last chunk's decode is then exposed
        fv = pick_fast_variant(c, L, (B + 5) / 6);
        const int cpb = kFastVariants[fv].wpb * (32 >> kFastVariants[fv].wlog);     // codewords per block
        chunk = (((long long)(B + 5) / 6 + cpb - 1) / cpb) * cpb;
        if (chunk > B) chunk = B;
    }
    int nchunks = (int)((B + chunk - 1) / chunk);
    if (nchunks > polar_b200_ctx::kMaxChunks) { nchunks = polar_b200_ctx::kMaxChunks; chunk = ((long long)B + nchunks - 1) / nchunks; }
    // chunk boundaries (in codewords). Equal chunks by default. With one codeword per warp (lists 17..32) decoding a
    // round takes longer than copying it in, so the chunks grow geometrically instead (1, 2, 4, then 6 rounds each):
    // only the first, small copy is exposed, every later copy is shorter than the decode it hides behind as long as the
    // host delivers about 11 GB/s to this GPU (measured: 55 GB/s with one GPU copying, 23 GB/s per GPU with eight at once
    // -- round 1's 1, 4, 23 schedule left the last 446 MB copy exposed there and cost 17 % at eight GPUs), and there are
    // still few chunk boundaries, each of which is a grid-wide join where the fastest warps wait for the slowest.
    long long bounds[polar_b200_ctx::kMaxChunks + 1];
    bounds[0] = 0;
    const bool geometric = fv >= 0 && kFastVariants[fv].wlog == 5 && env_int("POLAR_B200_HOST_CHUNKS", 1) == 1;
    if (geometric) {
        const long long per_round = round_size(fv);
        long long max_rounds = env_int("POLAR_B200_HOST_MAX_ROUNDS", 6);
        if (max_rounds < 1) max_rounds = 1;
        long long rounds_in_chunk = 1;
        nchunks = 0;
        while (bounds[nchunks] < B && nchunks < polar_b200_ctx::kMaxChunks) {
            long long hi = bounds[nchunks] + rounds_in_chunk * per_round;
            if (hi > B || nchunks == polar_b200_ctx::kMaxChunks - 1) hi = B;
            if (B - hi < per_round * rounds_in_chunk / 2) hi = B;       // no small leftover chunk
            bounds[++nchunks] = hi;
            rounds_in_chunk = rounds_in_chunk * 2 < max_rounds ? rounds_in_chunk * 2 : max_rounds;
        }
    } else {
        for (int i = 1; i <= nchunks; ++i) bounds[i] = ((long long)i * chunk < B) ? (long long)i * chunk : B;
    }
    c->last_chunks = nchunks;
    const bool final_copy = strict && redecode_on_device;     // the re-decode rewrites rows of earlier chunks
    if (strict && (rc = ensure_flags(c, B))) return rc;       // one flag list for all chunks: sized before the first launch
    for (int i = 0; i < nchunks; ++i) {
        const long long lo = bounds[i];
        const int nb = (int)(bounds[i + 1] - lo);
        float* d_in = c->d_llr_stage + (size_t)lo * c->N;
        uint32_t* d_o = c->d_out_stage + (size_t)lo * c->KW;
        CU_TRY(cudaMemcpyAsync(d_in, llr_host + (size_t)lo * c->N, (size_t)nb * c->N * sizeof(float),
                               cudaMemcpyHostToDevice, c->st_h2d));
        CU_TRY(cudaEventRecord(c->ev_in[i], c->st_h2d));
        CU_TRY(cudaStreamWaitEvent(c->st_run, c->ev_in[i], 0));
        rc = decode_mode(c, d_in, nb, L, d_o, mode, nullptr, c->st_run, fv, (int)lo, i == 0, false);
        if (rc) return rc;
        if (final_copy) continue;
        CU_TRY(cudaEventRecord(c->ev_done[i], c->st_run));
        CU_TRY(cudaStreamWaitEvent(c->st_d2h, c->ev_done[i], 0));
        CU_TRY(cudaMemcpyAsync(info_packed_host + (size_t)lo * c->KW, d_o, (size_t)nb * c->KW * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, c->st_d2h));
    }
    if (final_copy) {
        rc = redecode_flagged(c, c->d_llr_stage, B, L, c->d_out_stage, c->st_run);
        if (rc) return rc;
        CU_TRY(cudaMemcpyAsync(info_packed_host, c->d_out_stage, (size_t)B * c->KW * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->st_run));
        CU_TRY(cudaStreamSynchronize(c->st_run));
    }
    CU_TRY(cudaStreamSynchronize(c->st_d2h));
    return POLAR_B200_OK;
}

}  // namespace

extern "C" {

int polar_b200_abi_version(void) { return POLAR_B200_ABI_VERSION; }

const char* polar_b200_strerror(int code) {
    switch (code) {
        case POLAR_B200_OK: return "ok";
        case POLAR_B200_E_ARG: return "polar_b200: invalid argument";
        case POLAR_B200_E_UNSUPPORTED: return "polar_b200: parameter outside what this build supports (n <= 15, list <= 127)";
        case POLAR_B200_E_NOGPU: return "polar_b200: no usable CUDA device (there is no CPU fallback)";
        case POLAR_B200_E_BATCH: return "polar_b200: batch larger than the ctx's max_batch";
        case POLAR_B200_E_LIST: return "polar_b200: list size must be in 1..min(max_list, 127)";
        case POLAR_B200_E_NONCCL: return "polar_b200: libnccl.so.2 could not be loaded (needed only by polar_b200_comm_*)";
        case POLAR_B200_E_NCCL: return "polar_b200: an NCCL call failed";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "polar_b200: unknown error";
}

int polar_b200_info_words(int K) { return (K + 31) / 32; }

void* polar_b200_host_alloc(size_t bytes, int write_combined) {
    void* p = nullptr;
    const unsigned flags = cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0u);
    if (bytes == 0 || cudaHostAlloc(&p, bytes, flags) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
int polar_b200_host_free(void* p) {
    if (!p) return POLAR_B200_E_ARG;
    return (int)cudaFreeHost(p);
}

// test hook (no GPU needed): the pruned-tree schedule and output map sc_ssc.cuh builds for a code
int polar_b200_ssc_schedule(int n, const uint8_t* frozen_mask, uint32_t* ops_out, int cap) {
    if (!frozen_mask || cap < 0) return POLAR_B200_E_ARG;
    std::vector<uint32_t> ops;
    if (!ssc::build_schedule(n, frozen_mask, ops)) return 0;
    if ((int)ops.size() > cap || !ops_out) return (int)ops.size();
    memcpy(ops_out, ops.data(), ops.size() * 4);
    return (int)ops.size();
}
int polar_b200_ssc_positions(int n, const uint16_t* info_order, int K, uint16_t* pos_out) {
    if (!info_order || !pos_out || K < 1 || n < 2 || n > 15) return POLAR_B200_E_ARG;
    std::vector<uint16_t> pos;
    ssc::build_positions(n, info_order, K, pos);
    memcpy(pos_out, pos.data(), (size_t)K * 2);
    return POLAR_B200_OK;
}

int polar_b200_fast_variant_count(void) { return kNumFastVariants; }
int polar_b200_fast_variant_desc(int index, int* nlog, int* lanes_log2, int* warps_per_block) {
    if (index < 0 || index >= kNumFastVariants) return POLAR_B200_E_ARG;
    if (nlog) *nlog = kFastVariants[index].nlog;
    if (lanes_log2) *lanes_log2 = kFastVariants[index].wlog;
    if (warps_per_block) *warps_per_block = kFastVariants[index].wpb;
    return POLAR_B200_OK;
}

int polar_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int polar_b200_create(polar_b200_ctx** out, int device, int n, int K, int crc_bits,
                      const uint8_t* frozen_mask, const uint16_t* info_order,
                      const uint8_t* crc_matrix, int max_list, int max_batch) {
    if (!out) return POLAR_B200_E_ARG;
    *out = nullptr;
    if (!frozen_mask || !info_order) return POLAR_B200_E_ARG;
    if (n < 1 || K < 1 || crc_bits < 0 || max_batch < 1) return POLAR_B200_E_ARG;
    if (n > kMaxN) return POLAR_B200_E_UNSUPPORTED;
    const int N = 1 << n;
    if (K + crc_bits > N) return POLAR_B200_E_ARG;
    if (crc_bits > 0 && !crc_matrix) return POLAR_B200_E_ARG;
    if (max_list < 1) return POLAR_B200_E_LIST;
    if (max_list > kMaxList) return POLAR_B200_E_UNSUPPORTED;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return POLAR_B200_E_NOGPU; }
    if (device < 0 || device >= ndev) return POLAR_B200_E_ARG;
    // consistency of the tables
    int n_unfrozen = 0;
    for (int i = 0; i < N; ++i) n_unfrozen += frozen_mask[i] ? 0 : 1;
    if (n_unfrozen != K + crc_bits) return POLAR_B200_E_ARG;
    for (int j = 0; j < K + crc_bits; ++j)
        if (info_order[j] >= N || frozen_mask[info_order[j]]) return POLAR_B200_E_ARG;

    polar_b200_ctx* c = new (std::nothrow) polar_b200_ctx;
    if (!c) return POLAR_B200_E_ARG;
    c->device = device; c->n = n; c->N = N; c->K = K; c->crc = crc_bits;
    c->max_list = max_list; c->max_batch = max_batch;
    c->strict_tau = POLAR_B200_DEFAULT_STRICT_TAU;
    c->KW = (K + 31) / 32; c->NW = (N + 31) / 32;
    int rc = 0;
    auto fail = [&](int code) { polar_b200_destroy(c); return code; };
    if ((rc = (int)cudaSetDevice(device)) != 0) return fail(rc);
    if ((rc = (int)cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device)) != 0) return fail(rc);

    c->first_info = N;
    for (int i = N - 1; i >= 0; --i) if (!frozen_mask[i]) c->first_info = i;
    std::vector<uint32_t> fw(c->NW, 0);
    for (int i = 0; i < N; ++i) if (frozen_mask[i]) fw[i >> 5] |= 1u << (i & 31);
    // parity row r over decoding positions: ones at order[j] where M[r][j] = 1, plus the
    // parity bit's own position order[K + r] (PolarCode.cpp:95-101 rearranged to "XOR == 0").
    std::vector<uint32_t> cm((size_t)(crc_bits ? crc_bits : 1) * c->NW, 0);
    for (int r = 0; r < crc_bits; ++r) {
        for (int j = 0; j < K; ++j)
            if (crc_matrix[(size_t)r * K + j] & 1) cm[(size_t)r * c->NW + (info_order[j] >> 5)] ^= 1u << (info_order[j] & 31);
        const int p = info_order[K + r];
        cm[(size_t)r * c->NW + (p >> 5)] ^= 1u << (p & 31);
    }
    if ((rc = (int)cudaMalloc(&c->d_frozen, fw.size() * 4)) != 0) return fail(rc);
    if ((rc = (int)cudaMalloc(&c->d_order, (size_t)(K + crc_bits) * 2)) != 0) return fail(rc);
    if ((rc = (int)cudaMalloc(&c->d_crc_masks, cm.size() * 4)) != 0) return fail(rc);
    if ((rc = (int)cudaMemcpy(c->d_frozen, fw.data(), fw.size() * 4, cudaMemcpyHostToDevice)) != 0) return fail(rc);
    if ((rc = (int)cudaMemcpy(c->d_order, info_order, (size_t)(K + crc_bits) * 2, cudaMemcpyHostToDevice)) != 0) return fail(rc);
    if ((rc = (int)cudaMemcpy(c->d_crc_masks, cm.data(), cm.size() * 4, cudaMemcpyHostToDevice)) != 0) return fail(rc);
    {
        std::vector<uint16_t> inv(N, 0xFFFF);
        for (int j = 0; j < K + crc_bits; ++j) inv[info_order[j]] = (uint16_t)j;
        std::vector<uint32_t> rows((size_t)(crc_bits ? crc_bits : 1) * c->KW, 0);
        for (int r = 0; r < crc_bits; ++r)
            for (int j = 0; j < K; ++j)
                if (crc_matrix[(size_t)r * K + j] & 1) rows[(size_t)r * c->KW + (j >> 5)] |= 1u << (j & 31);
        if ((rc = (int)cudaMalloc(&c->d_inv_order, (size_t)N * 2)) != 0) return fail(rc);
        if ((rc = (int)cudaMalloc(&c->d_crc_rows, rows.size() * 4)) != 0) return fail(rc);
        if ((rc = (int)cudaMalloc(&c->d_amp, 64 * sizeof(double))) != 0) return fail(rc);
        if ((rc = (int)cudaMemcpy(c->d_inv_order, inv.data(), (size_t)N * 2, cudaMemcpyHostToDevice)) != 0) return fail(rc);
        if ((rc = (int)cudaMemcpy(c->d_crc_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice)) != 0) return fail(rc);
    }
    {
        // list size 1 in strict mode: schedule of the pruned tree and the output map (sc_ssc.cuh)
        std::vector<uint32_t> ops;
        std::vector<uint16_t> pos;
        if (ssc::build_schedule(n, frozen_mask, ops)) {
            ssc::build_positions(n, info_order, K, pos);
            c->ssc_lay = ssc::make_layout(n);
            if ((rc = (int)cudaMalloc(&c->d_ssc_sched, ops.size() * 4)) != 0) return fail(rc);
            if ((rc = (int)cudaMalloc(&c->d_ssc_pos, pos.size() * 2)) != 0) return fail(rc);
            if ((rc = (int)cudaMemcpy(c->d_ssc_sched, ops.data(), ops.size() * 4, cudaMemcpyHostToDevice)) != 0) return fail(rc);
            if ((rc = (int)cudaMemcpy(c->d_ssc_pos, pos.data(), pos.size() * 2, cudaMemcpyHostToDevice)) != 0) return fail(rc);
            c->ssc_ok = c->ssc_lay.warps >= 1;
        }
    }
    if ((rc = (int)cudaMalloc(&c->d_llr_stage, (size_t)max_batch * N * sizeof(float))) != 0) return fail(rc);
    if ((rc = (int)cudaMalloc(&c->d_out_stage, (size_t)max_batch * c->KW * sizeof(uint32_t))) != 0) return fail(rc);
    *out = c;
    return POLAR_B200_OK;
}

int polar_b200_destroy(polar_b200_ctx* c) {
    if (!c) return POLAR_B200_E_ARG;
    cudaSetDevice(c->device);
    cudaFree(c->d_frozen); cudaFree(c->d_order); cudaFree(c->d_crc_masks);
    cudaFree(c->d_inv_order); cudaFree(c->d_crc_rows); cudaFree(c->d_amp);
    cudaFree(c->d_gx); cudaFree(c->d_gs); cudaFree(c->d_llr_stage); cudaFree(c->d_out_stage);
    cudaFree(c->d_fgx); cudaFree(c->d_fgs); cudaFree(c->d_llr64_stage); cudaFree(c->d_wgx); cudaFree(c->d_wgs); cudaFree(c->d_prob_stage);
    cudaFree(c->d_flag_list); cudaFree(c->d_flag_count); cudaFree(c->d_cvt); cudaFree(c->d_ex_gx);
    cudaFree(c->d_ssc_sched); cudaFree(c->d_ssc_pos);
    cudaFree(c->d_sw_llr); cudaFree(c->d_sw_truth); cudaFree(c->d_sw_out); cudaFree(c->d_sw_err);
    cudaFreeHost(c->h_f32); cudaFreeHost(c->h_list); cudaFreeHost(c->h_gather); cudaFreeHost(c->h_out2);
    if (c->ev_last) cudaEventDestroy(c->ev_last);
    if (c->st_h2d) {
        cudaStreamDestroy(c->st_h2d); cudaStreamDestroy(c->st_run); cudaStreamDestroy(c->st_d2h);
        for (int i = 0; i < polar_b200_ctx::kMaxChunks; ++i) { cudaEventDestroy(c->ev_in[i]); cudaEventDestroy(c->ev_done[i]); }
    }
    delete c;
    return POLAR_B200_OK;
}

int polar_b200_decode_scl_llr(polar_b200_ctx* c, const float* llr, int B, int L,
                              uint32_t* info_packed, void* cuda_stream) {
    return polar_b200_decode_scl_llr_ex(c, llr, B, L, info_packed, POLAR_B200_MODE_FP32, nullptr, cuda_stream);
}

int polar_b200_decode_scl_llr_ex(polar_b200_ctx* c, const float* llr, int B, int L, uint32_t* info_packed,
                                 int mode, float* margin, void* cuda_stream) {
    if (!c || !llr || !info_packed || B < 0) return POLAR_B200_E_ARG;
    if (mode < POLAR_B200_MODE_FP32 || mode > POLAR_B200_MODE_MINSUM) return POLAR_B200_E_ARG;
    if (L < 1 || L > c->max_list || L > kMaxList) return POLAR_B200_E_LIST;
    if (B == 0) return POLAR_B200_OK;
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = call_begin(c, st);
    if (rc) return rc;
    rc = decode_mode(c, llr, B, L, info_packed, mode, margin, st);
    if (rc) return rc;
    return call_end(c, st);
}

int polar_b200_decode_scl_llr_host(polar_b200_ctx* c, const float* llr_host, int B, int L,
                                   uint32_t* info_packed_host, void* cuda_stream) {
    return polar_b200_decode_scl_llr_host_ex(c, llr_host, B, L, info_packed_host, POLAR_B200_MODE_FP32, cuda_stream);
}

int polar_b200_decode_scl_llr_host_ex(polar_b200_ctx* c, const float* llr_host, int B, int L,
                                      uint32_t* info_packed_host, int mode, void* cuda_stream) {
    if (!c || !llr_host || !info_packed_host || B < 0) return POLAR_B200_E_ARG;
    if (mode < POLAR_B200_MODE_FP32 || mode > POLAR_B200_MODE_MINSUM) return POLAR_B200_E_ARG;
    if (L < 1 || L > c->max_list || L > kMaxList) return POLAR_B200_E_LIST;
    if (B == 0) return POLAR_B200_OK;
    CU_TRY(cudaSetDevice(c->device));
    int rc = polar_b200_reserve(c, B);
    if (rc) return rc;
    rc = host_pipeline(c, llr_host, B, L, info_packed_host, mode, (cudaStream_t)cuda_stream, true);
    if (rc) drain(c);                  // no copy into the caller's buffers may still be in flight when an error is returned
    return rc;
}

int polar_b200_decode_scl_llr_f64_strict_host(polar_b200_ctx* c, const double* llr_host, int B, int L,
                                              uint32_t* info_packed_host, void* cuda_stream) {
    if (!c || !llr_host || !info_packed_host || B < 0) return POLAR_B200_E_ARG;
    if (L < 1 || L > c->max_list || L > kMaxList) return POLAR_B200_E_LIST;
    if (B == 0) return POLAR_B200_OK;
    CU_TRY(cudaSetDevice(c->device));
    if (pick_fast_variant(c, L, B) < 0)  // no margin-reporting kernel for this (n, list): everything in double
        return polar_b200_decode_scl_llr_f64_host(c, llr_host, B, L, info_packed_host, cuda_stream);
    int rc = polar_b200_reserve(c, B);
    if (rc) return rc;
    const size_t nel = (size_t)B * c->N;
    if ((rc = grow_host((void**)&c->h_f32, &c->h_f32_bytes, nel * sizeof(float)))) return rc;
    for (size_t i = 0; i < nel; ++i) c->h_f32[i] = (float)llr_host[i];
    // stage 1: fp32 decode of everything, flags only (the re-decode needs the caller's doubles)
    rc = host_pipeline(c, c->h_f32, B, L, info_packed_host, POLAR_B200_MODE_STRICT, (cudaStream_t)cuda_stream, false);
    if (rc) { drain(c); return rc; }
    int nf = 0;
    CU_TRY(cudaMemcpy(&nf, c->d_flag_count, sizeof(int), cudaMemcpyDeviceToHost));
    c->last_flagged = nf; c->flagged_pending = false;
    if (nf == 0) return POLAR_B200_OK;
    if ((rc = grow_host((void**)&c->h_list, &c->h_list_bytes, (size_t)nf * sizeof(int)))) return rc;
    if ((rc = grow_host((void**)&c->h_gather, &c->h_gather_bytes, (size_t)nf * c->N * sizeof(double)))) return rc;
    if ((rc = grow_host((void**)&c->h_out2, &c->h_out2_bytes, (size_t)nf * c->KW * sizeof(uint32_t)))) return rc;
    CU_TRY(cudaMemcpy(c->h_list, c->d_flag_list, (size_t)nf * sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < nf; ++i)
        memcpy(c->h_gather + (size_t)i * c->N, llr_host + (size_t)c->h_list[i] * c->N, (size_t)c->N * sizeof(double));
    // stage 2: the flagged codewords again, in double on the caller's double LLRs
    rc = polar_b200_decode_scl_llr_f64_host(c, c->h_gather, nf, L, c->h_out2, cuda_stream);
    if (rc) return rc;
    for (int i = 0; i < nf; ++i)
        memcpy(info_packed_host + (size_t)c->h_list[i] * c->KW, c->h_out2 + (size_t)i * c->KW, (size_t)c->KW * sizeof(uint32_t));
    return POLAR_B200_OK;
}

int polar_b200_reserve(polar_b200_ctx* c, int max_batch) {
    if (!c || max_batch < 1) return POLAR_B200_E_ARG;
    if (max_batch <= c->max_batch) return POLAR_B200_OK;
    CU_TRY(cudaSetDevice(c->device));
    drain(c);                                             // the old staging buffers may still be in use
    float* nl = nullptr; uint32_t* no = nullptr;
    CU_TRY(cudaMalloc(&nl, (size_t)max_batch * c->N * sizeof(float)));
    cudaError_t e = cudaMalloc(&no, (size_t)max_batch * c->KW * sizeof(uint32_t));
    if (e != cudaSuccess) { cudaFree(nl); return (int)e; }
    cudaFree(c->d_llr_stage); cudaFree(c->d_out_stage);
    c->d_llr_stage = nl; c->d_out_stage = no;
    // staging of the double / probability-domain host entry points is sized by max_batch too: reallocated on next use
    cudaFree(c->d_llr64_stage); c->d_llr64_stage = nullptr;
    cudaFree(c->d_prob_stage); c->d_prob_stage = nullptr;
    c->max_batch = max_batch;
    return POLAR_B200_OK;
}

int polar_b200_set_strict_tau(polar_b200_ctx* c, float tau) {
    if (!c || !(tau >= 0.0f)) return POLAR_B200_E_ARG;
    c->strict_tau = tau;
    return POLAR_B200_OK;
}

int polar_b200_decode_scl_llr_f64(polar_b200_ctx* c, const double* llr, int B, int L,
                                  uint32_t* info_packed, void* cuda_stream) {
    if (!c || !llr || !info_packed || B < 0) return POLAR_B200_E_ARG;
    if (L < 1 || L > c->max_list || L > kMaxList) return POLAR_B200_E_LIST;
    if (B == 0) return POLAR_B200_OK;
    CU_TRY(cudaSetDevice(c->device));
    int rc = call_begin(c, (cudaStream_t)cuda_stream);
    if (rc) return rc;
    rc = decode_any<double>(c, llr, B, L, info_packed, (cudaStream_t)cuda_stream);
    if (rc) return rc;
    return call_end(c, (cudaStream_t)cuda_stream);
}

int polar_b200_decode_scl_llr_f64_host(polar_b200_ctx* c, const double* llr_host, int B, int L,
                                       uint32_t* info_packed_host, void* cuda_stream) {
    if (!c || !llr_host || !info_packed_host || B < 0) return POLAR_B200_E_ARG;
    if (L < 1 || L > c->max_list || L > kMaxList) return POLAR_B200_E_LIST;
    if (B == 0) return POLAR_B200_OK;
    CU_TRY(cudaSetDevice(c->device));
    { const int rr = polar_b200_reserve(c, B); if (rr) return rr; }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = call_begin(c, st);
    if (rc) return rc;
    if (!c->d_llr64_stage) CU_TRY(cudaMalloc(&c->d_llr64_stage, (size_t)c->max_batch * c->N * sizeof(double)));
    CU_TRY(cudaMemcpyAsync(c->d_llr64_stage, llr_host, (size_t)B * c->N * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = decode_any<double>(c, c->d_llr64_stage, B, L, c->d_out_stage, st);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    CU_TRY(cudaMemcpyAsync(info_packed_host, c->d_out_stage, (size_t)B * c->KW * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    c->have_last = false;
    return POLAR_B200_OK;
}

int polar_b200_decode_scl_p1(polar_b200_ctx* c, const double* p1, const double* p0, int B, int L,
                             uint32_t* info_packed, void* cuda_stream) {
    if (!c || !p1 || !p0 || !info_packed || B < 0) return POLAR_B200_E_ARG;
    if (L < 1 || L > c->max_list || L > kMaxList) return POLAR_B200_E_LIST;
    if (B == 0) return POLAR_B200_OK;
    CU_TRY(cudaSetDevice(c->device));
    int rc = call_begin(c, (cudaStream_t)cuda_stream);
    if (rc) return rc;
    rc = decode_prob(c, p0, p1, B, L, info_packed, (cudaStream_t)cuda_stream);
    if (rc) return rc;
    return call_end(c, (cudaStream_t)cuda_stream);
}

int polar_b200_decode_scl_p1_host(polar_b200_ctx* c, const double* p1_host, const double* p0_host, int B, int L,
                                  uint32_t* info_packed_host, void* cuda_stream) {
    if (!c || !p1_host || !p0_host || !info_packed_host || B < 0) return POLAR_B200_E_ARG;
    if (L < 1 || L > c->max_list || L > kMaxList) return POLAR_B200_E_LIST;
    if (B == 0) return POLAR_B200_OK;
    CU_TRY(cudaSetDevice(c->device));
    { const int rr = polar_b200_reserve(c, B); if (rr) return rr; }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t per = (size_t)c->max_batch * c->N;
    int rc = call_begin(c, st);
    if (rc) return rc;
    if (!c->d_prob_stage) CU_TRY(cudaMalloc(&c->d_prob_stage, 2 * per * sizeof(double)));
    double* d_p0 = c->d_prob_stage;
    double* d_p1 = c->d_prob_stage + per;
    CU_TRY(cudaMemcpyAsync(d_p0, p0_host, (size_t)B * c->N * sizeof(double), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_p1, p1_host, (size_t)B * c->N * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = decode_prob(c, d_p0, d_p1, B, L, c->d_out_stage, st);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    CU_TRY(cudaMemcpyAsync(info_packed_host, c->d_out_stage, (size_t)B * c->KW * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    c->have_last = false;
    return POLAR_B200_OK;
}

int polar_b200_count_errors(polar_b200_ctx* c, const uint32_t* info_packed, const uint32_t* truth_packed,
                            int B, uint8_t* block_err, unsigned long long* n_err, void* cuda_stream) {
    if (!c || !info_packed || !truth_packed || B < 0) return POLAR_B200_E_ARG;
    if (B == 0) return POLAR_B200_OK;
    CU_TRY(cudaSetDevice(c->device));
    count_errors_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(info_packed, truth_packed, B, c->KW, block_err, n_err);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return POLAR_B200_OK;
}

int polar_b200_synthesize(polar_b200_ctx* c, unsigned long long seed, long long first_index, int B,
                          const double* ebno_db, int n_ebno, float* llr, uint32_t* truth_packed, void* cuda_stream) {
    if (!c || !ebno_db || !llr || !truth_packed || B < 0 || n_ebno < 1 || n_ebno > 64 || first_index < 0) return POLAR_B200_E_ARG;
    if (c->K > 2048 || c->N > 8192 || c->crc > 32) return POLAR_B200_E_UNSUPPORTED;
    if (B == 0) return POLAR_B200_OK;
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = upload_amplitudes(c, ebno_db, n_ebno, st);
    if (rc) return rc;
    return synth_launch(c, seed, first_index, B, n_ebno, llr, truth_packed, st);
}

int polar_b200_bler_sweep(polar_b200_ctx* c, unsigned long long seed, long long first_index, long long count,
                          const double* ebno_db, int n_ebno, const int* lists, int n_list, int mode,
                          long long* counts, void* cuda_stream) {
    if (!c || !ebno_db || !lists || !counts || count < 0 || first_index < 0 || n_ebno < 1 || n_ebno > 64 || n_list < 1)
        return POLAR_B200_E_ARG;
    if (mode < POLAR_B200_MODE_FP32 || mode > POLAR_B200_MODE_MINSUM) return POLAR_B200_E_ARG;
    for (int i = 0; i < n_list; ++i)
        if (lists[i] < 1 || lists[i] > c->max_list || lists[i] > kMaxList) return POLAR_B200_E_LIST;
    if (c->K > 2048 || c->N > 8192 || c->crc > 32) return POLAR_B200_E_UNSUPPORTED;
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = call_begin(c, st);
    if (rc) return rc;
    long long chunk = env_int("POLAR_B200_SWEEP_CHUNK", 65536);
    if (chunk < 1) chunk = 1;
    if (chunk > count && count > 0) chunk = count;
    const int cells = n_list * n_ebno;
    if ((int)chunk > c->sw_chunk) {
        drain(c);
        cudaFree(c->d_sw_llr); cudaFree(c->d_sw_truth); cudaFree(c->d_sw_out);
        c->d_sw_llr = nullptr; c->d_sw_truth = nullptr; c->d_sw_out = nullptr; c->sw_chunk = 0;
        CU_TRY(cudaMalloc(&c->d_sw_llr, (size_t)chunk * c->N * sizeof(float)));
        CU_TRY(cudaMalloc(&c->d_sw_truth, (size_t)chunk * c->KW * sizeof(uint32_t)));
        CU_TRY(cudaMalloc(&c->d_sw_out, (size_t)chunk * c->KW * sizeof(uint32_t)));
        c->sw_chunk = (int)chunk;
    }
    if (cells > c->sw_cells) {
        drain(c);
        cudaFree(c->d_sw_err); c->d_sw_err = nullptr; c->sw_cells = 0;
        CU_TRY(cudaMalloc(&c->d_sw_err, (size_t)cells * sizeof(unsigned long long)));
        c->sw_cells = cells;
    }
    CU_TRY(cudaMemsetAsync(c->d_sw_err, 0, (size_t)cells * sizeof(unsigned long long), st));
    if ((rc = upload_amplitudes(c, ebno_db, n_ebno, st))) return rc;
    // per chunk: one synthesis launch, then per list size one decode launch with the block-error count fused into its
    // tail (plus strict mode's second pass over the few flagged codewords); nothing but the counters leaves the device
    for (long long first = first_index; first < first_index + count; first += chunk) {
        const int nb = (int)((first_index + count - first) < chunk ? (first_index + count - first) : chunk);
        if ((rc = synth_launch(c, seed, first, nb, n_ebno, c->d_sw_llr, c->d_sw_truth, st))) return rc;
        for (int il = 0; il < n_list; ++il) {
            CountSpec cs;
            cs.truth = c->d_sw_truth; cs.err = c->d_sw_err + (size_t)il * n_ebno; cs.first_index = first; cs.n_ebno = n_ebno;
            if ((rc = decode_mode(c, c->d_sw_llr, nb, lists[il], c->d_sw_out, mode, nullptr, st, -2, 0, true, true, &cs))) return rc;
        }
    }
    std::vector<unsigned long long> err((size_t)cells);
    CU_TRY(cudaMemcpyAsync(err.data(), c->d_sw_err, (size_t)cells * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    c->have_last = false;
    for (int il = 0; il < n_list; ++il)
        for (int ie = 0; ie < n_ebno; ++ie) {
            // codewords g in [first_index, first_index + count) with g % n_ebno == ie
            const long long lo = first_index, hi = first_index + count;
            auto upto = [&](long long x) { return x <= ie ? 0 : (x - ie + n_ebno - 1) / n_ebno; };   // #{g < x : g % n_ebno == ie}
            counts[((size_t)il * n_ebno + ie) * 2 + 0] = (long long)err[(size_t)il * n_ebno + ie];
            counts[((size_t)il * n_ebno + ie) * 2 + 1] = upto(hi) - upto(lo);
        }
    return POLAR_B200_OK;
}

// ---- NCCL: only the (num_err, num_run) counters of a sharded sweep cross GPUs ----
// libnccl is loaded at run time (the copy a host framework already has in the process, else the system one), so this
// library has no link-time dependency on it.
struct polar_b200_comm {
    ncclComm_t comm = nullptr;
    int device = 0, nranks = 1, rank = 0;
    long long* d_buf = nullptr;
    size_t cap = 0;
    cudaStream_t st = nullptr;
};

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    bool ok = false;
};
const NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        // the copy already in the process (a host framework's) wins; POLAR_B200_NCCL_LIB names another one explicitly
        const char* forced = getenv("POLAR_B200_NCCL_LIB");
        if (forced && *forced) a.lib = dlopen(forced, RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
        a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.lib, "ncclCommInitAll");
        a.AllReduce = (decltype(a.AllReduce))dlsym(a.lib, "ncclAllReduce");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
        a.GroupStart = (decltype(a.GroupStart))dlsym(a.lib, "ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.lib, "ncclGroupEnd");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.AllReduce && a.CommDestroy && a.GroupStart && a.GroupEnd;
        return a;
    }();
    return api;
}
int comm_buffer(polar_b200_comm* m, size_t n) {
    if (n <= m->cap) return 0;
    if (m->d_buf) cudaFree(m->d_buf);
    m->d_buf = nullptr; m->cap = 0;
    CU_TRY(cudaMalloc(&m->d_buf, n * sizeof(long long)));
    m->cap = n;
    return 0;
}
}  // namespace

int polar_b200_comm_unique_id(unsigned char* id128) {
    if (!id128) return POLAR_B200_E_ARG;
    const NcclApi& api = nccl_api();
    if (!api.ok) return POLAR_B200_E_NONCCL;
    ncclUniqueId id;
    if (api.GetUniqueId(&id) != ncclSuccess) return POLAR_B200_E_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return POLAR_B200_OK;
}

int polar_b200_comm_init_rank(polar_b200_comm** out, int device, int nranks, int rank, const unsigned char* id128) {
    if (!out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return POLAR_B200_E_ARG;
    *out = nullptr;
    const NcclApi& api = nccl_api();
    if (!api.ok) return POLAR_B200_E_NONCCL;
    CU_TRY(cudaSetDevice(device));
    polar_b200_comm* m = new (std::nothrow) polar_b200_comm;
    if (!m) return POLAR_B200_E_ARG;
    m->device = device; m->nranks = nranks; m->rank = rank;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    if (api.CommInitRank(&m->comm, nranks, id, rank) != ncclSuccess) { delete m; return POLAR_B200_E_NCCL; }
    cudaError_t e = cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { api.CommDestroy(m->comm); delete m; return (int)e; }
    *out = m;
    return POLAR_B200_OK;
}

int polar_b200_comm_init_all(polar_b200_comm** out, int ndev, const int* devices) {
    if (!out || ndev < 1 || !devices) return POLAR_B200_E_ARG;
    const NcclApi& api = nccl_api();
    if (!api.ok) return POLAR_B200_E_NONCCL;
    std::vector<ncclComm_t> comms((size_t)ndev);
    if (api.CommInitAll(comms.data(), ndev, devices) != ncclSuccess) return POLAR_B200_E_NCCL;
    for (int i = 0; i < ndev; ++i) {
        polar_b200_comm* m = new (std::nothrow) polar_b200_comm;
        if (!m) return POLAR_B200_E_ARG;
        m->comm = comms[i]; m->device = devices[i]; m->nranks = ndev; m->rank = i;
        CU_TRY(cudaSetDevice(devices[i]));
        CU_TRY(cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking));
        out[i] = m;
    }
    return POLAR_B200_OK;
}

int polar_b200_comm_allreduce_i64(polar_b200_comm* m, long long* values, int n) {
    return polar_b200_comm_allreduce_i64_group(&m, 1, &values, n);
}

// one call for all communicators this process owns (ncclGroupStart/End): values[i] is communicator i's host vector
int polar_b200_comm_allreduce_i64_group(polar_b200_comm** ms, int ncomm, long long** values, int n) {
    if (!ms || !values || ncomm < 1 || n < 1) return POLAR_B200_E_ARG;
    const NcclApi& api = nccl_api();
    if (!api.ok) return POLAR_B200_E_NONCCL;
    for (int i = 0; i < ncomm; ++i) {
        if (!ms[i] || !values[i]) return POLAR_B200_E_ARG;
        CU_TRY(cudaSetDevice(ms[i]->device));
        int rc = comm_buffer(ms[i], (size_t)n);
        if (rc) return rc;
        CU_TRY(cudaMemcpyAsync(ms[i]->d_buf, values[i], (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, ms[i]->st));
    }
    bool bad = api.GroupStart() != ncclSuccess;
    for (int i = 0; i < ncomm && !bad; ++i)
        bad = api.AllReduce(ms[i]->d_buf, ms[i]->d_buf, (size_t)n, ncclInt64, ncclSum, ms[i]->comm, ms[i]->st) != ncclSuccess;
    if (api.GroupEnd() != ncclSuccess || bad) return POLAR_B200_E_NCCL;
    for (int i = 0; i < ncomm; ++i) {
        CU_TRY(cudaSetDevice(ms[i]->device));
        CU_TRY(cudaMemcpyAsync(values[i], ms[i]->d_buf, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, ms[i]->st));
        CU_TRY(cudaStreamSynchronize(ms[i]->st));
    }
    return POLAR_B200_OK;
}

int polar_b200_comm_destroy(polar_b200_comm* m) {
    if (!m) return POLAR_B200_E_ARG;
    cudaSetDevice(m->device);
    if (m->comm && nccl_api().ok) nccl_api().CommDestroy(m->comm);
    cudaFree(m->d_buf);
    if (m->st) cudaStreamDestroy(m->st);
    delete m;
    return POLAR_B200_OK;
}

long long polar_b200_get_info(polar_b200_ctx* c, int key) {
    if (!c) return -1;
    switch (key) {
        case POLAR_B200_INFO_KERNEL_LAUNCHES: return c->launches;
        case POLAR_B200_INFO_SM_COUNT: return c->sm_count;
        case POLAR_B200_INFO_WARPS_PER_BLOCK: return c->last_wpb;
        case POLAR_B200_INFO_BLOCKS: return c->last_blocks;
        case POLAR_B200_INFO_SMEM_BYTES: return c->last_smem;
        case POLAR_B200_INFO_SCRATCH_BYTES: return (long long)c->scratch_bytes;
        case POLAR_B200_INFO_KERNEL_KIND: return c->last_kernel;
        case POLAR_B200_INFO_HOST_CHUNKS: return c->last_chunks;
        case POLAR_B200_INFO_LAST_FLAGGED:
            if (c->flagged_pending) {
                int nf = 0;
                cudaSetDevice(c->device);
                drain(c);
                if (cudaMemcpy(&nf, c->d_flag_count, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
                c->last_flagged = nf; c->flagged_pending = false;
            }
            return c->last_flagged;
        default: return -1;
    }
}

}  // extern "C"
