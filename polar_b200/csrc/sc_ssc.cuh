// sc_ssc.cuh -- plain SC (list size 1), the first pass of STRICT mode and FP32 mode, N = 2^8 .. 2^12: a second mapping of the same
// decoder (the reference's decode_scl_llr with list size 1, PolarC/PolarCode.cpp:130-190, 422-473, 489-553) that keeps
// every byte of a codeword's state on the SM.
//
// What is different from scl_fast.cuh's list-1 variants (32 codewords per warp, lane = codeword, state in an HBM/L2
// scratch, 103 KB of DRAM traffic per codeword):
//
//   * FOUR LANES PER CODEWORD, EIGHT CODEWORDS PER WARP. A node of the decoding tree with M LLRs is split into four
//     quarters; lane q owns entries [q M/4, (q+1) M/4). The reference pairs entries (2b, 2b+1) of a layer into entry b
//     of the next one (:431-451), so a quarter maps onto the same quarter of both children: every f / g / partial-sum
//     step down to nodes of 8 entries touches lane-private data only, and no synchronisation of any kind is needed.
//     Nodes of 32 entries and below are finished in registers (sub_node), those of 4 entries (one per lane) with shuffles.
//   * Layer 0 is the channel row itself and layer 1 is never stored: an entry of layer 2 is computed straight from one
//     float4 of the channel row (its two layer-1 parents on the fly; the row is streamed four times per codeword, out of
//     L2). Layer 2 lives in TENSOR MEMORY (tcgen05.st/ld 32x32b: N/16 lane-private columns per warp), layers 3.. in shared
//     memory as [entry/4][lane][4] (one LDS.128 feeds two nodes, one STS.128 stores four), partial sums as packed
//     words [word][lane]. 21 KB of shared memory per warp at N = 2048: ten warps = 80 codewords per SM (with layer 1
//     stored in tensor memory and layer 2 in shared memory it was six warps, and the kernel is latency-bound: measured
//     31 % issue utilisation at 1.5 warps per scheduler, profiles/r02_ssc_v1_ncu_summary.txt).
//   * THE TREE IS PRUNED. With one path there is no path metric, so a subtree whose leaves are all frozen needs no LLR
//     at all (its bits are 0), and a subtree without frozen leaves is decided by the signs of its root LLRs: successive
//     cancellation below such a node reproduces exactly those signs -- sign f(a,b) = sign a * sign b, and g adds
//     magnitudes once the left half agrees with the signs (Alamdar-Yazdi & Kschischang's rate-0 / rate-1 nodes). That
//     holds in any arithmetic that gets the signs right and breaks only where an entry is zero or so small that its
//     sign is in doubt -- which is what STRICT mode's margin is for: the smallest |LLR| over every entry that decided a
//     bit is the codeword's margin, and a codeword below tau is decoded again by the double-precision second pass,
//     leaf by leaf. (FP32 mode, which has no second pass, lists the codewords with an exactly-zero deciding LLR -- ties,
//     none on a real channel -- the same way and hands them to the leaf-by-leaf generic fp32 kernel.)
//     At N = 2048, K = 1024: 6 370 check nodes instead of 11 264 (+ 1 024 for the layer-1 entries that are computed
//     twice), 328 visited nodes instead of 4 095.
//   * The walk over the tree is a per-code SCHEDULE built once on the host (build_schedule): a few hundred 32-bit
//     operations that every lane of every warp interprets in lock step (no divergence: the frozen set is the same for
//     every codeword). The warps of a block start every round of 8 codewords together (block barrier): they then share
//     the instruction caches instead of evicting each other's code.
//   * Measured on B200: N = 2048, K = 1024: 41.7 / 43.5 M codewords/s at batch 65 536 / 262 144
//     against 27.1 / 25.4 M for the leaf-by-leaf kernel; 13.5 KB of DRAM traffic per codeword (algorithmic: 8.3 KB).
//
// Arithmetic: fast::f_rule2 / g_top2 (scl_fast.cuh), i.e. the same fp32 check / variable node as every other fp32 kernel.
#pragma once
#include <vector>
#include "polar_dev.cuh"
#include "scl_fast.cuh"

namespace ssc {

enum : uint32_t { OP_END = 0, OP_F = 1, OP_G = 2, OP_G0 = 3, OP_C = 4, OP_R0 = 5, OP_R1 = 6, OP_SUB = 7 };
// op word: bits 0-2 type, 3-6 m (log2 of the node size: the parent for F / G / G0 / C, the node itself otherwise),
// 7 side (result slot of the node: 0 = left child of its parent, 1 = right), 8-9 for the operations that read layer 1
// (F / G / G0 / R1 of the root's children): which half of the codeword -- 1 = left (layer 1 = f of the channel pairs),
// 2 = right (g with the left half's partial sums), 3 = right with an all-frozen left half (g = a + b); 12 C: left child is
// all frozen (its partial sums are zero and were never written), 13 C: right child is all frozen. The root itself has no
// F / G: its children compute their layer-1 entries on the fly. SUB (a node of 32 entries that has frozen and
// unfrozen leaves, finished in registers) is followed by one word holding the frozen pattern of its 32 leaves.
constexpr uint32_t kSide = 1u << 7, kLeftZero = 1u << 12, kRightZero = 1u << 13;
constexpr int kNLogMin = 8, kNLogMax = 12;

// ---- host: the schedule of one code ----
// half: 0 below layer 1; for the root's children 1 / 2 / 3 as above; -1 for the root
inline void emit_node(const uint8_t* frozen, int lo, int m, uint32_t side, std::vector<uint32_t>& ops, int half) {
    const int M = 1 << m;
    const uint32_t hb = half > 0 ? (uint32_t)half << 8 : 0u;
    int nf = 0;
    for (int i = 0; i < M; ++i) nf += frozen[lo + i] ? 1 : 0;
    if (nf == M) { ops.push_back(OP_R0 | (uint32_t)m << 3 | side); return; }
    if (nf == 0) { ops.push_back(OP_R1 | (uint32_t)m << 3 | side | hb); return; }
    if (m == 5) {
        uint32_t mask = 0;
        for (int i = 0; i < 32; ++i) mask |= (frozen[lo + i] ? 1u : 0u) << i;
        ops.push_back(OP_SUB | 5u << 3 | side);
        ops.push_back(mask);
        return;
    }
    const int h = M / 2;
    int nl = 0, nr = 0;
    for (int i = 0; i < h; ++i) { nl += frozen[lo + i] ? 1 : 0; nr += frozen[lo + h + i] ? 1 : 0; }
    const bool left_zero = nl == h, right_zero = nr == h;
    if (!left_zero) {
        if (half >= 0) ops.push_back(OP_F | (uint32_t)m << 3 | hb);
        emit_node(frozen, lo, m - 1, 0u, ops, half < 0 ? 1 : 0);
    }
    if (!right_zero) {
        if (half >= 0) ops.push_back((left_zero ? OP_G0 : OP_G) | (uint32_t)m << 3 | hb);
        emit_node(frozen, lo + h, m - 1, kSide, ops, half < 0 ? (left_zero ? 3 : 2) : 0);
    }
    ops.push_back(OP_C | (uint32_t)m << 3 | side | (left_zero ? kLeftZero : 0u) | (right_zero ? kRightZero : 0u));
}

// frozen: one byte per decoding position. Returns false when this kernel does not serve the code (block length out of
// range, or a code without frozen / without unfrozen positions).
inline bool build_schedule(int n, const uint8_t* frozen, std::vector<uint32_t>& ops) {
    ops.clear();
    if (n < kNLogMin || n > kNLogMax) return false;
    const int N = 1 << n;
    int nf = 0;
    for (int i = 0; i < N; ++i) nf += frozen[i] ? 1 : 0;
    if (nf == 0 || nf == N) return false;
    emit_node(frozen, 0, n, 0u, ops, -1);
    for (int i = 0; i < 3; ++i) ops.push_back(OP_END);          // the interpreter reads two words ahead
    return true;
}

// where output bit j comes from: u-hat[order[j]] = y[bitrev(order[j])], y = the re-encoded codeword after n butterfly
// stages, held as (quarter, word, bit) -- see the kernel's tail. Encoded (word << 7) | (quarter << 5) | bit.
inline void build_positions(int n, const uint16_t* order, int K, std::vector<uint16_t>& pos) {
    pos.assign((size_t)((K + 31) / 32) * 32, 0);
    for (int j = 0; j < K; ++j) {
        unsigned p = 0;
        for (int b = 0; b < n; ++b) if (order[j] & (1u << b)) p |= 1u << (n - 1 - b);
        const unsigned q = p >> (n - 2), local = p & ((1u << (n - 2)) - 1u);
        pos[j] = (uint16_t)(((local >> 5) << 7) | (q << 5) | (local & 31u));
    }
}

// ---- layout of one warp's shared memory ----
struct Layout {
    int n;
    int x_rows4;          // float4 rows ([lane][4] floats, 512 bytes) of the LLR layers 3 .. n-5
    int s_rows;           // word rows (128 bytes) of the partial-sum slots of layers 1 .. n-2
    int xrow[16];         // word row of layer lam's left slot (the right slot follows it)
    int bytes;            // per warp
    int warps;            // per block (= per SM)
};
inline Layout make_layout(int n) {
    Layout l{};
    l.n = n;
    const int N = 1 << n;
    l.x_rows4 = N / 64 < 2 ? 2 : N / 64;   // layers 3 .. n-5 (N/32 .. 8 entries per lane): N/64 - 2 rows
    int row = 0;
    for (int lam = 1; lam <= n - 2; ++lam) {
        l.xrow[lam] = row;
        const int cnt = 1 << (n - lam - 2);
        row += 2 * (cnt >= 32 ? cnt / 32 : 1);
    }
    l.s_rows = row;
    l.bytes = l.x_rows4 * 512 + l.s_rows * 128;
    int w = (227 * 1024 - 1024) / l.bytes;
    const int tm_cols = N / 16;                      // tensor-memory columns of layer 2 per warp; 512 per lane quadrant
    const int tm_warps = 4 * (512 / tm_cols);
    if (w > tm_warps) w = tm_warps;
    if (w > 16) w = 16;
    l.warps = w;
    return l;
}

struct Args {
    fastcommon::Args a;
    const uint32_t* sched;
    const uint16_t* pos;
    Layout lay;
    int sync_rounds;      // block barrier at the start of every round
};

#ifdef __CUDACC__
using POLAR_FAST_NS::node4;
using POLAR_FAST_NS::tm_ld4;
using POLAR_FAST_NS::tm_st4;
using POLAR_FAST_NS::tm_wait_st;

// tcgen05.wait::ld with the loaded registers as operands: nothing that uses them can be scheduled above the wait
__device__ __forceinline__ void tm_wait_ld16(float (&v)[4][4]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0][0]), "+f"(v[0][1]), "+f"(v[0][2]), "+f"(v[0][3]), "+f"(v[1][0]), "+f"(v[1][1]), "+f"(v[1][2]), "+f"(v[1][3]),
                   "+f"(v[2][0]), "+f"(v[2][1]), "+f"(v[2][2]), "+f"(v[2][3]), "+f"(v[3][0]), "+f"(v[3][1]), "+f"(v[3][2]), "+f"(v[3][3])
                 :: "memory");
}
__device__ __forceinline__ void tm_wait_ld4(float (&v)[4]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]) :: "memory");
}

// bits of a (even positions) and b (odd positions), 16 + 16 -> 32
__device__ __forceinline__ uint32_t spread16(uint32_t x) {
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
__device__ __forceinline__ uint32_t interleave16(uint32_t a, uint32_t b) { return spread16(a & 0xFFFFu) | (spread16(b & 0xFFFFu) << 1); }

// the same for H + H -> 2H bits, H <= 4 (inputs hold nothing above bit H-1)
template <int H>
__device__ __forceinline__ uint32_t interleave_small(uint32_t a, uint32_t b) {
    uint32_t x = a | (b << 16);
    if constexpr (H > 2) x = (x | (x << 2)) & 0x33333333u;
    if constexpr (H > 1) x = (x | (x << 1)) & 0x55555555u;
    return (x & 0xFFFFu) | ((x >> 16) << 1);
}

__device__ __forceinline__ void note(uint32_t& mg, float x) { mg = min(mg, __float_as_uint(fabsf(x))); }

// float offset (within the warp's LLR area, before adding 4 * lane) of layer lam, 3 <= lam <= n-5: [entry / 4][lane][4];
// sum_{mu=3}^{lam-1} cnt(mu)/4 = N/64 - cnt(lam)/2 rows of 128 floats lie before it
__device__ __forceinline__ int x_base(int n, int lam) { return ((1 << (n - 6)) - (1 << (n - lam - 3))) * 128; }

// four children from eight consecutive parent entries (two float4): child e = f or g of entries (2e, 2e+1)
template <bool ISG>
__device__ __forceinline__ void four(const float4& v0, const float4& v1, uint32_t word, int pos, float (&y)[4]) {
    const float pa[4] = {v0.x, v0.z, v1.x, v1.z}, pb[4] = {v0.y, v0.w, v1.y, v1.w};
    node4<ISG>(pa, pb, word, pos, y);
}

// the two layer-1 entries that one float4 of the channel row yields (PolarCode.cpp:431-451 applied to layer 0): left half
// of the codeword f, right half g with the left half's partial sums (bits k, k+1 of ab; ab = 0 when that half is all frozen)
__device__ __forceinline__ void layer1_pair(int half, const float4& v, uint32_t ab, int k, float& e0, float& e1) {
    if (half == 1) POLAR_FAST_NS::f_rule2(v.x, v.y, v.z, v.w, e0, e1);
    else POLAR_FAST_NS::g_top2(v.x, v.y, ab << (31 - k), v.z, v.w, ab << (30 - k), e0, e1);
}

// ---- child LLRs of a node: entry j = f or g of the parent's entries (2j, 2j+1) (PolarCode.cpp:431-451); cnt = entries
// per lane of the child (>= 8). has_bits = false: every partial sum of the left child is 0 (it is all frozen). Loads run
// one stage ahead of the arithmetic: with ten warps per SM there is little else to hide their latency behind. ----
template <bool ISG, int ST>
__device__ __forceinline__ void child_llrs(int n, int lam, int cnt, bool has_bits, int half, const float4* chan, uint32_t tm,
                                           float* X, const uint32_t* sw, const uint32_t* sa) {
    uint32_t word = 0;
    if (lam == 1) {
        // channel row -> (layer 1 on the fly) -> layer 2 in tensor memory: child j comes from float4 j of the lane's quarter;
        // ST children per stage, the next stage's loads issued before this stage's arithmetic
        float4 buf[ST];
#pragma unroll
        for (int k = 0; k < ST; ++k) buf[k] = __ldg(chan + k);
        uint32_t aw = 0;
#pragma unroll 1
        for (int j = 0; j < cnt; j += ST) {
            float4 nx[ST];
            const bool more = j + ST < cnt;
            if (more) {
#pragma unroll
                for (int k = 0; k < ST; ++k) nx[k] = __ldg(chan + j + ST + k);
            }
            if (ISG && has_bits && (j & 31) == 0) word = sw[(j >> 5) * 32];
            if (half == 2 && (j & 15) == 0) aw = sa[(j >> 4) * 32];      // layer-1 entries 2j .. 2j+31 of this lane
#pragma unroll
            for (int h = 0; h < ST; h += 4) {
                const uint32_t ab = aw >> ((2 * (j + h)) & 31);
                float pa[4], pb[4], y[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) layer1_pair(half, buf[h + e], ab, 2 * e, pa[e], pb[e]);
                node4<ISG>(pa, pb, word, (j + h) & 31, y);
                tm_st4(tm + j + h, y);
            }
            if (more) {
#pragma unroll
                for (int k = 0; k < ST; ++k) buf[k] = nx[k];
            }
        }
        tm_wait_st();
    } else if (lam == 2) {
        // tensor memory -> shared memory (layer 3), 8 children (16 columns) per stage, two register sets in turn
        float* dst = X + x_base(n, 3);
        float va[4][4], vb[4][4];
        auto load = [&](float (&v)[4][4], int j) {
#pragma unroll
            for (int k = 0; k < 4; ++k) tm_ld4(tm + 2 * j + 4 * k, v[k]);
        };
        auto work = [&](float (&v)[4][4], int j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 v0 = make_float4(v[2 * h][0], v[2 * h][1], v[2 * h][2], v[2 * h][3]);
                const float4 v1 = make_float4(v[2 * h + 1][0], v[2 * h + 1][1], v[2 * h + 1][2], v[2 * h + 1][3]);
                float y[4];
                four<ISG>(v0, v1, word, (j + 4 * h) & 31, y);
                *reinterpret_cast<float4*>(dst + ((j >> 2) + h) * 128) = make_float4(y[0], y[1], y[2], y[3]);
            }
        };
        load(va, 0);
#pragma unroll 1
        for (int j = 0; j < cnt; j += 16) {
            if (ISG && has_bits && (j & 31) == 0) word = sw[(j >> 5) * 32];
            tm_wait_ld16(va);
            const bool second = j + 8 < cnt;
            if (second) load(vb, j + 8);
            work(va, j);
            if (second) {
                tm_wait_ld16(vb);
                if (j + 16 < cnt) load(va, j + 16);
                work(vb, j + 8);
            }
        }
    } else {
        const float* src = X + x_base(n, lam);
        float* dst = X + x_base(n, lam + 1);
#pragma unroll 1
        for (int j = 0; j < cnt; j += 8) {
            const float4 v0 = *reinterpret_cast<const float4*>(src + (j >> 1) * 128);
            const float4 v1 = *reinterpret_cast<const float4*>(src + ((j >> 1) + 1) * 128);
            const float4 v2 = *reinterpret_cast<const float4*>(src + ((j >> 1) + 2) * 128);
            const float4 v3 = *reinterpret_cast<const float4*>(src + ((j >> 1) + 3) * 128);
            if (ISG && has_bits && (j & 31) == 0) word = sw[(j >> 5) * 32];
            float y0[4], y1[4];
            four<ISG>(v0, v1, word, j & 31, y0);
            four<ISG>(v2, v3, word, (j + 4) & 31, y1);
            *reinterpret_cast<float4*>(dst + (j >> 2) * 128) = make_float4(y0[0], y0[1], y0[2], y0[3]);
            *reinterpret_cast<float4*>(dst + ((j >> 2) + 1) * 128) = make_float4(y1[0], y1[1], y1[2], y1[3]);
        }
    }
}

// ---- a node of four entries, one per lane (lane q holds entry q), some of its leaves frozen (fm: bit i = leaf i): leaf by
// leaf exactly as the reference does, with shuffles between the four lanes of the codeword. Returns this lane's partial sum. ----
__device__ __forceinline__ uint32_t leaf4(float al, uint32_t fm, uint32_t& mg, int q) {
    const float ap = __shfl_xor_sync(FULL_MASK, al, 1);
    const float a0 = (q & 1) ? ap : al, a1 = (q & 1) ? al : ap;        // entries 2b, 2b+1 of this lane's pair b = q >> 1
    uint32_t xl = 0;                                                   // left child's partial sum of pair b
    if ((fm & 3u) != 3u) {
        const float l_mine = f_rule(a0, a1);                           // left child, entry b
        const float l_other = __shfl_xor_sync(FULL_MASK, l_mine, 2);
        const float l0 = (q & 2) ? l_other : l_mine, l1 = (q & 2) ? l_mine : l_other;
        uint32_t u0 = 0, u1 = 0;
        if (!(fm & 1u)) { const float t = f_rule(l0, l1); note(mg, t); u0 = t < 0.0f ? 1u : 0u; }
        if (!(fm & 2u)) { const float t = POLAR_FAST_NS::g_rule(l0, l1, u0); note(mg, t); u1 = t < 0.0f ? 1u : 0u; }
        xl = (q & 2) ? u1 : (u0 ^ u1);
    }
    uint32_t xr = 0;
    if ((fm & 12u) != 12u) {
        const float r_mine = POLAR_FAST_NS::g_rule(a0, a1, xl);        // right child, entry b
        const float r_other = __shfl_xor_sync(FULL_MASK, r_mine, 2);
        const float r0 = (q & 2) ? r_other : r_mine, r1 = (q & 2) ? r_mine : r_other;
        uint32_t u2 = 0, u3 = 0;
        if (!(fm & 4u)) { const float t = f_rule(r0, r1); note(mg, t); u2 = t < 0.0f ? 1u : 0u; }
        if (!(fm & 8u)) { const float t = POLAR_FAST_NS::g_rule(r0, r1, u2); note(mg, t); u3 = t < 0.0f ? 1u : 0u; }
        xr = (q & 2) ? u3 : (u2 ^ u3);
    }
    return (q & 1) ? xr : (xl ^ xr);                                   // entry 2b = left ^ right, entry 2b+1 = right
}

// ---- a node of 4 E entries (E = 8, 4, 2, 1 per lane) finished in registers: the same walk as the schedule's (rate-0 /
// rate-1 shortcuts, children, combine), with the frozen pattern fm of its 4 E leaves deciding every branch -- uniformly
// for the whole warp. Returns this lane's E partial-sum bits. ----
template <int E>
__device__ __forceinline__ uint32_t sub_node(const float (&a)[E], uint32_t fm, uint32_t& mg, int q) {
    constexpr int M = 4 * E;
    constexpr uint32_t full = M == 32 ? 0xFFFFFFFFu : ((1u << (M & 31)) - 1u);
    if (fm == full) return 0u;
    if (fm == 0u) {
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < E; ++j) { note(mg, a[j]); bits |= (a[j] < 0.0f ? 1u : 0u) << j; }
        return bits;
    }
    if constexpr (E == 1) {
        return leaf4(a[0], fm, mg, q);
    } else {
        constexpr int H = E / 2;
        constexpr uint32_t half = (1u << (M / 2)) - 1u;
        const uint32_t fl = fm & half, fr = fm >> (M / 2);
        uint32_t xl = 0, xr = 0;
        if (fl != half) {
            float l[H];
            if constexpr (H >= 2) {
#pragma unroll
                for (int j = 0; j < H; j += 2) POLAR_FAST_NS::f_rule2(a[2 * j], a[2 * j + 1], a[2 * j + 2], a[2 * j + 3], l[j], l[j + 1]);
            } else {
                l[0] = f_rule(a[0], a[1]);
            }
            xl = sub_node<H>(l, fl, mg, q);
        }
        if (fr != half) {
            float r[H];
#pragma unroll
            for (int j = 0; j < H; ++j) r[j] = POLAR_FAST_NS::g_rule(a[2 * j], a[2 * j + 1], (xl >> j) & 1u);
            xr = sub_node<H>(r, fr, mg, q);
        }
        return interleave_small<H>(xl ^ xr, xr);
    }
}

// MAXW: most warps per block this build is launched with (the fewer, the more registers per thread); ST: channel float4s
// per prefetch stage of the top operations (4 or 8)
template <int MAXW, int ST>
__global__ void __launch_bounds__(MAXW * 32, 1) sc_ssc_kernel(const Args A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t tm_base;
    const fastcommon::Args& a = A.a;
    const int n = A.lay.n, N = 1 << n;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    const int q = lane & 3, cwl = lane >> 2;
    unsigned char* my = smem_raw + (size_t)A.lay.bytes * wib;
    float* X = reinterpret_cast<float*>(my) + 4 * lane;                 // LLR layers: entry j of a layer at base + (j/4)*128 + (j&3)
    uint32_t* S = reinterpret_cast<uint32_t*>(my + A.lay.x_rows4 * 512) + lane;   // partial-sum words: row r at S[r * 32]
    uint32_t* R = reinterpret_cast<uint32_t*>(my) + lane;               // root partial sums / u-hat words (LLR area, dead by then)
    if (wib == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&tm_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tm_base + ((uint32_t)((wib & 3) * 32) << 16) + (uint32_t)((wib >> 2) * (N >> 4));

    const int KW = (a.K + 31) >> 5;
    const int groups = (a.B + 7) >> 3;
    const int total_warps = gridDim.x * WPB;
    // Every warp runs the same number of rounds and the warps of a block start each round together: they then walk through
    // the same code at about the same time and share the instruction caches (the walk is identical for every codeword).
    // Unsynchronised, ten warps drift apart over the rounds and instruction fetch becomes the largest stall (measured:
    // 26 M codewords/s at 22 rounds against 33 M at 6).
    const int rounds = (groups + total_warps - 1) / total_warps;
    for (int rnd = 0; rnd < rounds; ++rnd) {
        const int grp = rnd * total_warps + blockIdx.x * WPB + wib;
        if (A.sync_rounds) __syncthreads();
        if (grp >= groups) continue;
        const int cw = grp * 8 + cwl;
        const bool valid = cw < a.B;
        const float4* chan = reinterpret_cast<const float4*>(a.llr + (size_t)(valid ? cw : a.B - 1) * N + (size_t)q * (N >> 2));
        uint32_t mg = 0x7F800000u;                   // smallest |LLR| that decided a bit (float bits; +inf = none yet)
        int pc = 0;
        uint32_t op = __ldg(A.sched);
#pragma unroll 1
        for (;;) {
            const uint32_t n1 = __ldg(A.sched + pc + 1), n2 = __ldg(A.sched + pc + 2);   // the schedule ends in three END words
            const uint32_t type = op & 7u;
            if (type == OP_END) break;
            const int m = (int)((op >> 3) & 15u);
            const int lam = n - m;                   // layer of the node (F / G / G0 / C: of the parent)
            const int side = (int)((op >> 7) & 1u);
            if (type <= OP_G0) {
                const int cnt = 1 << (m - 3);
                const uint32_t* sw = S + A.lay.xrow[lam + 1] * 32;      // left child's partial sums
                const uint32_t* sa = S + A.lay.xrow[1] * 32;            // partial sums of the left half of the codeword
                const int half = (int)((op >> 8) & 3u);
                if (type == OP_F) child_llrs<false, ST>(n, lam, cnt, false, half, chan, tm, X, sw, sa);
                else child_llrs<true, ST>(n, lam, cnt, type == OP_G, half, chan, tm, X, sw, sa);
            } else if (type == OP_C) {
                // partial sums of the parent: entry 2j = left[j] ^ right[j], entry 2j+1 = right[j] (:463-468)
                const int cl = 1 << (m - 3);         // bits per lane of each child (>= 8)
                const uint32_t* sl = S + A.lay.xrow[lam + 1] * 32;
                const int wc = cl >= 32 ? cl >> 5 : 1;
                const uint32_t* sr = sl + wc * 32;
                uint32_t* out = (lam == 0) ? R : S + (A.lay.xrow[lam] + side * (cl >= 16 ? cl >> 4 : 1)) * 32;
                const bool lz = (op & kLeftZero) != 0, rz = (op & kRightZero) != 0;
                // the root's partial sums go to the start of the LLR area as [word][lane] rows, which cuts across the other
                // lanes' [lane][4] LLR slots: every lane must be done reading LLRs first
                if (lam == 0) __syncwarp();
#pragma unroll 1
                for (int w = 0; w < wc; ++w) {
                    const uint32_t r = rz ? 0u : sr[w * 32];
                    const uint32_t l = (lz ? 0u : sl[w * 32]) ^ r;
                    if (cl >= 32) {
                        out[(2 * w) * 32] = interleave16(l, r);
                        out[(2 * w + 1) * 32] = interleave16(l >> 16, r >> 16);
                    } else {
                        out[0] = interleave16(l, r);
                    }
                }
            } else if (type == OP_R0) {
                const int cnt = 1 << (m - 2);
                const int wc = cnt >= 32 ? cnt >> 5 : 1;
                uint32_t* out = S + (A.lay.xrow[lam] + side * wc) * 32;
                for (int w = 0; w < wc; ++w) out[w * 32] = 0u;
            } else if (type == OP_R1) {
                // no frozen leaf below: the bits are the signs of the node's LLRs (cnt >= 8 entries per lane)
                const int cnt = 1 << (m - 2);
                const int wc = cnt >= 32 ? cnt >> 5 : 1;
                uint32_t* out = S + (A.lay.xrow[lam] + side * wc) * 32;
                uint32_t word = 0;
                if (lam == 1) {
                    // a whole half of the codeword without a frozen leaf: its layer-1 entries on the fly
                    const int half = (int)((op >> 8) & 3u);
                    const uint32_t* sa = S + A.lay.xrow[1] * 32;
                    uint32_t aw = 0;
#pragma unroll 1
                    for (int j = 0; j < cnt; j += 4) {
                        const float4 c0 = __ldg(chan + (j >> 1)), c1 = __ldg(chan + (j >> 1) + 1);
                        if (half == 2 && (j & 31) == 0) aw = sa[(j >> 5) * 32];
                        const uint32_t ab = aw >> (j & 31);
                        float v[4];
                        layer1_pair(half, c0, ab, 0, v[0], v[1]);
                        layer1_pair(half, c1, ab, 2, v[2], v[3]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) { note(mg, v[e]); word |= (v[e] < 0.0f ? 1u : 0u) << ((j + e) & 31); }
                        if (((j + 4) & 31) == 0) { out[(j >> 5) * 32] = word; word = 0; }
                    }
                } else if (lam == 2) {
#pragma unroll 1
                    for (int j = 0; j < cnt; j += 4) {
                        float v[4];
                        tm_ld4(tm + j, v);
                        tm_wait_ld4(v);
#pragma unroll
                        for (int e = 0; e < 4; ++e) { note(mg, v[e]); word |= (v[e] < 0.0f ? 1u : 0u) << ((j + e) & 31); }
                        if (((j + 4) & 31) == 0) { out[(j >> 5) * 32] = word; word = 0; }
                    }
                } else {
                    const float* src = X + x_base(n, lam);
#pragma unroll 1
                    for (int j = 0; j < cnt; j += 4) {
                        const float4 v = *reinterpret_cast<const float4*>(src + (j >> 2) * 128);
                        note(mg, v.x); note(mg, v.y); note(mg, v.z); note(mg, v.w);
                        word |= ((v.x < 0.0f ? 1u : 0u) | (v.y < 0.0f ? 2u : 0u) | (v.z < 0.0f ? 4u : 0u) | (v.w < 0.0f ? 8u : 0u)) << (j & 31);
                        if (((j + 4) & 31) == 0 || j + 4 == cnt) { out[(j >> 5) * 32] = word; word = 0; }
                    }
                }
            } else {
                // OP_SUB: a node of 32 entries with frozen and unfrozen leaves, finished in registers; its frozen pattern is
                // the next schedule word
                const float* src = X + x_base(n, n - 5);
                const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 128);
                const float e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                S[(A.lay.xrow[n - 5] + side) * 32] = sub_node<8>(e, n1, mg, q);
                op = n2; pc += 2;
                continue;
            }
            op = n1; ++pc;
        }

        // ---- u-hat: n butterfly stages over the re-encoded codeword (lane q holds positions [q N/4, (q+1) N/4) as
        // N/128 words); u-hat[i] is then bit bitrev(i) of the result (the encoder's last step, PolarCode.cpp:76-87) ----
        const int NWL = N >> 7;
        for (int w = 0; w < NWL; ++w) {
            uint32_t v = R[w * 32];
            v ^= (v >> 16) & 0x0000FFFFu;
            v ^= (v >> 8) & 0x00FF00FFu;
            v ^= (v >> 4) & 0x0F0F0F0Fu;
            v ^= (v >> 2) & 0x33333333u;
            v ^= (v >> 1) & 0x55555555u;
            R[w * 32] = v;
        }
        for (int s = 1; s < NWL; s <<= 1)
            for (int w = 0; w < NWL; ++w)
                if ((w & s) == 0) R[w * 32] ^= R[(w + s) * 32];
        for (int w = 0; w < NWL; ++w) {
            uint32_t v = R[w * 32];
            const uint32_t t1 = __shfl_xor_sync(FULL_MASK, v, 1);
            if ((q & 1) == 0) v ^= t1;
            const uint32_t t2 = __shfl_xor_sync(FULL_MASK, v, 2);
            if ((q & 2) == 0) v ^= t2;
            R[w * 32] = v;
        }
        // the codeword's margin and flag
        mg = min(mg, __shfl_xor_sync(FULL_MASK, mg, 1));
        mg = min(mg, __shfl_xor_sync(FULL_MASK, mg, 2));
        const uint32_t mgq = POLAR_FAST_NS::q_of(__uint_as_float(mg));
        const bool flagme = a.flag_list != nullptr && mgq < a.tauq_flag;
        if (q == 0 && valid) {
            if (a.margin != nullptr)
                a.margin[a.cw_base + cw] = (mgq == POLAR_FAST_NS::kQSat) ? CUDART_INF_F : (float)mgq * (1.0f / POLAR_FAST_NS::kQScale);
            if (flagme) a.flag_list[atomicAdd(a.flag_count, 1)] = a.cw_base + cw;
        }
        __syncwarp();
        // decoded[j] = u-hat[order[j]] (:171-174): the four lanes of a codeword gather its output words in turn
        const uint32_t* Y = reinterpret_cast<const uint32_t*>(my) + 4 * cwl;
        bool differs = false;
        for (int t = q; t < KW; t += 4) {
            const int jmax = min(32, a.K - 32 * t);
            const uint16_t* pp = A.pos + 32 * t;
            uint32_t word = 0;
#pragma unroll
            for (int i8 = 0; i8 < 4; ++i8) {
                const uint4 p8 = __ldg(reinterpret_cast<const uint4*>(pp) + i8);          // eight 16-bit entries
                const uint32_t pw[4] = {p8.x, p8.y, p8.z, p8.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t e = (pw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
                    const uint32_t v = Y[(e >> 7) * 32 + ((e >> 5) & 3u)];
                    if (8 * i8 + k < jmax) word |= ((v >> (e & 31u)) & 1u) << (8 * i8 + k);
                }
            }
            if (valid) {
                if (a.out != nullptr) a.out[(size_t)cw * KW + t] = word;
                if (a.truth != nullptr) differs |= word != a.truth[(size_t)cw * KW + t];
            }
        }
        if (a.truth != nullptr) {                    // block-error count of the BLER loop (:758-769); flagged codewords are counted by the second pass
            differs |= __shfl_xor_sync(FULL_MASK, (int)differs, 1) != 0;
            differs |= __shfl_xor_sync(FULL_MASK, (int)differs, 2) != 0;
            if (q == 0 && valid && differs && !flagme)
                atomicAdd(a.err + (int)((unsigned long long)(a.first_index + a.cw_base + cw) % (unsigned)a.n_ebno), 1ull);
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wib == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm_base), "r"(512u) : "memory");
}
#endif  // __CUDACC__

}  // namespace ssc
