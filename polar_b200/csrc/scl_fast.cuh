// scl_fast.cuh -- the throughput kernel (STRICT mode's first pass): block lengths 2^8..2^13, list sizes 17..32 with one codeword per warp
// (lane = list path) and list sizes 1..16 with 2..32 codewords per warp (W = list size rounded up to a power of
// two lanes per codeword; list 1 = plain SC with lane = codeword).
//
// Same algorithm and the same arithmetic contract as the generic kernel in polar_b200.cu (the
// reference's decode_scl_llr, PolarC/PolarCode.cpp:130-190, 422-644). What differs is purely where
// data lives and how much code runs per tree node:
//
//   * everything is a template of (NLOG, T, LAMS): layer sizes, row offsets and memory spaces are
//     compile-time constants and the per-bit descent is a jump into straight-line code;
//   * the TOP T layers of the LLR tree -- 1 - 2^-T of all per-path bytes -- are never stored per
//     path. Their f-only prefixes are computed once while a single path exists (shared, compact
//     arrays XS_1..XS_T), and every other node of layer T is recomputed on the fly from the
//     channel LLRs (or from a shared array) and the path's own partial-sum bits: a g node costs one
//     add, and the few f nodes that get evaluated twice add about 5% to the f count at T = 3.
//     Per-path storage therefore starts at layer T: N/2^T rows instead of N;
//   * layers T..LAMS-1 go to a per-warp scratch in HBM that is small enough to stay L2-resident,
//     layers LAMS..NLOG-5 to shared memory; the four deepest layers (a 16-leaf subtree: 16 + 8 + 4
//     + 2 LLRs per path) are registers, so the per-bit work at the bottom of the tree touches no
//     memory at all and cloning a path there is a shuffle of those registers;
//   * partial sums of the five deepest layers are one packed register per lane (cloning a path =
//     one shuffle); the larger ones are packed words [word][lane] behind 5-bit column pointers;
//   * the 2L -> L prune is "promote the best unlikely fork, demote the worst likely fork, until
//     the best unlikely fork no longer beats the worst likely one", two warp REDUX per round;
//     its first round is the common no-fork-survives exit. It yields exactly the reference's
//     sort / threshold / index-order selection (PolarCode.cpp:528-553);
//   * with 16 or 32 codewords per warp (lists 1..2) the channel LLRs and the XS arrays of the warp's codewords
//     are staged transposed ([position/4][codeword][4], [index][codeword]) so that the lanes of a warp read
//     neighbouring words (top_solo_transposed);
//   * path metrics are fixed-point numbers relative to the best path of the list (Q8.24, renormalised every 16 leaves,
//     saturating): exact additions, keys that feed the warp REDUX directly, no float granularity at metric ~ 700;
//     plain SC keeps no metric at all;
//   * every keep/drop decision leaves its margin (best dropped - worst kept metric; |LLR| for plain SC; runner-up gap
//     of the final pick) and the smallest one per codeword decides whether STRICT mode decodes the codeword again in
//     double (scl_exact.cuh); the block-error comparison of the BLER loop can be fused into the tail;
//   * compiled configurations are listed in fast_parts.cu (see fast_variants.cuh); DESIGN.md section 3 has the
//     placement of every layer and section 5 the measurements behind each choice.
#pragma once
#include <type_traits>

// POLAR_MINSUM=1 (with POLAR_FAST_NS=fastms) compiles the same kernels in the opt-in, NON-PARITY fast arithmetic of
// pure min-sum check nodes (the reference's own fallback branch, PolarC/PolarCode.cpp:442-446,
// applied everywhere) and the hardware-friendly path-metric update of Balatsoukas-Stimming et al. (PM += |LLR| when
// the decision contradicts the LLR's sign, nothing otherwise) -- no transcendental at all. Checked bit for bit against
// the oracle's minsum_only = 2 mode, never against the reference.
#ifndef POLAR_MINSUM
#define POLAR_MINSUM 0
#endif
#ifndef POLAR_FAST_NS
#define POLAR_FAST_NS fast
#endif

namespace fastcommon {
struct Args {
    const float* llr;
    uint32_t* out;
    const uint32_t* frozen_words;
    const uint16_t* info_order;
    const uint32_t* crc_masks;
    float* gx;
    uint32_t* gs;
    int B, K, crc, L;
    int PA;            // leading frozen leaves decoded cooperatively (multiple of 16, < N >> T)
    // decision margins (strict mode): the smallest gap any keep/drop decision of a codeword was
    // taken with. margin (may be null): [B] floats out. flag_list / flag_count (may be null): codewords whose margin
    // is below tau are appended (atomic counter) for re-decoding in reference precision.
    float* margin;
    int* flag_list;
    int* flag_count;
    float tau;         // plain SC: |LLR| below tau is recorded;  lists: gaps below tauq (Q8.24)
    uint32_t tauq;
    uint32_t tauq_flag;// codewords whose smallest recorded margin is below this are appended to flag_list
    int cw_base;       // index of llr's row 0 in the caller's batch (margin / flag_list are indexed by batch position)
    // block-error counting fused into the tail (the comparison loop of the BLER harness, PolarCode.cpp:758-769):
    // truth (may be null): [B][KW] packed info bits; err: one counter per Eb/N0 point, codeword with global index g
    // belongs to point g % n_ebno. Codewords that strict mode flags are counted by the second pass instead.
    // out may be null when only the counters are wanted.
    const uint32_t* truth;
    unsigned long long* err;
    long long first_index;
    int n_ebno;
};
}  // namespace fastcommon

namespace POLAR_FAST_NS {

// channel LLRs: read-only path. (Streaming / evict-first loads were measured 5% slower: each codeword's
// LLRs are re-read by four top-layer nodes and those re-reads do hit L2.)
#ifndef POLAR_LDCHAN
#define POLAR_LDCHAN(p) __ldg(p)
#endif

// lists with at least this many codewords per warp stage their channel LLRs / XS arrays transposed (0 = never)
#ifndef POLAR_TRANSPOSED_MIN_G
#define POLAR_TRANSPOSED_MIN_G 16
#endif

template <int NLOG_, int T_, int LAMS_, int WLOG_ = 5, int SGMIN_ = 16, int TM_ = 0>
struct Cfg {
    static constexpr int NLOG = NLOG_, T = T_, LAMS = LAMS_, WLOG = WLOG_;
    static constexpr int W = 1 << WLOG;               // lanes per codeword (list size rounded up to a power of two)
    static constexpr int G = 32 / W;                  // codewords per warp
    static_assert(WLOG >= 0 && WLOG <= 5, "1..32 lanes per codeword");
    static constexpr int N = 1 << NLOG;
    static constexpr int MT = N >> T;                 // rows of the first per-path layer
    static constexpr int NW = N / 32;
    static constexpr int LB = NLOG - 4;               // layer whose 16-entry array (and everything deeper) is registers
    static_assert(NLOG - T >= 5, "layer T must have at least 32 rows");
    static_assert(LAMS >= T && LAMS <= LB, "bad shared-memory split");
    // rows of per-path layers [a, b)
    static constexpr __host__ __device__ int rows(int a, int b) { return (N >> (a - 1)) - (N >> (b - 1)); }
    // TM: the last layer below the shared-memory ones (LT = LAMS-1) lives in tensor memory (TMEM), used as a
    // per-warp scratch: one 32-lane quadrant x (N >> LT) columns = that layer's [beta][lane] rows.
    static constexpr bool TM = TM_ != 0;
    static constexpr int LT = LAMS - 1;
    static constexpr int TM_COLS = TM ? ((N >> LT) < 32 ? 32 : (N >> LT)) : 0;
    static_assert(!TM || (LT >= T && (N >> LT) <= 256), "TMEM layer must lie in [T, LAMS) and have <= 256 rows");
    static constexpr int GX_ROWS = rows(T, TM ? LT : LAMS);   // HBM scratch rows (32 floats each)
    static constexpr int SX_ROWS = rows(LAMS, LB);     // shared rows (layers LAMS..NLOG-5)
    static constexpr int XS_FLOATS = N - MT;           // shared compact arrays XS_1..XS_T
    static constexpr __host__ __device__ int xs_off(int lev) { return N - (N >> (lev - 1)); }   // XS_lev at this float offset
    // partial-sum word layers: 1..NLOG-5 (>= 32 bits). Layers with >= 16 words live in HBM scratch.
    static constexpr int SWL = NLOG - 5;               // last word layer
    static constexpr __host__ __device__ int swords(int lam) { return (N >> lam) / 32; }
    // partial-sum word layers with at least SGMIN words live in the HBM scratch, smaller ones in shared memory
    static constexpr __host__ __device__ bool s_global(int lam) { return lam == 0 || swords(lam) >= SGMIN_; }
    static constexpr __host__ __device__ int s_off(int lam) {              // row offset within its space
        int off = 0;
        for (int j = 0; j < lam; ++j)
            if (s_global(j) == s_global(lam)) off += (j == 0 ? NW : swords(j));
        return off;
    }
    static constexpr __host__ __device__ int gs_rows() { int r = NW; for (int j = 1; j <= SWL; ++j) if (s_global(j)) r += swords(j); return r; }
    static constexpr __host__ __device__ int ss_rows() { int r = 0; for (int j = 1; j <= SWL; ++j) if (!s_global(j)) r += swords(j); return r; }
    static constexpr int GS_ROWS = gs_rows();
    static constexpr int SS_ROWS = ss_rows();
    // + 64 bytes of clone scatter / free-path stack, + 16 bytes for the decision-margin slot of the one-codeword-per-warp
    // kernels, which have no register to spare (the others keep the margin in a register; anything more per warp would
    // cost the N = 2048 variants their fourth block per SM)
    static constexpr int SMEM_PER_WARP = (SX_ROWS + SS_ROWS) * 128 + 64 + (WLOG_ == 5 ? 16 : 0);
    static constexpr int PA_FLOATS = MT + MT / 2;       // phase A: two walk buffers + one subtree buffer
    // lists <= 16 (G > 1 codewords per warp): the channel LLRs of the warp's codewords are staged TRANSPOSED
    // ([position / 4][codeword][4], so that a lane still reads float4) behind the XS arrays, which are stored
    // [index][codeword]
    static constexpr bool TRANSPOSED = POLAR_TRANSPOSED_MIN_G > 1 && G >= POLAR_TRANSPOSED_MIN_G;
    static constexpr int CHT_FLOATS = TRANSPOSED ? G * N : 0;
    static constexpr size_t XST_OFF = (size_t)GX_ROWS * 32;                       // XS arrays (compact per codeword, or transposed)
    static constexpr size_t CHT_OFF = XST_OFF + (size_t)G * XS_FLOATS + PA_FLOATS;
    static constexpr size_t GX_FLOATS = (size_t)GX_ROWS * 32 + (size_t)G * XS_FLOATS + PA_FLOATS + CHT_FLOATS;
    static constexpr size_t GS_WORDS = (size_t)GS_ROWS * 32;
};

using fastcommon::Args;

struct Warp {          // per-warp pointers
    float* sx;         // shared LLR rows
    uint32_t* ss;      // shared partial-sum rows
    unsigned char* srcof;   // clone scatter: srcof[new lane] = parent lane
    unsigned char* stack;   // free-path stack (PolarCode.h:60 _inactivePathIndices)
    float* gx;         // HBM LLR rows, followed by the XS arrays of the warp's codewords
    float* xs;         // compact shared arrays of THIS LANE's codeword
    uint32_t* gs;      // HBM partial-sum rows
    const float* chan; // channel LLRs of THIS LANE's codeword
    uint32_t tm;       // TMEM address of this warp's quadrant (TM configurations)
    int lane;
    float* xst;        // lists <= 16: XS arrays of the warp's codewords, transposed [index][codeword]
    float* cht;        // lists <= 16: channel LLRs of the warp's codewords, transposed [position][codeword]
    int g;             // lists <= 16: this lane's codeword within the warp
    uint32_t* mg;      // shared: smallest decision margin so far (float bits, >= 0), one slot per codeword of the warp
};

struct Lane {          // per-path state
    uint32_t pm;       // path metric above the list's best, Q8.24 (q_of / q_add)
    uint32_t mg;       // lists <= 16: smallest decision margin so far (Q8.24; plain SC: float bits of the smallest |LLR|)
    bool active;
    unsigned long long px;   // column pointers of LLR layers T.. (index lam - T)
    unsigned long long ps;   // column pointers of partial-sum word layers 1..SWL (index lam - 1)
    uint32_t sreg;           // packed partial sums of layers NLOG-k, k = 0..4, at bit 2^k - 1
};

__device__ __forceinline__ unsigned brev_bits(unsigned x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }
constexpr __host__ __device__ unsigned cbrev(unsigned x, int bits) {
    unsigned r = 0;
    for (int b = 0; b < bits; ++b) if (x & (1u << b)) r |= 1u << (bits - 1 - b);
    return r;
}
constexpr __host__ __device__ int lead_zeros(int node, int bits) {
    int z = 0;
    for (int b = bits - 1; b >= 0; --b) { if (node & (1 << b)) break; ++z; }
    return z;
}

// reductions / ballots over the W lanes of one codeword
template <int W> __device__ __forceinline__ unsigned gmin(unsigned v) {
    if constexpr (W == 32) return __reduce_min_sync(FULL_MASK, v);
    else {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL_MASK, v, o));
        return v;
    }
}
template <int W> __device__ __forceinline__ unsigned gmax(unsigned v) {
    if constexpr (W == 32) return __reduce_max_sync(FULL_MASK, v);
    else {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL_MASK, v, o));
        return v;
    }
}
template <int W> __device__ __forceinline__ unsigned gballot(bool p, int gbase) {
    const unsigned b = __ballot_sync(FULL_MASK, p);
    if constexpr (W == 32) return b;
    else return (b >> gbase) & ((1u << W) - 1u);
}
template <class P> __device__ __forceinline__ P* shfl_ptr(P* p, int src) {
    return reinterpret_cast<P*>(__shfl_sync(FULL_MASK, reinterpret_cast<unsigned long long>(p), src));
}

// ---- path metrics in fixed point (Q8.24: 2^-24 ~ 6e-8 resolution, saturating at 256) ----
// The reference keeps metrics in double; a float metric of magnitude ~700 (every frozen bit adds up to ln 2) resolves
// only 6e-5, which alone made 4 % of the list-32 codewords tie somewhere. Only metric DIFFERENCES within a codeword
// decide anything (PolarCode.cpp:528-553, 609-644), so the kernel keeps metric - (smallest metric of the list), renormalised
// once per 16 leaves, as an unsigned fixed-point number: additions are exact, keys need no conversion for the warp
// REDUX, and a saturated key (>= 256 above the best path, or the reference's +inf) still orders correctly against any
// unsaturated one. Two saturated keys compare equal, which shows up as a zero margin.
constexpr float kQScale = 16777216.0f;
constexpr uint32_t kQSat = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t q_of(float x) { return __float2uint_rn(x * kQScale); }        // x >= 0; saturates
__device__ __forceinline__ uint32_t q_add(uint32_t a, uint32_t b) { const uint32_t r = a + b; return r < a ? kQSat : r; }

// Decision margin: the gap (>= 0) between the worst fork that was kept and the best fork that was dropped. A gap
// below the arithmetic's own error means the double-precision reference may decide otherwise; such codewords are
// re-decoded in double (strict mode). Only gaps below tauq are recorded (the smallest one per codeword).
template <int W> __device__ __forceinline__ void note_gap(const Warp& w, Lane& s, uint32_t gapq, uint32_t tauq) {
    if constexpr (W == 32) { if (gapq < tauq && w.lane == 0) atomicMin(w.mg, gapq); }    // rare: only sub-threshold gaps
    else s.mg = min(s.mg, gapq);                     // uniform over the lanes of a codeword
}

// ---- tensor memory as a per-warp scratch (tcgen05.ld/st, shape 32x32b: lane i of the warp <-> TMEM lane i of
// the warp's quadrant, one 32-bit value per column) ----
__device__ __forceinline__ void tm_st4(uint32_t taddr, const float (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                    "r"(__float_as_uint(v[3])) : "memory");
}
__device__ __forceinline__ void tm_st1(uint32_t taddr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(taddr), "r"(__float_as_uint(v)) : "memory");
}
__device__ __forceinline__ void tm_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <class C, int LAM>
__device__ __forceinline__ float* xbase(const Warp& w) {
    if constexpr (LAM >= C::LAMS) return w.sx + C::rows(C::LAMS, LAM) * 32;
    else return w.gx + C::rows(C::T, LAM) * 32;
}
template <class C, int LAM>
__device__ __forceinline__ uint32_t* sbase(const Warp& w) {
    if constexpr (C::s_global(LAM)) return w.gs + C::s_off(LAM) * 32;
    else return w.ss + C::s_off(LAM) * 32;
}
template <class C>
__device__ __forceinline__ uint32_t* sbase_rt(const Warp& w, int lam) {
    uint32_t* p = nullptr;
#define POLAR_SB(L_) if constexpr (L_ <= C::SWL) { if (lam == L_) p = sbase<C, L_>(w); }
    POLAR_SB(0) POLAR_SB(1) POLAR_SB(2) POLAR_SB(3) POLAR_SB(4) POLAR_SB(5) POLAR_SB(6) POLAR_SB(7) POLAR_SB(8)
#undef POLAR_SB
    return p;
}

template <class C>
__device__ __forceinline__ float* xbase_rt(const Warp& w, int lam) {
    float* p = nullptr;
#define POLAR_XB(L_) if constexpr (L_ >= C::T && L_ < C::LB && !(C::TM && L_ == C::LT)) { if (lam == L_) p = xbase<C, L_>(w); }
    POLAR_XB(2) POLAR_XB(3) POLAR_XB(4) POLAR_XB(5) POLAR_XB(6) POLAR_XB(7) POLAR_XB(8) POLAR_XB(9)
#undef POLAR_XB
    return p;
}

__device__ __forceinline__ float g_rule(float a, float b, uint32_t bit) {
    // (1 - 2u) a + b, PolarCode.cpp:448-451
    return b + __int_as_float(__float_as_int(a) ^ (int)(bit << 31));
}

// ---- two nodes per instruction: Blackwell's packed fp32 pipe (FADD2 / FMUL2 / FFMA2 on a register pair).
// Every packed operation is the IEEE round-to-nearest operation of its two halves, so f_rule2 / g_top2 return
// exactly the bits of two f_rule / g_rule calls; what changes is the instruction count (14 instead of 18
// issue slots per check node, 2.5 instead of 4 per variable node). POLAR_PACKED=0 compiles the scalar forms.
#ifndef POLAR_PACKED
#define POLAR_PACKED 1
#endif
#ifndef POLAR_TM_PIPE
#define POLAR_TM_PIPE 1
#endif
#ifndef POLAR_SW_PIPE
#define POLAR_SW_PIPE 1
#endif
#ifndef POLAR_X4_UNCOND
#define POLAR_X4_UNCOND 1
#endif
// leaves per iteration of the per-bit loop (0 = rolled, 2 or 4), for one codeword per warp and for plain SC
#ifndef POLAR_UNROLL2
#define POLAR_UNROLL2 2
#endif
#ifndef POLAR_PS_SHORTCUT
#define POLAR_PS_SHORTCUT 1
#endif
#ifndef POLAR_TAIL_REGS
#define POLAR_TAIL_REGS 1
#endif
__device__ __forceinline__ float sign_min(float a, float b) {
    return __int_as_float(__float_as_int(fminf(fabsf(a), fabsf(b))) |
                          ((__float_as_int(a) ^ __float_as_int(b)) & (int)0x80000000));
}
#if POLAR_MINSUM
// these hide the reference-rule helpers of polar_dev.cuh inside this namespace
__device__ __forceinline__ float f_rule(float a, float b) { return sign_min(a, b); }
__device__ __forceinline__ float softplus_ref(float x) { return fmaxf(x, 0.0f); }
__device__ __forceinline__ float log1p_exp_neg(float) { return 0.0f; }
#endif
// log(1 + e^x) for the fixed-point metric: the reference's corner cases (softplus_ref) fall out of q_of by themselves
__device__ __forceinline__ float softplus_q(float x) { return fmaxf(x, 0.0f) + log1p_exp_neg(fabsf(x)); }
__device__ __forceinline__ void f_rule2(float a0, float b0, float a1, float b1, float& y0, float& y1) {
#if POLAR_MINSUM
    y0 = sign_min(a0, b0); y1 = sign_min(a1, b1);
#elif POLAR_PACKED
    const float2 A = make_float2(a0, a1);
    const float2 S = __fadd2_rn(A, make_float2(b0, b1));
    const float2 D = __fadd2_rn(A, make_float2(-b0, -b1));
    const float2 nl = make_float2(-kLog2e, -kLog2e), one = make_float2(1.0f, 1.0f);
    const float2 ES = __fmul2_rn(make_float2(fabsf(S.x), fabsf(S.y)), nl);
    const float2 ED = __fmul2_rn(make_float2(fabsf(D.x), fabsf(D.y)), nl);
    const float2 P = __fadd2_rn(make_float2(ex2_approx(ES.x), ex2_approx(ES.y)), one);
    const float2 Q = __fadd2_rn(make_float2(ex2_approx(ED.x), ex2_approx(ED.y)), one);
    const float2 LQ = make_float2(-lg2_approx(Q.x), -lg2_approx(Q.y));
    const float2 DIFF = __fadd2_rn(make_float2(lg2_approx(P.x), lg2_approx(P.y)), LQ);
    const float2 SC = make_float2((fmaxf(fabsf(a0), fabsf(b0)) < 40.0f) ? kLn2 : 0.0f,
                                  (fmaxf(fabsf(a1), fabsf(b1)) < 40.0f) ? kLn2 : 0.0f);
    const float2 R = __ffma2_rn(DIFF, SC, make_float2(sign_min(a0, b0), sign_min(a1, b1)));
    y0 = R.x; y1 = R.y;
#else
    y0 = f_rule(a0, b0); y1 = f_rule(a1, b1);
#endif
}
// variable node with the partial-sum bit delivered in bit 31 of w (the other bits of w are ignored)
__device__ __forceinline__ float g_top(float a, float b, uint32_t w) {
    return b + __int_as_float(__float_as_int(a) ^ (int)(w & 0x80000000u));
}
__device__ __forceinline__ void g_top2(float a0, float b0, uint32_t w0, float a1, float b1, uint32_t w1, float& y0,
                                       float& y1) {
#if POLAR_PACKED
    const float2 R = __fadd2_rn(make_float2(b0, b1),
                                make_float2(__int_as_float(__float_as_int(a0) ^ (int)(w0 & 0x80000000u)),
                                            __int_as_float(__float_as_int(a1) ^ (int)(w1 & 0x80000000u))));
    y0 = R.x; y1 = R.y;
#else
    y0 = g_top(a0, b0, w0); y1 = g_top(a1, b1, w1);
#endif
}
// four nodes of one layer: y[j] = f(a[j], b[j]) or g(a[j], b[j], bit pos + j of word)
template <bool ISG>
__device__ __forceinline__ void node4(const float (&a)[4], const float (&b)[4], uint32_t word, int pos, float (&y)[4]) {
    if constexpr (ISG) {
        const uint32_t w = word << (28 - pos);                 // bit pos + j -> bit 28 + j
        g_top2(a[0], b[0], w << 3, a[1], b[1], w << 2, y[0], y[1]);
        g_top2(a[2], b[2], w << 1, a[3], b[3], w, y[2], y[3]);
    } else {
        f_rule2(a[0], b[0], a[1], b[1], y[0], y[1]);
        f_rule2(a[2], b[2], a[3], b[3], y[2], y[3]);
    }
}

// ---- one per-path layer in memory: X_LAM = f / g (X_{LAM-1}), T < LAM <= NLOG-5 ----
template <class C, int LAM, bool ISG>
__device__ __forceinline__ void layer_step(const Warp& w, Lane& s) {
    constexpr int M = C::N >> LAM;
    static_assert(M >= 32, "memory layers have at least 32 rows");
    constexpr bool SRC_TM = C::TM && (LAM - 1 == C::LT), DST_TM = C::TM && (LAM == C::LT);
    const int pcol = get_ptr(s.px, LAM - 1 - C::T);
    if constexpr (SRC_TM || DST_TM) {
        // tcgen05.ld/st are warp-collective: every lane runs the loop (idle lanes work on don't-care data)
        const uint32_t* sw = nullptr;
        if constexpr (ISG) sw = sbase<C, LAM>(w) + get_ptr(s.ps, LAM - 1);
        const float* src = nullptr;
        if constexpr (!SRC_TM) src = xbase<C, LAM - 1>(w) + pcol;
        float* dst = nullptr;
        if constexpr (!DST_TM) dst = xbase<C, LAM>(w) + w.lane;
        uint32_t word = 0;
#if POLAR_TM_PIPE
        if constexpr (!SRC_TM) {
            // source rows are in the HBM/L2 scratch, destination is tensor memory: the loads of group i+1 are issued
            // before group i is computed, so that their latency overlaps the MUFU work of the check nodes
            float a[4], b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { a[j] = src[j * 32]; b[j] = src[(j + M) * 32]; }
#pragma unroll 1
            for (int i0 = 0; i0 < M; i0 += 4) {
                if constexpr (ISG) { if ((i0 & 31) == 0) word = sw[(i0 >> 5) * 32]; }
                float na[4], nb[4], y[4];
                const int nx = (i0 + 4 < M) ? i0 + 4 : i0;          // last group re-reads itself (harmless)
#pragma unroll
                for (int j = 0; j < 4; ++j) { na[j] = src[(nx + j) * 32]; nb[j] = src[(nx + j + M) * 32]; }
                node4<ISG>(a, b, word, i0 & 31, y);
                tm_st4(w.tm + i0, y);
#pragma unroll
                for (int j = 0; j < 4; ++j) { a[j] = na[j]; b[j] = nb[j]; }
            }
            tm_wait_st();
            s.px = set_ptr(s.px, LAM - C::T, w.lane);
            return;
        }
#endif
#pragma unroll 1
        for (int i0 = 0; i0 < M; i0 += 4) {
            if constexpr (ISG) { if ((i0 & 31) == 0) word = sw[(i0 >> 5) * 32]; }
            float a[4], b[4], y[4];
            if constexpr (SRC_TM) {
                tm_ld4(w.tm + i0, a);
                tm_ld4(w.tm + i0 + M, b);
                tm_wait_ld();
#pragma unroll
                for (int j = 0; j < 4; ++j) {             // the path's rows may belong to another lane's column
                    a[j] = __shfl_sync(FULL_MASK, a[j], pcol);
                    b[j] = __shfl_sync(FULL_MASK, b[j], pcol);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { a[j] = src[(i0 + j) * 32]; b[j] = src[(i0 + j + M) * 32]; }
            }
            node4<ISG>(a, b, word, i0 & 31, y);
            if constexpr (DST_TM) tm_st4(w.tm + i0, y);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[(i0 + j) * 32] = y[j];
            }
        }
        if constexpr (DST_TM) tm_wait_st();
        s.px = set_ptr(s.px, LAM - C::T, w.lane);
        return;
    }
    const float* src = xbase<C, LAM - 1>(w) + pcol;
    if (s.active) {
        float* dst = xbase<C, LAM>(w) + w.lane;
        const uint32_t* sw = nullptr;
        if constexpr (ISG) sw = sbase<C, LAM>(w) + get_ptr(s.ps, LAM - 1);
        if constexpr (LAM - 1 < C::LAMS) {
            // source rows are in the HBM/L2 scratch: groups of 4 nodes, the loads of group i+1 are issued
            // before group i is computed so that their latency overlaps the MUFU work
            float a[4], b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { a[j] = src[j * 32]; b[j] = src[(j + M) * 32]; }
            uint32_t word = 0;
#pragma unroll 1
            for (int i0 = 0; i0 < M; i0 += 4) {
                if constexpr (ISG) { if ((i0 & 31) == 0) word = sw[(i0 >> 5) * 32]; }
                float na[4], nb[4];
                const int nx = (i0 + 4 < M) ? i0 + 4 : i0;          // last group re-reads itself (harmless)
#pragma unroll
                for (int j = 0; j < 4; ++j) { na[j] = src[(nx + j) * 32]; nb[j] = src[(nx + j + M) * 32]; }
                float y[4];
                node4<ISG>(a, b, word, i0 & 31, y);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[(i0 + j) * 32] = y[j];
#pragma unroll
                for (int j = 0; j < 4; ++j) { a[j] = na[j]; b[j] = nb[j]; }
            }
        } else {
            // source rows are in shared memory: plain groups of 4
            uint32_t word = 0;
#pragma unroll 1
            for (int i0 = 0; i0 < M; i0 += 4) {
                if constexpr (ISG) { if ((i0 & 31) == 0) word = sw[(i0 >> 5) * 32]; }
                float a[4], b[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { a[j] = src[(i0 + j) * 32]; b[j] = src[(i0 + j + M) * 32]; }
                float y[4];
                node4<ISG>(a, b, word, i0 & 31, y);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[(i0 + j) * 32] = y[j];
            }
        }
    }
    s.px = set_ptr(s.px, LAM - C::T, w.lane);
}

// ---- the register subtree: layers NLOG-4 .. NLOG of one path ----
struct Sub {
    float x4[16], x3[8], x2[4], x1[2];
};
template <int K> __device__ __forceinline__ float* sub_arr(Sub& r) {
    if constexpr (K == 4) return r.x4;
    else if constexpr (K == 3) return r.x3;
    else if constexpr (K == 2) return r.x2;
    else return r.x1;
}

// layer NLOG-4 (16 entries) straight into registers from the 32-row layer above it
template <class C, bool ISG>
__device__ __forceinline__ void layer_to_regs(const Warp& w, const Lane& s, Sub& r) {
    constexpr int LAM = C::LB;
    const float* src = xbase<C, LAM - 1>(w) + get_ptr(s.px, LAM - 1 - C::T);
    const uint32_t field = s.sreg >> 15;                  // packed partial sums of layer NLOG-4
    // POLAR_X4_UNCOND: idle lanes compute too (on don't-care rows of their own column), so that x4 is dead across
    // the whole descent and its 16 registers are free for the prefetches up there
    if (POLAR_X4_UNCOND || s.active) {
#pragma unroll
        for (int j0 = 0; j0 < 16; j0 += 4) {
            float a[4], b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { a[j] = src[(j0 + j) * 32]; b[j] = src[(j0 + j + 16) * 32]; }
            float y[4];
            node4<ISG>(a, b, field, j0, y);
#pragma unroll
            for (int j = 0; j < 4; ++j) r.x4[j0 + j] = y[j];
        }
    }
}

// level K of the subtree (2^K entries) from level K+1; K = 0 yields the decision LLR
template <int K>
__device__ __forceinline__ void sub_step(Sub& r, uint32_t sreg, bool isg, float& lam_n) {
    constexpr int M = 1 << K;
    const float* in = sub_arr<K + 1>(r);
    const uint32_t field = sreg >> (M - 1);
    if constexpr (K == 0) {
        if (isg) lam_n = g_rule(in[0], in[1], field & 1u);
        else lam_n = f_rule(in[0], in[1]);
    } else if (isg) {
#pragma unroll
        for (int i = 0; i < M; i += 2)
            g_top2(in[i], in[i + M], field << (31 - i), in[i + 1], in[i + 1 + M], field << (30 - i),
                   sub_arr<K>(r)[i], sub_arr<K>(r)[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < M; i += 2)
            f_rule2(in[i], in[i + M], in[i + 1], in[i + 1 + M], sub_arr<K>(r)[i], sub_arr<K>(r)[i + 1]);
    }
}

// compile-time loop: f(integral_constant<int, I>) for I in [B, E)
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// inputs of layer-T node NODE for the betas q0 + 2H and q0 + 2H + 1 (q0 a multiple of 4): the bit-reversed channel
// position of q0 + c is that of q0 plus a constant, so the four betas of a quad share one address computation
template <class C, int NODE, int H>
__device__ __forceinline__ void top_load_pair(const Warp& w, int q0, float (&v)[2][1 << (C::T - lead_zeros(NODE, C::T))]) {
    constexpr int T = C::T, MT = C::MT, BITS = C::NLOG - T;
    constexpr int S0 = lead_zeros(NODE, T);
    constexpr int CNT = 1 << (T - S0);
    if constexpr (C::TRANSPOSED) {
        // several codewords per warp: transposed staging, lanes of different codewords read neighbouring words
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int c = 2 * H + p;
            if constexpr (S0 == 0) {
                const float4* src = reinterpret_cast<const float4*>(w.cht) +
                                    ((size_t)(CNT / 4) * (brev_bits(q0, BITS) + (cbrev(c, 2) << (BITS - 2)))) * C::G + w.g;
#pragma unroll
                for (int q = 0; q < CNT / 4; ++q) {
                    const float4 t4 = src[q * C::G];
                    v[p][4 * q] = t4.x; v[p][4 * q + 1] = t4.y; v[p][4 * q + 2] = t4.z; v[p][4 * q + 3] = t4.w;
                }
            } else {
                const float* src = w.xst + (size_t)(C::xs_off(S0) + q0 + c) * C::G + w.g;
#pragma unroll
                for (int i = 0; i < CNT; ++i) v[p][i] = src[(MT * (int)cbrev(i, T - S0)) * C::G];
            }
        }
        return;
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int c = 2 * H + p;
        if constexpr (S0 == 0) {
            const float4* c4 = reinterpret_cast<const float4*>(w.chan + (size_t)CNT * brev_bits(q0, BITS)) +
                               (CNT / 4) * (int)(cbrev(c, 2) << (BITS - 2));
#pragma unroll
            for (int q = 0; q < CNT / 4; ++q) {
                const float4 t4 = POLAR_LDCHAN(c4 + q);
                v[p][4 * q] = t4.x; v[p][4 * q + 1] = t4.y; v[p][4 * q + 2] = t4.z; v[p][4 * q + 3] = t4.w;
            }
        } else {
            const float* xs = w.xs + C::xs_off(S0) + q0 + c;
#pragma unroll
            for (int i = 0; i < CNT; ++i) v[p][i] = xs[MT * (int)cbrev(i, T - S0)];
        }
    }
}

// ---- layer T node NODE (1 .. 2^T - 1) for every path, from the channel / shared arrays ----
template <class C, int NODE>
__device__ __forceinline__ void top_node(const Warp& w, Lane& s) {
    constexpr int T = C::T, MT = C::MT, NLOG = C::NLOG, BITS = NLOG - T;
    constexpr int S0 = lead_zeros(NODE, T);          // values enter at level S0 (0 = channel)
    constexpr int CNT = 1 << (T - S0);
    constexpr bool DST_TM = C::TM && (C::LT == T);     // layer T itself lives in tensor memory
    float* dst = nullptr;
    if constexpr (!DST_TM) dst = xbase<C, T>(w) + w.lane;
    if (DST_TM || s.active) {                           // tcgen05.st is warp-collective
        // partial-sum words of this path for the g levels: level lev, local element i sits at
        // position beta + MT * brev(i) of layer lev; flat index (1 << (T - lev)) + i
        uint32_t sw[1 << T] = {};
        auto load_sw = [&](int wd, uint32_t (&dstw)[1 << T]) {
            static_for<S0 + 1, T + 1>([&](auto lev_c) {
                constexpr int lev = decltype(lev_c)::value;
                if constexpr ((NODE >> (T - lev)) & 1) {
                    const uint32_t* base = sbase<C, lev>(w) + get_ptr(s.ps, lev - 1);
                    static_for<0, (1 << (T - lev))>([&](auto i_c) {
                        constexpr int i = decltype(i_c)::value;
                        dstw[(1 << (T - lev)) + i] = base[(wd + (MT / 32) * (int)cbrev(i, T - lev)) * 32];
                    });
                }
            });
        };
#if POLAR_SW_PIPE
        load_sw(0, sw);
#endif
#pragma unroll 1
        for (int wd = 0; wd < MT / 32; ++wd) {
#if POLAR_SW_PIPE
            // the words of the next 32 betas are requested now (most of them come from the L2-resident scratch)
            uint32_t swn[1 << T] = {};
            load_sw(wd + 1 < MT / 32 ? wd + 1 : wd, swn);
#else
            load_sw(wd, sw);
#endif
            // two betas at a time (wd * 32 + bi and + 1), so that every check node has a partner for the packed pipe
            auto compute_pair = [&](int bi, float (&v)[2][CNT]) {
                static_for<S0 + 1, T + 1>([&](auto lev_c) {
                    constexpr int lev = decltype(lev_c)::value;
                    constexpr bool isg = (NODE >> (T - lev)) & 1;
                    static_for<0, (1 << (T - lev))>([&](auto i_c) {
                        constexpr int i = decltype(i_c)::value;
                        if constexpr (isg) {
                            const uint32_t word = sw[(1 << (T - lev)) + i];
                            g_top2(v[0][2 * i], v[0][2 * i + 1], word << (31 - bi), v[1][2 * i], v[1][2 * i + 1],
                                   word << (30 - bi), v[0][i], v[1][i]);
                        } else {
                            f_rule2(v[0][2 * i], v[0][2 * i + 1], v[1][2 * i], v[1][2 * i + 1], v[0][i], v[1][i]);
                        }
                    });
                });
                const int beta = wd * 32 + bi;
                if constexpr (DST_TM) { tm_st1(w.tm + beta, v[0][0]); tm_st1(w.tm + beta + 1, v[1][0]); }
                else { dst[beta * 32] = v[0][0]; dst[(beta + 1) * 32] = v[1][0]; }
            };
            float va[2][CNT], vb[2][CNT];
            top_load_pair<C, NODE, 0>(w, wd * 32, va);
#pragma unroll 1
            for (int bi = 0; bi < 32; bi += 4) {
                top_load_pair<C, NODE, 1>(w, wd * 32 + bi, vb);
                compute_pair(bi, va);
                top_load_pair<C, NODE, 0>(w, wd * 32 + ((bi + 4) & 31), va);   // wraps harmlessly
                compute_pair(bi + 2, vb);
            }
#if POLAR_SW_PIPE
            static_for<0, (1 << T)>([&](auto i_c) { sw[decltype(i_c)::value] = swn[decltype(i_c)::value]; });
#endif
        }
        if constexpr (DST_TM) tm_wait_st();
    }
    s.px = set_ptr(s.px, 0, w.lane);
}

// ---- node 0 of layer T: one path exists, the warp works across beta; fills XS_1..XS_T ----
template <class C>
__device__ __forceinline__ void top_solo_one(const Warp& w, const float* chan, float* xs, int c0) {
    constexpr int T = C::T, N = C::N, NLOG = C::NLOG;
    float* xs1 = xs + C::xs_off(1);
    for (int k = w.lane; k < N / 2; k += 32) {
        const float2 c = POLAR_LDCHAN(reinterpret_cast<const float2*>(chan) + k);
        xs1[brev_bits(k, NLOG - 1)] = f_rule(c.x, c.y);
    }
    __syncwarp();
#pragma unroll
    for (int lev = 2; lev <= T; ++lev) {
        const float* in = xs + C::xs_off(lev - 1);
        float* out = xs + C::xs_off(lev);
        const int M = N >> lev;
        for (int b = w.lane; b < M; b += 32) out[b] = f_rule(in[b], in[b + M]);
        __syncwarp();
    }
    if constexpr (!(C::TM && C::LT == T)) {
        const float* xt = xs + C::xs_off(T);
        float* col = xbase<C, T>(w) + c0;
        for (int b = w.lane; b < C::MT; b += 32) col[b * 32] = xt[b];
        __syncwarp();
    }
}
// layer T in tensor memory: every lane stores its own codeword's XS_T rows into its column
template <class C>
__device__ __forceinline__ void top_solo_publish_tm(const Warp& w) {
    if constexpr (C::TM && C::LT == C::T) {
        const float* xt = w.xs + C::xs_off(C::T);
        for (int b = 0; b < C::MT; b += 4) {
            const float v[4] = {xt[b], xt[b + 1], xt[b + 2], xt[b + 3]};
            tm_st4(w.tm + b, v);
        }
        tm_wait_st();
    }
}
// ---- lists <= 16: stage the channel LLRs of the warp's G codewords transposed, then node 0 of layer T (and the
// compact arrays XS_1..XS_T, transposed too) with lane = (codeword, residue of beta mod W). `first` = channel row of
// the warp's first codeword, `last_g` = last valid codeword of the warp (rows beyond the batch repeat it).
template <class C>
__device__ __noinline__ void top_solo_transposed(const Warp w, const float* first, int last_g, int c0) {
    constexpr int T = C::T, N = C::N, NLOG = C::NLOG, G = C::G, W = C::W, MT = C::MT;
    const int lane = w.lane, g = w.g, slot = lane & (W - 1);
    // channel rows (coalesced float4 reads) -> [position / 4][codeword][4]: every lane writes whole 128-byte lines
    float4* cht4 = reinterpret_cast<float4*>(w.cht);
    constexpr int GB = G < 16 ? G : 16;               // codewords per batch of loads (all issued before the first store)
#pragma unroll 1
    for (int j0 = 0; j0 < N / 4; j0 += 32) {
#pragma unroll 1
        for (int q0 = 0; q0 < G; q0 += GB) {
            float4 v[GB];
#pragma unroll
            for (int q = 0; q < GB; ++q)
                v[q] = POLAR_LDCHAN(reinterpret_cast<const float4*>(first + (size_t)(q0 + q < last_g ? q0 + q : last_g) * N) + j0 + lane);
#pragma unroll
            for (int q = 0; q < GB; ++q) cht4[(size_t)(j0 + lane) * G + q0 + q] = v[q];
        }
    }
    __syncwarp();
    // XS_1[brev(k)] = f(chan[2k], chan[2k+1]) (PolarCode.cpp:438-446): one float4 = two check nodes (packed pipe).
    // Loads are issued in batches of four so that their latency overlaps (the stores may alias for the compiler).
    {
        float* xs1 = w.xst + (size_t)C::xs_off(1) * G + g;
        static_assert((N / 4) % (4 * W) == 0, "whole batches");
#pragma unroll 1
        for (int j = slot; j < N / 4; j += 4 * W) {
            float4 c[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = cht4[(size_t)(j + q * W) * G + g];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float y0, y1;
                f_rule2(c[q].x, c[q].y, c[q].z, c[q].w, y0, y1);
                xs1[(size_t)brev_bits(2 * (j + q * W), NLOG - 1) * G] = y0;
                xs1[(size_t)brev_bits(2 * (j + q * W) + 1, NLOG - 1) * G] = y1;
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int lev = 2; lev <= T; ++lev) {
        const float* in = w.xst + (size_t)C::xs_off(lev - 1) * G + g;
        float* out = w.xst + (size_t)C::xs_off(lev) * G + g;
        const int M = N >> lev;
#pragma unroll 1
        for (int b = slot; b < M; b += 4 * W) {
            float x[4], y[4], z[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { x[q] = in[(size_t)(b + q * W) * G]; y[q] = in[(size_t)(b + q * W + M) * G]; }
            f_rule2(x[0], y[0], x[1], y[1], z[0], z[1]);
            f_rule2(x[2], y[2], x[3], y[3], z[2], z[3]);
#pragma unroll
            for (int q = 0; q < 4; ++q) out[(size_t)(b + q * W) * G] = z[q];
        }
        __syncwarp();
    }
    const float* xt = w.xst + (size_t)C::xs_off(T) * G + g;
    if constexpr (C::TM && C::LT == T) {
        // layer T in tensor memory: every lane stores its own codeword's XS_T rows into its column
        for (int b = 0; b < MT; b += 4) {
            const float v[4] = {xt[(size_t)b * G], xt[(size_t)(b + 1) * G], xt[(size_t)(b + 2) * G], xt[(size_t)(b + 3) * G]};
            tm_st4(w.tm + b, v);
        }
        tm_wait_st();
    } else {
        float* col = xbase<C, T>(w) + (lane & ~(W - 1)) + c0;
        for (int b = slot; b < MT; b += W) col[b * 32] = xt[(size_t)b * G];
        __syncwarp();
    }
}

// all codewords of the warp, one after the other (c0 = local index of the first path)
template <class C>
__device__ __forceinline__ void top_solo(const Warp& w, int c0, bool valid) {
#pragma unroll 1
    for (int g = 0; g < C::G; ++g) {
        const float* chan = shfl_ptr(w.chan, g * C::W);
        float* xs = shfl_ptr(w.xs, g * C::W);
        const bool ok = __shfl_sync(FULL_MASK, (int)valid, g * C::W);
        if (ok) top_solo_one<C>(w, chan, xs, g * C::W + c0);
    }
    top_solo_publish_tm<C>(w);
}

// ---- phase A: the leading PA leaves are frozen and only one path exists, so the warp decodes them
// across beta instead of one lane doing all the work. All partial sums are zero there, which also
// removes the bit-serial dependency: a completely frozen subtree is evaluated level by level
// (f to the left half, a + b to the right half of every node), its leaf LLRs land in decoding order
// and the path metric is accumulated in exactly the reference's order (PolarCode.cpp:475-487).
// On return the lane-per-path state is what bit-serial decoding of leaves 0..PA-1 would have left:
// layers T..NLOG-5 of the path to leaf PA in column c0, zero partial sums, x4 = layer NLOG-4.
template <class C>
__device__ __noinline__ float phase_a(const Warp w, int PA, int c0) {
    constexpr int T = C::T, LB = C::LB, MT = C::MT, N = C::N;
    const int lane = w.lane;
    top_solo_one<C>(w, w.chan, w.xs, c0);
    top_solo_publish_tm<C>(w);
    float* bufA = w.xs + C::XS_FLOATS;
    float* bufB = bufA + MT / 2;
    float* V = bufB + MT / 2;
    const float* cur = w.xs + C::xs_off(T);
    float pm = 0.0f;
    int offset = 0;
#pragma unroll 1
    for (int lam = T; lam < LB; ++lam) {
        const int half = (N >> lam) >> 1;
        float* nxt = ((lam - T) & 1) ? bufB : bufA;
        if (PA >= offset + half) {
            // the left child's subtree [offset, offset + half) is completely frozen
            for (int b = lane; b < half; b += 32) V[b] = f_rule(cur[b], cur[b + half]);
            __syncwarp();
#pragma unroll 1
            for (int m = half; m >= 2; m >>= 1) {
                const int hm = m >> 1;
                for (int idx = lane; idx < half / 2; idx += 32) {
                    const int p = (idx / hm) * m + (idx % hm);
                    const float x = V[p], y = V[p + hm];
                    V[p] = f_rule(x, y);
                    V[p + hm] = y + x;                                   // g with u = 0
                }
                __syncwarp();
            }
            for (int b = lane; b < half; b += 32) V[b] = softplus_ref(-V[b]);
            __syncwarp();
#pragma unroll 4
            for (int i = 0; i < half; ++i) pm += V[i];                   // leaf order, as the reference adds them
            for (int b = lane; b < half; b += 32) nxt[b] = cur[b + half] + cur[b];
            if (lam + 1 <= C::SWL) {                                     // its partial sums: all zero
                uint32_t* sw = sbase_rt<C>(w, lam + 1) + c0;
                for (int x = lane; x < C::swords(lam + 1); x += 32) sw[x * 32] = 0u;
            }
            offset += half;
        } else {
            for (int b = lane; b < half; b += 32) nxt[b] = f_rule(cur[b], cur[b + half]);
        }
        __syncwarp();
        if (C::TM && lam + 1 == C::LT) {
            // publish into the tensor-memory layer: every lane stores the same rows (only column c0 matters)
            for (int b = 0; b < half; b += 4) {
                const float v[4] = {nxt[b], nxt[b + 1], nxt[b + 2], nxt[b + 3]};
                tm_st4(w.tm + b, v);
            }
            tm_wait_st();
        } else if (lam + 1 < LB) {
            float* col = xbase_rt<C>(w, lam + 1) + c0;
            for (int b = lane; b < half; b += 32) col[b * 32] = nxt[b];
        }
        cur = nxt;
    }
    __syncwarp();
    const float keep = cur[lane & 15];
    __syncwarp();
    if (lane < 16) V[lane] = keep;                       // layer NLOG-4 of the path, picked up by the caller
    __syncwarp();
    return pm;
}

// refresh everything above the register subtree for the 16-leaf block starting at phi0, ending with
// the 16 LLRs of layer NLOG-4 in registers (PolarCode.cpp:422-455 for the layers involved)
template <class C>
__device__ __forceinline__ void descend_block(const Warp& w, Lane& s, Sub& r, int phi0, int c0, bool valid,
                                              const float* first_row, int last_g) {
    constexpr int T = C::T, NLOG = C::NLOG, LB = C::LB;
    const int lam_top = (phi0 == 0) ? 0 : NLOG - (__ffs(phi0) - 1);      // <= LB
    bool first = true;
    __syncwarp();       // every lane is done reading the columns that are about to be overwritten
    if (lam_top <= T) {
        const int node = phi0 >> (NLOG - T);
        switch (node) {
            case 0:
                if constexpr (C::TRANSPOSED) top_solo_transposed<C>(w, first_row, last_g, c0);
                else top_solo<C>(w, c0, valid);
                s.px = set_ptr(s.px, 0, (w.lane & ~(C::W - 1)) + c0);
                break;
#define POLAR_TN(N_) case N_: if constexpr (N_ < (1 << T)) top_node<C, N_>(w, s); break;
            POLAR_TN(1) POLAR_TN(2) POLAR_TN(3) POLAR_TN(4) POLAR_TN(5) POLAR_TN(6) POLAR_TN(7)
            POLAR_TN(8) POLAR_TN(9) POLAR_TN(10) POLAR_TN(11) POLAR_TN(12) POLAR_TN(13) POLAR_TN(14) POLAR_TN(15)
#undef POLAR_TN
            default: break;
        }
        first = false;
    }
    const int entry = (lam_top <= T) ? T + 1 : lam_top;
    switch (entry) {
#define POLAR_LS(L_)                                                                           \
    case L_:                                                                                   \
        if constexpr (L_ > T && L_ < LB) {                                                     \
            if (first) layer_step<C, L_, true>(w, s);                                          \
            else layer_step<C, L_, false>(w, s);                                               \
            first = false;                                                                     \
        }                                                                                      \
        [[fallthrough]];
        POLAR_LS(2) POLAR_LS(3) POLAR_LS(4) POLAR_LS(5) POLAR_LS(6) POLAR_LS(7) POLAR_LS(8) POLAR_LS(9)
#undef POLAR_LS
        default: break;
    }
    if (first) layer_to_regs<C, true>(w, s, r);
    else layer_to_regs<C, false>(w, s, r);
}

// ---- fork / prune at an unfrozen bit (PolarCode.cpp:489-607); W lanes = the list of one codeword ----
// Returns the decided bit of this lane's (possibly new) path. `permuted` is warp-uniform.
// tauq / tau: decision gaps below this (fixed point / float for plain SC) are recorded in the warp's margin slots.
template <class C>
__device__ __forceinline__ uint32_t info_step(const Warp& w, Lane& s, float lam_n, int L, int& sp, bool& permuted,
                                              int& src_lane, uint32_t tauq, float tau) {
    constexpr int W = C::W;
    const int lane = w.lane;
    const int gbase = lane & ~(W - 1), slot = lane & (W - 1);
    const float ax = fabsf(lam_n);
    const bool neg = lam_n < 0.0f;
    permuted = false;
    src_lane = lane;
    if constexpr (W == 1) {
        // plain SC (list 1): the better of the two forks survives, fork 0 on a tie (index order, PolarCode.cpp:543-553),
        // i.e. the sign of the leaf LLR; no metric is kept at all (a single path needs none). Margin = |LLR|.
        s.mg = min(s.mg, __float_as_uint(ax));          // non-negative floats order like their bit patterns
        return neg ? 1u : 0u;
    }
    // both fork metrics from one log1p(exp(-|x|)) (softplus_ref(-|x|) and softplus_ref(+|x|)), as fixed-point increments
    const float t = log1p_exp_neg(ax);
#if POLAR_MINSUM
    const uint32_t klo = s.pm, khi = q_add(s.pm, q_of(ax));                       // hardware-friendly metric
#else
    // The double reference's corner cases need no code here: its log(1 + e^-|x|) rounds to exactly 0 from |x| = 36.74 on
    // (1e-16, far below the fixed point's 6e-8 step: q_of gives 0 too), and its +inf from |x| = 709.78 on (exp overflow)
    // lies beyond the saturation at 256, like every other hopeless fork.
    const uint32_t klo = q_add(s.pm, q_of(t));                                    // likely fork
    const uint32_t khi = q_add(s.pm, q_of(ax + t));                               // unlikely fork
#endif
    const uint32_t m0 = neg ? khi : klo, m1 = neg ? klo : khi;
    const unsigned act = gballot<W>(s.active, gbase);
    const int A = __popc(act);
    bool keep0 = s.active, keep1 = s.active;
    // keep the L best forks under (metric asc, fork index asc)
    const bool like1 = m1 < m0;                            // likely fork is bit 1
    if constexpr (W == 32) {
        if (2 * A > L) {
            if (A == L) {
                // common exit: every unlikely fork is strictly worse than every likely fork
                const unsigned kb = gmin<W>(s.active ? khi : 0xFFFFFFFFu);
                const unsigned ka = gmax<W>(s.active ? klo : 0u);
                if (kb > ka) {
                    if (s.active) s.pm = klo;
                    note_gap<W>(w, s, kb - ka, tauq);
                    return like1 ? 1u : 0u;
                }
            }
            const unsigned lk = __ballot_sync(FULL_MASK, like1);
            unsigned keptA = act, keptB = 0;
            int count = A;
            // margin bookkeeping: last promoted / last demoted metric, and the pair the loop stopped at
            // (0 and kQSat are the identities of the max / min below)
            unsigned kbl = 0u, kal = kQSat, kbn = kQSat, kan = 0u;
            while (true) {
                const unsigned candB = act & ~keptB;
                if (candB == 0) {                          // every unlikely fork was kept
                    kan = gmax<W>(((keptA >> lane) & 1u) ? klo : 0u);
                    break;
                }
                const unsigned kb = gmin<W>(((candB >> lane) & 1u) ? khi : 0xFFFFFFFFu);
                const unsigned eqb = __ballot_sync(FULL_MASK, ((candB >> lane) & 1u) && khi == kb);
                const int bl = __ffs(eqb) - 1;             // lowest lane = lowest fork index among equals
                if (count < L) { keptB |= 1u << bl; ++count; kbl = kb; continue; }
                const unsigned ka = gmax<W>(((keptA >> lane) & 1u) ? klo : 0u);
                const unsigned eqa = __ballot_sync(FULL_MASK, ((keptA >> lane) & 1u) && klo == ka);
                const int al = 31 - __clz(eqa);            // highest lane = highest fork index among equals
                const int idxb = 2 * bl + (((lk >> bl) & 1u) ? 0 : 1);
                const int idxa = 2 * al + (((lk >> al) & 1u) ? 1 : 0);
                const bool better = (kb < ka) || (kb == ka && idxb < idxa);
                if (!better) { kbn = kb; kan = ka; break; }
                keptB |= 1u << bl;
                keptA &= ~(1u << al);
                kbl = kb; kal = ka;
            }
            note_gap<W>(w, s, min(kbn, kal) - max(kan, kbl), tauq);  // best dropped fork - worst kept fork
            const bool ka_ = (keptA >> lane) & 1u, kb_ = (keptB >> lane) & 1u;
            keep0 = like1 ? kb_ : ka_;
            keep1 = like1 ? ka_ : kb_;
        }
    } else {
        // several codewords per warp: every lane runs the same rounds, a codeword that is finished idles
        const unsigned lk = gballot<W>(like1, gbase);
        unsigned keptA = act, keptB = 0;
        int count = A;
        bool done = !(2 * A > L);
        unsigned kbl = 0u, kal = kQSat, kbn = kQSat, kan = 0u;   // as above
        while (true) {
            const unsigned candB = act & ~keptB;
            if (!__any_sync(FULL_MASK, !done)) break;
            const bool cb = (candB >> slot) & 1u, ca = (keptA >> slot) & 1u;
            const unsigned kb = gmin<W>(cb ? khi : 0xFFFFFFFFu);
            const unsigned eqb = gballot<W>(cb && khi == kb, gbase);
            const unsigned ka = gmax<W>(ca ? klo : 0u);
            const unsigned eqa = gballot<W>(ca && klo == ka, gbase);
            if (!done) {
                if (candB == 0) { done = true; kan = ka; }       // every unlikely fork was kept
                else {
                    const int bl = __ffs(eqb) - 1;
                    if (count < L) { keptB |= 1u << bl; ++count; kbl = kb; }
                    else {
                        const int al = 31 - __clz(eqa);
                        const int idxb = 2 * bl + (((lk >> bl) & 1u) ? 0 : 1);
                        const int idxa = 2 * al + (((lk >> al) & 1u) ? 1 : 0);
                        const bool better = (kb < ka) || (kb == ka && idxb < idxa);
                        if (!better) { done = true; kbn = kb; kan = ka; }
                        else { keptB |= 1u << bl; keptA &= ~(1u << al); kbl = kb; kal = ka; }
                    }
                }
            }
        }
        if (2 * A > L) {
            note_gap<W>(w, s, min(kbn, kal) - max(kan, kbl), tauq);
            const bool ka_ = (keptA >> slot) & 1u, kb_ = (keptB >> slot) & 1u;
            keep0 = like1 ? kb_ : ka_;
            keep1 = like1 ? ka_ : kb_;
        }
    }
    const bool kill = s.active && !keep0 && !keep1;
    const bool clone = keep0 && keep1;
    uint32_t u = 0;
    if (!__any_sync(FULL_MASK, kill || clone)) {
        if (s.active) { u = keep1 ? 1u : 0u; s.pm = keep1 ? m1 : m0; }
        return u;
    }
    permuted = true;
    const unsigned Kg = gballot<W>(kill, gbase);
    const unsigned Cg = gballot<W>(clone, gbase);
    const int nk = __popc(Kg), nc = __popc(Cg);
    const unsigned lt = (1u << slot) - 1u;
    if (kill) w.stack[gbase + sp + __popc(Kg & lt)] = (unsigned char)slot;   // kills pushed in ascending path order
    const int sp2 = sp + nk;
    w.srcof[lane] = (unsigned char)lane;
    __syncwarp();
    if (clone) {                                                             // clones pop, for ascending l
        const int tgt = w.stack[gbase + sp2 - 1 - __popc(Cg & lt)];
        w.srcof[gbase + tgt] = (unsigned char)lane;
    }
    sp = sp2 - nc;
    __syncwarp();
    src_lane = w.srcof[lane];
    const bool is_new = (src_lane != lane);
    const uint32_t src_m1 = __shfl_sync(FULL_MASK, m1, src_lane);
    const unsigned long long src_px = __shfl_sync(FULL_MASK, s.px, src_lane);
    const unsigned long long src_ps = __shfl_sync(FULL_MASK, s.ps, src_lane);
    const uint32_t src_sreg = __shfl_sync(FULL_MASK, s.sreg, src_lane);
    if (is_new) {
        s.active = true; s.pm = src_m1; u = 1u; s.px = src_px; s.ps = src_ps; s.sreg = src_sreg;
    } else if (kill) {
        s.active = false; s.pm = 0u;
    } else if (s.active) {
        u = keep0 ? 0u : 1u;
        s.pm = keep0 ? m0 : m1;
    }
    return u;
}

// partial-sum chain above the register layers (runs after 1 bit in 32): out of line to keep the
// per-bit code small
template <class C>
__device__ __noinline__ unsigned long long deep_chain(uint32_t* gs, uint32_t* ss, int lane, unsigned long long ps,
                                                      uint32_t P, int t) {
    constexpr int NLOG = C::NLOG;
    Warp w;
    w.gs = gs; w.ss = ss; w.lane = lane;
    const int lam_end = NLOG - t;
    int lam = NLOG - 5;
    uint32_t* D = sbase_rt<C>(w, lam_end) + lane;
    const int Wd = 1 << (t - 5);
    __syncwarp();       // every lane is done reading the columns that are about to be overwritten
    D[(Wd - 1) * 32] = P;
    for (; lam > lam_end; --lam) {
        const int mw = 1 << (NLOG - lam - 5);
        const int base = Wd - mw;
        const uint32_t* S = sbase_rt<C>(w, lam) + get_ptr(ps, lam - 1);
        for (int x = 0; x < mw; ++x) D[(base - mw + x) * 32] = S[x * 32] ^ D[(base + x) * 32];
    }
    if (lam_end >= 1) ps = set_ptr(ps, lam_end - 1, lane);
    return ps;
}

// ---- partial sums after an odd bit (PolarCode.cpp:457-473) ----
template <class C>
__device__ __forceinline__ void update_partial_sums(const Warp& w, Lane& s, int phi, uint32_t u) {
    constexpr int NLOG = C::NLOG;
#if POLAR_PS_SHORTCUT
    if ((phi & 3) == 1) {
        // every other odd leaf closes only a pair: [left ^ right | right] into the 2-bit field of layer NLOG-1
        const uint32_t P2 = ((s.sreg ^ u) & 1u) | (u << 1);
        s.sreg = (s.sreg & ~6u) | (P2 << 1);
        return;
    }
#endif
    const int t = __ffs(~phi) - 1;               // trailing ones, 1..NLOG
    uint32_t P = u;
    const int kmax = t < 5 ? t : 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        if (k < kmax) {
            const uint32_t Sk = (s.sreg >> ((1 << k) - 1)) & ((1u << (1 << k)) - 1u);
            P = (Sk ^ P) | (P << (1 << k));
        }
    }
    if (t < 5) {
        const int sh = (1 << t) - 1;
        const uint32_t msk = ((1u << (1 << t)) - 1u) << sh;
        s.sreg = (s.sreg & ~msk) | (P << sh);
        return;
    }
    s.ps = deep_chain<C>(w.gs, w.ss, w.lane, s.ps, P, t);
}

template <class C, int WPB, int BPS>
__global__ void __launch_bounds__(WPB * 32, BPS) scl_fast_kernel(const Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int N = C::N, NW = C::NW;
    Warp w;
    w.lane = threadIdx.x & 31;
    const int lane = w.lane;
    const int wib = threadIdx.x >> 5;
    const int gwarp = blockIdx.x * WPB + wib;
    const int total_warps = gridDim.x * WPB;
    unsigned char* my = smem_raw + (size_t)C::SMEM_PER_WARP * wib;
    w.sx = reinterpret_cast<float*>(my);
    w.ss = reinterpret_cast<uint32_t*>(my + C::SX_ROWS * 128);
    w.srcof = my + (C::SX_ROWS + C::SS_ROWS) * 128;
    w.stack = w.srcof + 32;
    w.mg = reinterpret_cast<uint32_t*>(w.stack + 32);
    constexpr int W = C::W, G = C::G;
    w.tm = 0;
    // NQ warps share one SM sub-partition (warp index mod 4), i.e. one L0 instruction cache and one TMEM lane quadrant
    constexpr int NQ = WPB / 4;
    static_assert(WPB % 4 == 0, "whole groups of four warps (one per TMEM lane quadrant)");
    constexpr int TM_ALLOC = C::TM_COLS * NQ <= 32 ? 32 : C::TM_COLS * NQ <= 64 ? 64 : C::TM_COLS * NQ <= 128 ? 128
                             : C::TM_COLS * NQ <= 256 ? 256 : 512;
    if constexpr (C::TM) {
        static_assert(C::TM_COLS * NQ <= 512, "tensor memory has 512 columns");
        __shared__ uint32_t tm_base;
        if (wib == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"((uint32_t)__cvta_generic_to_shared(&tm_base)), "r"((uint32_t)TM_ALLOC) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        w.tm = tm_base + ((uint32_t)((wib & 3) * 32) << 16) + (uint32_t)((wib >> 2) * C::TM_COLS);
    }
    const int gbase = lane & ~(W - 1), slot = lane & (W - 1), grp_in_warp = lane / W;
    w.gx = a.gx + C::GX_FLOATS * gwarp;
    w.xs = w.gx + C::XST_OFF + (size_t)grp_in_warp * C::XS_FLOATS;
    w.xst = w.gx + C::XST_OFF;
    w.cht = w.gx + C::CHT_OFF;
    w.g = grp_in_warp;
    w.gs = a.gs + C::GS_WORDS * gwarp;
    const int L = a.L, KW = (a.K + 31) >> 5;
    const int c0 = L - 1;                          // first path popped from the free stack (PolarCode.cpp:250-263)

    // Every warp runs the same number of rounds. With more than one warp per sub-partition (NQ > 1) the warps that
    // share a sub-partition start each round together (named barrier 1 + quadrant): they then walk through the same
    // code at about the same time and share the sub-partition's small L0 instruction cache instead of evicting each
    // other's loops -- unsynchronised warps drift apart over the rounds and instruction fetch becomes the largest
    // single stall (profiles/r01_icache_*).
    const int groups = (a.B + G - 1) / G;
    const int rounds = (groups + total_warps - 1) / total_warps;
    for (int rnd = 0; rnd < rounds; ++rnd) {
        const int grp = gwarp + rnd * total_warps;
        if constexpr (NQ > 1) asm volatile("bar.sync %0, %1;" :: "r"(1 + (wib & 3)), "r"(NQ * 32) : "memory");
        if (grp >= groups) continue;
        const int cw = grp * G + grp_in_warp;
        const bool valid = cw < a.B;
        w.chan = a.llr + (size_t)(valid ? cw : a.B - 1) * N;
        const float* first_row = a.llr + (size_t)grp * G * N;     // the warp's codewords are consecutive rows
        const int last_g = (a.B - 1 - grp * G) < (G - 1) ? (a.B - 1 - grp * G) : (G - 1);
        Lane s;
        s.active = valid && (slot == c0);
        s.pm = 0u; s.px = 0; s.ps = 0; s.sreg = 0;
        w.stack[lane] = (unsigned char)slot;            // free stack 0..L-2 of every codeword (entries >= sp are don't-care)
        if constexpr (W == 32) { if (lane == 0) *w.mg = kQSat; }   // smallest decision margin so far: none recorded
        s.mg = (W == 1) ? 0x7F800000u : kQSat;
        __syncwarp();
        int sp = L - 1;
        float lam_n = 0.0f;
        Sub r;
#pragma unroll
        for (int i = 0; i < 16; ++i) r.x4[i] = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.x3[i] = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) r.x2[i] = 0.0f;
        r.x1[0] = r.x1[1] = 0.0f;

        bool have_x4 = false;
        if (W == 32 && a.PA > 0) {
            // out of line, runs once per codeword. The metric of the frozen prefix is common to every path that will ever
            // exist, so it is dropped (only its +inf case, PolarCode.cpp:483 with exp overflowing, is kept: saturated)
            s.pm = (phase_a<C>(w, a.PA, c0) < CUDART_INF_F) ? 0u : kQSat;
            const float* x4src = w.xs + C::XS_FLOATS + C::MT;
#pragma unroll
            for (int i = 0; i < 16; ++i) r.x4[i] = x4src[i];
            const unsigned long long rep = 0x0084210842108421ull * (unsigned long long)c0;   // c0 in every 5-bit field
            s.px = rep; s.ps = rep; s.sreg = 0u;
            have_x4 = true;
        }
#pragma unroll 1
        for (int phi0 = (W == 32 ? a.PA : 0); phi0 < N; phi0 += 16) {
            if (!have_x4) descend_block<C>(w, s, r, phi0, c0, valid, first_row, last_g);
            have_x4 = false;
            if constexpr (W != 1) {
                // renormalise: metrics are kept relative to the best path of the list (exact integer subtraction)
                const uint32_t base = gmin<W>(s.active ? s.pm : kQSat);
                if (s.active && s.pm != kQSat) s.pm -= base;
            }
            const uint32_t frozen16 = (a.frozen_words[phi0 >> 5] >> (phi0 & 31)) & 0xFFFFu;
            if constexpr (POLAR_UNROLL2 && (W == 32 || W == 1)) {
            // (measured: +3..4 % with one codeword per warp and for plain SC, -7 % at list 4, whose selection loop is larger)
            // one leaf; R = j mod POLAR_UNROLL2 is static (the loop below is unrolled), j is not
            auto leaf = [&](const int j, auto r_c) {
                constexpr int R = decltype(r_c)::value;
                constexpr bool ODD = (R & 1) != 0;
                // levels 3..0 of the register subtree: g at level ctz(j), f below it (all f for j = 0)
                if constexpr (ODD) {
                    sub_step<0>(r, s.sreg, true, lam_n);
                } else if constexpr (POLAR_UNROLL2 >= 4 && (R & 3) == 2) {
                    sub_step<1>(r, s.sreg, true, lam_n);
                    sub_step<0>(r, s.sreg, false, lam_n);
                } else {
                    const int e = j ? (__ffs(j) - 1) : 3;
                    bool fg = (j != 0);
                    if (e >= 3) { sub_step<3>(r, s.sreg, fg, lam_n); fg = false; }
                    if (e >= 2) { sub_step<2>(r, s.sreg, fg, lam_n); fg = false; }
                    sub_step<1>(r, s.sreg, fg, lam_n);
                    sub_step<0>(r, s.sreg, false, lam_n);
                }
                uint32_t u = 0;
                if ((frozen16 >> j) & 1u) {
                    if constexpr (W != 1) { if (s.active) s.pm = q_add(s.pm, q_of(softplus_q(-lam_n))); }   // PolarCode.cpp:475-487
                } else {
                    bool permuted; int src_lane;
                    u = info_step<C>(w, s, lam_n, L, sp, permuted, src_lane, a.tauq, a.tau);
                    if (permuted) {
                        // a cloned path takes over its parent's subtree registers -- those that are still live:
                        // x4 is read again at leaf 8, x3 at leaves 4 and 12, x2 at leaves 2 mod 4, x1 at odd leaves
                        if (j < 8) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) r.x4[i] = __shfl_sync(FULL_MASK, r.x4[i], src_lane);
                        }
                        if ((j & 7) < 4) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) r.x3[i] = __shfl_sync(FULL_MASK, r.x3[i], src_lane);
                        }
                        if ((j & 3) < 2) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) r.x2[i] = __shfl_sync(FULL_MASK, r.x2[i], src_lane);
                        }
                        if constexpr (!ODD) {
                            r.x1[0] = __shfl_sync(FULL_MASK, r.x1[0], src_lane);
                            r.x1[1] = __shfl_sync(FULL_MASK, r.x1[1], src_lane);
                        }
                    }
                }
                if constexpr (!ODD) s.sreg = (s.sreg & ~1u) | u;
                else update_partial_sums<C>(w, s, phi0 + j, u);
            };
#pragma unroll 1
            for (int j = 0; j < 16; j += (POLAR_UNROLL2 >= 4 ? 4 : 2)) {
                leaf(j, std::integral_constant<int, 0>{});
                leaf(j + 1, std::integral_constant<int, 1>{});
                if constexpr (POLAR_UNROLL2 >= 4) {
                    leaf(j + 2, std::integral_constant<int, 2>{});
                    leaf(j + 3, std::integral_constant<int, 3>{});
                }
            }
            } else {
#pragma unroll 1
            for (int j = 0; j < 16; ++j) {
                // levels 3..0 of the register subtree: g at level ctz(j), f below it (all f for j = 0)
                if (j & 1) {
                    sub_step<0>(r, s.sreg, true, lam_n);
                } else {
                    const int e = j ? (__ffs(j) - 1) : 3;
                    bool fg = (j != 0);
                    if (e >= 3) { sub_step<3>(r, s.sreg, fg, lam_n); fg = false; }
                    if (e >= 2) { sub_step<2>(r, s.sreg, fg, lam_n); fg = false; }
                    sub_step<1>(r, s.sreg, fg, lam_n);
                    sub_step<0>(r, s.sreg, false, lam_n);
                }
                uint32_t u = 0;
                if ((frozen16 >> j) & 1u) {
                    if constexpr (W != 1) { if (s.active) s.pm = q_add(s.pm, q_of(softplus_q(-lam_n))); }   // PolarCode.cpp:475-487
                } else {
                    bool permuted; int src_lane;
                    u = info_step<C>(w, s, lam_n, L, sp, permuted, src_lane, a.tauq, a.tau);
                    if (permuted) {
                        // a cloned path takes over its parent's live subtree registers
#pragma unroll
                        for (int i = 0; i < 16; ++i) r.x4[i] = __shfl_sync(FULL_MASK, r.x4[i], src_lane);
#pragma unroll
                        for (int i = 0; i < 8; ++i) r.x3[i] = __shfl_sync(FULL_MASK, r.x3[i], src_lane);
#pragma unroll
                        for (int i = 0; i < 4; ++i) r.x2[i] = __shfl_sync(FULL_MASK, r.x2[i], src_lane);
                        r.x1[0] = __shfl_sync(FULL_MASK, r.x1[0], src_lane);
                        r.x1[1] = __shfl_sync(FULL_MASK, r.x1[1], src_lane);
                    }
                }
                if ((j & 1) == 0) s.sreg = (s.sreg & ~1u) | u;
                else update_partial_sums<C>(w, s, phi0 + j, u);
            }
            }
        }
        __syncwarp();

        // ---- u-hat = packed polar transform of the re-encoded codeword (partial-sum layer 0) ----
        uint32_t* D = sbase<C, 0>(w) + lane;
        bool pass = true;
#if POLAR_TAIL_REGS
        {
            // chunks of up to 32 words are transformed in registers (the subtree registers are dead here): every
            // chunk is 32 independent loads, not a chain of read-modify-writes through L2
            constexpr int CH = NW < 32 ? NW : 32;
            for (int sw = NW >> 1; sw >= CH; sw >>= 1) {              // strides between chunks (N > 1024 only)
#pragma unroll 1
                for (int i0 = 0; i0 < NW; i0 += 8) {
                    if (i0 & sw) continue;
                    uint32_t lo[8], hi[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) { lo[j] = D[(i0 + j) * 32]; hi[j] = D[(i0 + j + sw) * 32]; }
#pragma unroll
                    for (int j = 0; j < 8; ++j) D[(i0 + j) * 32] = lo[j] ^ hi[j];
                }
            }
            unsigned long long odd = 0ull;                            // parity rows with an odd sum so far
            const int crc_fast = a.crc < 64 ? a.crc : 64;
#pragma unroll 1
            for (int base = 0; base < NW; base += CH) {
                uint32_t x[CH];
#pragma unroll
                for (int i = 0; i < CH; ++i) x[i] = D[(base + i) * 32];
#pragma unroll
                for (int sw = CH >> 1; sw >= 1; sw >>= 1)
#pragma unroll
                    for (int i = 0; i < CH; ++i)
                        if ((i & sw) == 0) x[i] ^= x[i + sw];
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    uint32_t v = x[i];
                    v ^= (v >> 16) & 0x0000FFFFu;
                    v ^= (v >> 8) & 0x00FF00FFu;
                    v ^= (v >> 4) & 0x0F0F0F0Fu;
                    v ^= (v >> 2) & 0x33333333u;
                    v ^= (v >> 1) & 0x55555555u;
                    x[i] = v;
                    D[(base + i) * 32] = v;
                }
#pragma unroll 1
                for (int r = 0; r < crc_fast; ++r) {                  // PolarCode.cpp:93-108
                    const uint32_t* m = a.crc_masks + r * NW + base;
                    uint32_t acc = 0;
#pragma unroll
                    for (int i = 0; i < CH; ++i) acc ^= x[i] & m[i];
                    odd ^= (unsigned long long)(__popc(acc) & 1) << r;
                }
            }
            if (odd) pass = false;
            for (int r = 64; r < a.crc; ++r) {                        // more than 64 parity rows: from memory
                uint32_t acc = 0;
                for (int i = 0; i < NW; ++i) acc ^= D[i * 32] & a.crc_masks[r * NW + i];
                if (__popc(acc) & 1) pass = false;
            }
        }
#else
        for (int sw = NW >> 1; sw >= 1; sw >>= 1)
            for (int i = 0; i < NW; ++i)
                if ((i & sw) == 0) D[i * 32] ^= D[(i + sw) * 32];
        for (int i = 0; i < NW; ++i) {
            uint32_t x = D[i * 32];
            x ^= (x >> 16) & 0x0000FFFFu;
            x ^= (x >> 8) & 0x00FF00FFu;
            x ^= (x >> 4) & 0x0F0F0F0Fu;
            x ^= (x >> 2) & 0x33333333u;
            x ^= (x >> 1) & 0x55555555u;
            D[i * 32] = x;
        }
        for (int r = 0; r < a.crc; ++r) {                 // PolarCode.cpp:93-108
            uint32_t acc = 0;
            for (int i = 0; i < NW; ++i) acc ^= D[i * 32] & a.crc_masks[r * NW + i];
            if (__popc(acc) & 1) pass = false;
        }
#endif
        // ---- final pick, PolarCode.cpp:609-644 ----
        const unsigned act = gballot<W>(s.active, gbase);
        const unsigned passm = gballot<W>(s.active && pass, gbase);
        const bool use_parity = (a.crc != 0) && (passm != 0);
        const bool eligible = s.active && (use_parity ? pass : true) && (W == 1 || s.pm != kQSat);
        const unsigned best = gmin<W>(eligible ? s.pm : 0xFFFFFFFFu);
        const unsigned cand = gballot<W>(eligible && s.pm == best, gbase);
        const int win = cand ? (__ffs(cand) - 1) : 0;
        const bool win_active = (act >> win) & 1u;
        __syncwarp();
        bool flagme = false;
        if (a.margin != nullptr || a.flag_list != nullptr) {
            // the final pick is a decision too: runner-up metric - winner's metric (0 when two paths tie, or when every
            // candidate is saturated and the reference's choice cannot be reproduced from these metrics)
            uint32_t mgq;
            if constexpr (W == 32) mgq = *w.mg;
            else if constexpr (W == 1) mgq = q_of(__uint_as_float(s.mg));
            else mgq = s.mg;
            if constexpr (W > 1) {
                const unsigned second = gmin<W>((eligible && slot != win) ? s.pm : 0xFFFFFFFFu);
                if (cand == 0) mgq = 0u;
                else if (second != 0xFFFFFFFFu) mgq = min(mgq, second - best);
            }
            flagme = a.flag_list != nullptr && mgq < a.tauq_flag;
            if (slot == 0 && valid) {
                if (a.margin != nullptr) a.margin[a.cw_base + cw] = (mgq == kQSat) ? CUDART_INF_F : (float)mgq * (1.0f / kQScale);
                if (flagme) a.flag_list[atomicAdd(a.flag_count, 1)] = a.cw_base + cw;
            }
        }
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
            const int cwg = grp * G + g;
            if (cwg >= a.B) break;
            const int wl = g * W + __shfl_sync(FULL_MASK, win, g * W);
            const bool wa = __shfl_sync(FULL_MASK, (int)win_active, g * W);
            const bool fl = __shfl_sync(FULL_MASK, (int)flagme, g * W);
            bool differs = false;                            // from the transmitted info bits (a.truth)
            const uint32_t* U = sbase<C, 0>(w) + wl;
#if POLAR_TAIL_REGS
            // the winner's u-hat spread over the lanes (word 32 q + lane in uw[q]); a bit lookup is then a shuffle,
            // not a second dependent trip to L2
            constexpr int NR = (NW + 31) / 32;
            uint32_t uw[NR];
#pragma unroll
            for (int q = 0; q < NR; ++q) uw[q] = (32 * q + lane < NW) ? U[(32 * q + lane) * 32] : 0u;
            for (int t0 = 0; t0 < KW; t0 += 32) {           // decoded[j] = u-hat[order[j]], PolarCode.cpp:171-174
                const int t = t0 + lane;
                const int jmax = (t < KW) ? min(32, a.K - 32 * t) : 0;
                uint32_t word = 0;
#pragma unroll 8
                for (int i = 0; i < 32; ++i) {
                    const int pos = (i < jmax) ? (int)a.info_order[32 * t + i] : 0;
                    const int wi = pos >> 5;
                    uint32_t v = __shfl_sync(FULL_MASK, uw[0], wi & 31);
#pragma unroll
                    for (int q = 1; q < NR; ++q) {
                        const uint32_t vq = __shfl_sync(FULL_MASK, uw[q], wi & 31);
                        if ((wi >> 5) == q) v = vq;
                    }
                    if (i < jmax) word |= ((v >> (pos & 31)) & 1u) << i;
                }
                if (t < KW) {
                    if (!wa) word = 0u;
                    if (a.out != nullptr) a.out[(size_t)cwg * KW + t] = word;
                    if (a.truth != nullptr) differs |= word != a.truth[(size_t)cwg * KW + t];
                }
            }
#else
            for (int t = lane; t < KW; t += 32) {           // decoded[j] = u-hat[order[j]], PolarCode.cpp:171-174
                uint32_t word = 0;
                if (wa) {
                    const int jmax = min(32, a.K - 32 * t);
                    for (int i = 0; i < jmax; ++i) {
                        const int pos = a.info_order[32 * t + i];
                        word |= ((U[(pos >> 5) * 32] >> (pos & 31)) & 1u) << i;
                    }
                }
                if (a.out != nullptr) a.out[(size_t)cwg * KW + t] = word;
                if (a.truth != nullptr) differs |= word != a.truth[(size_t)cwg * KW + t];
            }
#endif
            if (a.truth != nullptr) {                        // PolarCode.cpp:758-769
                const bool block_error = __any_sync(FULL_MASK, differs);
                if (lane == 0 && block_error && !fl)
                    atomicAdd(a.err + (int)((unsigned long long)(a.first_index + a.cw_base + cwg) % (unsigned)a.n_ebno), 1ull);
            }
        }
        __syncwarp();
    }
    if constexpr (C::TM) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (wib == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                         :: "r"(w.tm & 0x0000FFFFu), "r"((uint32_t)TM_ALLOC) : "memory");
    }
}

}  // namespace POLAR_FAST_NS
