// scl_wide.cuh -- list sizes 33..127: one thread BLOCK (64 or 128 threads) per codeword, thread = list path.
//
// The reference accepts any list size below 128 (its loop counters are uint8_t, PolarC/PolarCode.cpp:497-605);
// a warp only has 32 lanes, so for longer lists the organisation of scl_decode_kernel (polar_b200.cu) is kept
// but widened from a warp to a block: rows are [beta][W] (W = 64 or 128 columns), column pointers take 8 bits,
// and what the warp kernel does with shuffles and ballots (fork ranking, free-path stack, clone) goes through
// a small exchange area in shared memory between block barriers. Every rule of order is the reference's:
// fork selection = the rho best forks under (metric, fork index) (:528-553), kills pushed in ascending path
// order, clones popping in ascending path order (:555-570, :274-303), first path = L-1 (:250-263), final pick
// = strictly smaller metric / lowest index, parity filter with fall-through (:609-644).
//
// The kernel is written once over a "domain": LLR domain (decode_scl_llr, float or double) or the reference's
// probability domain (decode_scl_p1, PolarCode.cpp:110-128, 375-420: pairs of likelihoods per tree entry, every
// refreshed layer divided by the maximum over all live paths, forks ranked by the likelihood pair itself).
//
// Included by polar_b200.cu after Arith<>, rmin/rmax and kMaxN are defined.
#pragma once

namespace wide {

// ---- LLR domain: PolarCode.cpp:422-455 (f / g), :483, :505-506 (metrics) ----
template <class R>
struct LlrDom {
    using Real = R;
    using Val = R;
    static constexpr bool kNormalise = false;
    static __device__ __forceinline__ void load(const R* llr, const R*, size_t base, int i, Val& x0, Val& x1) {
        x0 = llr[base + 2 * i]; x1 = llr[base + 2 * i + 1];
    }
    static __device__ __forceinline__ Val f(Val a, Val b) { return Arith<R>::f(a, b); }
    static __device__ __forceinline__ Val g(Val a, Val b, uint32_t bit) { return b + (bit ? -a : a); }
    static __device__ __forceinline__ R vmax(Val) { return 0; }
    static __device__ __forceinline__ Val scale(Val v, R) { return v; }
    // fork metrics, smaller = better (m = -probForks)
    static __device__ __forceinline__ void forks(R pm, Val leaf, R& m0, R& m1) {
        m0 = pm + Arith<R>::softplus(-leaf);
        m1 = pm + Arith<R>::softplus(leaf);
    }
    static __device__ __forceinline__ R frozen(R pm, Val leaf) { return pm + Arith<R>::softplus(-leaf); }
    static __device__ __forceinline__ R pick_init() { return Arith<R>::inf(); }
};

// ---- probability domain (double, the reference's own type): PolarCode.cpp:375-420, :510-514, :631-637 ----
// Products and sums are written with explicit round-to-nearest intrinsics so that nvcc cannot contract
// a*b + c*d into an FMA: the reference (g++ -O2, x86-64) rounds every product, and the decisions of exact-tie
// inputs depend on that last bit.
struct P2 { double p0, p1; };
struct ProbDom {
    using Real = double;
    using Val = P2;
    static constexpr bool kNormalise = true;
    static __device__ __forceinline__ void load(const double* p0, const double* p1, size_t base, int i, Val& x0, Val& x1) {
        x0.p0 = p0[base + 2 * i]; x0.p1 = p1[base + 2 * i];             // p_0[2 beta] = p0[beta], p_0[2 beta + 1] = p1[beta] (:121-124)
        x1.p0 = p0[base + 2 * i + 1]; x1.p1 = p1[base + 2 * i + 1];
    }
    static __device__ __forceinline__ Val f(Val a, Val b) {               // :392-396
        Val y;
        y.p0 = __dmul_rn(0.5, __dadd_rn(__dmul_rn(a.p0, b.p0), __dmul_rn(a.p1, b.p1)));
        y.p1 = __dmul_rn(0.5, __dadd_rn(__dmul_rn(a.p1, b.p0), __dmul_rn(a.p0, b.p1)));
        return y;
    }
    static __device__ __forceinline__ Val g(Val a, Val b, uint32_t bit) { // :398-402
        Val y;
        y.p0 = __dmul_rn(__dmul_rn(0.5, bit ? a.p1 : a.p0), b.p0);
        y.p1 = __dmul_rn(__dmul_rn(0.5, bit ? a.p0 : a.p1), b.p1);
        return y;
    }
    static __device__ __forceinline__ double vmax(Val v) { return fmax(v.p0, v.p1); }
    static __device__ __forceinline__ Val scale(Val v, double sigma) {    // :415-416
        Val y; y.p0 = __ddiv_rn(v.p0, sigma); y.p1 = __ddiv_rn(v.p1, sigma); return y;
    }
    // "metric" = minus the likelihood, so that smaller = better as in the LLR domain; negation is exact
    static __device__ __forceinline__ void forks(double, Val leaf, double& m0, double& m1) { m0 = -leaf.p0; m1 = -leaf.p1; }
    static __device__ __forceinline__ double frozen(double, Val leaf) { return -leaf.p0; }
    static __device__ __forceinline__ double pick_init() { return 0.0; }  // p_p1 = 0, strictly larger wins (:612, :634)
};

template <class Dom>
struct Args {
    using Real = typename Dom::Real;
    using Val = typename Dom::Val;
    const Real* in0;             // [B][N] LLRs, or p0 in the probability domain
    const Real* in1;             // unused, or p1
    uint32_t* out;               // [B][KW]
    const uint32_t* frozen_words;
    const uint16_t* info_order;  // [K + crc]
    const uint32_t* crc_masks;   // [crc][NW] over phi
    Val* gx;                     // per-block LLR scratch rows (W values each)
    uint32_t* gs;                // per-block partial-sum scratch rows (W words each)
    unsigned long long gx_stride;// values per block
    unsigned long long gs_stride;// words per block
    int B, n, K, crc, L;
    int lamS;                    // first LLR layer kept in shared memory (1..n)
    int smem_x_rows, smem_s_rows;
    int s_off[kMaxN + 2];
};

// 8-bit column pointers, one per layer 1..n-1 (index lam-1), in two 64-bit registers.
struct Ptrs {
    unsigned long long lo, hi;
};
__device__ __forceinline__ int pget(const Ptrs& p, int idx) {
    return (int)(((idx < 8 ? p.lo : p.hi) >> (8 * (idx & 7))) & 255ull);
}
__device__ __forceinline__ void pset(Ptrs& p, int idx, unsigned col) {
    const int sh = 8 * (idx & 7);
    if (idx < 8) p.lo = (p.lo & ~(255ull << sh)) | ((unsigned long long)col << sh);
    else p.hi = (p.hi & ~(255ull << sh)) | ((unsigned long long)col << sh);
}

template <class Dom, int W>
__global__ void __launch_bounds__(W) scl_wide_kernel(const Args<Dom> a) {
    using Real = typename Dom::Real;
    using Val = typename Dom::Val;
    constexpr int NWARP = W / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, N = 1 << n, L = a.L, lamS = a.lamS;
    const int NW = (N + 31) >> 5, KW = (a.K + 31) >> 5;

    // ---- shared memory carve-up ----
    Val* sx = reinterpret_cast<Val*>(smem_raw);
    unsigned char* p = smem_raw + (size_t)a.smem_x_rows * W * sizeof(Val);
    Real* red = reinterpret_cast<Real*>(p); p += 2 * NWARP * sizeof(Real);   // block-max scratch, double-buffered
    Real* r_lo = reinterpret_cast<Real*>(p); p += NWARP * sizeof(Real);      // per-warp worst likely / best unlikely fork
    Real* r_hi = reinterpret_cast<Real*>(p); p += NWARP * sizeof(Real);
    Real* x_m0 = reinterpret_cast<Real*>(p); p += W * sizeof(Real);
    Real* x_m1 = reinterpret_cast<Real*>(p); p += W * sizeof(Real);
    unsigned long long* x_plo = reinterpret_cast<unsigned long long*>(p); p += W * 8;   // staged px.lo, px.hi, ps.lo, ps.hi
    unsigned long long* x_phi = reinterpret_cast<unsigned long long*>(p); p += W * 8;
    unsigned long long* x_slo = reinterpret_cast<unsigned long long*>(p); p += W * 8;
    unsigned long long* x_shi = reinterpret_cast<unsigned long long*>(p); p += W * 8;
    uint32_t* ss = reinterpret_cast<uint32_t*>(p); p += (size_t)a.smem_s_rows * W * 4;
    uint32_t* x_sn = reinterpret_cast<uint32_t*>(p); p += W * 4;
    int* stk = reinterpret_cast<int*>(p); p += W * 4;            // free-path stack (PolarCode.cpp:250-256)
    int* srcof = reinterpret_cast<int*>(p); p += W * 4;
    uint32_t* b_act = reinterpret_cast<uint32_t*>(p); p += NWARP * 4;
    uint32_t* b_kill = reinterpret_cast<uint32_t*>(p); p += NWARP * 4;
    uint32_t* b_clone = reinterpret_cast<uint32_t*>(p); p += NWARP * 4;
    uint32_t* b_misc = reinterpret_cast<uint32_t*>(p); p += NWARP * 4;

    Val* gx = a.gx + a.gx_stride * blockIdx.x;
    int red_parity = 0;
    // maximum over the block; one barrier per call (the scratch alternates between two buffers)
    auto block_max = [&](Real v) -> Real {
        for (int o = 16; o > 0; o >>= 1) v = rmax<Real>(v, __shfl_xor_sync(FULL_MASK, v, o));
        Real* buf = red + red_parity * NWARP;
        if (lane == 0) buf[warp] = v;
        __syncthreads();
        Real r = buf[0];
        for (int w = 1; w < NWARP; ++w) r = rmax<Real>(r, buf[w]);
        red_parity ^= 1;
        return r;
    };
    uint32_t* gs = a.gs + a.gs_stride * blockIdx.x;

    auto xrow = [&](int lam, int beta) -> Val* {
        if (lam >= lamS) return sx + (size_t)((1 << (n - lamS + 1)) - (1 << (n - lam + 1)) + beta) * W;
        return gx + (size_t)(N - (1 << (n - lam + 1)) + beta) * W;
    };
    const int lamSS = lamS < 1 ? 1 : lamS;
    auto srow = [&](int lam, int w) -> uint32_t* {
        if (lam >= lamSS) return ss + (size_t)(a.s_off[lam] + w) * W;
        return gs + (size_t)(a.s_off[lam] + w) * W;
    };
    // number of set bits below my thread / in total, over per-warp ballots staged in shared memory
    auto rank_below = [&](const uint32_t* words, uint32_t mine) -> int {
        int r = __popc(mine & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) r += __popc(words[w]);
        return r;
    };
    auto total = [&](const uint32_t* words) -> int {
        int r = 0;
        for (int w = 0; w < NWARP; ++w) r += __popc(words[w]);
        return r;
    };

    for (int cw = blockIdx.x; cw < a.B; cw += gridDim.x) {
        const size_t chan = (size_t)cw * N;
        bool active = (tid == L - 1);           // PolarCode.cpp:259-263
        Real pm = 0;
        Ptrs px = {0ull, 0ull}, ps = {0ull, 0ull};
        uint32_t s_n = 0;
        int sp = L - 1;                         // stack height (uniform)
        Val lam_n = Val();
        uint32_t frozen_word = 0;
        stk[tid] = tid;
        __syncthreads();

        for (int phi = 0; phi < N; ++phi) {
            // ---- LLR layers lam_top..n (PolarCode.cpp:422-455) ----
            const int lam_top = (phi == 0) ? 1 : n - (__ffs(phi) - 1);
            for (int lam = lam_top; lam <= n; ++lam) {
                const int M = 1 << (n - lam);
                const bool is_g = (lam == lam_top) && (phi != 0);
                const Val* src = nullptr;
                if (lam > 1) src = xrow(lam - 1, 0) + pget(px, lam - 2);
                const uint32_t* sw = nullptr;
                if (is_g && lam < n) sw = srow(lam, 0) + pget(ps, lam - 1);
                Val* dst = (lam < n) ? xrow(lam, 0) + tid : nullptr;
                Real vm = 0;
                if (active) {
                    for (int i = 0; i < M; ++i) {
                        Val x0, x1;
                        int beta = i;
                        if (lam == 1) {
                            // channel layer: reference pairs are (2k, 2k+1); the result lands at the bit-reversed position
                            Dom::load(a.in0, a.in1, chan, i, x0, x1);
                            beta = (n > 1) ? (int)(__brev((unsigned)i) >> (33 - n)) : 0;
                        } else {
                            x0 = src[(size_t)i * W];
                            x1 = src[(size_t)(i + M) * W];
                        }
                        Val y;
                        if (is_g) {
                            uint32_t bit;
                            if (lam == n) bit = s_n & 1u;
                            else bit = (sw[(size_t)(beta >> 5) * W] >> (beta & 31)) & 1u;
                            y = Dom::g(x0, x1, bit);                        // PolarCode.cpp:448-451 / :398-402
                        } else {
                            y = Dom::f(x0, x1);                             // PolarCode.cpp:438-446 / :392-396
                        }
                        if (Dom::kNormalise) vm = rmax<Real>(vm, Dom::vmax(y));
                        if (lam == n) lam_n = y; else dst[(size_t)beta * W] = y;
                    }
                }
                if (Dom::kNormalise) {
                    // PolarCode.cpp:404-418: divide the refreshed layer of every live path by the common maximum
                    const Real sigma = block_max(vm);
                    if (sigma != 0 && active) {
                        if (lam == n) lam_n = Dom::scale(lam_n, sigma);
                        else for (int i = 0; i < M; ++i) dst[(size_t)i * W] = Dom::scale(dst[(size_t)i * W], sigma);
                    }
                }
                if (lam < n) pset(px, lam - 1, tid);
            }

            // ---- leaf decision ----
            if ((phi & 31) == 0) frozen_word = a.frozen_words[phi >> 5];
            const bool frozen = (frozen_word >> (phi & 31)) & 1u;
            uint32_t u = 0;
            if (frozen) {
                if (active) pm = Dom::frozen(pm, lam_n);                    // PolarCode.cpp:475-487
            } else {
                // PolarCode.cpp:489-607; metrics kept positive (m = -probForks)
                Real m0, m1;
                Dom::forks(pm, lam_n, m0, m1);
                x_m0[tid] = m0; x_m1[tid] = m1;
                x_plo[tid] = px.lo; x_phi[tid] = px.hi; x_slo[tid] = ps.lo; x_shi[tid] = ps.hi; x_sn[tid] = s_n;
                srcof[tid] = tid;
                const uint32_t ab = __ballot_sync(FULL_MASK, active);
                {
                    // per-warp worst likely and best unlikely fork, for the common exit below
                    Real wl = active ? rmin<Real>(m0, m1) : -Arith<Real>::inf();
                    Real bu = active ? rmax<Real>(m0, m1) : Arith<Real>::inf();
                    for (int o = 16; o > 0; o >>= 1) {
                        wl = rmax<Real>(wl, __shfl_xor_sync(FULL_MASK, wl, o));
                        bu = rmin<Real>(bu, __shfl_xor_sync(FULL_MASK, bu, o));
                    }
                    if (lane == 0) { b_act[warp] = ab; r_lo[warp] = wl; r_hi[warp] = bu; }
                }
                __syncthreads();
                const int A = total(b_act);
                if (A == L) {
                    // common exit (block-uniform): the list is full and every unlikely fork is strictly worse than
                    // every likely fork, so each path keeps its likely fork; nothing is killed or cloned
                    Real wl = r_lo[0], bu = r_hi[0];
                    for (int w = 1; w < NWARP; ++w) { wl = rmax<Real>(wl, r_lo[w]); bu = rmin<Real>(bu, r_hi[w]); }
                    if (bu > wl) {
                        if (active) { u = (m1 < m0) ? 1u : 0u; pm = rmin<Real>(m0, m1); }
                        goto leaf_done;
                    }
                }
                bool keep0 = active, keep1 = active;
                if (2 * A > L && active) {
                    // keep the rho = L best of the 2A forks under (metric asc, fork index asc)
                    int r0 = 0, r1 = 0;
                    for (int w = 0; w < NWARP; ++w) {
                        uint32_t bits = b_act[w];
                        while (bits) {
                            const int j = 32 * w + (__ffs(bits) - 1);
                            bits &= bits - 1;
                            const Real o0 = x_m0[j], o1 = x_m1[j];
                            r0 += (o0 < m0) || (o0 == m0 && j < tid);
                            r0 += (o1 < m0) || (o1 == m0 && j < tid);
                            r1 += (o0 < m1) || (o0 == m1 && j <= tid);
                            r1 += (o1 < m1) || (o1 == m1 && j < tid);
                        }
                    }
                    keep0 = r0 < L;
                    keep1 = r1 < L;
                }
                const bool kill = active && !keep0 && !keep1;
                const bool clone = keep0 && keep1;
                const uint32_t kb = __ballot_sync(FULL_MASK, kill);
                const uint32_t cb = __ballot_sync(FULL_MASK, clone);
                if (lane == 0) { b_kill[warp] = kb; b_clone[warp] = cb; }
                __syncthreads();
                const int nk = total(b_kill), nc = total(b_clone);
                if ((nk | nc) == 0) {
                    if (active) { u = keep1 ? 1u : 0u; pm = keep1 ? m1 : m0; }
                } else {
                    // killPath pushes in ascending path order (PolarCode.cpp:555-560, :292)
                    if (kill) stk[sp + rank_below(b_kill, kb)] = tid;
                    __syncthreads();
                    const int sp2 = sp + nk;
                    // clonePath pops for ascending l (PolarCode.cpp:562-570, :275-276)
                    if (clone) srcof[stk[sp2 - 1 - rank_below(b_clone, cb)]] = tid;
                    sp = sp2 - nc;
                    __syncthreads();
                    const int s = srcof[tid];
                    if (s != tid) {
                        active = true; pm = x_m1[s]; u = 1u;
                        px.lo = x_plo[s]; px.hi = x_phi[s]; ps.lo = x_slo[s]; ps.hi = x_shi[s]; s_n = x_sn[s];
                    } else if (kill) {
                        active = false; pm = 0;
                    } else if (active) {
                        u = keep0 ? 0u : 1u;
                        pm = keep0 ? m0 : m1;
                    }
                }
            }

        leaf_done:
            // ---- partial sums (PolarCode.cpp:457-473), bit-packed, butterfly order ----
            if ((phi & 1) == 0) {
                s_n = u;
            } else {
                const int t = __ffs(~phi) - 1;
                const int lam_end = n - t;
                uint32_t P = u;
                int lam = n;
                while (lam > lam_end && (n - lam) < 5) {
                    const int M = 1 << (n - lam);
                    uint32_t Sw;
                    if (lam == n) Sw = s_n;
                    else Sw = srow(lam, 0)[pget(ps, lam - 1)];
                    P = ((Sw ^ P) & ((1u << M) - 1u)) | (P << M);
                    --lam;
                }
                if (lam == lam_end) {
                    srow(lam, 0)[tid] = P;
                } else {
                    const int Wd = 1 << (t - 5);
                    uint32_t* D = srow(lam_end, 0) + tid;
                    D[(size_t)(Wd - 1) * W] = P;
                    for (; lam > lam_end; --lam) {
                        const int mw = 1 << (n - lam - 5);
                        const int base = Wd - mw;
                        const uint32_t* S = srow(lam, 0) + pget(ps, lam - 1);
                        for (int w = 0; w < mw; ++w)
                            D[(size_t)(base - mw + w) * W] = S[(size_t)w * W] ^ D[(size_t)(base + w) * W];
                    }
                }
                if (lam_end >= 1) pset(ps, lam_end - 1, tid);
            }
            __syncthreads();
        }

        // ---- u-hat of every path: packed polar transform of the re-encoded codeword (layer 0), own column ----
        uint32_t* D = srow(0, 0) + tid;
        for (int sw = NW >> 1; sw >= 1; sw >>= 1)
            for (int i = 0; i < NW; ++i)
                if ((i & sw) == 0) D[(size_t)i * W] ^= D[(size_t)(i + sw) * W];
        bool pass = true;
        for (int i = 0; i < NW; ++i) {
            uint32_t w = D[(size_t)i * W];
            if (N > 16) w ^= (w >> 16) & 0x0000FFFFu;
            if (N > 8) w ^= (w >> 8) & 0x00FF00FFu;
            if (N > 4) w ^= (w >> 4) & 0x0F0F0F0Fu;
            if (N > 2) w ^= (w >> 2) & 0x33333333u;
            w ^= (w >> 1) & 0x55555555u;
            D[(size_t)i * W] = w;
        }
        for (int r = 0; r < a.crc; ++r) {                                   // PolarCode.cpp:93-108
            uint32_t acc = 0;
            for (int i = 0; i < NW; ++i) acc ^= D[(size_t)i * W] & a.crc_masks[(size_t)r * NW + i];
            if (__popc(acc) & 1) pass = false;
        }
        // ---- final pick, PolarCode.cpp:609-644 ----
        x_m0[tid] = pm;
        const uint32_t ab = __ballot_sync(FULL_MASK, active);
        const uint32_t pb = __ballot_sync(FULL_MASK, active && pass);
        if (lane == 0) { b_act[warp] = ab; b_misc[warp] = pb; }
        __syncthreads();
        const bool use_parity = (a.crc != 0) && (total(b_misc) != 0);
        int win = 0;
        {
            Real best = Dom::pick_init();
            bool found = false;
            for (int w = 0; w < NWARP; ++w) {
                uint32_t bits = use_parity ? b_misc[w] : b_act[w];
                while (bits) {
                    const int j = 32 * w + (__ffs(bits) - 1);
                    bits &= bits - 1;
                    const Real m = x_m0[j];
                    if (m < best) { best = m; win = j; found = true; }      // strictly smaller, ascending index
                }
            }
            if (!found) win = 0;
        }
        const bool win_active = (b_act[win >> 5] >> (win & 31)) & 1u;
        __syncthreads();      // every column of layer 0 is final; exchange area free again

        // ---- output gather: decoded[j] = u-hat[order[j]], PolarCode.cpp:171-174 ----
        const uint32_t* U = srow(0, 0) + win;
        for (int t = tid; t < KW; t += W) {
            uint32_t word = 0;
            if (win_active) {
                const int jmax = min(32, a.K - 32 * t);
                for (int i = 0; i < jmax; ++i) {
                    const int pos = a.info_order[32 * t + i];
                    word |= ((U[(size_t)(pos >> 5) * W] >> (pos & 31)) & 1u) << i;
                }
            }
            a.out[(size_t)cw * KW + t] = word;
        }
        __syncthreads();
    }
}

}  // namespace wide
