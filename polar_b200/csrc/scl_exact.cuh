// scl_exact.cuh -- the latency-oriented double-precision decoder behind STRICT mode: one thread BLOCK per codeword.
//
// STRICT mode decodes everything with the fp32 kernels and then decodes again, in the reference's own arithmetic
// (double, literal formulas, PolarC/PolarCode.cpp:438-446, 483, 505-506), the few codewords that took a decision on a
// margin the fp32 arithmetic cannot vouch for. There are only a few hundred of those per batch, so what matters is
// how long ONE double-precision decode takes, not how many run side by side: the warp-per-codeword generic kernel
// (polar_b200.cu) needs 10+ ms for one codeword regardless of how empty the GPU is.
//
// Same organisation as the generic kernel (path-interleaved rows [beta][32], 5-bit column pointers instead of lazy
// copies, packed partial sums, butterfly order, u-hat by polar transform -- see polar_b200.cu), with two changes:
//   * the refresh of a tree layer (PolarCode.cpp:422-455) is spread over all threads of the block, work item =
//     (path, beta), whenever the layer has more than 32 items; the small layers at the bottom of the tree, the leaf
//     decision (PolarCode.cpp:475-607) and the partial-sum chain (:457-473) are done by warp 0 alone, lane = list path.
//     Block barriers are needed only around the big layers; runs of leaves that touch small layers only are decoded by
//     warp 0 without any block-wide synchronisation;
//   * column pointers and the active mask live in shared memory (every warp needs them), the rest of the per-path state
//     in warp 0's registers.
// Every rule of order is the reference's, exactly as in the other kernels: fork selection = the rho best forks under
// (metric, fork index) (:528-553), kills pushed / clones popped in ascending path order (:555-570, :274-303), first
// path = L-1 (:250-263), final pick = strictly smaller metric, lowest index, parity filter with fall-through (:609-644).
//
// Included by polar_b200.cu after polar_dev.cuh.
#pragma once

namespace exact {

constexpr int NT = 256;          // threads per block at most (the launch may use fewer: blockDim.x)

// ---- double-precision check node and metric update with short dependency chains ----
// The reference's literal formulas cost three exp, a division and a log per check node with CUDA's general-purpose
// double routines: about 70 dependent double operations, and one decode is one long chain of them. What the second
// pass has to deliver is the reference's DECISIONS, i.e. the same real-valued functions to far better than the margins
// that matter (strict mode hands it decisions closer than ~1e-5; it is accurate to ~1e-15), so it evaluates
//     f(a, b)      = sign * min + log1p(e^-|a+b|) - log1p(e^-|a-b|)      (== log((e^(a+b) + 1) / (e^a + e^b)), :438-441)
//     softplus(x)  = max(x, 0) + log1p(e^-|x|)                           (== log(1 + e^x), :483, :505-506)
// with table-driven kernels: e^-s = 2^(k/64) * 2^(k div 64) * p5(r), log1p(e) = -log(c_j) + p6((1 + e) c_j - 1), about 20
// dependent operations per check node. Codewords that take a decision on a double-precision margin below 1e-9 (exact
// cancellations, whose outcome in the reference is decided by the last-bit rounding of its literal formulas) are passed
// on to the literal-formula kernel (list2).
struct Tables {
    double exp2j[64];        // 2^(j/64)
    double rcp[132];         // c_j = 1 / (1 + (j + 1/2) / 128)
    double mlog[132];        // -log(c_j)
};
__device__ __forceinline__ void build_tables(Tables* t, int tid, int nt) {
    for (int j = tid; j < 64; j += nt) t->exp2j[j] = exp2((double)j / 64.0);
    for (int j = tid; j < 129; j += nt) {
        const double c = 1.0 / (1.0 + ((double)j + 0.5) / 128.0);
        t->rcp[j] = c;
        t->mlog[j] = -log(c);
    }
}
// e^-s for s >= 0
__device__ __forceinline__ double exp_neg(double s, const Tables* t) {
    s = fmin(s, 700.0);                                     // e^-700 ~ 1e-304: below anything that matters, still normal
    const double kd = rint(s * -92.33248261689366);         // -s * 64 / ln 2
    double r = fma(kd, -0.010830424667801708, -s);          // -s - k ln2/64, ln2/64 split in two parts
    r = fma(kd, -2.8447437476627285e-11, r);
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);             // |r| <= ln2/128: r^6/720 < 4e-17
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int k = (int)kd;
    const double v = t->exp2j[k & 63] * p;                  // in [1, 2.02)
    return __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));   // * 2^(k div 64)
}
// log(1 + e) for 0 <= e <= 1
__device__ __forceinline__ double log1p_unit(double e, const Tables* t) {
    const double m = 1.0 + e;
    const int j = (int)((m - 1.0) * 128.0);                 // 0 .. 128
    const double r = fma(m, t->rcp[j], -1.0);               // |r| <= 2^-8: r^7/7 < 2e-18
    double q = fma(r, -1.0 / 6.0, 1.0 / 5.0);
    q = fma(q, r, -0.25);
    q = fma(q, r, 1.0 / 3.0);
    q = fma(q, r, -0.5);
    q = fma(q * r, r, r);
    return t->mlog[j] + q;
}
__device__ __forceinline__ double f_fast(double a, double b, const Tables* t) {
    const double ma = fabs(a), mb = fabs(b);
    const double mn = fmin(ma, mb);
    // sgn(a) sgn(b) min(|a|, |b|) with sgn(0) = 0 (the minimum is then 0 anyway): sign bits XORed onto the minimum
    double y = __hiloint2double(__double2hiint(mn) | ((__double2hiint(a) ^ __double2hiint(b)) & (int)0x80000000), __double2loint(mn));
    if (40.0 > fmax(ma, mb))                                // PolarCode.cpp:438: exact box-plus below 40, sign-min otherwise
        y += log1p_unit(exp_neg(fabs(a + b), t), t) - log1p_unit(exp_neg(fabs(a - b), t), t);
    return y;
}
__device__ __forceinline__ double softplus_fast(double x, const Tables* t) {
    if (x > 709.782712893384) return CUDART_INF;            // exp(x) overflows a double (:483 yields +inf)
    if (x < -36.7368005696771) return 0.0;                  // 1 + exp(x) rounds to 1
    return fmax(x, 0.0) + log1p_unit(exp_neg(fabs(x), t), t);
}
constexpr double kTieMargin = 1e-9;
__device__ __forceinline__ unsigned long long wmax64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(FULL_MASK, v, o); v = t > v ? t : v; }
    return v;
}
__device__ __forceinline__ unsigned long long wmin64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(FULL_MASK, v, o); v = t < v ? t : v; }
    return v;
}

template <class In>
struct Args {
    const In* llr;               // [B][N], converted to double on load
    uint32_t* out;               // [B][KW]
    const int* list;             // null, or the codeword indices to decode
    const int* count;            // with `list`: number of entries (device memory)
    int* list2;                  // null, or: codewords decided on a double-precision margin below kTieMargin are appended
    int* count2;                 //   here (to be decoded once more with the literal formulas) and are neither stored nor counted
    const uint32_t* frozen_words;
    const uint16_t* info_order;
    const uint32_t* crc_masks;
    double* gx;                  // per-block scratch rows (32 doubles each) of layers 1 .. lamS-1
    unsigned long long gx_stride;// doubles per block
    int B, n, K, crc, L;
    int W;                       // list size rounded up to a power of two (work is spread over W paths x beta)
    int lamS;                    // first layer kept in shared memory
    int big;                     // layers with more than this many (path, beta) items are refreshed by the whole block
    int smem_x_rows, smem_s_rows;
    int s_off[kMaxN + 2];        // word-row offset of partial-sum layer lam (all of them in shared memory)
    // fused block-error counting, as in fast::Args
    const uint32_t* truth;
    unsigned long long* err;
    long long first_index;
    int n_ebno;
};

template <class In>
__global__ void __launch_bounds__(NT) scl_exact_kernel(const Args<In> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5, nt = blockDim.x;
    const bool lead = wib == 0;                  // the warp that decodes the leaves (rotating it per block was measured: no gain)
    const int n = a.n, N = 1 << n, L = a.L, W = a.W, lamS = a.lamS;
    int wsh = 0;
    while ((1 << wsh) < W) ++wsh;
    const int NW = (N + 31) >> 5, KW = (a.K + 31) >> 5;
    const int slot = lane;                       // warp 0: lane = list path (lanes >= W never become active)
    const unsigned gmask_lo = (W == 32) ? FULL_MASK : ((1u << W) - 1u);

    double* sx = reinterpret_cast<double*>(smem_raw);
    uint32_t* ss = reinterpret_cast<uint32_t*>(sx + (size_t)a.smem_x_rows * 32);
    unsigned char* px = reinterpret_cast<unsigned char*>(ss + (size_t)a.smem_s_rows * 32);   // [16][32] LLR column pointers
    unsigned char* ps = px + 16 * 32;                                                        // [16][32] partial-sum pointers
    unsigned char* srcof = ps + 16 * 32;                                                     // [32] clone scatter
    volatile uint32_t* actp = reinterpret_cast<volatile uint32_t*>(srcof + 32);              // active-path mask
    Tables* tb = reinterpret_cast<Tables*>(srcof + 48);
    build_tables(tb, tid, nt);
    double* gx = a.gx + a.gx_stride * blockIdx.x;

    auto xrow = [&](int lam, int beta) -> double* {
        if (lam >= lamS) return sx + ((size_t)((1 << (n - lamS + 1)) - (1 << (n - lam + 1)) + beta) << 5);
        return gx + ((size_t)(N - (1 << (n - lam + 1)) + beta) << 5);
    };
    auto srow = [&](int lam, int w) -> uint32_t* { return ss + ((size_t)(a.s_off[lam] + w) << 5); };

    const int nB = a.list ? *a.count : a.B;
    for (int item = blockIdx.x; item < nB; item += gridDim.x) {
        const int cw = a.list ? a.list[item] : item;
        const In* chan = a.llr + (size_t)cw * N;

        // per-path state of warp 0 (PolarCode.cpp:250-263: free stack 0..L-1, first path = L-1)
        bool active = lead && (slot == L - 1);
        double pm = 0;
        uint32_t s_n = 0;
        int stk = slot, sp = L - 1;
        double lam_n = 0;
        double dmin = CUDART_INF;                          // smallest decision margin (warp-uniform)
        uint32_t frozen_word = 0;
        __syncthreads();                                   // the previous codeword's output gather is done
        for (int i = tid; i < 2 * 16 * 32; i += nt) px[i] = 0;          // px and ps
        if (tid == 0) *actp = 1u << (L - 1);
        __syncthreads();

        // one tree layer for every active path: items (path, i), spread over the threads t0, t0 + T, ... (T a multiple of W):
        // a thread keeps its path and strides over i, so everything that depends on the path alone is hoisted
        auto refresh = [&](int lam, bool is_g, int t0, int T) {
            const int M = 1 << (n - lam);
            const int path = t0 & (W - 1), i0 = t0 >> wsh, di = T >> wsh;
            if (i0 >= M || !((*actp >> path) & 1u)) return;
            const uint32_t* sw = (is_g && lam < n) ? srow(lam, 0) + ps[(lam - 1) * 32 + path] : nullptr;
            if (lam == n) {
                // the decision LLR (one item per path; warp 0, lane = path)
                double x0, x1;
                if (n > 1) {
                    const double* src = xrow(lam - 1, 0) + px[(lam - 2) * 32 + path];
                    x0 = src[0]; x1 = src[32];
                } else {
                    x0 = (double)chan[0]; x1 = (double)chan[1];
                }
                lam_n = is_g ? x1 + ((s_n & 1u) ? -x0 : x0) : f_fast(x0, x1, tb);
                return;
            }
            double* dst = xrow(lam, 0) + path;
            if (lam == 1) {
                // channel layer: reference pairs are (2k, 2k+1); the result lands at the bit-reversed position
                for (int i = i0; i < M; i += di) {
                    const double x0 = (double)chan[2 * i], x1 = (double)chan[2 * i + 1];
                    const int beta = (int)(__brev((unsigned)i) >> (33 - n));
                    double y;
                    if (is_g) y = x1 + (((sw[(size_t)(beta >> 5) << 5] >> (beta & 31)) & 1u) ? -x0 : x0);
                    else y = f_fast(x0, x1, tb);
                    dst[(size_t)beta << 5] = y;
                }
                return;
            }
            const double* src = xrow(lam - 1, 0) + px[(lam - 2) * 32 + path];
            if (is_g) {
                for (int i = i0; i < M; i += di) {                             // PolarCode.cpp:448-451
                    const double x0 = src[(size_t)i << 5], x1 = src[(size_t)(i + M) << 5];
                    dst[(size_t)i << 5] = x1 + (((sw[(size_t)(i >> 5) << 5] >> (i & 31)) & 1u) ? -x0 : x0);
                }
            } else {
                for (int i = i0; i < M; i += di)                               // PolarCode.cpp:438-446
                    dst[(size_t)i << 5] = f_fast(src[(size_t)i << 5], src[(size_t)(i + M) << 5], tb);
            }
        };

        for (int phi = 0; phi < N; ++phi) {
            // ---- refresh LLR layers lam_top..n (PolarCode.cpp:422-455): big layers by the block, small ones by warp 0 ----
            const int lam_top = (phi == 0) ? 1 : n - (__ffs(phi) - 1);
            const bool any_big = ((1 << (n - lam_top)) << wsh) > a.big;
            if (any_big) __syncthreads();                  // warp 0's decisions of the leaves before this one are visible
            for (int lam = lam_top; lam <= n; ++lam) {
                const bool big = ((1 << (n - lam)) << wsh) > a.big;
                const bool is_g = (lam == lam_top) && (phi != 0);
                if (big) {
                    refresh(lam, is_g, tid, nt);
                    if (lam < n && tid < 32) px[(lam - 1) * 32 + tid] = (unsigned char)tid;
                    __syncthreads();
                } else if (lead) {
                    refresh(lam, is_g, lane, 32);
                    if (lam < n) px[(lam - 1) * 32 + lane] = (unsigned char)lane;
                    __syncwarp();
                }
            }
            if (!lead) continue;                           // the other warps wait at the next leaf with a big layer

            // ---- leaf decision (warp 0, lane = path) ----
            if ((phi & 31) == 0) frozen_word = a.frozen_words[phi >> 5];
            const bool frozen = (frozen_word >> (phi & 31)) & 1u;
            uint32_t u = 0;
            if (frozen) {
                if (active) pm += softplus_fast(-lam_n, tb);                   // PolarCode.cpp:475-487
            } else {
                // PolarCode.cpp:489-607. Metrics are kept positive (m = -probForks).
                // both fork metrics share log1p(e^-|x|): softplus(-x) and softplus(x) differ by max(-+x, 0)
                const double al = fabs(lam_n);
                double lo = 0.0, hi = CUDART_INF;                              // likely / unlikely increment (:505-506)
                if (!(al > 36.7368005696771)) lo = log1p_unit(exp_neg(al, tb), tb);
                if (!(al > 709.782712893384)) hi = al + lo;
                const double m0 = pm + (lam_n < 0 ? hi : lo);
                const double m1 = pm + (lam_n < 0 ? lo : hi);
                const unsigned act_g = __ballot_sync(FULL_MASK, active) & gmask_lo;
                const int A = __popc(act_g);
                bool keep0 = active, keep1 = active;
                bool slow = 2 * A > L;                                         // warp-uniform
                if (slow && A == L) {
                    // common case: every likely fork strictly beats every unlikely fork and the list is full
                    const double lo = rmin<double>(m0, m1), hi = rmax<double>(m0, m1);
                    const double worst_likely = group_max<double>(active ? lo : -CUDART_INF, 32);
                    const double best_unlikely = group_min<double>(active ? hi : CUDART_INF, 32);
                    if (best_unlikely > worst_likely) {
                        dmin = fmin(dmin, best_unlikely - worst_likely);
                        slow = false;
                        keep0 = active && (m0 <= m1);    // m0 == m1 cannot happen here (would not be strict)
                        keep1 = active && !keep0;
                    }
                }
                if (slow) {
                    // keep the L best of the 2A forks under (metric asc, fork index asc) == PolarCode.cpp:528-553 (sort,
                    // threshold, '>' pass then '==' pass in index order): start from "every likely fork", promote the best
                    // unlikely fork while the list has room, then swap it for the worst kept likely fork while that improves
                    // the kept set. Metrics are non-negative doubles, so their bit patterns order like the values.
                    const bool like1 = m1 < m0;                                // likely fork is bit 1
                    const unsigned long long klo = (unsigned long long)__double_as_longlong(like1 ? m1 : m0);
                    const unsigned long long khi = (unsigned long long)__double_as_longlong(like1 ? m0 : m1);
                    const unsigned lk = __ballot_sync(FULL_MASK, like1);
                    unsigned keptA = act_g, keptB = 0;
                    int count = A;
                    while (true) {
                        const unsigned candB = act_g & ~keptB;
                        if (candB == 0) break;
                        const unsigned long long kb = wmin64(((candB >> lane) & 1u) ? khi : ~0ull);
                        const unsigned eqb = __ballot_sync(FULL_MASK, ((candB >> lane) & 1u) && khi == kb);
                        const int bl = __ffs(eqb) - 1;                         // lowest lane = lowest fork index among equals
                        if (count < L) { keptB |= 1u << bl; ++count; continue; }
                        const unsigned long long ka = wmax64(((keptA >> lane) & 1u) ? klo : 0ull);
                        const unsigned eqa = __ballot_sync(FULL_MASK, ((keptA >> lane) & 1u) && klo == ka);
                        const int al = 31 - __clz(eqa);                        // highest lane = highest fork index among equals
                        const int idxb = 2 * bl + (((lk >> bl) & 1u) ? 0 : 1);
                        const int idxa = 2 * al + (((lk >> al) & 1u) ? 1 : 0);
                        if (!((kb < ka) || (kb == ka && idxb < idxa))) break;
                        keptB |= 1u << bl;
                        keptA &= ~(1u << al);
                    }
                    const bool ka_ = (keptA >> lane) & 1u, kb_ = (keptB >> lane) & 1u;
                    keep0 = like1 ? kb_ : ka_;
                    keep1 = like1 ? ka_ : kb_;
                    // margin of this selection: best dropped fork - worst kept fork
                    const double kept = group_max<double>(rmax<double>(keep0 ? m0 : -CUDART_INF, keep1 ? m1 : -CUDART_INF), 32);
                    const double dropped = group_min<double>(rmin<double>((active && !keep0) ? m0 : CUDART_INF,
                                                                          (active && !keep1) ? m1 : CUDART_INF), 32);
                    dmin = fmin(dmin, dropped - kept);
                }
                const bool kill = active && !keep0 && !keep1;
                const bool clone = keep0 && keep1;
                const unsigned Kg = __ballot_sync(FULL_MASK, kill) & gmask_lo;
                const unsigned Cg = __ballot_sync(FULL_MASK, clone) & gmask_lo;
                if ((Kg | Cg) == 0) {
                    if (active) { u = keep1 ? 1u : 0u; pm = keep1 ? m1 : m0; }
                } else {
                    const int nk = __popc(Kg), nc = __popc(Cg);
                    // killPath pushes in ascending path order (PolarCode.cpp:555-560, :292)
                    if (slot >= sp && slot < sp + nk) stk = (int)__fns(Kg, 0, slot - sp + 1);
                    const int sp2 = sp + nk;
                    // clonePath pops for ascending l (PolarCode.cpp:562-570, :275-276)
                    const int ci = __popc(Cg & ((1u << slot) - 1u));
                    const int tgt = __shfl_sync(FULL_MASK, stk, (sp2 - 1 - ci) & 31);
                    sp = sp2 - nc;
                    srcof[lane] = (unsigned char)lane;
                    __syncwarp();
                    if (clone) srcof[tgt] = (unsigned char)lane;
                    __syncwarp();
                    const int src_lane = srcof[lane];
                    __syncwarp();
                    const bool is_new = (src_lane != lane);
                    const double src_m1 = __shfl_sync(FULL_MASK, m1, src_lane);
                    const uint32_t src_sn = __shfl_sync(FULL_MASK, s_n, src_lane);
                    // a clone takes over its parent's column pointers (both tables: 2 x 16 rows of 32 bytes)
                    unsigned char cp[32];
#pragma unroll
                    for (int r = 0; r < 32; ++r) cp[r] = px[r * 32 + src_lane];
                    __syncwarp();
                    if (is_new) {
#pragma unroll
                        for (int r = 0; r < 32; ++r) px[r * 32 + lane] = cp[r];
                        active = true; pm = src_m1; u = 1u; s_n = src_sn;
                    } else if (kill) {
                        active = false; pm = 0;
                    } else if (active) {
                        u = keep0 ? 0u : 1u;
                        pm = keep0 ? m0 : m1;
                    }
                    const unsigned now = __ballot_sync(FULL_MASK, active);
                    if (lane == 0) *actp = now;
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
                    else Sw = srow(lam, 0)[ps[(lam - 1) * 32 + lane]];
                    P = ((Sw ^ P) & ((1u << M) - 1u)) | (P << M);
                    --lam;
                }
                __syncwarp();       // every lane has read the columns that are about to be overwritten
                if (lam == lam_end) {
                    srow(lam, 0)[lane] = P;
                } else {
                    const int Wd = 1 << (t - 5);          // words of the destination vector
                    uint32_t* D = srow(lam_end, 0) + lane;
                    D[(size_t)(Wd - 1) << 5] = P;
                    for (; lam > lam_end; --lam) {
                        const int mw = 1 << (n - lam - 5);
                        const int base = Wd - mw;
                        const uint32_t* S = srow(lam, 0) + ps[(lam - 1) * 32 + lane];
                        for (int w = 0; w < mw; ++w)
                            D[(size_t)(base - mw + w) << 5] = S[(size_t)w << 5] ^ D[(size_t)(base + w) << 5];
                    }
                }
                if (lam_end >= 1) ps[(lam_end - 1) * 32 + lane] = (unsigned char)lane;
            }
            __syncwarp();
        }

        if (lead) {
            // ---- u-hat of every path: packed polar transform of the re-encoded codeword (layer 0) ----
            uint32_t* D = srow(0, 0) + lane;
            for (int sw = NW >> 1; sw >= 1; sw >>= 1)
                for (int i = 0; i < NW; ++i)
                    if ((i & sw) == 0) D[(size_t)i << 5] ^= D[(size_t)(i + sw) << 5];
            bool pass = true;
            for (int i = 0; i < NW; ++i) {
                uint32_t w = D[(size_t)i << 5];
                if (N > 16) w ^= (w >> 16) & 0x0000FFFFu;
                if (N > 8) w ^= (w >> 8) & 0x00FF00FFu;
                if (N > 4) w ^= (w >> 4) & 0x0F0F0F0Fu;
                if (N > 2) w ^= (w >> 2) & 0x33333333u;
                w ^= (w >> 1) & 0x55555555u;
                D[(size_t)i << 5] = w;
            }
            for (int r = 0; r < a.crc; ++r) {            // PolarCode.cpp:93-108
                uint32_t acc = 0;
                for (int i = 0; i < NW; ++i) acc ^= D[(size_t)i << 5] & a.crc_masks[(size_t)r * NW + i];
                if (__popc(acc) & 1) pass = false;
            }
            // ---- final pick, PolarCode.cpp:609-644 ----
            const unsigned act_all = __ballot_sync(FULL_MASK, active);
            const unsigned pass_g = __ballot_sync(FULL_MASK, active && pass);
            const bool use_parity = (a.crc != 0) && (pass_g != 0);
            const bool eligible = active && (use_parity ? pass : true) && (pm < CUDART_INF);
            const double best = group_min<double>(eligible ? pm : CUDART_INF, 32);
            const unsigned cand = __ballot_sync(FULL_MASK, eligible && pm == best);
            const int win = cand ? (__ffs(cand) - 1) : 0;
            const bool wa = (act_all >> win) & 1u;
            const double second = group_min<double>((eligible && lane != win) ? pm : CUDART_INF, 32);
            if (cand != 0) dmin = fmin(dmin, second - best);
            __syncwarp();
            if (a.list2 != nullptr && dmin < kTieMargin) {
                // an exact cancellation somewhere: only the literal formulas resolve it the way the reference does
                if (lane == 0) a.list2[atomicAdd(a.count2, 1)] = cw;
                continue;
            }
            // ---- output gather: decoded[j] = u-hat[order[j]], PolarCode.cpp:171-174 ----
            const uint32_t* U = srow(0, 0) + win;
            bool differs = false;
            for (int t = lane; t < KW; t += 32) {
                uint32_t word = 0;
                if (wa) {
                    const int jmax = min(32, a.K - 32 * t);
                    for (int i = 0; i < jmax; ++i) {
                        const int pos = a.info_order[32 * t + i];
                        word |= ((U[(size_t)(pos >> 5) << 5] >> (pos & 31)) & 1u) << i;
                    }
                }
                if (a.out != nullptr) a.out[(size_t)cw * KW + t] = word;
                if (a.truth != nullptr) differs |= word != a.truth[(size_t)cw * KW + t];
            }
            if (a.truth != nullptr) {                        // PolarCode.cpp:758-769
                const bool block_error = __any_sync(FULL_MASK, differs);
                if (lane == 0 && block_error)
                    atomicAdd(a.err + (int)((unsigned long long)(a.first_index + cw) % (unsigned)a.n_ebno), 1ull);
            }
        }
    }
}

}  // namespace exact
