// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// CPU restatement ("port" oracle) of the reference's LLR-domain SC / SCL polar
// decoder, probability-domain decoder (decode_scl_p1), code construction, encoder and AWGN BLER harness
// (reference: /root/reference/PolarC/PolarCode.{h,cpp}, cited per function below).
//
// It is written from the algorithm, not from the reference's data structures: the
// reference keeps ref-counted, lazily copied per-layer array pools (Tal-Vardy
// Alg. 5-9, PolarCode.cpp:195-373); this restatement gives every list path its own
// eagerly copied tree. The two are observably identical because the lazy store only
// ever changes *where* a path's data lives, never its contents; the one thing that
// does leak out of the store -- the LIFO order in which path indices are recycled,
// which decides exact metric ties -- is reproduced explicitly (see PathPool).
//
// Parity pin: the reference ships no golden vectors or tests, so
// this file is pinned against the UNMODIFIED reference compiled here into
// oracle/_ref/libpolar_ref.so (tests/test_oracle.py::test_port_equals_compiled_reference,
// ::test_port_probability_domain_equals_compiled_reference) and against the fixtures that
// binary generated (tests/golden/, made by tests/golden/make_golden*.py).
//
// Language note: C++ rather than plain C because the reference's reliability order
// is whatever libstdc++'s unstable std::sort makes of exact ties
// (PolarCode.cpp:38-40); reproducing that permutation means calling the same
// std::sort with the same comparator on the same sequence.
//
// Besides the reference's rules it can evaluate two variants that are NOT the reference's (Decoder::minsum_only):
// min-sum check nodes everywhere (a sensitivity probe) and, on top, the hardware-friendly metric update max(x, 0) --
// the checker of the product's opt-in MINSUM mode (tests/test_gpu_parity.py::test_minsum_mode_matches_the_minsum_oracle).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <thread>
#include <vector>

namespace {

// instrumentation (single-threaded use only): info steps, info steps with >= 1 clone, clones, kills
long long g_stats[4] = {0, 0, 0, 0};

struct Code {
    int n = 0, N = 0, K = 0, crc = 0;
    double eps = 0.0;
    std::vector<uint8_t> frozen;     // [N], index = phi (decoding order)
    std::vector<uint16_t> order;     // [N], most reliable first
    std::vector<uint16_t> bitrev;    // [N]
    std::vector<uint8_t> crcm;       // [crc][K] row-major
};

// PolarCode.cpp:647-656 -- n-bit reversal table.
void make_bitrev(Code& c) {
    c.bitrev.assign(c.N, 0);
    for (int i = 0; i < c.N; ++i) {
        unsigned r = 0;
        for (int b = 0; b < c.n; ++b)
            if (i & (1 << b)) r |= 1u << (c.n - 1 - b);
        c.bitrev[i] = (uint16_t)r;
    }
}

// PolarCode.cpp:17-58 -- BEC Bhattacharyya recursion, reliability sort, frozen set,
// random parity ("CRC") matrix drawn from rand().
void construct(Code& c, bool draw_crc) {
    std::vector<double> z(c.N, c.eps);
    // :23-33  stage `it` pairs (i+j, i+j+2^it): worse channel to the low index.
    for (int it = 0; it < c.n; ++it) {
        const int inc = 1 << it;
        for (int j = 0; j < inc; ++j)
            for (int i = 0; i < c.N; i += 2 * inc) {
                const double a = z[i + j], b = z[i + j + inc];
                z[i + j] = a + b - a * b;
                z[i + j + inc] = a * b;
            }
    }
    // :35-40  ascending sort of indices by z[bitrev[i]]; ties are resolved by
    // libstdc++'s introsort, hence the identical call shape (uint16_t sequence,
    // comparator taking ints).
    c.order.resize(c.N);
    for (int i = 0; i < c.N; ++i) c.order[i] = (uint16_t)i;
    const std::vector<uint16_t>& br = c.bitrev;
    std::sort(c.order.begin(), c.order.end(),
              [&](int i1, int i2) { return z[br.at(i1)] < z[br.at(i2)]; });
    // :42-49  first K+crc of the order carry data, the rest are frozen to 0.
    c.frozen.assign(c.N, 1);
    for (int i = 0; i < c.K + c.crc; ++i) c.frozen[c.order[i]] = 0;
    // :51-56  parity matrix, row = parity bit, rand()%2 per entry, row-major draw order.
    c.crcm.assign((size_t)c.crc * c.K, 0);
    if (draw_crc)
        for (int r = 0; r < c.crc; ++r)
            for (int j = 0; j < c.K; ++j) c.crcm[(size_t)r * c.K + j] = (uint8_t)(rand() % 2);
}

// PolarCode.cpp:60-91 -- place info + parity bits, n XOR-butterfly stages, bit reversal.
void encode(const Code& c, const uint8_t* info, uint8_t* coded) {
    std::vector<uint8_t> u(c.N, 0);
    for (int i = 0; i < c.K; ++i) u[c.order[i]] = info[i] & 1;               // :65-67
    for (int r = 0; r < c.crc; ++r) {                                        // :68-74
        unsigned acc = 0;
        for (int j = 0; j < c.K; ++j) acc ^= (unsigned)(c.crcm[(size_t)r * c.K + j] & info[j]);
        u[c.order[c.K + r]] = (uint8_t)(acc & 1);
    }
    for (int it = 0; it < c.n; ++it) {                                       // :76-83
        const int inc = 1 << it;
        for (int j = 0; j < inc; ++j)
            for (int i = 0; i < c.N; i += 2 * inc) u[i + j] ^= u[i + j + inc];
    }
    for (int i = 0; i < c.N; ++i) coded[i] = u[c.bitrev[i]];                 // :85-87
}

template <class R> inline R r_exp(R x);
template <> inline double r_exp<double>(double x) { return std::exp(x); }
template <> inline float r_exp<float>(float x) { return expf(x); }
template <class R> inline R r_log(R x);
template <> inline double r_log<double>(double x) { return std::log(x); }
template <> inline float r_log<float>(float x) { return logf(x); }

// PolarCode.cpp:437-446 -- the reference's check-node rule: exact box-plus while both
// magnitudes are below 40 (strictly), sign-min otherwise, with sgn(0) = 0.
template <class R>
inline R f_rule(R a, R b, int minsum_only) {
    const R ma = std::fabs(a), mb = std::fabs(b);
    if (!minsum_only && R(40) > std::max(ma, mb))
        return r_log<R>((r_exp<R>(a + b) + R(1)) / (r_exp<R>(a) + r_exp<R>(b)));
    const R sa = (a < 0) ? R(-1) : R(a > 0);
    const R sb = (b < 0) ? R(-1) : R(b > 0);
    return sa * sb * std::min(ma, mb);
}

// log(1 + exp(x)) exactly as written at PolarCode.cpp:483,505-506,580-601: no
// stabilisation, so it returns +inf once exp overflows and 0 once 1+exp rounds to 1.
template <class R>
inline R softplus_literal(R x) { return r_log<R>(R(1) + r_exp<R>(x)); }

// Path-index recycling exactly as the reference's _inactivePathIndices stack
// (PolarCode.cpp:250-256 fill 0..L-1 so the first pop is L-1; :292 push on kill;
// :275-276 pop on clone).
struct PathPool {
    std::vector<int> lifo;
    void reset(int L) { lifo.clear(); for (int l = 0; l < L; ++l) lifo.push_back(l); }
    int take() { int l = lifo.back(); lifo.pop_back(); return l; }
    void give(int l) { lifo.push_back(l); }
};

template <class R>
struct Decoder {
    const Code& c;
    int L;
    int minsum_only;     // 0: the reference's rules. 1: min-sum check nodes everywhere (sensitivity probe). 2: min-sum and
                         // the hardware-friendly metric update max(x, 0) in place of log(1 + exp(x)) -- the checker of the
                         // product's opt-in MINSUM mode, NOT the reference's arithmetic
    R metric_inc(R x) const { return minsum_only >= 2 ? std::max(x, R(0)) : softplus_literal<R>(x); }
    // per path: llr tree (layer lam at offset off[lam], 2^(n-lam) entries), partial sums
    // (layer lam at 2*off[lam], 2 phase slots per entry), decided bits, metric, alive flag.
    std::vector<size_t> off;
    size_t tree = 0;
    std::vector<R> llr;
    std::vector<uint8_t> ps;
    std::vector<uint8_t> bits;
    std::vector<R> pm;
    std::vector<uint8_t> alive;
    PathPool pool;

    Decoder(const Code& code, int list, int ms) : c(code), L(list), minsum_only(ms) {
        off.resize(c.n + 1);
        for (int lam = 0; lam <= c.n; ++lam) { off[lam] = tree; tree += (size_t)1 << (c.n - lam); }
        llr.resize(tree * L);
        ps.resize(2 * tree * L);
        bits.resize((size_t)c.N * L);
        pm.resize(L);
        alive.resize(L);
    }
    R* A(int l, int lam) { return &llr[(size_t)l * tree + off[lam]]; }
    uint8_t* C(int l, int lam) { return &ps[2 * ((size_t)l * tree + off[lam])]; }
    uint8_t* U(int l) { return &bits[(size_t)l * c.N]; }

    // PolarCode.cpp:422-455 -- refresh layers needed for bit phi, shallowest first.
    void calc_llr(int lam, int phi) {
        if (lam == 0) return;
        if ((phi & 1) == 0) calc_llr(lam - 1, phi >> 1);
        const int M = 1 << (c.n - lam);
        for (int l = 0; l < L; ++l) {
            if (!alive[l]) continue;
            R* y = A(l, lam);
            const R* x = A(l, lam - 1);
            const uint8_t* cs = C(l, lam);
            for (int b = 0; b < M; ++b) {
                if ((phi & 1) == 0) y[b] = f_rule<R>(x[2 * b], x[2 * b + 1], minsum_only);
                else y[b] = R(1 - 2 * (int)cs[2 * b]) * x[2 * b] + x[2 * b + 1];        // :448-451
            }
        }
    }

    // PolarCode.cpp:457-473 -- push the finished pair of partial sums one layer up.
    void update_c(int lam, int phi) {
        const int psi = phi >> 1;
        const int M = 1 << (c.n - lam);
        for (int l = 0; l < L; ++l) {
            if (!alive[l]) continue;
            const uint8_t* cs = C(l, lam);
            uint8_t* up = C(l, lam - 1);
            for (int b = 0; b < M; ++b) {
                up[2 * (2 * b) + (psi & 1)] = cs[2 * b] ^ cs[2 * b + 1];
                up[2 * (2 * b + 1) + (psi & 1)] = cs[2 * b + 1];
            }
        }
        if (psi & 1) update_c(lam - 1, psi);
    }

    // PolarCode.cpp:475-487
    void frozen_step(int phi) {
        for (int l = 0; l < L; ++l) {
            if (!alive[l]) continue;
            C(l, c.n)[phi & 1] = 0;
            pm[l] += metric_inc(-A(l, c.n)[0]);
            U(l)[phi] = 0;
        }
    }

    // PolarCode.cpp:274-288 (+ the copy the lazy store would do on first write).
    int clone(int l) {
        const int lp = pool.take();
        alive[lp] = 1;
        pm[lp] = pm[l];
        std::copy(A(l, 0), A(l, 0) + tree, A(lp, 0));
        std::copy(C(l, 0), C(l, 0) + 2 * tree, C(lp, 0));
        return lp;
    }
    // PolarCode.cpp:290-303
    void kill(int l) { alive[l] = 0; pool.give(l); pm[l] = 0; }

    // PolarCode.cpp:489-607
    void info_step(int phi) {
        const R nan = std::numeric_limits<R>::quiet_NaN();
        std::vector<R> fork(2 * L, nan), pool_sorted;
        int n_alive = 0;
        for (int l = 0; l < L; ++l) {                                         // :497-519
            if (!alive[l]) continue;
            const R lam = A(l, c.n)[0];
            fork[2 * l] = -(pm[l] + metric_inc(-lam));
            fork[2 * l + 1] = -(pm[l] + metric_inc(lam));
            pool_sorted.push_back(fork[2 * l]);
            pool_sorted.push_back(fork[2 * l + 1]);
            ++n_alive;
        }
        const int rho = std::min(2 * n_alive, L);                             // :521-523
        std::sort(pool_sorted.begin(), pool_sorted.end(), std::greater<R>()); // :528
        const R thr = pool_sorted.at(rho - 1);                                // :530
        std::vector<uint8_t> keep(2 * L, 0);
        int kept = 0;
        for (int i = 0; i < 2 * L && kept < rho; ++i)                         // :533-541
            if (fork[i] > thr) { keep[i] = 1; ++kept; }
        for (int i = 0; i < 2 * L && kept < rho; ++i)                         // :543-553
            if (fork[i] == thr) { keep[i] = 1; ++kept; }
        {
            int nclone = 0, nkill = 0;
            for (int l = 0; l < L; ++l) {
                if (alive[l] && keep[2 * l] && keep[2 * l + 1]) ++nclone;
                if (alive[l] && !keep[2 * l] && !keep[2 * l + 1]) ++nkill;
            }
            g_stats[0] += 1; g_stats[1] += (nclone > 0); g_stats[2] += nclone; g_stats[3] += nkill;
        }
        for (int l = 0; l < L; ++l)                                           // :555-560
            if (alive[l] && !keep[2 * l] && !keep[2 * l + 1]) kill(l);
        for (int l = 0; l < L; ++l) {                                         // :562-605
            if (!keep[2 * l] && !keep[2 * l + 1]) continue;
            const R lam = A(l, c.n)[0];
            if (keep[2 * l] && keep[2 * l + 1]) {
                C(l, c.n)[phi & 1] = 0;
                const int lp = clone(l);
                C(lp, c.n)[phi & 1] = 1;
                std::copy(U(l), U(l) + phi, U(lp));
                U(l)[phi] = 0;
                U(lp)[phi] = 1;
                pm[l] += metric_inc(-lam);
                pm[lp] += metric_inc(lam);
            } else if (keep[2 * l]) {
                C(l, c.n)[phi & 1] = 0;
                U(l)[phi] = 0;
                pm[l] += metric_inc(-lam);
            } else {
                C(l, c.n)[phi & 1] = 1;
                U(l)[phi] = 1;
                pm[l] += metric_inc(lam);
            }
        }
    }

    // PolarCode.cpp:93-108
    bool parity_ok(const uint8_t* u) const {
        for (int r = 0; r < c.crc; ++r) {
            unsigned acc = 0;
            for (int j = 0; j < c.K; ++j) acc ^= (unsigned)(c.crcm[(size_t)r * c.K + j] & u[c.order[j]]);
            if ((acc & 1) != u[c.order[c.K + r]]) return false;
        }
        return true;
    }

    // PolarCode.cpp:609-644 -- strictly smaller metric wins, first index wins ties, index 0
    // if nothing beats DBL_MAX; retry without the parity filter if no path passes it.
    int pick(bool use_parity) {
        int best = 0;
        R best_pm = std::numeric_limits<R>::max();
        bool any = false;
        for (int l = 0; l < L; ++l) {
            if (!alive[l]) continue;
            if (use_parity && !parity_ok(U(l))) continue;
            any = true;
            if (pm[l] < best_pm) { best_pm = pm[l]; best = l; }
        }
        if (any) return best;
        return pick(false);
    }

    // PolarCode.cpp:130-190
    void run(const R* channel, uint8_t* info_out) {
        std::fill(llr.begin(), llr.end(), R(0));
        std::fill(ps.begin(), ps.end(), 0);
        std::fill(bits.begin(), bits.end(), 0);
        std::fill(pm.begin(), pm.end(), R(0));
        std::fill(alive.begin(), alive.end(), 0);
        pool.reset(L);
        const int l0 = pool.take();                                            // :138, :259-272
        alive[l0] = 1;
        std::copy(channel, channel + c.N, A(l0, 0));                           // :140-144
        for (int phi = 0; phi < c.N; ++phi) {                                  // :152-168
            calc_llr(c.n, phi);
            if (c.frozen[phi]) frozen_step(phi); else info_step(phi);
            if (phi & 1) update_c(c.n, phi);
        }
        const int w = pick(c.crc != 0);                                        // :169
        for (int j = 0; j < c.K; ++j) info_out[j] = U(w)[c.order[j]];          // :171-174
    }
};

// ---- probability-domain decoder: PolarCode::decode_scl_p1 (PolarCode.cpp:110-128) -> decode_scl (:150-190)
// with recursivelyCalcP (:375-420) in place of the LLR recursion. Double only, like the reference. Layer
// arrays hold pairs (P(..|0), P(..|1)); after every layer refresh all values of all live paths are divided by
// their common maximum (:392-418), which keeps path likelihoods comparable across the list; there is no
// separate path metric: forks are ranked by the pair at the decision layer (:510-514) and the final pick takes
// the largest likelihood of the last decided bit (:631-637).
struct ProbDecoder {
    const Code& c;
    int L;
    std::vector<size_t> off;
    size_t tree = 0;
    std::vector<double> pr;      // per path: 2 doubles per tree entry
    std::vector<uint8_t> ps, bits, alive;
    PathPool pool;

    ProbDecoder(const Code& code, int list) : c(code), L(list) {
        off.resize(c.n + 1);
        for (int lam = 0; lam <= c.n; ++lam) { off[lam] = tree; tree += (size_t)1 << (c.n - lam); }
        pr.resize(2 * tree * L);
        ps.resize(2 * tree * L);
        bits.resize((size_t)c.N * L);
        alive.resize(L);
    }
    double* P(int l, int lam) { return &pr[2 * ((size_t)l * tree + off[lam])]; }
    uint8_t* C(int l, int lam) { return &ps[2 * ((size_t)l * tree + off[lam])]; }
    uint8_t* U(int l) { return &bits[(size_t)l * c.N]; }

    // PolarCode.cpp:375-420
    void calc_p(int lam, int phi) {
        if (lam == 0) return;
        if ((phi & 1) == 0) calc_p(lam - 1, phi >> 1);
        const int M = 1 << (c.n - lam);
        double sigma = 0.0;
        for (int l = 0; l < L; ++l) {
            if (!alive[l]) continue;
            double* y = P(l, lam);
            const double* x = P(l, lam - 1);
            const uint8_t* cs = C(l, lam);
            for (int b = 0; b < M; ++b) {
                if ((phi & 1) == 0) {                                                   // :392-396
                    y[2 * b] = 0.5f * (x[2 * (2 * b)] * x[2 * (2 * b + 1)] + x[2 * (2 * b) + 1] * x[2 * (2 * b + 1) + 1]);
                    y[2 * b + 1] = 0.5f * (x[2 * (2 * b) + 1] * x[2 * (2 * b + 1)] + x[2 * (2 * b)] * x[2 * (2 * b + 1) + 1]);
                } else {                                                                // :398-402
                    const uint8_t u = cs[2 * b];
                    y[2 * b] = 0.5f * x[2 * (2 * b) + (u % 2)] * x[2 * (2 * b + 1)];
                    y[2 * b + 1] = 0.5f * x[2 * (2 * b) + ((u + 1) % 2)] * x[2 * (2 * b + 1) + 1];
                }
                sigma = std::max(sigma, y[2 * b]);
                sigma = std::max(sigma, y[2 * b + 1]);
            }
        }
        if (sigma == 0) return;                                                          // :410-411 (underflow)
        for (int l = 0; l < L; ++l) {
            if (!alive[l]) continue;
            double* y = P(l, lam);
            for (int b = 0; b < M; ++b) { y[2 * b] = y[2 * b] / sigma; y[2 * b + 1] = y[2 * b + 1] / sigma; }
        }
    }
    // PolarCode.cpp:457-473 (same as the LLR decoder)
    void update_c(int lam, int phi) {
        const int psi = phi >> 1;
        const int M = 1 << (c.n - lam);
        for (int l = 0; l < L; ++l) {
            if (!alive[l]) continue;
            const uint8_t* cs = C(l, lam);
            uint8_t* up = C(l, lam - 1);
            for (int b = 0; b < M; ++b) {
                up[2 * (2 * b) + (psi & 1)] = cs[2 * b] ^ cs[2 * b + 1];
                up[2 * (2 * b + 1) + (psi & 1)] = cs[2 * b + 1];
            }
        }
        if (psi & 1) update_c(lam - 1, psi);
    }
    int clone(int l) {                                                                   // :274-288
        const int lp = pool.take();
        alive[lp] = 1;
        std::copy(P(l, 0), P(l, 0) + 2 * tree, P(lp, 0));
        std::copy(C(l, 0), C(l, 0) + 2 * tree, C(lp, 0));
        return lp;
    }
    void kill(int l) { alive[l] = 0; pool.give(l); }                                     // :290-303
    void frozen_step(int phi) {                                                          // :475-487 (no metric here)
        for (int l = 0; l < L; ++l) {
            if (!alive[l]) continue;
            C(l, c.n)[phi & 1] = 0;
            U(l)[phi] = 0;
        }
    }
    void info_step(int phi) {                                                            // :489-607
        const double nan = std::numeric_limits<double>::quiet_NaN();
        std::vector<double> fork(2 * L, nan), sorted;
        int n_alive = 0;
        for (int l = 0; l < L; ++l) {
            if (!alive[l]) continue;
            fork[2 * l] = P(l, c.n)[0];                                                  // :510-514
            fork[2 * l + 1] = P(l, c.n)[1];
            sorted.push_back(fork[2 * l]);
            sorted.push_back(fork[2 * l + 1]);
            ++n_alive;
        }
        const int rho = std::min(2 * n_alive, L);
        std::sort(sorted.begin(), sorted.end(), std::greater<double>());
        const double thr = sorted.at(rho - 1);
        std::vector<uint8_t> keep(2 * L, 0);
        int kept = 0;
        for (int i = 0; i < 2 * L && kept < rho; ++i)
            if (fork[i] > thr) { keep[i] = 1; ++kept; }
        for (int i = 0; i < 2 * L && kept < rho; ++i)
            if (fork[i] == thr) { keep[i] = 1; ++kept; }
        for (int l = 0; l < L; ++l)
            if (alive[l] && !keep[2 * l] && !keep[2 * l + 1]) kill(l);
        for (int l = 0; l < L; ++l) {
            if (!keep[2 * l] && !keep[2 * l + 1]) continue;
            if (keep[2 * l] && keep[2 * l + 1]) {
                C(l, c.n)[phi & 1] = 0;
                const int lp = clone(l);
                C(lp, c.n)[phi & 1] = 1;
                std::copy(U(l), U(l) + phi, U(lp));
                U(l)[phi] = 0;
                U(lp)[phi] = 1;
            } else if (keep[2 * l]) {
                C(l, c.n)[phi & 1] = 0;
                U(l)[phi] = 0;
            } else {
                C(l, c.n)[phi & 1] = 1;
                U(l)[phi] = 1;
            }
        }
    }
    bool parity_ok(const uint8_t* u) const {                                             // :93-108
        for (int r = 0; r < c.crc; ++r) {
            unsigned acc = 0;
            for (int j = 0; j < c.K; ++j) acc ^= (unsigned)(c.crcm[(size_t)r * c.K + j] & u[c.order[j]]);
            if ((acc & 1) != u[c.order[c.K + r]]) return false;
        }
        return true;
    }
    int pick(bool use_parity) {                                                          // :609-644
        int best = 0;
        double p_best = 0;
        bool any = false;
        for (int l = 0; l < L; ++l) {
            if (!alive[l]) continue;
            if (use_parity && !parity_ok(U(l))) continue;
            any = true;
            const double p = P(l, c.n)[C(l, c.n)[1]];                                    // :631-637
            if (p_best < p) { best = l; p_best = p; }
        }
        if (any) return best;
        return pick(false);
    }
    void run(const double* p1, const double* p0, uint8_t* info_out) {                    // :110-128, :150-175
        std::fill(pr.begin(), pr.end(), 0.0);
        std::fill(ps.begin(), ps.end(), 0);
        std::fill(bits.begin(), bits.end(), 0);
        std::fill(alive.begin(), alive.end(), 0);
        pool.reset(L);
        const int l0 = pool.take();
        alive[l0] = 1;
        double* ch = P(l0, 0);
        for (int b = 0; b < c.N; ++b) { ch[2 * b] = p0[b]; ch[2 * b + 1] = p1[b]; }      // :121-124
        for (int phi = 0; phi < c.N; ++phi) {
            calc_p(c.n, phi);
            if (c.frozen[phi]) frozen_step(phi); else info_step(phi);
            if (phi & 1) update_c(c.n, phi);
        }
        const int w = pick(c.crc != 0);
        for (int j = 0; j < c.K; ++j) info_out[j] = U(w)[c.order[j]];
    }
};

template <class R>
void decode_many(const Code& c, const float* llr, int lo, int hi, int L, int minsum_only, uint8_t* out) {
    Decoder<R> d(c, L, minsum_only);
    std::vector<R> in(c.N);
    for (int b = lo; b < hi; ++b) {
        for (int i = 0; i < c.N; ++i) in[i] = (R)llr[(size_t)b * c.N + i];
        d.run(in.data(), out + (size_t)b * c.K);
    }
}

}  // namespace

extern "C" {

// reseed != 0: srand(1) first (= glibc start-of-process rand() state), so the parity
// matrix equals the one a fresh reference process draws.
void* oracle_create(int n, int K, double epsilon, int crc, int reseed) {
    Code* c = new Code;
    c->n = n; c->N = 1 << n; c->K = K; c->crc = crc; c->eps = epsilon;
    make_bitrev(*c);
    if (reseed) srand(1);
    construct(*c, true);
    return c;
}

// Build from explicit tables (e.g. the committed golden construction), bypassing
// rand()/std::sort.
void* oracle_create_from_tables(int n, int K, int crc, const uint8_t* frozen, const uint16_t* order,
                                const uint8_t* crc_matrix) {
    Code* c = new Code;
    c->n = n; c->N = 1 << n; c->K = K; c->crc = crc; c->eps = 0;
    make_bitrev(*c);
    c->frozen.assign(frozen, frozen + c->N);
    c->order.assign(order, order + c->N);
    c->crcm.assign((size_t)crc * K, 0);
    if (crc) c->crcm.assign(crc_matrix, crc_matrix + (size_t)crc * K);
    return c;
}

void oracle_destroy(void* h) { delete static_cast<Code*>(h); }

// read-and-reset the instrumentation counters (meaningful after single-threaded decodes only)
void oracle_stats(long long* out4) { for (int i = 0; i < 4; ++i) { out4[i] = g_stats[i]; g_stats[i] = 0; } }

void oracle_get_construction(void* h, uint8_t* frozen, uint16_t* order, uint8_t* crc_matrix, uint16_t* bitrev) {
    const Code& c = *static_cast<Code*>(h);
    if (frozen) memcpy(frozen, c.frozen.data(), c.N);
    if (order) memcpy(order, c.order.data(), c.N * sizeof(uint16_t));
    if (bitrev) memcpy(bitrev, c.bitrev.data(), c.N * sizeof(uint16_t));
    if (crc_matrix && c.crc) memcpy(crc_matrix, c.crcm.data(), c.crcm.size());
}

void oracle_encode(void* h, const uint8_t* info, uint8_t* coded) { encode(*static_cast<Code*>(h), info, coded); }

void oracle_decode_scl_llr(void* h, const double* llr, int L, uint8_t* info_out) {
    const Code& c = *static_cast<Code*>(h);
    Decoder<double> d(c, L, 0);
    d.run(llr, info_out);
}

// PolarCode.h:31 -- decode_scl_p1(p1, p0, list_size): note the argument order (p1 first).
void oracle_decode_scl_p1(void* h, const double* p1, const double* p0, int L, uint8_t* info_out) {
    const Code& c = *static_cast<Code*>(h);
    ProbDecoder d(c, L);
    d.run(p1, p0, info_out);
}

// B codewords, [B][N] doubles each; returns wall seconds.
double oracle_decode_p1_batch(void* h, const double* p1, const double* p0, int B, int L, uint8_t* info_out, int nthreads) {
    const Code& c = *static_cast<Code*>(h);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B > 0 ? B : 1;
    auto work = [&](int t) {
        const int lo = (int)((long long)B * t / nthreads), hi = (int)((long long)B * (t + 1) / nthreads);
        ProbDecoder d(c, L);
        for (int b = lo; b < hi; ++b) d.run(p1 + (size_t)b * c.N, p0 + (size_t)b * c.N, info_out + (size_t)b * c.K);
    };
    auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// B codewords of float LLRs, widened to `double` (precision 0) or kept in float with the
// same literal formulas (precision 1, a sensitivity probe); minsum_only != 0 swaps the
// reference f rule for pure min-sum (non-parity, for the opt-in fast mode's baseline).
// Returns wall seconds of the decode region.
double oracle_decode_batch(void* h, const float* llr, int B, int L, uint8_t* info_out, int nthreads,
                           int precision, int minsum_only) {
    const Code& c = *static_cast<Code*>(h);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B > 0 ? B : 1;
    auto work = [&](int t) {
        const int lo = (int)((long long)B * t / nthreads), hi = (int)((long long)B * (t + 1) / nthreads);
        if (precision == 1) decode_many<float>(c, llr, lo, hi, L, minsum_only, info_out);
        else decode_many<double>(c, llr, lo, hi, L, minsum_only, info_out);
    };
    auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// PolarCode.cpp:658-785 -- the BPSK/AWGN Monte-Carlo harness with its early-stop and
// "decoded at a lower Eb/N0" shortcut, same RNG objects in the same call order
// (rand() for info bits every 100th run, libstdc++ default_random_engine +
// normal_distribution<double> for noise). Silent (no progress lines).
// counts_out (may be NULL) is [n_list][n_ebno][2] = (num_err, num_run).
void oracle_get_bler_quick(void* h, const double* ebno, int n_ebno, const uint8_t* lists, int n_list,
                           int max_err, int max_runs, double* bler_out, double* counts_out) {
    const Code& c = *static_cast<Code*>(h);
    std::vector<double> nerr((size_t)n_list * n_ebno, 0), nrun((size_t)n_list * n_ebno, 0);
    std::vector<uint8_t> info(c.K, 0), coded(c.N), dec(c.K);
    std::vector<double> noise(c.N), bpsk(c.N), llr(c.N);
    const double N0 = 1.0;
    std::normal_distribution<double> gauss(0.0f, N0);                           // :688
    std::default_random_engine gen;                                             // :689
    for (int run = 0; run < max_runs; ++run) {
        if (run % 100 == 0)                                                     // :703-707
            for (int i = 0; i < c.K; ++i) info[i] = (uint8_t)(rand() % 2);
        for (int i = 0; i < c.N; ++i) noise[i] = gauss(gen);                    // :708-710
        encode(c, info.data(), coded.data());                                   // :712
        for (int i = 0; i < c.N; ++i) bpsk[i] = 2.0f * ((double)coded[i]) - 1.0f;
        for (int li = 0; li < n_list; ++li) {
            bool ok_lower = false;                                              // any prev_decoded
            for (int ei = 0; ei < n_ebno; ++ei) {
                const size_t cell = (size_t)li * n_ebno + ei;
                if (nerr[cell] > max_err) continue;                             // :725-726
                nrun[cell] += 1;                                                // :728
                if (ok_lower) continue;                                         // :730-742
                const double a = std::pow(10.0f, ebno[ei] / 20) * std::sqrt(((double)c.K) / ((double)c.N));
                for (int i = 0; i < c.N; ++i) {
                    const double r = a * bpsk[i] + std::sqrt(N0 / 2) * noise[i]; // :747
                    llr[i] = -4 * r * a / N0;                                    // :752
                }
                Decoder<double> d(c, lists[li], 0);
                d.run(llr.data(), dec.data());
                if (memcmp(dec.data(), info.data(), c.K) != 0) nerr[cell] += 1;  // :758-767
                else ok_lower = true;                                            // :769
            }
        }
    }
    for (size_t i = 0; i < nerr.size(); ++i) {
        bler_out[i] = nerr[i] / nrun[i];                                         // :777-781
        if (counts_out) { counts_out[2 * i] = nerr[i]; counts_out[2 * i + 1] = nrun[i]; }
    }
}

}  // extern "C"
