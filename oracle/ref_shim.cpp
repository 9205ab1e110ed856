// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-ABI shim around the UNMODIFIED reference class, compiled from the sources
// where they lie under /root/reference/PolarC (never copied into this repo).
// The reference translation unit is compiled with -DPolarCode=RefPolarCode so
// it can live next to the product's own `PolarCode`; this file includes the
// reference header under the same define, with `private` opened so the
// construction tables (PolarCode.h:43-49) can be read for parity checks.
//
// Built by oracle/Makefile into oracle/_ref/libpolar_ref.so (git-ignored).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <stack>
#include <thread>
#include <chrono>
#include <math.h>

#define PolarCode RefPolarCode
#define private public
#include "PolarCode.h"   // /root/reference/PolarC/PolarCode.h via -I
#undef private

extern "C" {

// reseed != 0: srand(1) first, i.e. glibc's start-of-process rand() state, so the
// random parity ("CRC") matrix equals the one a fresh process would draw
// (PolarCode.cpp:51-56 is the first rand() user in main.cpp).
void* ref_create(int n, int K, double epsilon, int crc, int reseed) {
    if (reseed) srand(1);
    return new RefPolarCode((uint8_t)n, (uint16_t)K, epsilon, (uint16_t)crc);
}

void ref_destroy(void* h) { delete static_cast<RefPolarCode*>(h); }

// frozen[N], order[N], crc_matrix[crc*K] row-major, bitrev[N]; any may be NULL.
void ref_get_construction(void* h, uint8_t* frozen, uint16_t* order, uint8_t* crc_matrix, uint16_t* bitrev) {
    RefPolarCode* p = static_cast<RefPolarCode*>(h);
    const int N = p->_block_length;
    if (frozen) for (int i = 0; i < N; ++i) frozen[i] = p->_frozen_bits[i];
    if (order) for (int i = 0; i < N; ++i) order[i] = p->_channel_order_descending[i];
    if (bitrev) for (int i = 0; i < N; ++i) bitrev[i] = p->_bit_rev_order[i];
    if (crc_matrix)
        for (int r = 0; r < p->_crc_size; ++r)
            for (int j = 0; j < p->_info_length; ++j)
                crc_matrix[r * p->_info_length + j] = p->_crc_matrix[r][j];
}

void ref_encode(void* h, const uint8_t* info, uint8_t* coded) {
    RefPolarCode* p = static_cast<RefPolarCode*>(h);
    std::vector<uint8_t> in(info, info + p->_info_length);
    std::vector<uint8_t> out = p->encode(in);
    memcpy(coded, out.data(), out.size());
}

void ref_decode_scl_llr(void* h, const double* llr, int L, uint8_t* info_out) {
    RefPolarCode* p = static_cast<RefPolarCode*>(h);
    std::vector<double> in(llr, llr + p->_block_length);
    std::vector<uint8_t> out = p->decode_scl_llr(in, (uint16_t)L);
    memcpy(info_out, out.data(), out.size());
}

void ref_decode_scl_p1(void* h, const double* p1, const double* p0, int L, uint8_t* info_out) {
    RefPolarCode* p = static_cast<RefPolarCode*>(h);
    std::vector<double> a(p1, p1 + p->_block_length), b(p0, p0 + p->_block_length);
    std::vector<uint8_t> out = p->decode_scl_p1(a, b, (uint16_t)L);
    memcpy(info_out, out.data(), out.size());
}

// Decode B codewords given as float LLRs (widened to double, so the reference sees
// exactly the values the GPU path sees) with `nthreads` worker threads. The class is
// not re-entrant (PolarCode.h:56-68) so every worker owns a clone of the code object
// (same tables, incl. the random parity matrix). info_out is [B][K] bytes.
// Returns wall seconds of the decode region (slowest worker).
double ref_decode_batch(void* h, const float* llr, int B, int L, uint8_t* info_out, int nthreads) {
    RefPolarCode* master = static_cast<RefPolarCode*>(h);
    const int N = master->_block_length, K = master->_info_length;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B > 0 ? B : 1;
    std::vector<RefPolarCode*> objs(nthreads);
    for (int t = 0; t < nthreads; ++t) objs[t] = new RefPolarCode(*master);
    auto work = [&](int t) {
        std::vector<double> in(N);
        const int lo = (int)((long long)B * t / nthreads), hi = (int)((long long)B * (t + 1) / nthreads);
        for (int b = lo; b < hi; ++b) {
            for (int i = 0; i < N; ++i) in[i] = (double)llr[(size_t)b * N + i];
            std::vector<uint8_t> out = objs[t]->decode_scl_llr(in, (uint16_t)L);
            memcpy(info_out + (size_t)b * K, out.data(), K);
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    for (auto* o : objs) delete o;
    return std::chrono::duration<double>(t1 - t0).count();
}

// The reference Monte-Carlo harness (PolarCode.cpp:658-785); prints its progress
// lines to stdout like the original. bler_out is [n_list][n_ebno].
void ref_get_bler_quick(void* h, const double* ebno, int n_ebno, const uint8_t* lists, int n_list, double* bler_out) {
    RefPolarCode* p = static_cast<RefPolarCode*>(h);
    std::vector<double> e(ebno, ebno + n_ebno);
    std::vector<uint8_t> l(lists, lists + n_list);
    std::vector<std::vector<double>> r = p->get_bler_quick(e, l);
    for (int i = 0; i < n_list; ++i)
        for (int j = 0; j < n_ebno; ++j) bler_out[i * n_ebno + j] = r[i][j];
}

}  // extern "C"
