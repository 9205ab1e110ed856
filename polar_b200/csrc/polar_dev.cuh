// polar_dev.cuh -- device helpers shared by every kernel of the library (the arithmetic contract of DESIGN.md
// section 2: the reference's check-node rule and metric updates, PolarC/PolarCode.cpp:438-446, 483, 505-506) and the
// column-pointer packing. Everything sits in an anonymous namespace: each translation unit gets its own copy.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <math.h>

#define FULL_MASK 0xffffffffu

namespace {

constexpr int kMaxList = 127;       // the reference's own limit (uint8_t loop counters, PolarCode.cpp:497-605)
constexpr int kMaxNWarp = 13;      // log2 block length of the warp kernels' pointer packing (12 layers x 5 bits)
constexpr int kMaxN = 15;          // the reference's own limit: block length is a uint16_t (PolarCode.h:40); n = 14, 15
                                   // run on the block-per-codeword kernel (8-bit pointers, 16 layers)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

template <class Real, class In = Real>
struct DecodeArgsT {
    const In* llr;               // [B][N] (converted to Real on load)
    uint32_t* out;               // [B][KW]
    const int* list;             // null, or the codeword indices to decode (strict mode's re-decode of flagged codewords)
    const int* count;            // with `list`: number of entries (device memory, written by the kernel before this one)
    const uint32_t* frozen_words;// [max(1,N/32)], bit phi set = frozen
    const uint16_t* info_order;  // [K + crc]
    const uint32_t* crc_masks;   // [crc][NW] over phi
    Real* gx;                    // per-warp LLR scratch rows (32 values each)
    uint32_t* gs;                // per-warp partial-sum scratch rows (32 words each)
    unsigned long long gx_stride;// values per warp
    unsigned long long gs_stride;// words per warp
    int B, n, K, crc, L;
    int W;                       // lanes per codeword (power of two >= L)
    int lamS;                    // first LLR layer kept in shared memory (1..n)
    int smem_x_rows;             // rows of 32 floats per warp
    int smem_s_rows;             // rows of 32 words per warp
    int s_off[kMaxN + 2];        // word-row offset of S layer lam (global for lam < lamS, shared otherwise)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// log(1 + exp(-x)) for x >= 0, in (0, ln 2].
__device__ __forceinline__ float log1p_exp_neg(float x) {
    return kLn2 * lg2_approx(1.0f + ex2_approx(-kLog2e * x));
}

// The reference's check-node rule (PolarCode.cpp:438-446): box-plus when both magnitudes
// are strictly below 40, sign * min otherwise (sgn(0) = 0, which min() already yields).
// Box-plus is evaluated as sign*min + log1p(e^-|a+b|) - log1p(e^-|a-b|): algebraically the
// reference's log((e^(a+b)+1)/(e^a+e^b)), but with an absolute error of a few 1e-7 at every
// magnitude (the literal form loses that much *relative* to e^40 in fp32).
__device__ __forceinline__ float f_rule(float a, float b) {
    const float ma = fabsf(a), mb = fabsf(b);
    const float mn = fminf(ma, mb);
    const float r = __int_as_float(__float_as_int(mn) | ((__float_as_int(a) ^ __float_as_int(b)) & 0x80000000));
    const float s = fabsf(a + b), d = fabsf(a - b);
    // branch-free: the correction is always evaluated (4 MUFU) and scaled by 0 above the threshold
    const float diff = lg2_approx(1.0f + ex2_approx(-kLog2e * s)) - lg2_approx(1.0f + ex2_approx(-kLog2e * d));
    const float scale = (fmaxf(ma, mb) < 40.0f) ? kLn2 : 0.0f;
    return fmaf(diff, scale, r);
}

// log(1 + exp(x)) with the double-precision reference's corner behaviour
// (PolarCode.cpp:483,505-506): +inf once exp(x) overflows a double (x > 709.78...),
// exactly 0 once 1 + exp(x) rounds to 1 in double (x < -36.7368...).
__device__ __forceinline__ float softplus_ref(float x) {
    float r = fmaxf(x, 0.0f) + log1p_exp_neg(fabsf(x));
    if (x >= 709.78271484375f) r = CUDART_INF_F;
    if (x <= -36.7368f) r = 0.0f;
    return r;
}


// ---- arithmetic by evaluation type. float: the throughput contract above. double: the reference's
// literal formulas (PolarCode.cpp:438-446, 483, 505-506) evaluated in double like the reference itself.
template <class Real> struct Arith;
template <> struct Arith<float> {
    static __device__ __forceinline__ float f(float a, float b) { return f_rule(a, b); }
    static __device__ __forceinline__ float softplus(float x) { return softplus_ref(x); }
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
};
template <> struct Arith<double> {
    static __device__ __forceinline__ double f(double a, double b) {
        const double ma = fabs(a), mb = fabs(b);
        if (40.0 > fmax(ma, mb)) return log((exp(a + b) + 1.0) / (exp(a) + exp(b)));
        const double sa = (a < 0) ? -1.0 : (double)(a > 0), sb = (b < 0) ? -1.0 : (double)(b > 0);
        return sa * sb * fmin(ma, mb);
    }
    static __device__ __forceinline__ double softplus(double x) { return log(1.0 + exp(x)); }
    static __device__ __forceinline__ double inf() { return CUDART_INF; }
};
template <class Real> __device__ __forceinline__ Real rmin(Real a, Real b) { return a < b ? a : b; }
template <class Real> __device__ __forceinline__ Real rmax(Real a, Real b) { return a > b ? a : b; }

template <class Real>
__device__ __forceinline__ Real group_min(Real v, int W) {
    for (int o = W >> 1; o > 0; o >>= 1) v = rmin<Real>(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
template <class Real>
__device__ __forceinline__ Real group_max(Real v, int W) {
    for (int o = W >> 1; o > 0; o >>= 1) v = rmax<Real>(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

__device__ __forceinline__ unsigned long long set_ptr(unsigned long long p, int idx, unsigned lane) {
    const int sh = 5 * idx;
    return (p & ~(31ull << sh)) | ((unsigned long long)lane << sh);
}
__device__ __forceinline__ int get_ptr(unsigned long long p, int idx) { return (int)((p >> (5 * idx)) & 31ull); }

}  // namespace

