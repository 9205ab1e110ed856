// fast_variants.cuh -- the table of compiled scl_fast_kernel configurations. The table is built from several
// translation units (fast_parts.cu compiled once per POLAR_PART) so that the template instantiations compile in
// parallel; polar_b200.cu concatenates the parts in order, so a variant's index is its position in the lists below.
#pragma once
#include <string.h>
#include "polar_dev.cuh"
#include "scl_fast.cuh"

// ---- fast-kernel variants (scl_fast.cuh) ----
struct FastVariant {
    int nlog, T, lamS, wlog, wpb, bps;
    size_t gx_floats, gs_words;
    int smem_per_warp;
    cudaError_t (*launch)(const fastcommon::Args&, int blocks, cudaStream_t st, const cudaLaunchAttribute* attrs, int nattrs);
    cudaError_t (*prepare)();
};

#define POLAR_NUM_PARTS 4
extern const FastVariant kFastPart0[], kFastPart1[], kFastPart2[], kFastPart3[];
extern const int kFastPartN0, kFastPartN1, kFastPartN2, kFastPartN3;
// the min-sum build of a few of them (fast_parts.cu compiled with -DPOLAR_MINSUM=1 -DPOLAR_FAST_NS=fastms)
extern const FastVariant kFastMsPart[];
extern const int kFastMsPartN;

namespace {

template <class C, int WPB, int BPS>
cudaError_t launch_fast(const fastcommon::Args& a, int blocks, cudaStream_t st, const cudaLaunchAttribute* attrs, int nattrs) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(WPB * 32);
    cfg.dynamicSmemBytes = C::SMEM_PER_WARP * WPB; cfg.stream = st;
    cfg.attrs = const_cast<cudaLaunchAttribute*>(attrs); cfg.numAttrs = nattrs;
    return cudaLaunchKernelEx(&cfg, POLAR_FAST_NS::scl_fast_kernel<C, WPB, BPS>, a);
}
template <class C, int WPB, int BPS>
cudaError_t prepare_fast() {
    return cudaFuncSetAttribute(POLAR_FAST_NS::scl_fast_kernel<C, WPB, BPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                C::SMEM_PER_WARP * WPB);
}
#define POLAR_FAST(NLOG, T, LAMS, WLOG, WPB, BPS)                                                                   \
    { NLOG, T, LAMS, WLOG, WPB, BPS, POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG>::GX_FLOATS, POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG>::GS_WORDS, \
      POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG>::SMEM_PER_WARP, launch_fast<POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG>, WPB, BPS>,               \
      prepare_fast<POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG>, WPB, BPS> }

#define POLAR_FAST_SG(NLOG, T, LAMS, WLOG, SG, WPB, BPS)                                                              \
    { NLOG, T, LAMS, WLOG, WPB, BPS, POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, SG>::GX_FLOATS, POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, SG>::GS_WORDS, \
      POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, SG>::SMEM_PER_WARP, launch_fast<POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, SG>, WPB, BPS>,               \
      prepare_fast<POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, SG>, WPB, BPS> }

#define POLAR_FAST_TM(NLOG, T, LAMS, WLOG, WPB, BPS)                                                                \
    { NLOG, T, LAMS, WLOG, WPB, BPS, POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, 16, 1>::GX_FLOATS, POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, 16, 1>::GS_WORDS, \
      POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, 16, 1>::SMEM_PER_WARP, launch_fast<POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, 16, 1>, WPB, BPS>,               \
      prepare_fast<POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, 16, 1>, WPB, BPS> }

#define POLAR_FAST_TM_SG(NLOG, T, LAMS, WLOG, SG, WPB, BPS)                                                            \
    { NLOG, T, LAMS, WLOG, WPB, BPS, POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, SG, 1>::GX_FLOATS, POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, SG, 1>::GS_WORDS, \
      POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, SG, 1>::SMEM_PER_WARP, launch_fast<POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, SG, 1>, WPB, BPS>,               \
      prepare_fast<POLAR_FAST_NS::Cfg<NLOG, T, LAMS, WLOG, SG, 1>, WPB, BPS> }

}  // namespace
