// agg_inst.cu -- instantiates aggregate_kernel<S2PB_LPL, TSGM> for TSGM = 1..4.
#include "agg_dispatch.h"
#ifndef S2PB_LPL
#error "compile with -DS2PB_LPL=<labels per lane>"
#endif

namespace s2pb {

static constexpr int kLPL = S2PB_LPL;
static constexpr size_t kSmem = (size_t)kNW * kRing * 32 * kLPL * sizeof(float) + (size_t)kNW * kRing * sizeof(float);
static constexpr int kCtaPerSm = (kLPL <= 4) ? 2 : 1;

template <> int agg_configure_lpl<kLPL>()
{
    cudaError_t e = cudaSuccess;
    e = cudaFuncSetAttribute(aggregate_kernel<kLPL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem); if (e) return -1;
    e = cudaFuncSetAttribute(aggregate_kernel<kLPL, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem); if (e) return -1;
    e = cudaFuncSetAttribute(aggregate_kernel<kLPL, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem); if (e) return -1;
    e = cudaFuncSetAttribute(aggregate_kernel<kLPL, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem); if (e) return -1;
    return 0;
}

template <> int agg_launch_lpl<kLPL>(int tsgm, const AggParams &P, int sm_count, cudaStream_t st)
{
    // persistent CTAs, one (or two) per SM; bands are pulled from a global queue in dependency order
    int grid = sm_count * kCtaPerSm;
    int total = P.maxBands * P.nPV;
    if (grid > total) grid = total;
    dim3 block(kNW * 32);
    switch (tsgm) {
    case 1: aggregate_kernel<kLPL, 1><<<grid, block, kSmem, st>>>(P); break;
    case 2: aggregate_kernel<kLPL, 2><<<grid, block, kSmem, st>>>(P); break;
    case 3: aggregate_kernel<kLPL, 3><<<grid, block, kSmem, st>>>(P); break;
    case 4: aggregate_kernel<kLPL, 4><<<grid, block, kSmem, st>>>(P); break;
    default: return -2;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace s2pb
