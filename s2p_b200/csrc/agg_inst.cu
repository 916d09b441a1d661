// agg_inst.cu -- instantiates aggregate_kernel<S2PB_LPL, TSGM> for TSGM = 1..4.
#include "agg_dispatch.h"
#ifndef S2PB_LPL
#error "compile with -DS2PB_LPL=<labels per lane>"
#endif

namespace s2pb {

static constexpr int kLPL = S2PB_LPL;
static constexpr size_t kSmem = AggSmem<kLPL>::bytes;
static constexpr int kCtaPerSm = (kLPL <= 4) ? 2 : 1;

template <> int agg_configure_lpl<kLPL>()
{
    cudaError_t e = cudaSuccess;
#define CFG(T, S) e = cudaFuncSetAttribute(aggregate_kernel<kLPL, T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem); if (e) return -1;
    CFG(1, false) CFG(2, false) CFG(3, false) CFG(4, false) CFG(1, true) CFG(2, true) CFG(3, true) CFG(4, true)
#undef CFG
    return 0;
}

template <> int agg_launch_lpl<kLPL>(int tsgm, const AggParams &P, int sm_count, cudaStream_t st)
{
    // persistent CTAs, one (or two) per SM; bands are pulled from a global queue in dependency order
    int grid = sm_count * kCtaPerSm;
    int total = P.maxBands * P.nPV;
    if (grid > total) grid = total;
    dim3 block(kAggThreads);
    const bool scaled = P.lut != nullptr;
#define GO(T) do { if (scaled) aggregate_kernel<kLPL, T, true><<<grid, block, kSmem, st>>>(P); \
                   else aggregate_kernel<kLPL, T, false><<<grid, block, kSmem, st>>>(P); } while (0)
    switch (tsgm) {
    case 1: GO(1); break;
    case 2: GO(2); break;
    case 3: GO(3); break;
    case 4: GO(4); break;
    default: return -2;
    }
#undef GO
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace s2pb
static_assert(s2pb::AggSmem<S2PB_LPL>::bytes <= 227 * 1024, "aggregation CTA exceeds the 227 KB of shared memory of an sm_100 SM");
