// agg_dispatch.h -- host-side dispatch to the per-LPL instantiations of aggregate_kernel.
// Each agg_inst.cu object (compiled with -DS2PB_LPL=n) carries the four TSGM variants for one
// labels-per-lane value, so the instantiations build in parallel.
#pragma once
#include "agg_kernel.cuh"

namespace s2pb {
template <int LPL> int agg_launch_lpl(int tsgm, const AggParams &P, int sm_count, cudaStream_t st, int ctas_per_sm);
template <int LPL> int agg_configure_lpl();
int agg_configure();                                   // 0 ok
// ctas_per_sm > 0 caps the persistent CTAs per SM (two launches meant to share the SMs); 0 ok, -1 CUDA error, -2 unsupported
int agg_launch(int LPL, int tsgm, const AggParams &P, int sm_count, cudaStream_t st, int ctas_per_sm = 0);
// experimental chunk-skipping aggregation (agg_chunked.cuh); -2 = shape not served, use the dense kernel
struct ChunkedParams;
int agg_chunked_configure();
int agg_chunked_launch(int tsgm, const ChunkedParams &P, int sm_count, cudaStream_t st);
}  // namespace s2pb
