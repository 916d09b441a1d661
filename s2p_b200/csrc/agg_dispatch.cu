// agg_dispatch.cu -- run-time switch over the labels-per-lane instantiations.
#include "agg_dispatch.h"

namespace s2pb {

#define S2PB_FOR_EACH_LPL(X) X(1) X(2) X(3) X(4) X(5) X(6) X(8) X(12) X(16)

#define DECL(n) template <> int agg_launch_lpl<n>(int, const AggParams &, int, cudaStream_t, int); template <> int agg_configure_lpl<n>();
S2PB_FOR_EACH_LPL(DECL)
#undef DECL

int agg_configure()
{
#define CFG(n) if (agg_configure_lpl<n>() != 0) return -1;
    S2PB_FOR_EACH_LPL(CFG)
#undef CFG
    return 0;
}

int agg_launch(int LPL, int tsgm, const AggParams &P, int sm_count, cudaStream_t st, int ctas_per_sm)
{
    switch (LPL) {
#define CASE(n) case n: return agg_launch_lpl<n>(tsgm, P, sm_count, st, ctas_per_sm);
    S2PB_FOR_EACH_LPL(CASE)
#undef CASE
    default: return -2;
    }
}

}  // namespace s2pb
