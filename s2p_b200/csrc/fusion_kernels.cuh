// fusion_kernels.cuh -- n-view height-map merge.
//
// Behavioural reference: s2p/fusion.py:16-68.  merge_n stacks n float32 rasters as float64, subtracts one
// offset per raster, reduces every pixel's n values with an operator ('average_if_close': NaN when
// nanmax - nanmin > threshold, else nanmedian), adds the mean offset and stores float32.  The reference runs
// the operator through np.apply_along_axis, i.e. one Python call per pixel; here it is one thread per pixel
// in float64, bit-identical.
#pragma once
#include <cuda_runtime.h>

namespace s2pb {

constexpr int kMaxFusion = 16;
enum { FUSE_AVERAGE_IF_CLOSE = 0, FUSE_NANMEDIAN = 1, FUSE_NANMEAN = 2, FUSE_NANMIN = 3, FUSE_NANMAX = 4 };
struct FusionParams {
    const float *in[kMaxFusion];
    double offset[kMaxFusion];
    int n, op;
    double threshold, mean_offset;
    size_t npix;
    float *out;
};

__global__ void fusion_kernel(const FusionParams P)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.npix) return;
    double v[kMaxFusion];
    int m = 0;
    double mx = 0, mn = 0, sum = 0;
    for (int k = 0; k < P.n; k++) {
        const double x = (double)P.in[k][i] - P.offset[k];
        if (x == x) {                       // insertion sort of the non-NaN values
            int q = m++;
            while (q > 0 && v[q - 1] > x) { v[q] = v[q - 1]; q--; }
            v[q] = x;
            sum += x;                       // np.nanmean adds in index order with NaN replaced by 0
        }
    }
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double r = nan;
    if (m > 0) {
        mn = v[0]; mx = v[m - 1];
        const double med = (m & 1) ? v[m / 2] : 0.5 * (v[m / 2 - 1] + v[m / 2]);   // np.nanmedian: mean of the two middle values
        switch (P.op) {
        case FUSE_AVERAGE_IF_CLOSE: r = (mx - mn > P.threshold) ? nan : med; break;
        case FUSE_NANMEDIAN: r = med; break;
        case FUSE_NANMEAN: r = sum / (double)m; break;
        case FUSE_NANMIN: r = mn; break;
        default: r = mx; break;
        }
    }
    P.out[i] = (float)(r + P.mean_offset);
}

}  // namespace s2pb
