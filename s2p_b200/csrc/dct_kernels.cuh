// dct_kernels.cuh -- the DCT round trip the reference applies to the MATCHED image of every cost volume, bit for bit.
//
// Behavioural reference: shift() (3rdparty/mgm_multi/mgm_costvolume.cc:23-43) sends the matched image through
// image_shear (shear.c:28-101): per row REDFT10, /n, phase factors cos(ka) / sin(ka) with a = -(pi/n) q, REDFT01 +
// RODFT01, out = 0.5 (sym + antisym) cast to float -- even for q = 0 (mgm_costvolume.cc:50-60, `us[0]`), where it is
// the identity up to double rounding.  That rounding is invisible after the cast back to float32 EXCEPT on pixels that
// are (nearly) zero relative to their row: no-data pixels (NaN -> 0, main_mgm.cc:172-173) come back as +-1e-13-ish
// noise, and the census transform then compares noise with noise.  Two reference builds that differ only in the
// summation order of their DCT agree with each other within rounding-noise statistics, but the identity does not
// (profiles/r02_nodata_spread.md), so the engine reproduces the round trip with the reference's own arithmetic:
//   * transforms are dense matrix products in double, terms added in ascending index order, multiply and add rounded
//     separately (what the oracle build's DCT does; real fftw builds use other orders and differ among themselves);
//   * the coefficient tables are computed on the HOST with libm's cos / sin (device cos differs in the last bit);
//   * q = 0: only rows that hold a pixel with |x| <= rowmax * n * 2^-24 are transformed, and only those pixels are
//     replaced; every other pixel provably returns to its float32 value (margin > 10x over 1e6 random rows,
//     tests/test_oracle.py::test_roundtrip_flag_rule), so a tile without no-data costs one pass over its pixels;
//   * q = 1/2 (SUBPIX = 2 of mgm_multi): every row, both inverse transforms.
// Nothing here is translated from the reference; the transforms are FFTW's published definitions (REDFT10, REDFT01,
// RODFT01, unnormalised).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace s2pb {

// per-image bookkeeping of the round trip, in device memory (so that nothing has to synchronise with the host)
struct RtState {
    int nrows;       // rows that need the transform
    int ncols;       // columns that hold a flagged pixel in any of those rows
    int pad[2];
};

// One block per image row: copy the row to `rt`, find rowmax, flag the pixels that the round trip may change and append
// the row to the list when it has any.  A row of zeros returns exact zeros and is skipped.
__global__ void rt_flag_kernel(const float *__restrict__ img, int w, int h, float *__restrict__ rt, RtState *st, int *__restrict__ rowlist,
                               float *__restrict__ rowthr, unsigned char *__restrict__ colflag)
{
    const int row = blockIdx.x;
    const float *src = img + (size_t)row * w;
    float *dst = rt + (size_t)row * w;
    __shared__ float red[32];
    float m = 0.f;
    for (int i = threadIdx.x; i < w; i += blockDim.x) { float v = src[i]; dst[i] = v; m = fmaxf(m, fabsf(v)); }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) red[0] = m;
    }
    __syncthreads();
    const float rowmax = red[0];
    const float thr = rowmax * ((float)w * (1.f / 16777216.f));
    bool any = false;
    if (rowmax > 0.f)
        for (int i = threadIdx.x; i < w; i += blockDim.x)
            if (fabsf(src[i]) <= thr) { any = true; colflag[i] = 1; }      // (every writer stores the same value)
    const int cnt = __syncthreads_or(any ? 1 : 0);
    if (threadIdx.x == 0) {
        rowthr[row] = (cnt && rowmax > 0.f) ? thr : -1.f;
        if (cnt && rowmax > 0.f) rowlist[atomicAdd(&st->nrows, 1)] = row;
    }
}

// Y[r][o] = (sum_{i=0}^{n-1} T[o][i] * X[r][i]) * scale-by-division, r over the listed rows (or all rows when
// rowlist == nullptr).  The sum runs in ascending i with separately rounded multiply and add, from 0.0, like a plain C
// loop compiled without contraction.  T is [n][n] row-major.  64 outputs x 32 rows per block, 4 x 2 per thread (the FP64
// pipe issues one warp instruction every two cycles, so a small register tile already keeps it busy; the smaller tile
// gives 512 work items on a 1024 x 1024 image for the 148 SMs).
// IN = float (image rows) or double.  DIVN: divide the sum by n (the forward transform's normalisation, shear.c:62-63).
// `mul` (optional, [n]): the output is multiplied by mul[o] (the phase factor cos(o a));
// `mul2`/`Y2` (optional): a second output Y2[r][o-1] = y * mul2[o] for o >= 1 and Y2[r][n-1] = 0 (the antisymmetric
// part's input, shear.c:71-79).
template <typename IN, bool DIVN>
__global__ void __launch_bounds__(256) dct_gemm_kernel(const double *__restrict__ T, const IN *__restrict__ X, int n, int nrows_all,
                                                       const RtState *st, const int *__restrict__ rowlist,
                                                       double *__restrict__ Y, const double *__restrict__ mul,
                                                       const double *__restrict__ mul2, double *__restrict__ Y2)
{
    constexpr int TO = 64, TR = 32, TK = 16, QA = 4, QB = 2;
    __shared__ double Ts[TK][TO + 2];
    __shared__ double Xs[TK][TR + 2];
    __shared__ int rows_s[TR];
    const int nrows = rowlist ? st->nrows : nrows_all;
    const int otiles = (n + TO - 1) / TO, rtiles = (nrows + TR - 1) / TR;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int item = blockIdx.x; item < otiles * rtiles; item += gridDim.x) {
        const int o0 = (item % otiles) * TO, r0 = (item / otiles) * TR;
        __syncthreads();
        if (threadIdx.x < TR) {
            const int rr = r0 + threadIdx.x;
            rows_s[threadIdx.x] = rr < nrows ? (rowlist ? rowlist[rr] : rr) : -1;
        }
        double acc[QA][QB];
#pragma unroll
        for (int a = 0; a < QA; a++)
#pragma unroll
            for (int b = 0; b < QB; b++) acc[a][b] = 0.0;
        for (int k0 = 0; k0 < n; k0 += TK) {
            __syncthreads();
            {   // T tile: 64 outputs x 16 inputs; thread t loads output (t >> 2), inputs 4 (t & 3) .. +3
                const int o = o0 + (threadIdx.x >> 2), kk = 4 * (threadIdx.x & 3);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int k = k0 + kk + q;
                    Ts[kk + q][threadIdx.x >> 2] = (o < n && k < n) ? T[(size_t)o * n + k] : 0.0;
                }
                // X tile: 32 rows x 16 inputs; thread t loads row (t >> 3), inputs 2 (t & 7), + 1
                const int row = rows_s[threadIdx.x >> 3], kx = 2 * (threadIdx.x & 7);
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int k = k0 + kx + q;
                    Xs[kx + q][threadIdx.x >> 3] = (row >= 0 && k < n) ? (double)X[(size_t)row * n + k] : 0.0;
                }
            }
            __syncthreads();
            const int kmax = (n - k0 < TK) ? n - k0 : TK;
            for (int kk = 0; kk < kmax; kk++) {
                double a[QA], b[QB];
#pragma unroll
                for (int q = 0; q < QA; q++) a[q] = Ts[kk][tx + 16 * q];
#pragma unroll
                for (int q = 0; q < QB; q++) b[q] = Xs[kk][ty + 16 * q];
#pragma unroll
                for (int qa = 0; qa < QA; qa++)
#pragma unroll
                    for (int qb = 0; qb < QB; qb++) acc[qa][qb] = __dadd_rn(acc[qa][qb], __dmul_rn(a[qa], b[qb]));
            }
        }
#pragma unroll
        for (int qb = 0; qb < QB; qb++) {
            const int row = rows_s[ty + 16 * qb];
            if (row < 0) continue;
#pragma unroll
            for (int qa = 0; qa < QA; qa++) {
                const int o = o0 + tx + 16 * qa;
                if (o >= n) continue;
                double y = acc[qa][qb];
                if (DIVN) y = __ddiv_rn(y, (double)n);
                if (Y2) {
                    if (o >= 1) Y2[(size_t)row * n + o - 1] = __dmul_rn(y, mul2[o]);
                    if (o == n - 1) Y2[(size_t)row * n + n - 1] = 0.0;
                }
                Y[(size_t)row * n + o] = mul ? __dmul_rn(y, mul[o]) : y;
            }
        }
    }
}

// the columns that hold a flagged pixel, compacted (one block; the order is irrelevant)
__global__ void rt_collist_kernel(const unsigned char *__restrict__ colflag, int w, RtState *st, int *__restrict__ collist)
{
    __shared__ int count;
    if (threadIdx.x == 0) count = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < w; i += blockDim.x)
        if (colflag[i]) collist[atomicAdd(&count, 1)] = i;
    __syncthreads();
    if (threadIdx.x == 0) st->ncols = count;
}

// q = 0, inverse: rt[r][i] = (float)(0.5 * (sum_k T01[i][k] Y[r][k] + 0.0)) for the FLAGGED pixels of the listed rows (the
// antisymmetric part is identically +0: its inputs are Y[k] sin(0) = +-0).  The same tiled product as dct_gemm_kernel, over
// the listed rows x the listed columns only -- a no-data margin is a few percent of the columns, so the inverse costs a few
// percent of the forward transform; with a ragged mask the column list grows towards the full width and the cost towards
// that of the forward one.  T01 is [n][n] row-major (output pixel, coefficient).
__global__ void __launch_bounds__(256) rt_inverse_gemm_kernel(const double *__restrict__ T01, const double *__restrict__ Y, int n,
                                                              const RtState *st, const int *__restrict__ rowlist, const int *__restrict__ collist,
                                                              const float *__restrict__ rowthr, const float *__restrict__ img, float *__restrict__ rt)
{
    constexpr int TO = 64, TR = 32, TK = 16, QA = 4, QB = 2;
    __shared__ double Ts[TK][TO + 2];
    __shared__ double Xs[TK][TR + 2];
    __shared__ int rows_s[TR], cols_s[TO];
    const int nrows = st->nrows, ncols = st->ncols;
    const int otiles = (ncols + TO - 1) / TO, rtiles = (nrows + TR - 1) / TR;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int item = blockIdx.x; item < otiles * rtiles; item += gridDim.x) {
        const int o0 = (item % otiles) * TO, r0 = (item / otiles) * TR;
        __syncthreads();
        if (threadIdx.x < TR) rows_s[threadIdx.x] = (r0 + (int)threadIdx.x < nrows) ? rowlist[r0 + threadIdx.x] : -1;
        else if (threadIdx.x < TR + TO) { const int q = threadIdx.x - TR; cols_s[q] = (o0 + q < ncols) ? collist[o0 + q] : -1; }
        double acc[QA][QB];
#pragma unroll
        for (int a = 0; a < QA; a++)
#pragma unroll
            for (int b = 0; b < QB; b++) acc[a][b] = 0.0;
        for (int k0 = 0; k0 < n; k0 += TK) {
            __syncthreads();
            {
                const int col = cols_s[threadIdx.x >> 2], kk = 4 * (threadIdx.x & 3);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int k = k0 + kk + q;
                    Ts[kk + q][threadIdx.x >> 2] = (col >= 0 && k < n) ? T01[(size_t)col * n + k] : 0.0;
                }
                const int row = rows_s[threadIdx.x >> 3], kx = 2 * (threadIdx.x & 7);
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int k = k0 + kx + q;
                    Xs[kx + q][threadIdx.x >> 3] = (row >= 0 && k < n) ? Y[(size_t)row * n + k] : 0.0;
                }
            }
            __syncthreads();
            const int kmax = (n - k0 < TK) ? n - k0 : TK;
            for (int kk = 0; kk < kmax; kk++) {
                double a[QA], b[QB];
#pragma unroll
                for (int q = 0; q < QA; q++) a[q] = Ts[kk][tx + 16 * q];
#pragma unroll
                for (int q = 0; q < QB; q++) b[q] = Xs[kk][ty + 16 * q];
#pragma unroll
                for (int qa = 0; qa < QA; qa++)
#pragma unroll
                    for (int qb = 0; qb < QB; qb++) acc[qa][qb] = __dadd_rn(acc[qa][qb], __dmul_rn(a[qa], b[qb]));
            }
        }
#pragma unroll
        for (int qb = 0; qb < QB; qb++) {
            const int row = rows_s[ty + 16 * qb];
            if (row < 0) continue;
            const float thr = rowthr[row];
#pragma unroll
            for (int qa = 0; qa < QA; qa++) {
                const int col = cols_s[tx + 16 * qa];
                if (col < 0) continue;
                const size_t p = (size_t)row * n + col;
                if (fabsf(img[p]) <= thr) rt[p] = (float)__dmul_rn(0.5, __dadd_rn(acc[qa][qb], 0.0));
            }
        }
    }
}

// q != 0: out = (float)(0.5 * (sym + antisym)) on every pixel (shear.c:88-89)
__global__ void dct_combine_kernel(const double *__restrict__ sym, const double *__restrict__ anti, size_t n, float *__restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)__dmul_rn(0.5, __dadd_rn(sym[i], anti[i]));
}

// census of the image and of its round-tripped copy in one pass; when no row needed the transform the second code is
// the first (saves the second evaluation on tiles without no-data)
__global__ void census_pair_kernel(const float *__restrict__ img, const float *__restrict__ rt, const RtState *st, int w, int h, int r,
                                   uint64_t *__restrict__ codes, uint64_t *__restrict__ codes_rt)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const bool both = st->nrows > 0;
    float c = img[(size_t)y * w + x], c2 = both ? rt[(size_t)y * w + x] : 0.f;
    uint64_t code = 0, code2 = 0;
    for (int j = -r; j <= r; j++)
        for (int i = -r; i <= r; i++) {
            if (i == 0 && j == 0) continue;
            int xx = x + i, yy = y + j;
            unsigned bit = 0, bit2 = 0;
            if (xx >= 0 && xx < w && yy >= 0 && yy < h) {
                bit = c < img[(size_t)yy * w + xx];
                if (both) bit2 = c2 < rt[(size_t)yy * w + xx];
            }
            code = (code << 1) | bit;
            code2 = (code2 << 1) | bit2;
        }
    codes[(size_t)y * w + x] = code;
    codes_rt[(size_t)y * w + x] = both ? code2 : code;
}

}  // namespace s2pb
