// agg_kernel.cuh -- shared device helpers and the persistent MGM aggregation kernel (sm_100a).
//
// Behavioural reference (what must be reproduced bit for bit), paths under
// /root/reference/3rdparty/mgm_multi: census_tools.cc:127-153 (census),
// mgm_costvolume.cc:74-174 (cost volume), mgm_core.cc:75-124,829-1074 (MGM
// recursion, WTA, consensus), mgm_refine.h:45-90 + refine.h:40-92 (sub-pixel).
// The architecture below is ours; nothing is translated from those files.
//
// HBM layout (per view):
//   census  u64  [H][W]
//   C       f16  [H][W][DP]   popcount of the census XOR (exact in f16), +INF = label
//                             outside the pixel's range or outside the image
//   L_p     f32  [H][W][DP]   one volume per scan pass p (kept separate so that the
//                             passes run concurrently and are still summed in the
//                             reference's 1-thread order 0..7)
//   Lmin_p  f32  [H][W], arg_p i16 [H][W]   per-pass minimum and LAST arg-minimum
// DP = 32*LPL slots per pixel (LPL labels per lane); a warp owns one pixel, lane l owns
// slots [l*LPL, (l+1)*LPL), so every access to a pixel's vector is one fully coalesced
// 64*LPL- or 128*LPL-byte request whatever the scan direction of the pass.
//
// All float arithmetic follows the reference's operation order; the translation unit is
// compiled with -fmad=false and every fused operation below is written explicitly.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace s2pb {

#define S2PB_INF __int_as_float(0x7f800000)

constexpr int kMaxPasses = 8;
constexpr int kMaxPV = 16;    // pass-views handled by one aggregation launch (2 views x 8 passes)
constexpr int kNW = 16;       // warps per CTA = scanlines per band
constexpr int kRing = 4;      // ring slots per warp for handing vectors to the next scanline
constexpr int kPublish = 8;   // a band publishes its progress every kPublish pixels

// ------------------------------------------------------------------ small helpers

__device__ __forceinline__ float fmin3f(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));   // FMNMX3
    return r;
}
__device__ __forceinline__ float warp_min_f32(float v)
{
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));  // CREDUX.MIN.F32
    return r;
}
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// per-lane vector load/store of LPL consecutive floats / halfs (16-byte requests when possible)
template <int LPL> __device__ __forceinline__ void ld_vec_cg(const float *p, float (&v)[LPL])
{
    if constexpr (LPL % 4 == 0) {
#pragma unroll
        for (int q = 0; q < LPL / 4; q++) {
            float4 t = __ldcg(reinterpret_cast<const float4 *>(p) + q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else if constexpr (LPL % 2 == 0) {
#pragma unroll
        for (int q = 0; q < LPL / 2; q++) {
            float2 t = __ldcg(reinterpret_cast<const float2 *>(p) + q);
            v[2 * q] = t.x; v[2 * q + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < LPL; q++) v[q] = __ldcg(p + q);
    }
}
template <int LPL> __device__ __forceinline__ void ld_vec(const float *p, float (&v)[LPL])
{
    if constexpr (LPL % 4 == 0) {
#pragma unroll
        for (int q = 0; q < LPL / 4; q++) {
            float4 t = reinterpret_cast<const float4 *>(p)[q];
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else if constexpr (LPL % 2 == 0) {
#pragma unroll
        for (int q = 0; q < LPL / 2; q++) {
            float2 t = reinterpret_cast<const float2 *>(p)[q];
            v[2 * q] = t.x; v[2 * q + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < LPL; q++) v[q] = p[q];
    }
}
template <int LPL> __device__ __forceinline__ void st_vec(float *p, const float (&v)[LPL])
{
    if constexpr (LPL % 4 == 0) {
#pragma unroll
        for (int q = 0; q < LPL / 4; q++)
            reinterpret_cast<float4 *>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else if constexpr (LPL % 2 == 0) {
#pragma unroll
        for (int q = 0; q < LPL / 2; q++) reinterpret_cast<float2 *>(p)[q] = make_float2(v[2 * q], v[2 * q + 1]);
    } else {
#pragma unroll
        for (int q = 0; q < LPL; q++) p[q] = v[q];
    }
}
// raw f16 bits of a lane's LPL costs
template <int LPL> struct HalfPack { unsigned short h[LPL]; };
template <int LPL> __device__ __forceinline__ HalfPack<LPL> ld_cost(const __half *p)
{
    HalfPack<LPL> r;
    if constexpr (LPL % 8 == 0) {
#pragma unroll
        for (int q = 0; q < LPL / 8; q++) {
            uint4 t = __ldg(reinterpret_cast<const uint4 *>(p) + q);
            unsigned w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int k = 0; k < 4; k++) { r.h[8 * q + 2 * k] = w[k] & 0xffff; r.h[8 * q + 2 * k + 1] = w[k] >> 16; }
        }
    } else if constexpr (LPL % 4 == 0) {
#pragma unroll
        for (int q = 0; q < LPL / 4; q++) {
            uint2 t = __ldg(reinterpret_cast<const uint2 *>(p) + q);
            r.h[4 * q] = t.x & 0xffff; r.h[4 * q + 1] = t.x >> 16; r.h[4 * q + 2] = t.y & 0xffff; r.h[4 * q + 3] = t.y >> 16;
        }
    } else if constexpr (LPL % 2 == 0) {
#pragma unroll
        for (int q = 0; q < LPL / 2; q++) {
            unsigned t = __ldg(reinterpret_cast<const unsigned *>(p) + q);
            r.h[2 * q] = t & 0xffff; r.h[2 * q + 1] = t >> 16;
        }
    } else {
#pragma unroll
        for (int q = 0; q < LPL; q++) r.h[q] = __ldg(reinterpret_cast<const unsigned short *>(p) + q);
    }
    return r;
}
// cost value of a stored f16: the popcount itself (census 5x5: ratio = 1, one code word) or
// the reference's scaled cost through a 64-entry table (mgm_costvolume.h:90-91)
__device__ __forceinline__ float cost_value(unsigned short hbits, const float *__restrict__ lut)
{
    float c = __half2float(__ushort_as_half(hbits));
    if (lut != nullptr && c < 64.f) c = lut[(int)c];
    return c;
}

// ------------------------------------------------------------------ MGM aggregation

// A scan pass in "scan coordinates": scanline s = 0..nS-1 in processing order, position
// i = 0..nI-1 along the scanline in processing order; pixel index = base + s*strideS + i*strideI.
// In these coordinates every pass of the reference's table (mgm_core.cc:884-891) has the same
// four neighbours -- A=(i-1,s) in-line, B=(i-1,s-1), Cn=(i,s-1), E=(i+1,s-1) -- listed in one of
// two orders: type 0 (passes 0-3) = A,Cn,B,E ; type 1 (passes 4-7) = E,B,Cn,A.  TSGM takes
// the first TSGM of them.
struct PassDesc {
    int nS, nI;
    long long base;
    int strideS, strideI;
    int type;
    int nBands;
    const __half *C;
    float *L;
    float *Lmin;
    short *arg;
    int *progress;   // [nBands] pixels of the band's last scanline visible in global memory
};
struct AggParams {
    PassDesc pv[kMaxPV];
    int nPV, maxBands;
    float P1, P2;
    int *next_item;
    const int *abort_flag;
    const float *lut;
};

template <int LPL> struct NbVec {
    float v[LPL];
    float l, r;   // slots just left / right of this lane's, from the neighbouring lanes (INF at the ends)
    float m;      // minimum of the whole vector
};
template <int LPL> __device__ __forceinline__ void fill_edges(NbVec<LPL> &n, int lane)
{
    float l = __shfl_up_sync(0xffffffffu, n.v[LPL - 1], 1);
    float r = __shfl_down_sync(0xffffffffu, n.v[0], 1);
    n.l = (lane == 0) ? S2PB_INF : l;
    n.r = (lane == 31) ? S2PB_INF : r;
}
// one neighbour's contribution for slot e: update_costW, mgm_core.cc:92-121 (unit weights)
template <int LPL> __device__ __forceinline__ float nb_term(const NbVec<LPL> &n, int e, float P1, float mP2)
{
    float a = (e == 0) ? n.l : n.v[e - 1];
    float b = (e == LPL - 1) ? n.r : n.v[e + 1];
    float v1 = fminf(a, b) + P1;
    return fmin3f(n.v[e], v1, mP2) - n.m;
}

template <int LPL, int TSGM, int TYPE>
__device__ __forceinline__ void run_band(const PassDesc &pd, int band, float P1, float P2, const float *__restrict__ lut,
                                         const int *abort_flag, float *ring, float *ringmin)
{
    constexpr int DP = 32 * LPL;
    constexpr bool useA = (TYPE == 0) ? true : (TSGM == 4);
    constexpr bool useCn = (TYPE == 0) ? (TSGM >= 2) : (TSGM >= 3);
    constexpr bool useB = (TYPE == 0) ? (TSGM >= 3) : (TSGM >= 2);
    constexpr bool useE = (TYPE == 0) ? (TSGM == 4) : true;
    constexpr bool usePrev = useCn || useB || useE;
    constexpr int SKEW = useE ? 2 : 1;    // scanline s trails scanline s-1 by SKEW pixels
    constexpr int LEAD = useE ? 1 : 0;    // newest previous-scanline pixel needed at position i is i+LEAD

    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    const int s = band * kNW + k;
    const int nI = pd.nI;
    const bool live = s < pd.nS;
    const bool from_global = (k == 0);                       // previous scanline belongs to the previous band
    const bool publish = live && (k == kNW - 1) && (s + 1 < pd.nS);
    const bool has_prev = usePrev && live && s > 0;
    const long long rowbase = pd.base + (long long)s * pd.strideS;
    const long long prevbase = rowbase - pd.strideS;
    float *myring = ring + (size_t)k * kRing * DP;
    float *myringmin = ringmin + k * kRing;
    const float *srcring = ring + (size_t)(k - 1) * kRing * DP;      // only dereferenced when k > 0
    const float *srcringmin = ringmin + (k - 1) * kRing;
    const int *prev_progress = (band > 0) ? pd.progress + (band - 1) : nullptr;
    int avail = 0;                                            // cached progress of the previous band

    NbVec<LPL> wA, wB, wC, wE, nxt;                           // window on the previous scanline + in-line neighbour
#pragma unroll
    for (int e = 0; e < LPL; e++) wA.v[e] = wB.v[e] = wC.v[e] = wE.v[e] = nxt.v[e] = 0.f;
    wA.l = wA.r = wA.m = wB.l = wB.r = wB.m = wC.l = wC.r = wC.m = wE.l = wE.r = wE.m = nxt.l = nxt.r = nxt.m = 0.f;

    // fetch pixel j of the previous scanline (vector + its minimum); edges are filled by the caller
    auto fetch_prev = [&](int j, NbVec<LPL> &dst) {
        if (from_global) {
            while (avail < j + 1) {
                avail = ld_acquire(prev_progress);
                if (*(volatile const int *)abort_flag) break;
            }
            long long q = prevbase + (long long)j * pd.strideI;
            ld_vec_cg<LPL>(pd.L + q * DP + lane * LPL, dst.v);
            dst.m = __ldcg(pd.Lmin + q);
        } else {
            ld_vec<LPL>(srcring + (j & (kRing - 1)) * DP + lane * LPL, dst.v);
            dst.m = srcringmin[j & (kRing - 1)];
        }
    };

    HalfPack<LPL> cnext;
    if (live) cnext = ld_cost<LPL>(pd.C + rowbase * DP + lane * LPL);

    const int nsteps = nI + (kNW - 1) * SKEW;
    for (int t = 0; t < nsteps; t++) {
        const int i = t - k * SKEW;
        if (live && i >= 0 && i < nI) {
            const long long p = rowbase + (long long)i * pd.strideI;
            // ---- this pixel's matching costs (prefetched), prefetch the next pixel's
            float c[LPL];
#pragma unroll
            for (int e = 0; e < LPL; e++) c[e] = cost_value(cnext.h[e], lut);
            if (i + 1 < nI) cnext = ld_cost<LPL>(pd.C + (p + pd.strideI) * DP + lane * LPL);

            // ---- slide the window over the previous scanline
            if (has_prev) {
                if (from_global) {
                    // software pipelined by one step: `nxt` was requested during the previous step
                    if (i == 0) { if (LEAD == 1) { fetch_prev(0, wE); fill_edges<LPL>(wE, lane); } fetch_prev(LEAD, nxt); }
                    wB = wC;
                    if (useE) { wC = wE; wE = nxt; fill_edges<LPL>(wE, lane); }
                    else { wC = nxt; fill_edges<LPL>(wC, lane); }
                    if (i + 1 + LEAD < nI) fetch_prev(i + 1 + LEAD, nxt);
                } else {
                    if (i == 0 && LEAD == 1) { fetch_prev(0, wE); fill_edges<LPL>(wE, lane); }
                    wB = wC;
                    if (useE) { wC = wE; if (i + 1 < nI) { fetch_prev(i + 1, wE); fill_edges<LPL>(wE, lane); } }
                    else { fetch_prev(i, wC); fill_edges<LPL>(wC, lane); }
                }
            }

            // ---- the recursion (border pixels keep L = C: mgm_core.cc:953-960)
            float L[LPL];
            const bool border = (s == 0) || (i == 0) || (i == nI - 1);
            if (border) {
#pragma unroll
                for (int e = 0; e < LPL; e++) L[e] = c[e];
            } else {
                const float mA = wA.m + P2, mB = wB.m + P2, mC = wC.m + P2, mE = wE.m + P2;
#pragma unroll
                for (int e = 0; e < LPL; e++) {
                    float acc = 0.f;
                    if constexpr (TYPE == 0) {
                        if (useA)  { float tt = nb_term<LPL>(wA, e, P1, mA); acc += (TSGM == 2) ? tt * 0.5f : tt; }
                        if (useCn) { float tt = nb_term<LPL>(wC, e, P1, mC); acc += (TSGM == 2) ? tt * 0.5f : tt; }
                        if (useB)  { float tt = nb_term<LPL>(wB, e, P1, mB); acc += tt; }
                        if (useE)  { float tt = nb_term<LPL>(wE, e, P1, mE); acc += tt; }
                    } else {
                        if (useE)  { float tt = nb_term<LPL>(wE, e, P1, mE); acc += (TSGM == 2) ? tt * 0.5f : tt; }
                        if (useB)  { float tt = nb_term<LPL>(wB, e, P1, mB); acc += (TSGM == 2) ? tt * 0.5f : tt; }
                        if (useCn) { float tt = nb_term<LPL>(wC, e, P1, mC); acc += tt; }
                        if (useA)  { float tt = nb_term<LPL>(wA, e, P1, mA); acc += tt; }
                    }
                    if constexpr (TSGM == 3) acc = __fdiv_rn(acc, 3.0f);
                    if constexpr (TSGM == 4) acc = acc * 0.25f;
                    L[e] = c[e] + acc;
                }
            }

            // ---- minimum, LAST arg-minimum (mgm_core.cc:1015-1019)
            float lm = L[0];
#pragma unroll
            for (int e = 1; e < LPL; e++) lm = fminf(lm, L[e]);
            const float m = warp_min_f32(lm);
            int am = -1;
#pragma unroll
            for (int e = 0; e < LPL; e++) if (L[e] == m) am = lane * LPL + e;
            am = __reduce_max_sync(0xffffffffu, am);

            // ---- hand over: registers (in-line), shared ring (next scanline of the band), HBM
#pragma unroll
            for (int e = 0; e < LPL; e++) wA.v[e] = L[e];
            wA.m = m;
            if (useA) fill_edges<LPL>(wA, lane);
            if (usePrev && k + 1 < kNW) {
                st_vec<LPL>(myring + (i & (kRing - 1)) * DP + lane * LPL, L);
                if (lane == 0) myringmin[i & (kRing - 1)] = m;
            }
            st_vec<LPL>(pd.L + p * DP + lane * LPL, L);
            if (lane == 0) { pd.Lmin[p] = m; pd.arg[p] = (short)am; }
            if (publish && (((i + 1) % kPublish) == 0 || i == nI - 1)) {
                __syncwarp();
                if (lane == 0) st_release(pd.progress + band, i + 1);
            }
        }
        __syncthreads();
    }
}

template <int LPL, int TSGM>
__global__ void __launch_bounds__(kNW * 32, (LPL <= 4) ? 2 : 1) aggregate_kernel(const __grid_constant__ AggParams P)
{
    extern __shared__ float smem[];
    float *ring = smem;
    float *ringmin = smem + (size_t)kNW * kRing * 32 * LPL;
    __shared__ int s_item;
    const int total = P.maxBands * P.nPV;
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(P.next_item, 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= total) return;
        if (*(volatile const int *)P.abort_flag) return;
        const int band = item / P.nPV, pvi = item - band * P.nPV;
        const PassDesc &pd = P.pv[pvi];
        if (band >= pd.nBands) continue;
        if (pd.type == 0) run_band<LPL, TSGM, 0>(pd, band, P.P1, P.P2, P.lut, P.abort_flag, ring, ringmin);
        else run_band<LPL, TSGM, 1>(pd, band, P.P1, P.P2, P.lut, P.abort_flag, ring, ringmin);
    }
}

}  // namespace s2pb
