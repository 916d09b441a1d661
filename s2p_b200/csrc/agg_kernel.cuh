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
//
// Execution model.  A band = kNW consecutive scanlines = one CTA, one warp per scanline, lanes over
// labels.  Scanline s trails scanline s-1 by SKEW pixels (1, or 2 when the neighbour E is used), the
// CTA advances in lock step (one __syncthreads per pixel step) and a warp hands its aggregated vector
// to the next scanline through a shared-memory ring.  Bands of one pass are chained through HBM: the
// last scanline of band b-1 is what warp 0 of band b reads back (from L2) behind a release/acquire
// progress counter.  Everything that comes from global memory -- the f16 costs of every scanline and
// the previous band's vectors -- is staged into shared memory with cp.async kStage steps ahead, so no
// global-memory latency sits on the lock-step critical path.  CTAs are persistent and pull
// (band, pass-view) items from a global queue ordered band-major, which keeps the chain deadlock free.
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

// cp.async pipeline depth in pixel steps (kStage) and slots of the previous-band ring (kR0 > kStage + 1,
// because pixel 0 is staged ahead of the pipeline); shallower for the widest volumes so the CTA fits 227 KB
template <int LPL> struct StageCfg {
    static constexpr int kStage = (LPL <= 12) ? 8 : 2;
    static constexpr int kR0 = (LPL <= 12) ? 16 : 4;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// copy NBYTES (multiple of 16) from global to shared with the whole warp
template <int NBYTES> __device__ __forceinline__ void warp_cp_async(void *smem, const void *gmem, int lane)
{
    constexpr int CH = NBYTES / 16;
#pragma unroll
    for (int q = 0; q < (CH + 31) / 32; q++) {
        int c = lane + 32 * q;
        if (CH % 32 == 0 || c < CH) cp_async16((char *)smem + 16 * c, (const char *)gmem + 16 * c);
    }
}
// a lane's LPL f16 costs from shared memory (widest aligned access)
template <int LPL> __device__ __forceinline__ HalfPack<LPL> lds_cost(const __half *p)
{
    HalfPack<LPL> r;
    if constexpr (LPL % 8 == 0) {
#pragma unroll
        for (int q = 0; q < LPL / 8; q++) {
            uint4 t = reinterpret_cast<const uint4 *>(p)[q];
            unsigned w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int k = 0; k < 4; k++) { r.h[8 * q + 2 * k] = w[k] & 0xffff; r.h[8 * q + 2 * k + 1] = w[k] >> 16; }
        }
    } else if constexpr (LPL % 4 == 0) {
#pragma unroll
        for (int q = 0; q < LPL / 4; q++) {
            uint2 t = reinterpret_cast<const uint2 *>(p)[q];
            r.h[4 * q] = t.x & 0xffff; r.h[4 * q + 1] = t.x >> 16; r.h[4 * q + 2] = t.y & 0xffff; r.h[4 * q + 3] = t.y >> 16;
        }
    } else if constexpr (LPL % 2 == 0) {
#pragma unroll
        for (int q = 0; q < LPL / 2; q++) {
            unsigned t = reinterpret_cast<const unsigned *>(p)[q];
            r.h[2 * q] = t & 0xffff; r.h[2 * q + 1] = t >> 16;
        }
    } else {
#pragma unroll
        for (int q = 0; q < LPL; q++) r.h[q] = reinterpret_cast<const unsigned short *>(p)[q];
    }
    return r;
}
// x / 3 with two fmas: bit-identical to IEEE division for every finite x >= 0
// (exhaustively verified on the CPU: tests/test_host_logic.py::test_div3_trick)
__device__ __forceinline__ float div3_exact(float x)
{
    const float r3 = 0.3333333432674407958984375f;   // RN(1/3)
    float q0 = x * r3;
    float r = fmaf(-3.0f, q0, x);
    return fmaf(r, r3, q0);
}

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

// shared memory carve-up of one CTA (floats first, then halfs; every block 16-byte aligned)
template <int LPL> struct AggSmem {
    static constexpr int DP = 32 * LPL;
    static constexpr int kStage = StageCfg<LPL>::kStage, kR0 = StageCfg<LPL>::kR0;
    static constexpr size_t ring_off = 0;                                           // float [kNW][kRing][DP]
    static constexpr size_t ringm_off = ring_off + sizeof(float) * kNW * kRing * DP; // float [kNW][kRing]
    static constexpr size_t r0_off = ringm_off + sizeof(float) * kNW * kRing;        // float [kR0][DP]   previous band
    static constexpr size_t r0m_off = r0_off + sizeof(float) * kR0 * DP;             // float [kR0]
    static constexpr size_t cst_off = r0m_off + sizeof(float) * kR0;                 // half [kNW][kStage][DP]
    static constexpr size_t bytes = cst_off + sizeof(__half) * kNW * kStage * DP;
};

// smem address (32-bit, shared state space) helpers for cp.async
__device__ __forceinline__ void cp_async16_s(unsigned sa, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4_s(unsigned sa, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
template <int NBYTES> __device__ __forceinline__ void warp_cp_async_s(unsigned sa, const char *gmem, int lane)
{
    constexpr int CH = NBYTES / 16;
#pragma unroll
    for (int q = 0; q < (CH + 31) / 32; q++) {
        int c = lane + 32 * q;
        if (CH % 32 == 0 || c < CH) cp_async16_s(sa + 16 * c, gmem + 16 * c);
    }
}

template <int LPL, int TSGM, int TYPE, bool SCALED>
__device__ __forceinline__ void run_band(const PassDesc &pd, int band, float P1, float P2, const float *__restrict__ lut,
                                         const int *abort_flag, unsigned char *smem)
{
    constexpr int DP = 32 * LPL;
    constexpr bool useA = (TYPE == 0) ? true : (TSGM == 4);
    constexpr bool useCn = (TYPE == 0) ? (TSGM >= 2) : (TSGM >= 3);
    constexpr bool useB = (TYPE == 0) ? (TSGM >= 3) : (TSGM >= 2);
    constexpr bool useE = (TYPE == 0) ? (TSGM == 4) : true;
    constexpr bool usePrev = useCn || useB || useE;
    constexpr int SKEW = useE ? 2 : 1;    // scanline s trails scanline s-1 by SKEW pixels
    constexpr int LEAD = useE ? 1 : 0;    // newest previous-scanline pixel needed at position i is i+LEAD
    constexpr int U = useE ? 3 : 2;       // the window registers rotate with period U: the step loop is unrolled by U
    using SM = AggSmem<LPL>;
    constexpr int kStage = SM::kStage, kR0 = SM::kR0;

    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    const int s = band * kNW + k;
    const int nI = pd.nI;
    const long long strideI = pd.strideI;
    const bool live = s < pd.nS;
    const bool from_global = (k == 0);                       // previous scanline belongs to the previous band
    const bool publish = live && (k == kNW - 1) && (s + 1 < pd.nS);
    const bool has_prev = usePrev && live && s > 0;
    const bool stage_prev = has_prev && from_global;
    const long long rowbase = pd.base + (long long)s * pd.strideS;
    const long long prevbase = rowbase - pd.strideS;

    // ---- shared memory (32-bit shared-space addresses for cp.async, generic pointers for ld/st)
    const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);
    float *ring = reinterpret_cast<float *>(smem + SM::ring_off);
    float *ringm = reinterpret_cast<float *>(smem + SM::ringm_off);
    float *myring = ring + (size_t)k * kRing * DP + lane * LPL;
    float *myringm = ringm + k * kRing;
    const float *srcring = (from_global ? reinterpret_cast<float *>(smem + SM::r0_off) : ring + (size_t)(k - 1) * kRing * DP) + lane * LPL;
    const float *srcringm = from_global ? reinterpret_cast<float *>(smem + SM::r0m_off) : ringm + (k - 1) * kRing;
    const int srcmask = from_global ? (kR0 - 1) : (kRing - 1);
    const __half *cst = reinterpret_cast<__half *>(smem + SM::cst_off) + (size_t)k * kStage * DP + lane * LPL;
    const unsigned cst_s = smem_s + (unsigned)SM::cst_off + (unsigned)(k * kStage * DP * 2);
    const unsigned r0_s = smem_s + (unsigned)SM::r0_off, r0m_s = smem_s + (unsigned)SM::r0m_off;

    // ---- global memory cursors (advance by one pixel of the scanline per step)
    const long long cstep = strideI * (DP * 2), lstep = strideI * (DP * 4);
    const char *c_stage = reinterpret_cast<const char *>(pd.C + rowbase * DP);            // next pixel to stage (costs)
    const char *p_stage = reinterpret_cast<const char *>(pd.L + prevbase * DP);           // next previous-band pixel to stage
    const float *pm_stage = pd.Lmin + prevbase;
    char *l_out = reinterpret_cast<char *>(pd.L + rowbase * DP + lane * LPL);             // this pixel's output vector
    long long p_cur = rowbase;                                                             // this pixel's index
    int j_stage = 0, jp_stage = 0;                                                         // indices of the two staging cursors

    const int *prev_progress = (band > 0) ? pd.progress + (band - 1) : nullptr;
    int avail = 0;                                            // cached progress of the previous band

    auto stage_cost = [&]() {            // stage pixel j_stage of this scanline's costs
        if (j_stage < nI) {
            warp_cp_async_s<DP * 2>(cst_s + (unsigned)((j_stage & (kStage - 1)) * (DP * 2)), c_stage, lane);
            c_stage += cstep;
        }
        j_stage++;
    };
    auto stage_prevband = [&]() {        // (warp 0) stage pixel jp_stage of the previous band's last scanline
        if (stage_prev && jp_stage < nI) {
            int spins = 0;
            while (avail < jp_stage + 1) {
                avail = ld_acquire(prev_progress);
                if (((++spins) & 1023) == 0 && *(volatile const int *)abort_flag) break;
            }
            const unsigned slot = (unsigned)(jp_stage & (kR0 - 1));
            warp_cp_async_s<DP * 4>(r0_s + slot * (DP * 4), p_stage, lane);
            if (lane == 0) cp_async4_s(r0m_s + slot * 4, pm_stage);
            p_stage += lstep;
            pm_stage += strideI;
        }
        jp_stage++;
    };
    auto fetch_prev = [&](int j, NbVec<LPL> &dst) {
        const int slot = j & srcmask;
        ld_vec<LPL>(srcring + slot * DP, dst.v);
        dst.m = srcringm[slot];
        fill_edges<LPL>(dst, lane);
    };

    NbVec<LPL> wA, x0, x1, x2;                                // in-line neighbour + rotating window on the previous scanline
#pragma unroll
    for (int e = 0; e < LPL; e++) wA.v[e] = x0.v[e] = x1.v[e] = x2.v[e] = 0.f;
    wA.l = wA.r = wA.m = x0.l = x0.r = x0.m = x1.l = x1.r = x1.m = x2.l = x2.r = x2.m = 0.f;

    // prologue: kStage-1 groups in flight; group g holds costs(g) and previous-band pixel g+LEAD (+ pixel 0 when LEAD = 1)
    if (live) {
        if (LEAD == 1) stage_prevband();
#pragma unroll
        for (int g = 0; g < kStage - 1; g++) { stage_cost(); stage_prevband(); cp_async_commit(); }
    }

    // one lock-step pixel step; wB / wC / wE are the window registers in their role for this step
    auto step = [&](const int t, NbVec<LPL> &wB, NbVec<LPL> &wC, NbVec<LPL> &wE) {
        const int i = t - k * SKEW;
        if (live && i >= 0 && i < nI) {
            stage_cost();
            stage_prevband();
            cp_async_commit();
            cp_async_wait<kStage - 1>();     // the group of step i has landed
            __syncwarp();
            // ---- this pixel's matching costs
            float c[LPL];
            {
                HalfPack<LPL> cp = lds_cost<LPL>(cst + (i & (kStage - 1)) * DP);
#pragma unroll
                for (int e = 0; e < LPL; e++) {
                    float cc = __half2float(__ushort_as_half(cp.h[e]));
                    if (SCALED) { if (cc < 64.f) cc = lut[(int)cc]; }
                    c[e] = cc;
                }
            }
            // ---- newest pixel of the previous scanline enters the window
            if (has_prev) {
                if (useE) { if (i == 0) fetch_prev(0, wC); if (i + 1 < nI) fetch_prev(i + 1, wE); }
                else fetch_prev(i, wC);
            }
            // ---- the recursion; border pixels keep L = C (mgm_core.cc:953-960)
            float L[LPL];
            const bool border = (s == 0) || (i == 0) || (i == nI - 1);
            {
                const float mA = wA.m + P2, mB = wB.m + P2, mC = wC.m + P2, mE = wE.m + P2;
#pragma unroll
                for (int e = 0; e < LPL; e++) {
                    float acc;
                    if constexpr (TYPE == 0) {
                        acc = nb_term<LPL>(wA, e, P1, mA);
                        if (TSGM == 2) acc *= 0.5f;
                        if (useCn) { float tt = nb_term<LPL>(wC, e, P1, mC); acc += (TSGM == 2) ? tt * 0.5f : tt; }
                        if (useB)  { float tt = nb_term<LPL>(wB, e, P1, mB); acc += tt; }
                        if (useE)  { float tt = nb_term<LPL>(wE, e, P1, mE); acc += tt; }
                    } else {
                        acc = nb_term<LPL>(wE, e, P1, mE);
                        if (TSGM == 2) acc *= 0.5f;
                        if (useB)  { float tt = nb_term<LPL>(wB, e, P1, mB); acc += (TSGM == 2) ? tt * 0.5f : tt; }
                        if (useCn) { float tt = nb_term<LPL>(wC, e, P1, mC); acc += tt; }
                        if (useA)  { float tt = nb_term<LPL>(wA, e, P1, mA); acc += tt; }
                    }
                    if constexpr (TSGM == 3) acc = div3_exact(acc);      // acc is finite and >= 0 on interior pixels
                    if constexpr (TSGM == 4) acc = acc * 0.25f;
                    L[e] = border ? c[e] : c[e] + acc;
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
            if (usePrev) st_vec<LPL>(myring + (i & (kRing - 1)) * DP, L);
            st_vec<LPL>(reinterpret_cast<float *>(l_out), L);
            if (lane == 0) {
                if (usePrev) myringm[i & (kRing - 1)] = m;
                pd.Lmin[p_cur] = m;
                pd.arg[p_cur] = (short)am;
            }
            l_out += lstep;
            p_cur += strideI;
            if (publish && (((i + 1) % kPublish) == 0 || i == nI - 1)) {
                __syncwarp();
                if (lane == 0) st_release(pd.progress + band, i + 1);
            }
        }
        __syncthreads();
    };

    const int nsteps = nI + (kNW - 1) * SKEW;
    for (int t = 0; t < nsteps; t += U) {
        if constexpr (U == 2) {          // window: C = newest, B = the one before
            step(t, x1, x0, x2);
            step(t + 1, x0, x1, x2);
        } else {                         // window: E = newest, then C, then B
            step(t, x1, x2, x0);
            step(t + 1, x2, x0, x1);
            step(t + 2, x0, x1, x2);
        }
    }
    cp_async_wait<0>();
}

template <int LPL, int TSGM, bool SCALED>
__global__ void __launch_bounds__(kNW * 32, (LPL <= 4) ? 2 : 1) aggregate_kernel(const __grid_constant__ AggParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_item;
    const int total = P.maxBands * P.nPV;
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(P.next_item, 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= total) return;
        const int band = item / P.nPV, pvi = item - band * P.nPV;
        const PassDesc &pd = P.pv[pvi];
        if (band >= pd.nBands) continue;
        if (pd.type == 0) run_band<LPL, TSGM, 0, SCALED>(pd, band, P.P1, P.P2, P.lut, P.abort_flag, smem);
        else run_band<LPL, TSGM, 1, SCALED>(pd, band, P.P1, P.P2, P.lut, P.abort_flag, smem);
    }
}

}  // namespace s2pb
