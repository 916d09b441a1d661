// triangulation_kernels.cuh -- disparity -> (lon, lat, alt) by RPC ray intersection.
//
// Behavioural reference: c/disp_to_h.c:70-141 (disp_to_lonlatalt) and c/rpc.c:279-297 (cubic RPC polynomial),
// :337-348,378-427 (inverse model; direct model evaluated iteratively from it when absent), :429-462 (scaling),
// :480-515 (rpc_height: iterative height search along the epipolar curve).  All float64; one thread per pixel.
// The reference is built -O3 -march=native, so its doubles depend on FMA contraction: parity is held to a
// tolerance far below any geometric meaning (tests/test_triangulation.py), not to the last bit.
#pragma once
#include <cuda_runtime.h>

namespace s2pb {

struct RpcModel {            // same layout as `struct rpc` (c/rpc.h:14-32) / s2p.triangulation.RPCStruct
    double numx[20], denx[20], numy[20], deny[20], scale[3], offset[3];
    double inumx[20], idenx[20], inumy[20], ideny[20], iscale[3], ioffset[3];
    double dmval[4], imval[4];
    double delta;
};

__device__ __forceinline__ double pol20(const double *c, double x, double y, double z)
{   // c/rpc.c:279-297 (note the x/y swap of the reference)
    const double col = y, lig = x, alt = z;
    const double m[20] = {1, lig, col, alt, lig * col, lig * alt, col * alt, lig * lig, col * col, alt * alt,
                          col * lig * alt, lig * lig * lig, lig * col * col, lig * alt * alt, lig * lig * col,
                          col * col * col, col * alt * alt, lig * lig * alt, col * col * alt, alt * alt * alt};
    double r = 0;
#pragma unroll
    for (int i = 0; i < 20; i++) r += c[i] * m[i];
    return r;
}
__device__ __forceinline__ void nrpci(double *res, const RpcModel *p, double x, double y, double z)
{   // c/rpc.c:337-348
    res[0] = pol20(p->inumx, x, y, z) / pol20(p->idenx, x, y, z);
    res[1] = pol20(p->inumy, x, y, z) / pol20(p->ideny, x, y, z);
}
__device__ void nrpc(double *res, const RpcModel *p, double x, double y, double z)
{   // c/rpc.c:414-427 ; iterative branch :378-411 (capped: the reference loops until convergence)
    if (isfinite(p->numx[0])) {
        res[0] = pol20(p->numx, x, y, z) / pol20(p->denx, x, y, z);
        res[1] = pol20(p->numy, x, y, z) / pol20(p->deny, x, y, z);
        return;
    }
    double x0[2], x1[2], x2[2];
    const double xf[2] = {x, y};
    double delta = 1.0;
    if (p->delta) delta = p->delta;
    double lon = -1 * delta, lat = -1 * delta, eps = 2 * delta;
    nrpci(x0, p, lon, lat, z);
    nrpci(x1, p, lon + eps, lat, z);
    nrpci(x2, p, lon, lat + eps, z);
    for (int it = 0; it < 1000; it++) {
        const double d0 = x0[0] - xf[0], d1 = x0[1] - xf[1];
        if (!(d0 * d0 + d1 * d1 > 1e-18)) break;
        const double u[2] = {xf[0] - x0[0], xf[1] - x0[1]};
        const double e1[2] = {x1[0] - x0[0], x1[1] - x0[1]};
        const double e2[2] = {x2[0] - x0[0], x2[1] - x0[1]};
        const double det = e1[0] * e2[1] - e1[1] * e2[0];
        double a0 = e2[1] * u[0] - e2[0] * u[1];
        double a1 = -e1[1] * u[0] + e1[0] * u[1];
        a0 /= det; a1 /= det;
        lon += a0 * eps;
        lat += a1 * eps;
        eps = 0.1;
        nrpci(x0, p, lon, lat, z);
        nrpci(x1, p, lon + eps, lat, z);
        nrpci(x2, p, lon, lat + eps, z);
    }
    res[0] = lon; res[1] = lat;
}
__device__ __forceinline__ void eval_rpc(double *res, const RpcModel *p, double x, double y, double z)
{   // c/rpc.c:429-439
    double t[2];
    nrpc(t, p, (x - p->offset[0]) / p->scale[0], (y - p->offset[1]) / p->scale[1], (z - p->offset[2]) / p->scale[2]);
    res[0] = t[0] * p->iscale[0] + p->ioffset[0];
    res[1] = t[1] * p->iscale[1] + p->ioffset[1];
}
__device__ __forceinline__ void eval_rpci(double *res, const RpcModel *p, double x, double y, double z)
{   // c/rpc.c:442-452
    double t[2];
    nrpci(t, p, (x - p->ioffset[0]) / p->iscale[0], (y - p->ioffset[1]) / p->iscale[1], (z - p->ioffset[2]) / p->iscale[2]);
    res[0] = t[0] * p->scale[0] + p->offset[0];
    res[1] = t[1] * p->scale[1] + p->offset[1];
}
__device__ __forceinline__ void rpc_pair(double *xp, const RpcModel *a, const RpcModel *b, double x, double y, double z)
{   // c/rpc.c:455-462
    double t[2];
    eval_rpc(t, a, x, y, z);
    eval_rpci(xp, b, t[0], t[1], z);
}
__device__ double rpc_height(const RpcModel *a, const RpcModel *b, double xa, double ya, double xb, double yb, double *outerr)
{   // c/rpc.c:480-515: RPCH_MAXIT 100, RPCH_HSTEP 1, RPCH_LAMBDA_STOP 1e-5
    double h = 0;
    for (int t = 0; t < 100; t++) {
        double p[2], q[2];
        rpc_pair(p, a, b, xa, ya, h);
        rpc_pair(q, a, b, xa, ya, h + 1);
        const double a0 = q[0] - p[0], a1 = q[1] - p[1], b0 = xb - p[0], b1 = yb - p[1];
        const double a2 = a0 * a0 + a1 * a1;
        const double lambda = (a0 * b0 + a1 * b1) / a2;
        const double z0 = p[0] + lambda * a0, z1 = p[1] + lambda * a1;
        *outerr = hypot(z0 - xb, z1 - yb);
        h += lambda * 1;
        if (fabs(lambda) < 0.00001) break;
    }
    return h;
}

struct TriParams {
    const float *dispx, *dispy, *msk, *msk_orig;
    int nx, ny, w, h;
    double ha_inv[9], hb_inv[9];
    const RpcModel *rpca, *rpcb;
    float col_min, col_max, row_min, row_max;
    double *lonlatalt;
    float *err;
};
__device__ __forceinline__ void hom_apply(double *y, const double *h, double x0, double x1)
{   // c/disp_to_h.c:14-23
    const double z = h[6] * x0 + h[7] * x1 + h[8];
    y[0] = (h[0] * x0 + h[1] * x1 + h[2]) / z;
    y[1] = (h[3] * x0 + h[4] * x1 + h[5]) / z;
}
// disp_to_lonlatalt, c/disp_to_h.c:70-141
__global__ void triangulate_kernel(const TriParams P)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y * blockDim.y + threadIdx.y;
    if (col >= P.nx || row >= P.ny) return;
    const size_t pix = (size_t)row * P.nx + col;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double o0 = nan, o1 = nan, o2 = nan;
    float oe = __int_as_float(0x7fc00000);
    if (P.msk[pix] != 0.f) {
        double p[2], q[2];
        hom_apply(p, P.ha_inv, (double)col, (double)row);
        const double r0 = round(p[0]), r1 = round(p[1]);
        bool ok = !(r0 < P.col_min || r0 > P.col_max || r1 < P.row_min || r1 > P.row_max);     // image-domain bounding box
        if (ok) {
            const int x = (int)((float)(int)r0 - P.col_min), y = (int)((float)(int)r1 - P.row_min);
            if (x < P.w && y < P.h && P.msk_orig[(size_t)y * P.w + x] == 0.f) ok = false;       // image-domain mask
        }
        if (ok) {
            const double dx = (double)P.dispx[pix], dy = (double)P.dispy[pix];
            hom_apply(q, P.hb_inv, (double)col + dx, (double)row + dy);
            double e = 0, ll[2];
            const double z = rpc_height(P.rpca, P.rpcb, p[0], p[1], q[0], q[1], &e);
            eval_rpc(ll, P.rpca, p[0], p[1], z);
            o0 = ll[0]; o1 = ll[1]; o2 = z; oe = (float)e;
        }
    }
    P.lonlatalt[3 * pix + 0] = o0;
    P.lonlatalt[3 * pix + 1] = o1;
    P.lonlatalt[3 * pix + 2] = o2;
    P.err[pix] = oe;
}

}  // namespace s2pb
