/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the s2p stereo-matching hot path.
 *
 * A plain-C restatement of what the reference binaries `mgm` / `mgm_multi`
 * (built from /root/reference/3rdparty/mgm_multi, pinned 06b262b9) compute for
 * the flags s2p passes (s2p/block_matching.py:155-186,269-308).  It is NOT part
 * of the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load liboracle.so.  The product
 * library (s2p_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks this file
 * bit-for-bit against the unmodified reference binary (oracle/_ref/mgm, built
 * by oracle/Makefile from the reference's own sources) at OMP_NUM_THREADS=1,
 * and tests/golden/ holds outputs of that binary.  The reference's own test
 * suite pins no disparity values for this path.
 *
 * Layout used here (ours, not the reference's): cost volumes are dense
 * [pixel][label] float arrays of D = gmax-gmin+1 slots per pixel with a
 * per-pixel valid label range [lo,hi]; slots outside the range hold +INF, which
 * is what the reference's bounds-checked vector read returns (dvec.cc:138).
 *
 * All arithmetic is float32 in the reference's operation order; this file is
 * compiled with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "fftw3.h" /* shim: DCT used by the reference's row shift */

typedef struct {
    int ndir;            /* -O  (2,4,8)                     main_mgm.cc:139            */
    int tsgm;            /* TSGM (1..4)                     mgm_multiscale.cc:120      */
    int census_win;      /* CENSUS_NCC_WIN (3,5,7)          mgm_costvolume.h:72        */
    float P1, P2;        /* -P1 -P2                         main_mgm.cc:140-141        */
    int median;          /* MEDIAN radius (0 = off)         mgm_core.h:11              */
    int lr_mode;         /* TESTLRRL (0,1,2)                mgm_core.h:8               */
    float lr_tau;        /* TESTLRRL_TAU                    mgm_core.h:9               */
    float mindiff;       /* MINDIFF (<0 = off)              mgm_multiscale.cc:126      */
    int remove_small_cc; /* REMOVESMALLCC (0 = off)         mgm_multiscale.cc:125      */
    int subpix;          /* SUBPIX (mgm_multi only)         main_mgm_multi.cc:73       */
    int scales;          /* -S ; <0 = single-scale `mgm`    main_mgm_multi.cc:113      */
    int refine;          /* -s : 0 none, 1 vfit, 2 parabola mgm_refine.h:19-22         */
    int fix_overcount;   /* TSGM_FIX_OVERCOUNT              mgm_multiscale.cc:121      */
    int dct_shift;       /* 1: reproduce the DCT round trip of shift() (mgm_costvolume.cc:23-43);
                            0: treat a zero translation as the identity it is meant to be */
    int cost;            /* -t : 0 census, 1 ad, 2 sd, 3 ncc, 4 btad, 5 btsd   mgm_costvolume.h:186-197 */
} orc_params;

#define ORC_INF INFINITY

static inline float fmin3_(float a, float b, float c)
{   /* mgm_core.cc:34-40 */
    float m = a;
    if (m > b) m = b;
    if (m > c) m = c;
    return m;
}
#define MINF(a, b) (((a) < (b)) ? (a) : (b)) /* mgm_core.cc:28 */
/* `x + P*w` of update_costW (mgm_core.cc:93-94,99-100,106-107,113-114): whether the reference build rounds it as
 * one fma or as mul + add depends on what gcc hoisted out of the label loop; bit 2n of the mask = the +-1 term of
 * neighbour n uses an fma, bit 2n+1 = its P2 term does.  Pinned against the reference binary (tests/test_oracle.py);
 * with w = 1 or power-of-two penalties every variant gives the same floats. */
int orc_fma_mask = 0xFA;   /* gcc 13.3 -O3 -march=x86-64-v3 build of mgm_core.cc: n0, n1: mul+add / fma; n2, n3: fma / fma */
void orc_set_fma_mask(int m) { orc_fma_mask = m; }
static inline float orc_pw(float x, float P, float w, int use_fma) { return use_fma ? fmaf(P, w, x) : x + P * w; }
#define ORC_W_V1(n, x, P, w) orc_pw((x), (P), (w), (orc_fma_mask >> (2 * (n))) & 1)
#define ORC_W_V2(n, x, P, w) orc_pw((x), (P), (w), (orc_fma_mask >> (2 * (n) + 1)) & 1)

/* ------------------------------------------------------------------ census */

/* census_tools.cc:38-57: window scanned row-major, centre skipped, bit = (centre < nb),
 * out-of-image neighbour is NaN so the bit is 0.  Bit order inside the code does not
 * matter for the Hamming distance; we put the first neighbour in the top bit. */
void orc_census(const float *img, int w, int h, int win, uint64_t *codes)
{
    int r = win / 2;
#pragma omp parallel for
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float c = img[y * w + x];
            uint64_t code = 0;
            for (int j = -r; j <= r; j++)
                for (int i = -r; i <= r; i++) {
                    if (!i && !j) continue;
                    int xx = x + i, yy = y + j;
                    int bit = 0;
                    if (xx >= 0 && xx < w && yy >= 0 && yy < h) bit = c < img[yy * w + xx];
                    code = (code << 1) | (uint64_t)bit;
                }
            codes[y * w + x] = code;
        }
}

/* -------------------------------------------------------- sub-pixel shift */

/* shear.c:28-101 with shear = 0: out(x) = in(x + q) via DCT phase shift, in double.
 * Called by shift() (mgm_costvolume.cc:23-43) with translation -q, even for q = 0. */
static void row_shift_dct(const float *in, float *out, int w, int h, float q)
{
    int n = w;
    double *li = malloc(sizeof(double) * n), *dct = malloc(sizeof(double) * n);
    double *dst = malloc(sizeof(double) * n), *os = malloc(sizeof(double) * n);
    double *oa = malloc(sizeof(double) * n);
    fftw_plan pf = fftw_plan_r2r_1d(n, li, dct, FFTW_REDFT10, FFTW_ESTIMATE);
    fftw_plan pc = fftw_plan_r2r_1d(n, dct, os, FFTW_REDFT01, FFTW_ESTIMATE);
    fftw_plan ps = fftw_plan_r2r_1d(n, dst, oa, FFTW_RODFT01, FFTW_ESTIMATE);
    float translation = -q;           /* mgm_costvolume.cc:35 passes (0., -q) as floats */
    for (int row = 0; row < h; row++) {
        for (int i = 0; i < n; i++) li[i] = in[row * n + i];
        fftw_execute(pf);
        for (int i = 0; i < n; i++) dct[i] /= n;
        double t = row * 0.0f + translation;
        double a = (M_PI / n) * t;
        dct[0] *= cos(0 * a);
        for (int k = 1; k <= n - 1; k++) {
            dst[k - 1] = dct[k];
            dst[k - 1] *= sin(k * a);
            dct[k] *= cos(k * a);
        }
        dst[n - 1] = 0;
        fftw_execute(pc);
        fftw_execute(ps);
        for (int i = 0; i < n; i++) out[row * n + i] = 0.5 * (os[i] + oa[i]);
    }
    fftw_destroy_plan(pf); fftw_destroy_plan(pc); fftw_destroy_plan(ps);
    free(li); free(dct); free(dst); free(os); free(oa);
}

void orc_shift(const float *in, float *out, int w, int h, float q, int dct_shift)
{
    if (!dct_shift && q == 0.f) memcpy(out, in, sizeof(float) * (size_t)w * h);
    else row_shift_dct(in, out, w, h, q);
}

/* ------------------------------------------------------------ cost volume */

static inline int popc64(uint64_t v) { return __builtin_popcountll(v); }
static inline int goodmod(int a, int b) { int r = a % b; return r < 0 ? r + b : r; }

/* Per-pixel label range from float bound images: allocate_costvolume, mgm_costvolume.cc:63-72 */
void orc_ranges(const float *dminI, const float *dmaxI, int npix, int zoom, int *lo, int *hi)
{
    for (int i = 0; i < npix; i++) {
        float a = dminI[i] * (float)zoom, b = dmaxI[i] * (float)zoom; /* mgm_multiscale.cc:218-219 */
        lo[i] = (int)floorf(a);
        hi[i] = (int)ceilf(b);
    }
}

/* allocate_and_fill_sgm_costvolume (mgm_costvolume.cc:74-174) for prefilter = distance = census.
 * u: reference image of this view, v: matched image (both NaN-free).  C is dense, D slots/pixel,
 * slot k <-> label gmin+k.  */
void orc_costvolume_census(const float *u, const float *v, int w, int h,
                           const int *lo, const int *hi, int gmin, int D,
                           int win, int zoom, int dct_shift, float *C)
{
    size_t npix = (size_t)w * h;
    uint64_t *cu = malloc(sizeof(uint64_t) * npix);
    uint64_t **cv = malloc(sizeof(uint64_t *) * zoom);
    float *tmp = malloc(sizeof(float) * npix);
    orc_census(u, w, h, win, cu);
    for (int z = 0; z < zoom; z++) {      /* alloc_prefiltered_fourier_subpix_interp :50-60 */
        cv[z] = malloc(sizeof(uint64_t) * npix);
        orc_shift(v, tmp, w, h, ((float)z) / ((float)zoom), dct_shift);
        orc_census(tmp, w, h, win, cv[z]);
    }
    int nbits = win * win - 1;
    int nch = (nbits / 8 + 3) / 4;        /* census_tools.cc:84 floats per code */
    const float ratio = 5 * 5 / ((double)win * win); /* mgm_costvolume.h:90 */
#pragma omp parallel for
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            size_t p = (size_t)y * w + x;
            float *Cp = C + p * D;
            for (int k = 0; k < D; k++) Cp[k] = ORC_INF;
            int allinvalid = 1;
            for (int o = lo[p]; o <= hi[p]; o++) {
                int qx = x + (int)floor((double)o / zoom); /* :148 */
                int z = goodmod(o, zoom);
                float e = ORC_INF;                            /* truncDist = inf */
                if (qx >= 0 && qx < w) {
                    float r = (float)popc64(cu[p] ^ cv[z][(size_t)y * w + qx]);
                    e = r * 1.0 * ratio / nch;                /* mgm_costvolume.h:91 */
                }
                Cp[o - gmin] = e;
                if (isfinite(e)) allinvalid = 0;
            }
            if (allinvalid)                                   /* :166-171 */
                for (int o = lo[p]; o <= hi[p]; o++) Cp[o - gmin] = 0;
        }
    for (int z = 0; z < zoom; z++) free(cv[z]);
    free(cv); free(cu); free(tmp);
}

/* The other distances of the reference's table (mgm_costvolume.h:186-197) on one channel, NaN-free images.
 * Where the reference's -O3 -march=native build contracts a*b+c into one fma (checked against the binary,
 * tests/test_oracle.py), the fma is written out; this file itself is compiled with -ffp-contract=off. */
static inline float img_at(const float *a, int w, int x, int y) { return a[(size_t)y * w + x]; }

static float cost_bt(const float *u, const float *v, int w, int x, int y, int qx)
{   /* BTAD, mgm_costvolume.h:96-124: half-sample interpolants in double ("/2.0"), then float min/max */
    float IL = img_at(u, w, x, y), ILp = IL, ILm = IL;
    if (x < w - 1) ILp = (IL + img_at(u, w, x + 1, y)) / 2.0;
    if (x >= 1) ILm = (IL + img_at(u, w, x - 1, y)) / 2.0;
    float IR = img_at(v, w, qx, y), IRp = IR, IRm = IR;
    if (qx < w - 1) IRp = (IR + img_at(v, w, qx + 1, y)) / 2.0;
    if (qx >= 1) IRm = (IR + img_at(v, w, qx - 1, y)) / 2.0;
    float IminR = fminf(IRm, fminf(IRp, IR)), ImaxR = fmaxf(IRm, fmaxf(IRp, IR));
    float IminL = fminf(ILm, fminf(ILp, IL)), ImaxL = fmaxf(ILm, fmaxf(ILp, IL));
    float dLR = fmaxf(0.f, fmaxf(IL - ImaxR, IminR - IL));
    float dRL = fmaxf(0.f, fmaxf(IR - ImaxL, IminL - IR));
    return fabsf(MINF(dLR, dRL));
}

static float cost_ncc(const float *u, const float *v, int w, int h, int x, int y, int qx, int win)
{   /* computeC_clippedNCC, mgm_costvolume.h:152-180: window scanned x-major (outer i = x offset) */
    int hw = win / 2;
    float mu1 = 0, mu2 = 0, s1 = 0, s2 = 0, prod = 0;
    int n = 0;
    for (int i = -hw; i <= hw; i++)
        for (int j = -hw; j <= hw; j++) {
            int px = x + i, py = y + j, rx = qx + i;
            if (px < 0 || px >= w || py < 0 || py >= h || rx < 0 || rx >= w) return ORC_INF;   /* valnan -> NaN */
            float v1 = img_at(u, w, px, py), v2 = img_at(v, w, rx, py);
            mu1 += v1; mu2 += v2;
            s1 = fmaf(v1, v1, s1); s2 = fmaf(v2, v2, s2); prod = fmaf(v1, v2, prod);
            n++;
        }
    mu1 /= n; mu2 /= n; s1 /= n; s2 /= n; prod /= n;
    float num = fmaf(-mu1, mu2, prod);
    float den = fmaf(-mu1, mu1, s1) * fmaf(-mu2, mu2, s2);
    double dd = (0.0000001 > den) ? 0.0000001 : (double)den;
    float NCC = 0;
    NCC += num / sqrt(dd);
    float cl = MINF(NCC, 1.f);
    cl = (0 > cl) ? 0 : cl;
    return (1 - cl) * 64;
}

/* allocate_and_fill_sgm_costvolume (mgm_costvolume.cc:74-174) for any distance, prefilter none (census implies the
 * census prefilter, :98-102).  cost: 0 census, 1 ad, 2 sd, 3 ncc, 4 btad, 5 btsd. */
void orc_costvolume(const float *u, const float *v, int w, int h, const int *lo, const int *hi, int gmin, int D,
                    int win, int zoom, int dct_shift, int cost, float *C)
{
    if (cost == 0) { orc_costvolume_census(u, v, w, h, lo, hi, gmin, D, win, zoom, dct_shift, C); return; }
    size_t npix = (size_t)w * h;
    float **vs = malloc(sizeof(float *) * zoom);
    for (int z = 0; z < zoom; z++) {      /* alloc_prefiltered_fourier_subpix_interp :50-60 */
        vs[z] = malloc(sizeof(float) * npix);
        orc_shift(v, vs[z], w, h, ((float)z) / ((float)zoom), dct_shift);
    }
#pragma omp parallel for
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            size_t p = (size_t)y * w + x;
            float *Cp = C + p * D;
            for (int k = 0; k < D; k++) Cp[k] = ORC_INF;
            int allinvalid = 1;
            for (int o = lo[p]; o <= hi[p]; o++) {
                int qx = x + (int)floor((double)o / zoom);
                const float *vz = vs[goodmod(o, zoom)];
                float e = ORC_INF;
                if (qx >= 0 && qx < w) {
                    float d = img_at(u, w, x, y) - img_at(vz, w, qx, y);
                    d = (d > -d) ? d : -d;
                    if (cost == 1) e = d;
                    else if (cost == 2) e = d * d;
                    else if (cost == 3) e = cost_ncc(u, vz, w, h, x, y, qx, win);
                    else { float b = cost_bt(u, vz, w, x, y, qx); e = (cost == 4) ? b : b * b; }
                }
                Cp[o - gmin] = e;
                if (isfinite(e)) allinvalid = 0;
            }
            if (allinvalid)
                for (int o = lo[p]; o <= hi[p]; o++) Cp[o - gmin] = 0;
        }
    for (int z = 0; z < zoom; z++) free(vs[z]);
    free(vs);
}

/* ------------------------------------------------------------ aggregation */

/* Pass table, mgm_core.cc:884-891: four neighbour offsets, scan orientation. */
typedef struct { int d[4][2]; int inc_x, inc_y, row_major; } orc_pass;
static const orc_pass PASSES[8] = {
    {{{-1, 0}, {0, -1}, {-1, -1}, {1, -1}}, 1, 1, 1},
    {{{1, 0}, {0, 1}, {1, 1}, {-1, 1}}, 0, 0, 1},
    {{{0, 1}, {-1, 0}, {-1, 1}, {-1, -1}}, 1, 0, 0},
    {{{0, -1}, {1, 0}, {1, -1}, {1, 1}}, 0, 1, 0},
    {{{-1, -1}, {1, -1}, {0, -1}, {1, 0}}, 0, 1, 1},
    {{{1, -1}, {1, 1}, {1, 0}, {0, 1}}, 0, 0, 0},
    {{{1, 1}, {-1, 1}, {0, 1}, {-1, 0}}, 1, 0, 1},
    {{{-1, 1}, {-1, -1}, {-1, 0}, {0, -1}}, 1, 1, 0},
};

/* One pass of mgm_naive_parallelism's scan (mgm_core.cc:910-1024) into its own volume L
 * (dense, INF outside the pixel's range).  Lmin[p] = min of L[p][*]; arg[p] = LAST label
 * attaining it (:1017-1019). */
static void orc_one_pass(const float *C, const int *lo, const int *hi, int w, int h, int gmin, int D,
                         float P1, float P2, int tsgm, int pass, const float *wgt, float *L, float *Lmin, float *arg)
{
    const orc_pass *ps = &PASSES[pass];
    int maxii = w, maxjj = h;
    if (!ps->row_major) { maxii = h; maxjj = w; }
    for (int jj = 0; jj < maxjj; jj++)
        for (int ii = 0; ii < maxii; ii++) {
            int x = ii, y = jj;
            if (!ps->row_major) { x = jj; y = ii; }
            if (!ps->inc_x) x = w - 1 - x;
            if (!ps->inc_y) y = h - 1 - y;
            size_t p = (size_t)y * w + x;
            const float *Cp = C + p * D;
            float *Lp = L + p * D;
            int inside = 1;
            size_t q[4];
            for (int k = 0; k < 4; k++) {                     /* :957-960, all four, whatever TSGM */
                int nx = x + ps->d[k][0], ny = y + ps->d[k][1];
                if (nx < 0 || nx >= w || ny < 0 || ny >= h) inside = 0;
                else q[k] = (size_t)ny * w + nx;
            }
            for (int k = 0; k < D; k++) Lp[k] = ORC_INF;
            if (!inside) {
                for (int o = lo[p]; o <= hi[p]; o++) Lp[o - gmin] = Cp[o - gmin];  /* :953 */
            } else {
                /* -wl / -wr weights: compute_mgm_weights_copyvalue gives every edge of pixel p the value w(p)
                 * (mgm_weights.h:92-110), read at the CURRENT pixel for each neighbour (mgm_core.cc:981-985) */
                const float wp = wgt ? wgt[p] : 1.0f;
                for (int o = lo[p]; o <= hi[p]; o++) {
                    int k0 = o - gmin;
                    float e = 0;
                    for (int n = 0; n < tsgm; n++) {          /* update_costW :88-121 / update_cost2 :46-70 */
                        const float *Lq = L + q[n] * D;
                        float mq = Lmin[q[n]];
                        float a = (k0 - 1 >= 0) ? Lq[k0 - 1] : ORC_INF;
                        float b = (k0 + 1 < D) ? Lq[k0 + 1] : ORC_INF;
                        float v1 = ORC_W_V1(n, MINF(a, b), P1, wp);
                        float v2 = ORC_W_V2(n, mq, P2, wp);
                        float t = fmin3_(Lq[k0], v1, v2) - mq;
                        if (tsgm == 2) t = t / 2;
                        e += t;
                    }
                    Lp[k0] = (tsgm == 2) ? Cp[k0] + e : Cp[k0] + e / tsgm;
                }
            }
            float m = ORC_INF;                                /* Dvec::get_minvalue, dvec.cc:81-88 */
            for (int o = lo[p]; o <= hi[p]; o++) if (Lp[o - gmin] < m) m = Lp[o - gmin];
            Lmin[p] = m;
            float am = 0;
            for (int o = lo[p]; o <= hi[p]; o++) if (Lp[o - gmin] == m) am = (float)o;
            arg[p] = am;
        }
}

/* mgm_naive_parallelism (mgm_core.cc:829-1074).  S (dense) receives sum_pass L - (ndir-1) C, passes
 * added in order 0..ndir-1 (the 1-thread order).  disp = first strict minimum over finite S,
 * cost = its value, conf = number of passes whose (last) argmin equals disp. */
void orc_aggregate_w(const float *C, const int *lo, const int *hi, int w, int h, int gmin, int D,
                     float P1, float P2, int ndir, int tsgm, int fix_overcount, const float *wgt,
                     float *S, float *disp, float *cost, float *conf);
void orc_aggregate(const float *C, const int *lo, const int *hi, int w, int h, int gmin, int D,
                   float P1, float P2, int ndir, int tsgm, int fix_overcount,
                   float *S, float *disp, float *cost, float *conf)
{
    orc_aggregate_w(C, lo, hi, w, h, gmin, D, P1, P2, ndir, tsgm, fix_overcount, NULL, S, disp, cost, conf);
}
/* wgt: per-pixel regularity weight of this view (-wl / -wr), or NULL = all ones */
void orc_aggregate_w(const float *C, const int *lo, const int *hi, int w, int h, int gmin, int D,
                     float P1, float P2, int ndir, int tsgm, int fix_overcount, const float *wgt,
                     float *S, float *disp, float *cost, float *conf)
{
    size_t npix = (size_t)w * h, nvox = npix * D;
    float **L = malloc(sizeof(float *) * ndir);
    float *args = malloc(sizeof(float) * npix * ndir);
    for (int p = 0; p < ndir; p++) L[p] = malloc(sizeof(float) * nvox);
#pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < ndir; p++) {
        float *Lmin = malloc(sizeof(float) * npix);
        orc_one_pass(C, lo, hi, w, h, gmin, D, P1, P2, tsgm, p, wgt, L[p], Lmin, args + (size_t)p * npix);
        free(Lmin);
    }
#pragma omp parallel for
    for (size_t i = 0; i < npix; i++) {
        float *Si = S + i * D;
        const float *Ci = C + i * D;
        for (int k = 0; k < D; k++) Si[k] = ORC_INF;
        float minP = 0, minL = ORC_INF;
        for (int o = lo[i]; o <= hi[i]; o++) {
            int k = o - gmin;
            float s = 0;
            for (int p = 0; p < ndir; p++) s += L[p][i * D + k];   /* dvec.cc:110-118, pass order */
            if (fix_overcount == 1) s = fmaf(-(float)(ndir - 1), Ci[k], s); /* :1041-1042; gcc contracts
                this to one fma at -O3 -march=native/x86-64-v3 (exact for integer costs, i.e. census 5x5) */
            Si[k] = s;
            if (isfinite(s) && minL > s) { minL = s; minP = (float)o; }
        }
        disp[i] = minP;
        cost[i] = minL;
        int c = 0;
        for (int p = 0; p < ndir; p++) if (args[(size_t)p * npix + i] == minP) c++;
        conf[i] = (float)c;
    }
    for (int p = 0; p < ndir; p++) free(L[p]);
    free(L); free(args);
}

/* ------------------------------------------------------------- refinement */

static void vfit(const float v[3], float *vmin, float *xmin)
{   /* refine.h:70-92 */
    if ((v[1] > v[0]) && (v[1] > v[2])) { *vmin = v[1]; *xmin = 0; return; }
    float slope = v[2] - v[1];
    if ((v[2] - v[1]) < (v[0] - v[1])) slope = v[0] - v[1];
    *xmin = (v[0] - v[2]) / (2 * slope);
    *vmin = fmaf(*xmin - 1, slope, v[2]);   /* `v[2] + (xmin - 1) * slope`: one fma in the reference build (pinned by the
                                               MINDIFF cases of scripts/fuzz_oracle.py: the cost image feeds mindiff's argmin) */
}
static void parabola(const float v[3], float *vmin, float *xmin)
{   /* refine.h:40-68 */
    if (v[1] > v[0] && v[1] > v[2]) { *xmin = 0; *vmin = v[1]; return; }
    float c = v[1];
    float b = (v[2] - v[0]) / 2;
    float a = (v[2] - 2 * v[1] + v[0]) / 2;
    float x = -b / (2 * a);
    if (x > 1) x = 1;
    if (x < -1) x = -1;
    *vmin = fmaf(fmaf(a, x, b), x, c);      /* `(a*x + b)*x + c`: two fmas in the reference build */
    *xmin = x;
}

/* subpixel_refinement_sgm, mgm_refine.h:45-90.  S entries outside the pixel's range read +INF. */
void orc_refine(const float *S, const int *lo, const int *hi, int npix, int gmin, int D,
                int refine, float *disp, float *cost)
{
    if (!refine) return;
#pragma omp parallel for
    for (int i = 0; i < npix; i++) {
        const float *Si = S + (size_t)i * D;
        float minP = disp[i], minL = cost[i];
        int o = (int)minP;
        if (o - 1 >= lo[i] && o + 2 <= hi[i]) {
            int k = o - gmin;
            float v[3] = {Si[k - 1], Si[k], Si[k + 1]};
            float dx = 0;
            (refine == 1 ? vfit : parabola)(v, &minL, &dx);
            minP = o + dx;
            float vr[3] = {Si[k + 1], Si[k], Si[k - 1]};
            float dxr = 0, minLr = disp[i];
            (refine == 1 ? vfit : parabola)(vr, &minLr, &dxr);
            if (minLr < minL) { minP = o - dxr; minL = minLr; }
        }
        disp[i] = minP;
        cost[i] = minL;
    }
}

/* ----------------------------------------------------------- post filters */

static int cmpf(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}
/* median_filter, img_tools.h:204-238: window clipped to the image, NaN skipped,
 * element size/2 of the sorted values (upper median); untouched if no value. */
void orc_median(const float *in, float *out, int w, int h, int radius)
{
    int side = 2 * radius + 1;
#pragma omp parallel
    {
        float *v = malloc(sizeof(float) * side * side);
#pragma omp for
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int n = 0;
                for (int j = -radius; j <= radius; j++) {
                    if (j + y < 0 || j + y >= h) continue;
                    for (int i = -radius; i <= radius; i++) {
                        if (i + x < 0 || i + x >= w) continue;
                        float t = in[(j + y) * w + i + x];
                        if (!isnan(t)) v[n++] = t;
                    }
                }
                if (n) { qsort(v, n, sizeof(float), cmpf); out[y * w + x] = v[n / 2]; }
                else out[y * w + x] = in[y * w + x];
            }
        free(v);
    }
}

/* leftright_test, stereo_utils.cc:9-32.  dx is modified in place using the OTHER view's
 * disparity other[] (size ow x h). */
void orc_lrcheck(float *dx, int w, int h, const float *other, int ow, float tau)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int i = x + y * w;
            float t = x + dx[i];
            if (!isfinite(t)) { dx[i] = NAN; continue; }  /* round(NaN)->int is out of range on x86 */
            int Lx = (int)roundf(t);
            if (Lx < ow && Lx >= 0) {
                float Rx = Lx + other[Lx + y * ow];
                if (fabs(Rx - x) > tau) dx[i] = NAN;
            } else dx[i] = NAN;
        }
}

/* mindiff, stereo_utils.cc:93-124 (corr = channel 0 of the other view's cost image). */
void orc_mindiff(float *disp, float *corr, int w, int h, int win, float tau)
{
    int wl = win / 2, wr = win / 2;
    if (win % 2 == 0) wr--;
    float *od = malloc(sizeof(float) * w * h), *oc = malloc(sizeof(float) * w * h);
    memcpy(od, disp, sizeof(float) * w * h);
    memcpy(oc, corr, sizeof(float) * w * h);
    for (int y = wl; y < h - wr; y++)
        for (int x = wl; x < w - wr; x++) {
            float mincorr = INFINITY, mindisp = 0;
            for (int i = -wl; i <= wr; i++)
                for (int j = -wl; j <= wr; j++) {
                    int q = (x + i) + (y + j) * w;
                    if (mincorr > corr[q] && isfinite(disp[q])) { mincorr = corr[q]; mindisp = disp[q]; }
                }
            if (fabsf(od[x + y * w] - mindisp) > tau) { oc[x + y * w] = INFINITY; od[x + y * w] = NAN; }
        }
    memcpy(disp, od, sizeof(float) * w * h);
    memcpy(corr, oc, sizeof(float) * w * h);
    free(od); free(oc);
}

static int uf_find(int *t, int a) { while (t[a] != a) { t[a] = t[t[a]]; a = t[a]; } return a; }
void orc_remove_small_cc(int w, int h, const float *in, float *out, int minarea, float thr);

/* ------------------------------------------------------- one-scale driver */

/* One view of mgm_call (mgm_multiscale.cc:161-305): volume, aggregation, refinement, /ZOOM. */
/* compute_PKR_confidence (mgm_costvolume.cc:199-214): ratio of the second minimum of S -- over the labels of the pixel's range
 * more than 2 away from the winner -- to the first, the winner being the integer WTA label (it runs before the sub-pixel
 * refinement, mgm_multiscale.cc:247).  Each mgm_call overwrites the image of the call before (img_dict, :247,:297): the caller gets
 * the last call's.  g_pkr[view] = destination or NULL. */
static float *g_pkr[2] = {NULL, NULL};
static int g_view = 0;
static void orc_pkr(const float *S, const float *disp, const int *lo, const int *hi, int npix, int gmin, int D, float *out)
{
    for (int i = 0; i < npix; i++) {
        float currdisp = roundf(disp[i]);
        int k0 = (int)currdisp - gmin;
        float firstmin = (k0 >= lo[i] - gmin && k0 <= hi[i] - gmin) ? S[(size_t)i * D + k0] : ORC_INF;   /* Dvec[]: +inf outside the range */
        float secondmin = ORC_INF;
        for (int o = lo[i]; o <= hi[i]; o++)
            if (fabsf((float)o - currdisp) > 2 && secondmin > S[(size_t)i * D + (o - gmin)]) secondmin = S[(size_t)i * D + (o - gmin)];
        out[i] = secondmin / fmax(firstmin, 0.01);
    }
}

static void orc_view(const float *u, const float *v, int w, int h, const float *dminI, const float *dmaxI,
                     const orc_params *P, int zoom, const float *wgt, float *disp, float *cost, float *conf)
{
    int npix = w * h;
    int *lo = malloc(sizeof(int) * npix), *hi = malloc(sizeof(int) * npix);
    orc_ranges(dminI, dmaxI, npix, zoom, lo, hi);
    {   /* analysis hook (scripts/range_width_analysis.py): dump the per-pixel label ranges of every mgm_call */
        const char *dp = getenv("ORC_DUMP_RANGES");
        if (dp) {
            static int cnt = 0;
            char fn[512];
            snprintf(fn, sizeof fn, "%s/ranges_%03d_z%d_%dx%d.bin", dp, cnt++, zoom, w, h);
            FILE *f = fopen(fn, "wb");
            if (f) { fwrite(lo, sizeof(int), npix, f); fwrite(hi, sizeof(int), npix, f); fclose(f); }
        }
    }
    int gmin = lo[0], gmax = hi[0];
    for (int i = 1; i < npix; i++) { if (lo[i] < gmin) gmin = lo[i]; if (hi[i] > gmax) gmax = hi[i]; }
    int D = gmax - gmin + 1;
    float *C = malloc(sizeof(float) * (size_t)npix * D), *S = malloc(sizeof(float) * (size_t)npix * D);
    orc_costvolume(u, v, w, h, lo, hi, gmin, D, P->census_win, zoom, P->dct_shift, P->cost, C);
    float P1 = P->P1 / zoom;                                  /* mgm_multiscale.cc:194-202 */
    orc_aggregate_w(C, lo, hi, w, h, gmin, D, P1, P->P2, P->ndir, P->tsgm, P->fix_overcount, wgt, S, disp, cost, conf);
    if (g_pkr[g_view]) orc_pkr(S, disp, lo, hi, npix, gmin, D, g_pkr[g_view]);
    orc_refine(S, lo, hi, npix, gmin, D, P->refine, disp, cost);
    for (int i = 0; i < npix; i++) disp[i] /= (float)zoom;     /* :253 */
    free(C); free(S); free(lo); free(hi);
}

/* mgm_call (mgm_multiscale.cc:161-335).  costR: channel 0 of cr (needed by mindiff). */
void orc_mgm_call_w(const float *u, const float *v, int w, int h,
                    const float *dminL, const float *dmaxL, const float *dminR, const float *dmaxR,
                    const orc_params *P, int zoom, const float *wl, const float *wr, float *dl, float *dr, float *confL);
void orc_mgm_call(const float *u, const float *v, int w, int h,
                  const float *dminL, const float *dmaxL, const float *dminR, const float *dmaxR,
                  const orc_params *P, int zoom, float *dl, float *dr, float *confL)
{
    orc_mgm_call_w(u, v, w, h, dminL, dmaxL, dminR, dmaxR, P, zoom, NULL, NULL, dl, dr, confL);
}
void orc_mgm_call_w(const float *u, const float *v, int w, int h,
                    const float *dminL, const float *dmaxL, const float *dminR, const float *dmaxR,
                    const orc_params *P, int zoom, const float *wl, const float *wr, float *dl, float *dr, float *confL)
{
    int npix = w * h;
    float *cl = malloc(sizeof(float) * npix), *cr = malloc(sizeof(float) * npix);
    float *confR = malloc(sizeof(float) * npix);
    g_view = 0; orc_view(u, v, w, h, dminL, dmaxL, P, zoom, wl, dl, cl, confL);
    g_view = 1; orc_view(v, u, w, h, dminR, dmaxR, P, zoom, wr, dr, cr, confR);
    g_view = 0;
    if (P->median) {                                          /* :312-315 */
        float *t = malloc(sizeof(float) * npix);
        orc_median(dl, t, w, h, P->median); memcpy(dl, t, sizeof(float) * npix);
        orc_median(dr, t, w, h, P->median); memcpy(dr, t, sizeof(float) * npix);
        free(t);
    }
    if (P->mindiff >= 0) orc_mindiff(dl, cr, w, h, P->census_win, P->mindiff);   /* :318-319 */
    if (P->lr_mode == 1) {                                    /* :322-327 */
        float *tl = malloc(sizeof(float) * npix), *tr = malloc(sizeof(float) * npix);
        memcpy(tl, dl, sizeof(float) * npix); memcpy(tr, dr, sizeof(float) * npix);
        orc_lrcheck(dr, w, h, tl, w, P->lr_tau);
        orc_lrcheck(dl, w, h, tr, w, P->lr_tau);
        free(tl); free(tr);
    }
    if (P->remove_small_cc > 0) {                             /* :330-334 */
        float *t = malloc(sizeof(float) * npix);
        memcpy(t, dl, sizeof(float) * npix); orc_remove_small_cc(w, h, t, dl, P->remove_small_cc, 5);
        memcpy(t, dr, sizeof(float) * npix); orc_remove_small_cc(w, h, t, dr, P->remove_small_cc, 5);
        free(t);
    }
    free(cl); free(cr); free(confR);
}

/* remove_small_cc.c:9-73.  Links are only made from pixels with i < w-1 and j < h-1
 * (so the last row has no horizontal links and the last column no vertical ones),
 * NaN pixels never join, and components with area <= minarea become NaN. */
void orc_remove_small_cc(int w, int h, const float *in, float *out, int minarea, float thr)
{
    int n = w * h;
    int *rep = malloc(sizeof(int) * n), *area = calloc(n, sizeof(int));
    for (int i = 0; i < n; i++) rep[i] = isnan(in[i]) ? -1 : i;
    for (int j = 0; j < h - 1; j++)
        for (int i = 0; i < w - 1; i++) {
            int p0 = j * w + i, nb[2] = {p0 + 1, p0 + w};
            for (int k = 0; k < 2; k++) {
                int p1 = nb[k];
                if (rep[p0] >= 0 && rep[p1] >= 0 && fabs(in[p0] - in[p1]) < thr) {
                    int a = uf_find(rep, p0), b = uf_find(rep, p1);
                    if (a < b) rep[b] = a; else if (b < a) rep[a] = b;
                }
            }
        }
    for (int i = 0; i < n; i++) if (rep[i] >= 0) rep[i] = uf_find(rep, i);
    for (int i = 0; i < n; i++) if (rep[i] >= 0) area[rep[i]]++;
    for (int i = 0; i < n; i++) out[i] = (rep[i] >= 0 && area[rep[i]] <= minarea) ? NAN : in[i];
    free(rep); free(area);
}

/* ------------------------------------------------------------- `mgm` main */

/* main() of mgm (main_mgm.cc:80-266) from "images in memory" to "disparity in memory".
 * im1/im2 may hold NaN (no data).  Outputs: disp (left, NaN = invalid), conf
 * (-confidence_consensusL), dispR (right view, -Rd; may be NULL). */
int orc_mgm_w(const float *im1, const float *im2, int w, int h, int dmin, int dmax,
              const orc_params *P, const float *wl, const float *wr, float *disp, float *conf, float *dispR);
int orc_mgm(const float *im1, const float *im2, int w, int h, int dmin, int dmax,
            const orc_params *P, float *disp, float *conf, float *dispR)
{
    return orc_mgm_w(im1, im2, w, h, dmin, dmax, P, NULL, NULL, disp, conf, dispR);
}
/* wl, wr: the -wl / -wr regularity weight images (both or neither, main_mgm.cc:219-222) */
int orc_mgm_w(const float *im1, const float *im2, int w, int h, int dmin, int dmax,
              const orc_params *P, const float *wl, const float *wr, float *disp, float *conf, float *dispR)
{
    if (!wl || !wr) wl = wr = NULL;
    int npix = w * h;
    float *u = malloc(sizeof(float) * npix), *v = malloc(sizeof(float) * npix);
    float *a = malloc(sizeof(float) * npix), *b = malloc(sizeof(float) * npix);
    float *c = malloc(sizeof(float) * npix), *d = malloc(sizeof(float) * npix);
    float *dr = malloc(sizeof(float) * npix);
    for (int i = 0; i < npix; i++) {
        u[i] = isfinite(im1[i]) ? im1[i] : 0;                 /* :172-173 */
        v[i] = isfinite(im2[i]) ? im2[i] : 0;
        a[i] = dmin; b[i] = dmax;                             /* :178 */
        c[i] = -dmax; d[i] = -dmin;                           /* :207 */
        if (isnan(im1[i])) { a[i] = dmin; b[i] = dmin + 1; }  /* :211-213 */
        if (isnan(im2[i])) { c[i] = dmin; d[i] = dmin + 1; }  /* :214-216 (sic: dmin, not -dmax) */
    }
    orc_params Q = *P;                                         /* P1,P2 *= nch with nch = 1 (:194-195) */
    orc_mgm_call_w(u, v, w, h, a, b, c, d, &Q, 1, wl, wr, disp, dr, conf);
    for (int i = 0; i < npix; i++) {                           /* :231-236 */
        if (isnan(im1[i])) disp[i] = NAN;
        if (isnan(im2[i])) dr[i] = NAN;
    }
    if (dispR) memcpy(dispR, dr, sizeof(float) * npix);
    free(u); free(v); free(a); free(b); free(c); free(d); free(dr);
    return 0;
}

/* -confidence_pkrL / -confidence_pkrR of either binary: run `multi ? orc_mgm_multi_w : orc_mgm_w` with the PKR images captured
 * (pkrL / pkrR: w*h floats).  Every mgm_call overwrites the images of the call before; in mgm_multi the ones written are those of
 * the full-resolution ZOOM = 1 call, because the SUBPIX pass runs on a `param` of its own (main_mgm_multi.cc:207,240-251). */
int orc_mgm_multi_w(const float *im1, const float *im2, int w, int h, int dmin, int dmax,
                    const orc_params *P, const float *wl, const float *wr, float *disp, float *conf, float *dispR);
int orc_mgm_pkr(const float *im1, const float *im2, int w, int h, int dmin, int dmax, const orc_params *P, int multi,
                float *disp, float *conf, float *dispR, float *pkrL, float *pkrR)
{
    g_pkr[0] = pkrL; g_pkr[1] = pkrR;
    int r = multi ? orc_mgm_multi_w(im1, im2, w, h, dmin, dmax, P, NULL, NULL, disp, conf, dispR)
                  : orc_mgm_w(im1, im2, w, h, dmin, dmax, P, NULL, NULL, disp, conf, dispR);
    g_pkr[0] = g_pkr[1] = NULL;
    return r;
}

/* ------------------------------------------------------------ `mgm_multi` */

/* downsample2x, mgm_multiscale.cc:57-95: 10x10 Gaussian taps centred between pixels, weights
 * renormalised over the in-image taps.  gcc contracts `acc += u*gg` into one fma at
 * -O3 -march=native (checked in the disassembly of the reference object), so do we. */
void orc_downsample2x(const float *u, int nx, int ny, float sigma, float *out)
{
    int onx = (nx + 1) / 2, ony = (ny + 1) / 2;
    float g[100];
    for (int j = 0; j < 10; j++)
        for (int i = 0; i < 10; i++) {
            float a = i - 4 - .5, b = j - 4 - .5;
            double sq = a * a + b * b;
            g[i + j * 10] = exp(-sq / (2.0 * sigma * sigma));
        }
#pragma omp parallel for
    for (int j = 0; j < ony; j++)
        for (int i = 0; i < onx; i++) {
            float acc = 0, norm = 0;
            for (int y = 0; y < 10; y++)
                for (int x = 0; x < 10; x++) {
                    int xx = i * 2 + x - 4, yy = j * 2 + y - 4;
                    if (xx >= 0 && yy >= 0 && xx < nx && yy < ny) {
                        float gg = g[x + y * 10];
                        acc = fmaf(u[xx + yy * nx], gg, acc);
                        norm += gg;
                    }
                }
            out[i + j * onx] = acc / norm;
        }
}

/* downsample2x_disp, mgm_multiscale.cc:98-117: 2x2 min (or max) pooling, halved */
void orc_downsample2x_disp(const float *u, int nx, int ny, int is_max, float *out)
{
    int onx = (nx + 1) / 2;
    for (int j = 0; j < ny; j += 2)
        for (int i = 0; i < nx; i += 2) {
            float vmin = INFINITY, vmax = -INFINITY;
            for (int k = 0; k < 2; k++)
                for (int l = 0; l < 2; l++) {
                    int x = i + k, y = j + l;
                    float t = (x < nx && y < ny) ? u[x + y * nx] : NAN;
                    vmin = fminf(vmin, t);
                    vmax = fmaxf(vmax, t);
                }
            out[i / 2 + j / 2 * onx] = is_max ? vmax / 2 : vmin / 2;
        }
}

static inline float clampget(const float *a, int nx, int ny, int x, int y)
{   /* valneumann, img_tools.h:77-85 */
    if (x < 0) x = 0; if (x >= nx) x = nx - 1;
    if (y < 0) y = 0; if (y >= ny) y = ny - 1;
    return a[x + y * nx];
}
/* update_dmin_dmax, stereo_utils.cc:134-175.  (dminI, dmaxI) have disp's size; the fall-back images
 * (dminP, dmaxP) may have another size: they are indexed with disp's coordinates, clamped to THEIR size
 * (this is what happens when upsample2x_disp passes the fine-level images, mgm_multiscale.cc:43-44). */
void orc_update_dmin_dmax(const float *disp, int nx, int ny, float *dminI, float *dmaxI,
                          const float *dminP, const float *dmaxP, int pnx, int pny, int slack, int radius)
{
    float *tmin = malloc(sizeof(float) * nx * ny), *tmax = malloc(sizeof(float) * nx * ny);
    memcpy(tmin, dminI, sizeof(float) * nx * ny);
    memcpy(tmax, dmaxI, sizeof(float) * nx * ny);
    if (slack < 0) slack = -slack;
#pragma omp parallel for
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            float dmin = INFINITY, dmax = -INFINITY;
            for (int dj = -radius; dj <= radius; dj++)
                for (int di = -radius; di <= radius; di++) {
                    float v = clampget(disp, nx, ny, i + di, j + dj);
                    float vminP = clampget(dminP, pnx, pny, i + di, j + dj);
                    float vmaxP = clampget(dmaxP, pnx, pny, i + di, j + dj);
                    if (isfinite(v)) { dmin = fminf(dmin, v - slack); dmax = fmaxf(dmax, v + slack); }
                    else { dmin = fminf(dmin, vminP); dmax = fmaxf(dmax, vmaxP); }
                }
            if (isfinite(dmin)) { tmin[i + j * nx] = dmin; tmax[i + j * nx] = dmax; }
        }
    memcpy(dminI, tmin, sizeof(float) * nx * ny);
    memcpy(dmaxI, tmax, sizeof(float) * nx * ny);
    free(tmin); free(tmax);
}

/* upsample2x_disp, mgm_multiscale.cc:36-48 (slack 8, radius 4, :33-34) + zoom_nn :16-31 */
void orc_upsample2x_disp(const float *sdisp, int snx, int sny, float *dmin, float *dmax, int nx, int ny)
{
    int n = snx * sny;
    float *d2 = malloc(sizeof(float) * n), *omin = calloc(n, sizeof(float)), *omax = calloc(n, sizeof(float));
    for (int i = 0; i < n; i++) d2[i] = sdisp[i] * 2.0;
    orc_update_dmin_dmax(d2, snx, sny, omin, omax, dmin, dmax, nx, ny, 8, 4);
    for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
            dmin[y * nx + x] = omin[(y / 2) * snx + x / 2];
            dmax[y * nx + x] = omax[(y / 2) * snx + x / 2];
        }
    free(d2); free(omin); free(omax);
}

/* recursive_multiscale, mgm_multiscale.cc:339-410 */
static void orc_recursive(const float *u, const float *v, int nx, int ny, float *dmin, float *dmax, float *dminR, float *dmaxR,
                          const orc_params *P, int zoom, int numscales, int scale, const float *wl, const float *wr,
                          float *dl, float *dr, float *confL)
{
    if (fmax(nx, ny) > 100 && fmin(nx, ny) > 50 && scale < numscales) {
        int sx = (nx + 1) / 2, sy = (ny + 1) / 2, sn = sx * sy;
        float *su = malloc(sizeof(float) * sn), *sv = malloc(sizeof(float) * sn);
        float *a = malloc(sizeof(float) * sn), *b = malloc(sizeof(float) * sn), *c = malloc(sizeof(float) * sn), *d = malloc(sizeof(float) * sn);
        float *sdl = malloc(sizeof(float) * sn), *sdr = malloc(sizeof(float) * sn), *sconf = malloc(sizeof(float) * sn);
        orc_downsample2x(u, nx, ny, 0.8f, su);
        orc_downsample2x(v, nx, ny, 0.8f, sv);
        orc_downsample2x_disp(dmin, nx, ny, 0, a);
        orc_downsample2x_disp(dmax, nx, ny, 1, b);
        orc_downsample2x_disp(dminR, nx, ny, 0, c);
        orc_downsample2x_disp(dmaxR, nx, ny, 1, d);
        float *swl = NULL, *swr = NULL;                        /* the weight maps follow the pyramid, :375-378 */
        if (wl && wr) {
            swl = malloc(sizeof(float) * sn); swr = malloc(sizeof(float) * sn);
            orc_downsample2x(wl, nx, ny, 0.8f, swl);
            orc_downsample2x(wr, nx, ny, 0.8f, swr);
        }
        orc_recursive(su, sv, sx, sy, a, b, c, d, P, zoom, numscales, scale + 1, swl, swr, sdl, sdr, sconf);
        orc_upsample2x_disp(sdl, sx, sy, dmin, dmax, nx, ny);
        orc_upsample2x_disp(sdr, sx, sy, dminR, dmaxR, nx, ny);
        free(su); free(sv); free(a); free(b); free(c); free(d); free(sdl); free(sdr); free(sconf); free(swl); free(swr);
    }
    orc_mgm_call_w(u, v, nx, ny, dmin, dmax, dminR, dmaxR, P, zoom, wl, wr, dl, dr, confL);
}

/* main() of mgm_multi (main_mgm_multi.cc:88-256), memory to memory.  The confidence written by
 * -confidence_consensusL is the one of the full-resolution ZOOM=1 call (the outer `param`, :197-209). */
int orc_mgm_multi_w(const float *im1, const float *im2, int w, int h, int dmin, int dmax,
                    const orc_params *P, const float *wl, const float *wr, float *disp, float *conf, float *dispR);
int orc_mgm_multi(const float *im1, const float *im2, int w, int h, int dmin, int dmax,
                  const orc_params *P, float *disp, float *conf, float *dispR)
{
    return orc_mgm_multi_w(im1, im2, w, h, dmin, dmax, P, NULL, NULL, disp, conf, dispR);
}
int orc_mgm_multi_w(const float *im1, const float *im2, int w, int h, int dmin, int dmax,
                    const orc_params *P, const float *wl, const float *wr, float *disp, float *conf, float *dispR)
{
    if (!wl || !wr) wl = wr = NULL;
    int npix = w * h;
    float *u = malloc(sizeof(float) * npix), *v = malloc(sizeof(float) * npix);
    float *a = malloc(sizeof(float) * npix), *b = malloc(sizeof(float) * npix);
    float *c = malloc(sizeof(float) * npix), *d = malloc(sizeof(float) * npix);
    float *dr = malloc(sizeof(float) * npix), *conf2 = malloc(sizeof(float) * npix);
    for (int i = 0; i < npix; i++) {
        u[i] = isfinite(im1[i]) ? im1[i] : 0;
        v[i] = isfinite(im2[i]) ? im2[i] : 0;
        a[i] = dmin; b[i] = dmax; c[i] = -dmax; d[i] = -dmin;
        if (isnan(im1[i])) { a[i] = dmin; b[i] = dmin + 1; }
        if (isnan(im2[i])) { c[i] = dmin; d[i] = dmin + 1; }
    }
    orc_recursive(u, v, w, h, a, b, c, d, P, 1, P->scales, 0, wl, wr, disp, dr, conf);
    g_pkr[0] = g_pkr[1] = NULL;      /* the SUBPIX pass runs on its own `param` (main_mgm_multi.cc:207): the images written come from the call above */
    if (P->subpix > 1) {                                       /* :203-209 */
        orc_update_dmin_dmax(disp, w, h, a, b, a, b, w, h, 2, 4);
        orc_update_dmin_dmax(dr, w, h, c, d, c, d, w, h, 2, 4);
        /* the half-pixel pass builds a fresh `param` without the weight images (:207): unit weights */
        orc_recursive(u, v, w, h, a, b, c, d, P, P->subpix, 0, 0, NULL, NULL, disp, dr, conf2);
    }
    if (P->lr_mode == 2) {                                     /* :212-217 */
        float *tl = malloc(sizeof(float) * npix), *tr = malloc(sizeof(float) * npix);
        memcpy(tl, disp, sizeof(float) * npix); memcpy(tr, dr, sizeof(float) * npix);
        orc_lrcheck(dr, w, h, tl, w, P->lr_tau);
        orc_lrcheck(disp, w, h, tr, w, P->lr_tau);
        free(tl); free(tr);
    }
    for (int i = 0; i < npix; i++) {
        if (isnan(im1[i])) disp[i] = NAN;
        if (isnan(im2[i])) dr[i] = NAN;
    }
    if (dispR) memcpy(dispR, dr, sizeof(float) * npix);
    free(u); free(v); free(a); free(b); free(c); free(d); free(dr); free(conf2);
    return 0;
}

/* ----------------------------------------------------- rejection mask (a10) */

static float cubic1(const float v[4], float x)
{   /* c/bicubic.c:8-13 (double arithmetic through the 0.5 / 2.0 literals) */
    return v[1] + 0.5 * x * (v[2] - v[0] + x * (2.0 * v[0] - 5.0 * v[1] + 4.0 * v[2] - v[3]
                 + x * (3.0 * (v[1] - v[2]) + v[3] - v[0])));
}
/* create_rejection_mask (s2p/block_matching.py:18-32): backflow (c/backflow.c:125-168) samples
 * im2 at (x + d, y) with bicubic_interpolation_boundary(.., 0) (c/bicubic.c:69-96, zero outside),
 * then mask = isfinite(disp) && isfinite(im1) && isfinite(warped). */
void orc_rejection_mask(const float *disp, const float *im1, const float *im2, int w, int h, uint8_t *mask)
{
#pragma omp parallel for
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            float px = i + disp[j * w + i], py = j + 0.0f;
            float x = px - 1, y = py - 1;
            int ix = (int)floorf(x), iy = (int)floorf(y);
            float r;
            if (!isfinite(x)) r = NAN;
            else {
                float cc[4][4];
                for (int jj = 0; jj < 4; jj++)
                    for (int ii = 0; ii < 4; ii++) {
                        int sx = ix + ii, sy = iy + jj;
                        cc[ii][jj] = (sx < 0 || sx >= w || sy < 0 || sy >= h) ? 0 : im2[sy * w + sx];
                    }
                float vv[4];
                for (int k = 0; k < 4; k++) vv[k] = cubic1(cc[k], y - iy);
                r = cubic1(vv, x - ix);
            }
            mask[j * w + i] = isfinite(disp[j * w + i]) && isfinite(im1[j * w + i]) && isfinite(r);
        }
}

/* exhaustive helper for tests: x/3 and x/4 computed the way the CUDA kernels do (reciprocal
 * multiply + one fma correction) must equal IEEE division for every float the volume can hold. */
float orc_div_by(float x, int n) { return x / n; }
