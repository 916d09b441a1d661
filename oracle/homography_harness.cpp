// TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference resampler from memory.
//
// The reference's `homography` binary (3rdparty/homography/main.cpp:65-177) needs GDAL for file I/O, which
// is not installed here.  This harness restates main()'s geometry (needed ROI of the source, crop, homography
// compensated by the crop: main.cpp:29-55,94-134) and then calls the reference's own runHomography()
// (LibHomography/Homography.cpp:28-48) compiled from the sources where they lie; rasters travel as PFM.
//
//   homography_ref src.pfm "h1 h2 ... h9" out.pfm width height
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "LibImages/LibImages.h"
#include "Utilities/Parameters.h"
#include "LibHomography/Homography.h"

static bool read_pfm(const char *path, std::vector<float> &a, int &w, int &h)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    char magic[8]; float scale;
    if (fscanf(f, "%7s %d %d %f", magic, &w, &h, &scale) != 4) { fclose(f); return false; }
    fgetc(f);
    a.resize((size_t)w * h);
    bool ok = fread(a.data(), sizeof(float), a.size(), f) == a.size();
    fclose(f);
    return ok;
}
static void write_pfm(const char *path, const float *a, int w, int h)
{
    FILE *f = fopen(path, "wb");
    fprintf(f, "Pf\n%d %d\n-1.0\n", w, h);
    fwrite(a, sizeof(float), (size_t)w * h, f);
    fclose(f);
}
static void inv33(double o[9], const double i[9])
{   // linalg.c:21-34
    double det = i[0]*i[4]*i[8] + i[2]*i[3]*i[7] + i[1]*i[5]*i[6] - i[2]*i[4]*i[6] - i[1]*i[3]*i[8] - i[0]*i[5]*i[7];
    o[0] = (i[4]*i[8] - i[5]*i[7]) / det; o[1] = (i[2]*i[7] - i[1]*i[8]) / det; o[2] = (i[1]*i[5] - i[2]*i[4]) / det;
    o[3] = (i[5]*i[6] - i[3]*i[8]) / det; o[4] = (i[0]*i[8] - i[2]*i[6]) / det; o[5] = (i[2]*i[3] - i[0]*i[5]) / det;
    o[6] = (i[3]*i[7] - i[4]*i[6]) / det; o[7] = (i[1]*i[6] - i[0]*i[7]) / det; o[8] = (i[0]*i[4] - i[1]*i[3]) / det;
}

int main(int c, char **v)
{
    if (c != 6) { fprintf(stderr, "usage: %s src.pfm \"h1..h9\" out.pfm w h\n", v[0]); return 1; }
    std::vector<float> src; int sw, sh;
    if (!read_pfm(v[1], src, sw, sh)) { fprintf(stderr, "cannot read %s\n", v[1]); return 1; }
    double H[9];
    { const char *s = v[2]; char *e; for (int k = 0; k < 9; k++) { H[k] = strtod(s, &e); s = e; } }
    int ow = atoi(v[4]), oh = atoi(v[5]);
    // needed ROI: pre-image of the output corners (main.cpp:39-55), integer bounding box (:29-36)
    double Hi[9]; inv33(Hi, H);
    double cx[4] = {0, (double)ow, (double)ow, 0}, cy[4] = {0, 0, (double)oh, (double)oh}, px[4], py[4];
    for (int k = 0; k < 4; k++) {
        double z = Hi[6]*cx[k] + Hi[7]*cy[k] + Hi[8];
        px[k] = (Hi[0]*cx[k] + Hi[1]*cy[k] + Hi[2]) / z; py[k] = (Hi[3]*cx[k] + Hi[4]*cy[k] + Hi[5]) / z;
    }
    double mnx = px[0], mxx = px[0], mny = py[0], mxy = py[0];
    for (int k = 1; k < 4; k++) { if (px[k] < mnx) mnx = px[k]; if (px[k] > mxx) mxx = px[k]; if (py[k] < mny) mny = py[k]; if (py[k] > mxy) mxy = py[k]; }
    int x = (int)floor(mnx), y = (int)floor(mny), w = (int)ceil(mxx - x), h = (int)ceil(mxy - y);
    if (x < 0) { w += x; x = 0; }                                  // main.cpp:112-127
    if (y < 0) { h += y; y = 0; }
    if (x + w > sw) w = sw - x;
    if (y + h > sh) h = sh - y;
    if (w <= 0 || h <= 0) { fprintf(stderr, "ERROR: empty roi\n"); return 1; }
    double T[9] = {1, 0, (double)x, 0, 1, (double)y, 0, 0, 1}, Hc[9];   // :132-134
    for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) Hc[3*r+q] = H[3*r]*T[q] + H[3*r+1]*T[3+q] + H[3*r+2]*T[6+q];
    std::vector<float> roi((size_t)w * h);
    for (int j = 0; j < h; j++) memcpy(&roi[(size_t)j * w], &src[(size_t)(y + j) * sw + x], sizeof(float) * w);
    Image in(roi.data(), w, h, 1), out(ow, oh, 1);
    Parameters params(0, ow, oh, true);                             // :158
    runHomography(in, Hc, out, params);
    std::vector<float> o((size_t)ow * oh);
    for (int j = 0; j < oh; j++) memcpy(&o[(size_t)j * ow], out.getPtr(0, j), sizeof(float) * ow);
    write_pfm(v[3], o.data(), ow, oh);
    return 0;
}
