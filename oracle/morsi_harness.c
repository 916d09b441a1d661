/* TEST INFRASTRUCTURE ONLY -- exposes the reference's own morphology (c/morsi.c, compiled in place) to ctypes.
 * morsi.c always compiles its command-line front end, which needs two iio entry points; they are stubbed here
 * because the harness feeds morsi_erosion() / build_disk() (c/morsi.c:54-66,280-298) from memory. */
#define HIDE_ALL_MAINS
#include "morsi.c"

float *iio_read_image_float_split(const char *f, int *w, int *h, int *pd) { (void)f; (void)w; (void)h; (void)pd; return NULL; }
void iio_write_image_float_split(char *f, float *x, int w, int h, int pd) { (void)f; (void)x; (void)w; (void)h; (void)pd; }

/* `morsi diskR erosion` on a float raster */
int s2pb_ref_disk_erosion(float *y, float *x, int w, int h, float radius)
{
    int *e = build_disk(radius);
    if (!e) return 1;
    morsi_erosion(y, x, w, h, e);
    free(e);
    return 0;
}
