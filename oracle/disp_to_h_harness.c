/* TEST INFRASTRUCTURE ONLY -- the six iio entry points that the command-line front ends inside the reference's
 * c/disp_to_h.c reference.  oracle/_ref/libdisp_to_h_ref.so is c/disp_to_h.c + c/rpc.c compiled in place (the
 * reference builds the same two files into lib/disp_to_h.so, makefile:119-122); the tests call its
 * disp_to_lonlatalt() through ctypes exactly as s2p/triangulation.py:118-143 does, so no image I/O is ever used. */
#include <stddef.h>
float *iio_read_image_float(const char *f, int *w, int *h) { (void)f; (void)w; (void)h; return NULL; }
float *iio_read_image_float_split(const char *f, int *w, int *h, int *pd) { (void)f; (void)w; (void)h; (void)pd; return NULL; }
double *iio_read_image_double_vec(const char *f, int *w, int *h, int *pd) { (void)f; (void)w; (void)h; (void)pd; return NULL; }
void iio_write_image_double_vec(char *f, double *x, int w, int h, int pd) { (void)f; (void)x; (void)w; (void)h; (void)pd; }
void iio_write_image_float_vec(char *f, float *x, int w, int h, int pd) { (void)f; (void)x; (void)w; (void)h; (void)pd; }
void iio_write_image_int(char *f, int *x, int w, int h) { (void)f; (void)x; (void)w; (void)h; }
