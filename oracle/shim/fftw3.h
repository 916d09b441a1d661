/* TEST INFRASTRUCTURE ONLY -- minimal stand-in for <fftw3.h>.
 *
 * The reference matcher (3rdparty/mgm_multi/shear.c:38-99) calls exactly six
 * FFTW entry points to phase-shift image rows with a DCT.  libfftw3 is not
 * installed in this image and there is no network, so the oracle build links
 * this header + fftw_shim.c instead.  It implements the three r2r kinds the
 * reference uses, with FFTW's unnormalised definitions, in double precision.
 * Nothing here is reference code; nothing here ships in the product library.
 */
#ifndef S2PB_FFTW3_SHIM_H
#define S2PB_FFTW3_SHIM_H
#include <stddef.h>
#include <stdio.h> /* the real fftw3.h pulls it in; shear.c relies on that */
#ifdef __cplusplus
extern "C" {
#endif
typedef struct s2pb_fftw_plan_s *fftw_plan;
typedef enum { FFTW_REDFT10 = 5, FFTW_REDFT01 = 4, FFTW_RODFT01 = 8 } fftw_r2r_kind;
#define FFTW_ESTIMATE (1U << 6)
void *fftw_malloc(size_t n);
void fftw_free(void *p);
fftw_plan fftw_plan_r2r_1d(int n, double *in, double *out, fftw_r2r_kind kind, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
#ifdef __cplusplus
}
#endif
#endif
