/* x/3 == fma(fma(-3, q0, x), r3, q0) with q0 = x * r3, r3 = RN(1/3), for non-negative finite floats.
 * usage: check_div3 [stride]   (stride 1 = exhaustive, ~20 s) */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static int bad(uint32_t u)
{
    float x; memcpy(&x, &u, 4);
    const float r3 = 1.0f / 3.0f;
    float q0 = x * r3, r = fmaf(-3.0f, q0, x), q = fmaf(r, r3, q0);
    return q != x / 3.0f;
}
int main(int argc, char **argv)
{
    uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
    long nbad = 0;
    for (uint64_t u = 0; u < 0x7f800000u; u += stride) nbad += bad((uint32_t)u);
    for (uint32_t e = 0; e < 255; e++)                 /* both sides of every exponent boundary */
        for (int d = -64; d <= 64; d++) {
            int64_t u = ((int64_t)e << 23) + d;
            if (u >= 0 && u < 0x7f800000) nbad += bad((uint32_t)u);
        }
    printf("mismatches: %ld\n", nbad);
    return nbad != 0;
}
