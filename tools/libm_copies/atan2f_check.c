#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline int32_t asi(float f){int32_t u;memcpy(&u,&f,4);return u;}
static inline float asf(int32_t u){float f;memcpy(&f,&u,4);return f;}
static float atanG(float x) {
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
                aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
                aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    const float one = 1.0f;
    const int32_t hx = asi(x), ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c000000) { if (ix > 0x7f800000) return x + x; return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3]; }
    if (ix < 0x3ee00000) { if (ix < 0x31000000) return x; id = -1; }
    else {
        x = fabsf(x);
        if (ix < 0x3f980000) { if (ix < 0x3f300000) { id = 0; x = (2.0f * x - one) / (2.0f + x); } else { id = 1; x = (x - one) / (x + one); } }
        else if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (one + 1.5f * x); }
        else { id = 3; x = -1.0f / x; }
    }
    const float z = x * x, w = z * z;
    const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return hx < 0 ? -r : r;
}
static int TH = 60, TL = 60;
static float atan2G(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int32_t hx = asi(x), ix = hx & 0x7fffffff, hy = asi(y), iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return atanG(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) return m < 2 ? y : (m == 2 ? pi + tiny : -pi - tiny);
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int k = (iy - ix) >> 23;
    float z;
    if (k > TH) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -TL) z = 0.0f;
    else z = atanG(fabsf(y / x));
    if (m == 0) return z;
    if (m == 1) return asf(asi(z) ^ (int)0x80000000);
    if (m == 2) return pi - (z - pi_lo);
    return (z - pi_lo) - pi;
}

int main() {
    TH = 24; TL = 26;
    srand48(7); long bad = 0, n = 200000000;
    for (long i = 0; i < n; i++) {
        float y = (float)(drand48() * 2 - 1), x = (float)(drand48() * 2 - 1);
        int e1 = (int)(lrand48() % 80) - 40, e2 = (int)(lrand48() % 80) - 40;
        if (i & 1) { y = ldexpf(y, e1); x = ldexpf(x, e2); }
        volatile float yv = y, xv = x;
        float a = atan2f(yv, xv), b = atan2G(y, x);
        bad += asi(a) != asi(b);
    }
    printf("mismatches %ld of %ld\n", bad, n);
}
