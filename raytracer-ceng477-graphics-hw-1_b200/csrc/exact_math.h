// exact_math.h — the two places where the kernels replace a libm call of the reference by a
// closed form, shared between device code and host-side test hooks so that CPU tests can pin them
// against glibc (tests/test_host.py).
#pragma once

#include <cmath>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

namespace rtb {

// The specular gate `acos(cos)*180/3.1415 <= 90.01` (raytracer.cpp:411-412: double acos, float
// theta, double comparison) is monotone in cos; with glibc's acos it is equivalent to
// kGateCos <= cos <= 1 (cos > 1 makes acos NaN, the comparison false).
constexpr unsigned kGateCosBits = 0xB90665D3u;  // -0.000128171683f

RT_HD bool specular_gate(float cos_theta) {
#if defined(__CUDA_ARCH__)
    const float thr = __uint_as_float(kGateCosBits);
#else
    union { unsigned u; float f; } c = {kGateCosBits};
    const float thr = c.f;
#endif
    return cos_theta >= thr && cos_theta <= 1.0f;
}

// pow((double)base, (double)e) narrowed to float (raytracer.cpp:414).  Integer exponents (every
// shipped scene: 1, 3, 50, 100) take a square-and-multiply chain in double: a handful of half-ulp
// double roundings, invisible after the narrowing to float except on a ~1e-7 sliver of inputs;
// anything else goes through double-precision pow.
RT_HD float pow_ref(float base, float e) {
    double b = (double) base;
    if (e >= 0.0f && e <= 1024.0f && e == truncf(e)) {
        unsigned n = (unsigned) e;
        double r = 1.0;
        while (n) {
            if (n & 1u) r = r * b;
            b = b * b;
            n >>= 1;
        }
        return (float) r;
    }
    return (float) pow(b, (double) e);
}

}  // namespace rtb
