// oracle/oracle_blur.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
//
// CPU restatement of apps/blur/halide_blur_generator.cpp:39-40.  Halide keeps u16 + u16 in u16
// (src/IROperator.cpp:769-816), so every sum wraps mod 2^16 before the unsigned divide by 3.
// The in-tree C reference (apps/blur/test.cpp:18-33) promotes to int instead; the two agree on
// inputs below 2^16/3, which is why the harness masks its input to 12 bits (test.cpp:169).
// Parity status: PINNED — tests compare this file with oracle/_ref/libref_blur.so, which is the
// reference's own test.cpp compiled from where it lies.
#include "oracle.h"

extern "C" int oracle_blur(const oracle_image_t *in, const oracle_image_t *out) {
    const uint16_t *ib = (const uint16_t *)in->base;
    uint16_t *ob = (uint16_t *)out->base;
    const int W = out->extent[0], H = out->extent[1];
    const int ox = out->min[0], oy = out->min[1];
    // the pipeline has no boundary condition: required region is [ox, ox+W+1] x [oy, oy+H+1]
    if (W > 0 && H > 0) {
        if (ox < in->min[0] || ox + W + 1 > in->min[0] + in->extent[0] - 1) return -4;
        if (oy < in->min[1] || oy + H + 1 > in->min[1] + in->extent[1] - 1) return -4;
    }
    auto I = [&](int x, int y) -> uint16_t {
        return ib[(int64_t)(x - in->min[0]) * in->stride[0] + (int64_t)(y - in->min[1]) * in->stride[1]];
    };
    auto bx = [&](int x, int y) -> uint16_t {
        uint16_t s = (uint16_t)(I(x, y) + I(x + 1, y));
        s = (uint16_t)(s + I(x + 2, y));
        return (uint16_t)(s / 3);
    };
#pragma omp parallel for schedule(static)
    for (int y = oy; y < oy + H; y++) {
        for (int x = ox; x < ox + W; x++) {
            uint16_t s = (uint16_t)(bx(x, y) + bx(x, y + 1));
            s = (uint16_t)(s + bx(x, y + 2));
            ob[(int64_t)(x - ox) * out->stride[0] + (int64_t)(y - oy) * out->stride[1]] = (uint16_t)(s / 3);
        }
    }
    return 0;
}
