// oracle/oracle_stencil_chain.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
//
// CPU restatement of apps/stencil_chain/stencil_chain_generator.cpp:16-34: `stencils` successive
// 5x5 stencils with weight (i+3)*(j+3), i (x offset) outer, j (y offset) inner, everything in
// uint16 with wrap-around (e starts as cast<uint16_t>(0); `+=` casts the RHS to uint16,
// src/IROperator.cpp:1857-1862).  Only the input is edge-clamped; stage s is evaluated on the
// output region grown by 2*(stencils-s) (bounds inference), which this file does literally.
// Exactly determined by the sources (integer arithmetic) — no float ambiguity.
#include <vector>

#include "halide_math.h"
#include "oracle.h"

extern "C" int oracle_stencil_chain(const oracle_image_t *in, const oracle_image_t *out, int stencils) {
    const uint16_t *ib = (const uint16_t *)in->base;
    uint16_t *ob = (uint16_t *)out->base;
    const int ox = out->min[0], oy = out->min[1], W = out->extent[0], H = out->extent[1];
    if (W <= 0 || H <= 0) return 0;
    const int R = 2 * stencils;
    // stage 0 = clamped input on the fully grown region
    int x0 = ox - R, y0 = oy - R, w = W + 2 * R, h = H + 2 * R;
    std::vector<uint16_t> cur((size_t)w * h), nxt;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            int cx = hl::clampi(x0 + x, in->min[0], in->min[0] + in->extent[0] - 1);
            int cy = hl::clampi(y0 + y, in->min[1], in->min[1] + in->extent[1] - 1);
            cur[(size_t)y * w + x] = ib[(int64_t)(cx - in->min[0]) * in->stride[0] + (int64_t)(cy - in->min[1]) * in->stride[1]];
        }
    }
    for (int s = 0; s < stencils; s++) {
        const int nw = w - 4, nh = h - 4;
        nxt.assign((size_t)nw * nh, 0);
#pragma omp parallel for schedule(static)
        for (int y = 0; y < nh; y++) {
            for (int x = 0; x < nw; x++) {
                uint16_t e = 0;
                for (int i = -2; i <= 2; i++) {
                    for (int j = -2; j <= 2; j++) {
                        uint16_t wgt = (uint16_t)((i + 3) * (j + 3));
                        e = (uint16_t)(e + (uint16_t)(wgt * cur[(size_t)(y + 2 + j) * w + (x + 2 + i)]));
                    }
                }
                nxt[(size_t)y * nw + x] = e;
            }
        }
        cur.swap(nxt);
        w = nw; h = nh;
    }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            ob[(int64_t)x * out->stride[0] + (int64_t)y * out->stride[1]] = cur[(size_t)y * w + x];
    return 0;
}
