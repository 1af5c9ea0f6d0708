// oracle/oracle_conv_layer.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
//
// CPU restatement of apps/conv_layer/conv_layer_generator.cpp:17-27:
//   conv(c,x,y,n) = bias(c);  conv += filter(c, r.y, r.z, r.x) * input(r.x, x + r.y, y + r.z, n)
//   with RDom r(0,CI, 0,3, 0,3) (r.x = ci innermost, then kx, then ky);  relu = max(0, conv).
// Parity status: UNPINNED by the reference.  Float pipeline: the CUDA path is compared within 1e-4 relative
// (it necessarily sums in a different order).
#include "oracle.h"

extern "C" int oracle_conv_layer(const float *input, const float *filter, const float *bias, float *out, int N, int CI, int CO,
                                 int W, int H) {
    // input (CI, W+2, H+2, N) ci-innermost; filter (CO, 3, 3, CI) co-innermost; out (CO, W, H, N) co-innermost
    const int64_t in_sx = CI, in_sy = (int64_t)CI * (W + 2), in_sn = in_sy * (H + 2);
    const int64_t f_skx = CO, f_sky = (int64_t)CO * 3, f_sci = (int64_t)CO * 9;
    const int64_t o_sx = CO, o_sy = (int64_t)CO * W, o_sn = o_sy * H;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; n++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                for (int co = 0; co < CO; co++) {
                    float acc = bias[co];
                    for (int ky = 0; ky < 3; ky++)
                        for (int kx = 0; kx < 3; kx++)
                            for (int ci = 0; ci < CI; ci++)
                                acc = acc + filter[co + kx * f_skx + ky * f_sky + ci * f_sci] *
                                                input[ci + (x + kx) * in_sx + (y + ky) * in_sy + n * in_sn];
                    out[co + x * o_sx + y * o_sy + n * o_sn] = acc > 0.0f ? acc : 0.0f;
                }
    return 0;
}
