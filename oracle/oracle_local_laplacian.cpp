// oracle/oracle_local_laplacian.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
//
// CPU restatement of apps/local_laplacian/local_laplacian_generator.cpp (algorithm section
// :19-87, downsample :266-273, upsample :276-282), op order per SURVEY.md Appendix B.
// Parity status: UNPINNED by the reference (no golden output exists; libHalide is unbuildable
// here) — this file *is* the definition the CUDA kernels are checked against.
//
// Deliberately different in structure from the CUDA path: every pyramid level is evaluated on
// the enlarged region G_j that bounds inference would give it (no per-level clamping tricks),
// so it independently checks the kernels' "clamp into the constant border" storage scheme.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "halide_math.h"
#include "oracle.h"

namespace {

struct Rng {
    int lo, hi;
    int n() const { return hi - lo + 1; }
};

struct Plane {  // K stacked f32 planes over X x Y
    Rng X, Y;
    int K;
    std::vector<float> v;
    void init(Rng x, Rng y, int k) {
        X = x; Y = y; K = k;
        v.assign((size_t)x.n() * y.n() * k, 0.f);
    }
    inline float &at(int x, int y, int k) {
        return v[((size_t)k * Y.n() + (y - Y.lo)) * X.n() + (x - X.lo)];
    }
    inline float at(int x, int y, int k) const {
        return v[((size_t)k * Y.n() + (y - Y.lo)) * X.n() + (x - X.lo)];
    }
};

// O_j: region where the output-side pyramids are needed; G_j: region of the Gaussian pyramids.
void regions(int a0, int b0, int J, std::vector<Rng> &O, std::vector<Rng> &G) {
    O.resize(J); G.resize(J);
    O[0] = {a0, b0};
    for (int j = 1; j < J; j++) {
        // upsample reads f((x+1)/2) and f((x-1)/2) with floor division (generator :279-280)
        O[j] = {hl::div_floor(O[j - 1].lo - 1, 2), hl::div_floor(O[j - 1].hi + 1, 2)};
    }
    G[J - 1] = O[J - 1];
    for (int j = J - 2; j >= 0; j--) {
        // downsample reads f(2x-1 .. 2x+2) (generator :270-271)
        G[j] = {std::min(O[j].lo, 2 * G[j + 1].lo - 1), std::max(O[j].hi, 2 * G[j + 1].hi + 2)};
    }
}

// downsample(f) over the region of `dst` (generator :266-273): y first, then x.
void downsample(const Plane &src, Plane &dst) {
    const int xlo = 2 * dst.X.lo - 1, xhi = 2 * dst.X.hi + 2;
#pragma omp parallel
    {
        std::vector<float> tmp(xhi - xlo + 1);
#pragma omp for collapse(2) schedule(static)
        for (int k = 0; k < dst.K; k++) {
            for (int y = dst.Y.lo; y <= dst.Y.hi; y++) {
                for (int x = xlo; x <= xhi; x++) {
                    tmp[x - xlo] = ((src.at(x, 2 * y - 1, k) + 3.0f * (src.at(x, 2 * y, k) + src.at(x, 2 * y + 1, k))) +
                                    src.at(x, 2 * y + 2, k)) * 0.125f;
                }
                for (int x = dst.X.lo; x <= dst.X.hi; x++) {
                    const float *t = &tmp[2 * x - 1 - xlo];
                    dst.at(x, y, k) = ((t[0] + 3.0f * (t[1] + t[2])) + t[3]) * 0.125f;
                }
            }
        }
    }
}

inline float upsample_at(const Plane &f, int x, int y, int k) {
    // upx(x,y) = lerp(f((x+1)/2,y), f((x-1)/2,y), ((x%2)*2+1)/4.0f); upy likewise (generator :279-280)
    const float wx = (float)(hl::mod_floor(x, 2) * 2 + 1) * 0.25f;
    const float wy = (float)(hl::mod_floor(y, 2) * 2 + 1) * 0.25f;
    const int xa = hl::div_floor(x + 1, 2), xb = hl::div_floor(x - 1, 2);
    const int ya = hl::div_floor(y + 1, 2), yb = hl::div_floor(y - 1, 2);
    const float upx_a = hl::lerpf(f.at(xa, ya, k), f.at(xb, ya, k), wx);
    const float upx_b = hl::lerpf(f.at(xa, yb, k), f.at(xb, yb, k), wx);
    return hl::lerpf(upx_a, upx_b, wy);
}

}  // namespace

extern "C" float oracle_ll_remap(int i, float alpha) {
    // remap(x) = alpha * fx * exp(-fx*fx/2), fx = x/256 (generator :24-25); Appendix B nesting.
    const float fx = (float)i * 0.00390625f;
    return (alpha * fx) * hl::halide_exp(((0.0f - fx) * fx) * 0.5f);
}

extern "C" int oracle_local_laplacian(const oracle_image_t *in, int levels, float alpha, float beta,
                                      const oracle_image_t *out, int J) {
    if (levels < 2 || J < 1 || J > 20) return -1;
    const uint16_t *ibase = (const uint16_t *)in->base;
    uint16_t *obase = (uint16_t *)out->base;
    const int ox = out->min[0], oy = out->min[1], oc = out->min[2];
    const int W = out->extent[0], H = out->extent[1], C = out->extent[2];
    if (W <= 0 || H <= 0 || C <= 0) return 0;

    auto in_clamped = [&](int x, int y, int c) -> uint16_t {  // repeat_edge (src/BoundaryConditions.cpp:15-35)
        x = hl::clampi(x, in->min[0], in->min[0] + in->extent[0] - 1);
        y = hl::clampi(y, in->min[1], in->min[1] + in->extent[1] - 1);
        c = hl::clampi(c, in->min[2], in->min[2] + in->extent[2] - 1);
        return ibase[(int64_t)(x - in->min[0]) * in->stride[0] + (int64_t)(y - in->min[1]) * in->stride[1] +
                     (int64_t)(c - in->min[2]) * in->stride[2]];
    };

    std::vector<Rng> OX, GX, OY, GY;
    regions(ox, ox + W - 1, J, OX, GX);
    regions(oy, oy + H - 1, J, OY, GY);

    const int K = levels;
    const int lut_half = 256 * (levels - 1);
    std::vector<float> lut(2 * lut_half + 1);
    for (int i = -lut_half; i <= lut_half; i++) lut[i + lut_half] = oracle_ll_remap(i, alpha);

    const float flm1 = (float)(levels - 1);
    const float inv_lm1 = 1.0f / flm1;

    // gray on G_0 (generator :32-36)
    Plane gray;
    gray.init(GX[0], GY[0], 1);
#pragma omp parallel for schedule(static)
    for (int y = GY[0].lo; y <= GY[0].hi; y++) {
        for (int x = GX[0].lo; x <= GX[0].hi; x++) {
            const float f0 = hl::div_const((float)in_clamped(x, y, 0), 65535.0f);
            const float f1 = hl::div_const((float)in_clamped(x, y, 1), 65535.0f);
            const float f2 = hl::div_const((float)in_clamped(x, y, 2), 65535.0f);
            gray.at(x, y, 0) = (0.299f * f0 + 0.587f * f1) + 0.114f * f2;
        }
    }

    // gPyramid[0](x,y,k) (generator :41-44), evaluated on demand
    auto gp0 = [&](int x, int y, int k) -> float {
        const float g = gray.at(x, y, 0);
        const float level = (float)k * inv_lm1;
        int idx = (int)((g * flm1) * 256.0f);
        idx = hl::clampi(idx, 0, (levels - 1) * 256);
        return ((beta * (g - level)) + level) + lut[idx - 256 * k + lut_half];
    };

    std::vector<Plane> gP(J), inG(J);
    // level 0 of the K-plane pyramid is never stored (it is 8x the frame); level 1 is built from gp0
    if (J > 1) {
        gP[1].init(GX[1], GY[1], K);
        const int xlo = 2 * GX[1].lo - 1, xhi = 2 * GX[1].hi + 2;
#pragma omp parallel
        {
            std::vector<float> tmp(xhi - xlo + 1);
#pragma omp for collapse(2) schedule(static)
            for (int k = 0; k < K; k++) {
                for (int y = GY[1].lo; y <= GY[1].hi; y++) {
                    for (int x = xlo; x <= xhi; x++) {
                        tmp[x - xlo] = ((gp0(x, 2 * y - 1, k) + 3.0f * (gp0(x, 2 * y, k) + gp0(x, 2 * y + 1, k))) +
                                        gp0(x, 2 * y + 2, k)) * 0.125f;
                    }
                    for (int x = GX[1].lo; x <= GX[1].hi; x++) {
                        const float *t = &tmp[2 * x - 1 - xlo];
                        gP[1].at(x, y, k) = ((t[0] + 3.0f * (t[1] + t[2])) + t[3]) * 0.125f;
                    }
                }
            }
        }
        inG[1].init(GX[1], GY[1], 1);
        downsample(gray, inG[1]);
    }
    for (int j = 2; j < J; j++) {
        gP[j].init(GX[j], GY[j], K);
        downsample(gP[j - 1], gP[j]);
        inG[j].init(GX[j], GY[j], 1);
        downsample(inG[j - 1], inG[j]);
    }

    // output pyramids (generator :63-79)
    std::vector<Plane> outG(J);
    for (int j = J - 1; j >= 0; j--) {
        outG[j].init(OX[j], OY[j], 1);
#pragma omp parallel for schedule(static)
        for (int y = OY[j].lo; y <= OY[j].hi; y++) {
            for (int x = OX[j].lo; x <= OX[j].hi; x++) {
                const float ing = (j == 0) ? gray.at(x, y, 0) : inG[j].at(x, y, 0);
                const float level = ing * flm1;
                const int li = hl::clampi((int)level, 0, levels - 2);
                const float lf = level - (float)li;
                float l0, l1;  // lPyramid[j](x,y,li), lPyramid[j](x,y,li+1)
                const float g0 = (j == 0) ? gp0(x, y, li) : gP[j].at(x, y, li);
                const float g1 = (j == 0) ? gp0(x, y, li + 1) : gP[j].at(x, y, li + 1);
                if (j == J - 1) {
                    l0 = g0;
                    l1 = g1;
                } else {
                    l0 = g0 - upsample_at(gP[j + 1], x, y, li);
                    l1 = g1 - upsample_at(gP[j + 1], x, y, li + 1);
                }
                const float outl = (1.0f - lf) * l0 + lf * l1;
                outG[j].at(x, y, 0) = (j == J - 1) ? outl : upsample_at(outG[j + 1], x, y, 0) + outl;
            }
        }
    }

    // colour reintroduction + cast (generator :82-87); input(x,y,c) here is NOT clamped
    const float eps = 0.01f;
#pragma omp parallel for collapse(2) schedule(static)
    for (int c = oc; c < oc + C; c++) {
        for (int y = oy; y < oy + H; y++) {
            for (int x = ox; x < ox + W; x++) {
                const uint16_t iv = ibase[(int64_t)(x - in->min[0]) * in->stride[0] + (int64_t)(y - in->min[1]) * in->stride[1] +
                                          (int64_t)(c - in->min[2]) * in->stride[2]];
                const float color = ((float)iv * (outG[0].at(x, y, 0) + eps)) / (gray.at(x, y, 0) + eps);
                const float cl = hl::clampf(color, 0.0f, 65535.0f);
                obase[(int64_t)(x - ox) * out->stride[0] + (int64_t)(y - oy) * out->stride[1] +
                      (int64_t)(c - oc) * out->stride[2]] = (uint16_t)cl;
            }
        }
    }
    return 0;
}
