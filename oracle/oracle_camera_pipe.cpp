// oracle/oracle_camera_pipe.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
//
// CPU restatement of apps/camera_pipe/camera_pipe_generator.cpp: shift(16,12) :412, hot-pixel
// suppression :240-249, GRBG deinterleave :251-261, gradient-directed demosaic :57-146, colour matrix
// :263-295, tone curve LUT :297-367, unsharp mask :369-404; op order per SURVEY.md Appendix B.
// All image arithmetic is integer with Halide's no-promotion typing (u16/u8 wrap, avg via widening,
// i16*u8 stays i16) and therefore exactly determined by the sources; only the 12-entry matrix and the
// 1024-entry curve involve floats (halide_pow = halide_exp(halide_log(x)*y)).
// Parity status: UNPINNED by the reference (no golden output; libHalide unbuildable here).
#include <algorithm>
#include <vector>

#include "halide_math.h"
#include "oracle.h"

namespace {

typedef uint16_t u16;
typedef int16_t i16;
typedef uint8_t u8;

inline u16 avg16(u16 a, u16 b) { return (u16)(((uint32_t)a + (uint32_t)b + 1u) / 2u); }  // generator :16-19
inline u16 absd16(u16 a, u16 b) { return a > b ? (u16)(a - b) : (u16)(b - a); }
inline u8 avg8(u8 a, u8 b) { return (u8)(((uint16_t)a + (uint16_t)b + 1u) / 2u); }
inline u8 blur121_8(u8 a, u8 b, u8 c) { return avg8(avg8(a, c), b); }  // generator :21-23

struct Plane16 {
    int x0, y0, w, h;
    std::vector<u16> v;
    void init(int x0_, int y0_, int x1, int y1) {
        x0 = x0_; y0 = y0_; w = x1 - x0_ + 1; h = y1 - y0_ + 1;
        v.assign((size_t)w * h, 0);
    }
    u16 &at(int x, int y) { return v[(size_t)(y - y0) * w + (x - x0)]; }
    u16 at(int x, int y) const { return v[(size_t)(y - y0) * w + (x - x0)]; }
};

}  // namespace

extern "C" void oracle_camera_pipe_tables(const float *m3200, const float *m7000, float color_temp, float gamma, float contrast,
                                          float sharpen_strength, int blackLevel, int whiteLevel, int16_t *matrix12,
                                          uint8_t *curve1024, uint8_t *s32_out) {
    // matrix(x,y) = i16((m3200*alpha + m7000*(1-alpha)) * 256) (generator :268-271); x in [0,4), y in [0,3)
    const float c3200 = 1.0f / 3200, c7000 = 1.0f / 7000;
    const float alpha = hl::div_const(1.0f / color_temp - c3200, c7000 - c3200);
    for (int i = 0; i < 12; i++) {
        float val = m3200[i] * alpha + m7000[i] * (1.0f - alpha);
        matrix12[i] = (int16_t)(val * 256.0f);
    }
    // tone curve (generator :301-332)
    const int minRaw = 0 + blackLevel, maxRaw = whiteLevel;
    const float invRange = 1.0f / (float)(maxRaw - minRaw);
    const float b = 2.0f - hl::halide_pow(2.0f, hl::div_const(contrast, 100.0f));
    const float a = 2.0f - 2.0f * b;
    const float inv_gamma = 1.0f / gamma;
    for (int x = 0; x < 1024; x++) {
        float xf = hl::clampf((float)(x - minRaw) * invRange, 0.0f, 1.0f);
        float g = hl::halide_pow(xf, inv_gamma);
        float z;
        if (g > 0.5f) z = 1.0f - ((a * (1.0f - g)) * (1.0f - g) + b * (1.0f - g));
        else z = (a * g) * g + b * g;
        uint8_t val = (uint8_t)hl::clampf(z * 255.0f + 0.5f, 0.0f, 255.0f);
        curve1024[x] = x <= minRaw ? 0 : (x > maxRaw ? 255 : val);
    }
    // sharpen_strength_x32 = u8_sat(sharpen_strength * 32) (generator :372; saturating float->u8 cast)
    float s = sharpen_strength * 32.0f;
    s = s > 0.0f ? s : 0.0f;
    *s32_out = s >= 255.0f ? 255 : (uint8_t)s;
}

extern "C" int oracle_camera_pipe(const oracle_image_t *in, const oracle_image_t *m3200, const oracle_image_t *m7000,
                                  float color_temp, float gamma, float contrast, float sharpen_strength, int blackLevel,
                                  int whiteLevel, const oracle_image_t *out) {
    const u16 *ib = (const u16 *)in->base;
    u8 *ob = (u8 *)out->base;
    const int ox = out->min[0], oy = out->min[1], oc = out->min[2];
    const int W = out->extent[0], H = out->extent[1], C = out->extent[2];
    if (W <= 0 || H <= 0 || C <= 0) return 0;

    float m32[12], m70[12];
    for (int y = 0; y < 3; y++)
        for (int x = 0; x < 4; x++) {
            m32[y * 4 + x] = ((const float *)m3200->base)[(int64_t)(x - m3200->min[0]) * m3200->stride[0] + (int64_t)(y - m3200->min[1]) * m3200->stride[1]];
            m70[y * 4 + x] = ((const float *)m7000->base)[(int64_t)(x - m7000->min[0]) * m7000->stride[0] + (int64_t)(y - m7000->min[1]) * m7000->stride[1]];
        }
    int16_t M[12];
    uint8_t curve[1024], s32;
    oracle_camera_pipe_tables(m32, m70, color_temp, gamma, contrast, sharpen_strength, blackLevel, whiteLevel, M, curve, &s32);
    auto Mx = [&](int x, int y) -> int32_t { return (int32_t)M[y * 4 + x]; };

    // regions: full-res demosaiced/curved on [ox-1, ox+W] x [oy-1, oy+H]; half-res sites hx on [(ox-1)>>1, (ox+W)>>1];
    // deinterleaved / g_r / g_b on half-res grown by 1; denoised on the matching full-res box; input grown by 2 more.
    const int fx0 = ox - 1, fx1 = ox + W, fy0 = oy - 1, fy1 = oy + H;
    const int hx0 = hl::div_floor(fx0, 2), hx1 = hl::div_floor(fx1, 2), hy0 = hl::div_floor(fy0, 2), hy1 = hl::div_floor(fy1, 2);
    const int dx0 = 2 * (hx0 - 1), dx1 = 2 * (hx1 + 1) + 1, dy0 = 2 * (hy0 - 1), dy1 = 2 * (hy1 + 1) + 1;
    // bounds check on the input (no boundary condition anywhere in this pipeline)
    {
        int need_x0 = dx0 - 2 + 16, need_x1 = dx1 + 2 + 16, need_y0 = dy0 - 2 + 12, need_y1 = dy1 + 2 + 12;
        if (need_x0 < in->min[0] || need_x1 > in->min[0] + in->extent[0] - 1 || need_y0 < in->min[1] ||
            need_y1 > in->min[1] + in->extent[1] - 1)
            return -4;
    }
    auto shifted = [&](int x, int y) -> u16 {  // generator :412
        return ib[(int64_t)(x + 16 - in->min[0]) * in->stride[0] + (int64_t)(y + 12 - in->min[1]) * in->stride[1]];
    };
    Plane16 den;
    den.init(dx0, dy0, dx1, dy1);
#pragma omp parallel for schedule(static)
    for (int y = dy0; y <= dy1; y++)
        for (int x = dx0; x <= dx1; x++) {
            u16 a = std::max(std::max(shifted(x - 2, y), shifted(x + 2, y)), std::max(shifted(x, y - 2), shifted(x, y + 2)));
            u16 v = shifted(x, y);
            den.at(x, y) = std::min(v, a);  // clamp(v, 0, a) on unsigned (generator :240-249)
        }
    // deinterleaved channels at half-res site (x,y) (generator :251-261, :57-60)
    auto g_gr = [&](int x, int y) { return den.at(2 * x, 2 * y); };
    auto r_r = [&](int x, int y) { return den.at(2 * x + 1, 2 * y); };
    auto b_b = [&](int x, int y) { return den.at(2 * x, 2 * y + 1); };
    auto g_gb = [&](int x, int y) { return den.at(2 * x + 1, 2 * y + 1); };
    // green at red / blue sites (generator :70-82), on the half-res box grown by 1 where the deinterleaved taps exist
    Plane16 g_r, g_b;
    g_r.init(hx0 - 1, hy0 - 1, hx1 + 1, hy1 + 1);
    g_b.init(hx0 - 1, hy0 - 1, hx1 + 1, hy1 + 1);
#pragma omp parallel for schedule(static)
    for (int y = hy0 - 1; y <= hy1 + 1; y++)
        for (int x = hx0 - 1; x <= hx1 + 1; x++) {
            // taps outside the denoised box are never consumed by the sites the output needs; guard the reads
            auto G = [&](bool gb, int xx, int yy) -> u16 {
                int fx = 2 * xx + (gb ? 1 : 0), fy = 2 * yy + (gb ? 1 : 0);
                if (fx < dx0 || fx > dx1 || fy < dy0 || fy > dy1) return 0;
                return den.at(fx, fy);
            };
            u16 gv_r = avg16(G(true, x, y - 1), G(true, x, y)), gvd_r = absd16(G(true, x, y - 1), G(true, x, y));
            u16 gh_r = avg16(G(false, x + 1, y), G(false, x, y)), ghd_r = absd16(G(false, x + 1, y), G(false, x, y));
            g_r.at(x, y) = ghd_r < gvd_r ? gh_r : gv_r;
            u16 gv_b = avg16(G(false, x, y + 1), G(false, x, y)), gvd_b = absd16(G(false, x, y + 1), G(false, x, y));
            u16 gh_b = avg16(G(true, x - 1, y), G(true, x, y)), ghd_b = absd16(G(true, x - 1, y), G(true, x, y));
            g_b.at(x, y) = ghd_b < gvd_b ? gh_b : gv_b;
        }
    // curved (u8) on the full-res box [fx0,fx1] x [fy0,fy1], 3 channels
    const int cw = fx1 - fx0 + 1, ch = fy1 - fy0 + 1;
    std::vector<u8> curved((size_t)3 * cw * ch);
    auto CV = [&](int x, int y, int c) -> u8 & { return curved[((size_t)c * ch + (y - fy0)) * cw + (x - fx0)]; };
#pragma omp parallel for schedule(static)
    for (int Y = fy0; Y <= fy1; Y++)
        for (int X = fx0; X <= fx1; X++) {
            const int x = hl::div_floor(X, 2), y = hl::div_floor(Y, 2);
            const bool xe = hl::mod_floor(X, 2) == 0, ye = hl::mod_floor(Y, 2) == 0;
            u16 r, g, b;
            if (ye && xe) {  // gr site (generator :89-96)
                g = g_gr(x, y);
                u16 corr = (u16)(g_gr(x, y) - avg16(g_r.at(x, y), g_r.at(x - 1, y)));
                r = (u16)(corr + avg16(r_r(x - 1, y), r_r(x, y)));
                corr = (u16)(g_gr(x, y) - avg16(g_b.at(x, y), g_b.at(x, y - 1)));
                b = (u16)(corr + avg16(b_b(x, y), b_b(x, y - 1)));
            } else if (ye && !xe) {  // r site (generator :121-130)
                r = r_r(x, y);
                g = g_r.at(x, y);
                u16 corr = (u16)(g_r.at(x, y) - avg16(g_b.at(x, y), g_b.at(x + 1, y - 1)));
                u16 bp = (u16)(corr + avg16(b_b(x, y), b_b(x + 1, y - 1)));
                u16 bpd = absd16(b_b(x, y), b_b(x + 1, y - 1));
                corr = (u16)(g_r.at(x, y) - avg16(g_b.at(x + 1, y), g_b.at(x, y - 1)));
                u16 bn = (u16)(corr + avg16(b_b(x + 1, y), b_b(x, y - 1)));
                u16 bnd = absd16(b_b(x + 1, y), b_b(x, y - 1));
                b = bpd < bnd ? bp : bn;
            } else if (!ye && xe) {  // b site (generator :111-119)
                b = b_b(x, y);
                g = g_b.at(x, y);
                u16 corr = (u16)(g_b.at(x, y) - avg16(g_r.at(x, y), g_r.at(x - 1, y + 1)));
                u16 rp = (u16)(corr + avg16(r_r(x, y), r_r(x - 1, y + 1)));
                u16 rpd = absd16(r_r(x, y), r_r(x - 1, y + 1));
                corr = (u16)(g_b.at(x, y) - avg16(g_r.at(x - 1, y), g_r.at(x, y + 1)));
                u16 rn = (u16)(corr + avg16(r_r(x - 1, y), r_r(x, y + 1)));
                u16 rnd = absd16(r_r(x - 1, y), r_r(x, y + 1));
                r = rpd < rnd ? rp : rn;
            } else {  // gb site (generator :98-102)
                g = g_gb(x, y);
                u16 corr = (u16)(g_gb(x, y) - avg16(g_r.at(x, y), g_r.at(x, y + 1)));
                r = (u16)(corr + avg16(r_r(x, y), r_r(x, y + 1)));
                corr = (u16)(g_gb(x, y) - avg16(g_b.at(x, y), g_b.at(x + 1, y)));
                b = (u16)(corr + avg16(b_b(x, y), b_b(x + 1, y)));
            }
            // demosaiced output is the i16 reinterpretation (generator :146); colour matrix in i32 (generator :277-292)
            const int32_t ir = (int32_t)(i16)r, ig = (int32_t)(i16)g, ib_ = (int32_t)(i16)b;
            for (int c = 0; c < 3; c++) {
                int32_t acc = ((Mx(3, c) + Mx(0, c) * ir) + Mx(1, c) * ig) + Mx(2, c) * ib_;
                i16 cc = (i16)hl::div_floor(acc, 256);
                CV(X, Y, c) = curve[hl::clampi((int)cc, 0, 1023)];
            }
        }
    // sharpen (generator :384-401)
#pragma omp parallel for collapse(2) schedule(static)
    for (int c = oc; c < oc + C; c++)
        for (int y = oy; y < oy + H; y++)
            for (int x = ox; x < ox + W; x++) {
                auto UY = [&](int xx) { return blur121_8(CV(xx, y - 1, c), CV(xx, y, c), CV(xx, y + 1, c)); };
                u8 unsharp = blur121_8(UY(x - 1), UY(x), UY(x + 1));
                i16 mask = (i16)((i16)CV(x, y, c) - (i16)unsharp);
                i16 prod = (i16)(mask * (i16)s32);  // i16 * u8 stays i16 and wraps (src/IROperator.cpp:769-816)
                i16 sum = (i16)((i16)CV(x, y, c) + (i16)hl::div_floor((int)prod, 32));
                ob[(int64_t)(x - ox) * out->stride[0] + (int64_t)(y - oy) * out->stride[1] + (int64_t)(c - oc) * out->stride[2]] =
                    (u8)hl::clampi((int)sum, 0, 255);
            }
    return 0;
}
