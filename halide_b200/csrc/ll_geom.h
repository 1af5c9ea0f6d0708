// ll_geom.h — pyramid geometry of local_laplacian shared by the host FFI code and the kernels.
//
// Halide evaluates every pyramid level as a pure function on an ENLARGED region derived by bounds
// inference (only the input is edge-clamped, SURVEY.md finding 5; generator
// apps/local_laplacian/local_laplacian_generator.cpp:28,266-282).  Because the input is constant
// outside its own bounds, each level is constant beyond a small border: if level j is constant
// for x <= L_j and for x >= R_j then level j+1 (which reads 2x-1 .. 2x+2) is constant for
// x <= floor((L_j-2)/2) and x >= ceil((R_j+1)/2).  The kernels therefore store level j only on
// S_j = G_j ∩ [L_j, R_j] and clamp coordinates into S_j when reading — bit-identical to the
// enlarged-region evaluation (the oracle evaluates the enlarged regions literally and the parity
// tests compare the two), at ~W/2^j + 3 columns per level instead of W/2^j + 2^(J-j+1).
#pragma once
#include <stdint.h>

namespace ll {

constexpr int kMaxJ = 8;  // pyramid_levels GeneratorParam of the shipped app (generator :10)

struct Span {
    int lo, hi;  // inclusive
    __host__ __device__ int n() const { return hi - lo + 1; }
};

__host__ __device__ inline int fdiv2(int a) { return a >> 1; }  // Euclidean /2 on signed ints

struct Level {
    Span sx, sy;   // stored region of the Gaussian-side planes (gPyramid[j] K planes + inGPyramid[j])
    Span ox, oy;   // region of outGPyramid[j] (stored)
    Span cy, coy;  // rows this device COMPUTES (== sy / oy on one GPU; the owned band when row-sharded)
    Span gy, goy;  // rows of the whole frame's S_j / O_j (== sy / oy on one GPU): the clamp range
    int gpitch;    // pixels per row of the Gaussian-side planes
    int opitch;    // pixels per row of the outG plane
};

struct Geom {
    int J;
    Level lv[kMaxJ];
};

// out_[xy]: output region; in_[xy]: bounds of the input buffer (for the clamp / constant borders).
inline void compute_axis(Span out, Span in, int J, Span *S, Span *O) {
    Span G[kMaxJ];
    O[0] = out;
    for (int j = 1; j < J; j++) O[j] = {fdiv2(O[j - 1].lo - 1), fdiv2(O[j - 1].hi + 1)};
    G[J - 1] = O[J - 1];
    for (int j = J - 2; j >= 0; j--) {
        int lo = 2 * G[j + 1].lo - 1, hi = 2 * G[j + 1].hi + 2;
        G[j] = {O[j].lo < lo ? O[j].lo : lo, O[j].hi > hi ? O[j].hi : hi};
    }
    int L = in.lo, R = in.hi;
    for (int j = 0; j < J; j++) {
        if (j > 0) {
            L = fdiv2(L - 2);
            R = fdiv2(R + 2);  // ceil((R+1)/2)
        }
        int lo = G[j].lo > L ? G[j].lo : L;
        if (lo > G[j].hi) lo = G[j].hi;
        int hi = G[j].hi < R ? G[j].hi : R;
        if (hi < lo) hi = lo;
        S[j] = {lo, hi};
    }
}

inline Geom make_geom(Span outx, Span outy, Span inx, Span iny, int J) {
    Geom g;
    g.J = J;
    Span SX[kMaxJ], OX[kMaxJ], SY[kMaxJ], OY[kMaxJ];
    compute_axis(outx, inx, J, SX, OX);
    compute_axis(outy, iny, J, SY, OY);
    for (int j = 0; j < J; j++) {
        g.lv[j].sx = SX[j]; g.lv[j].sy = SY[j];
        g.lv[j].ox = OX[j]; g.lv[j].oy = OY[j];
        g.lv[j].cy = SY[j]; g.lv[j].coy = OY[j];
        g.lv[j].gy = SY[j]; g.lv[j].goy = OY[j];
        g.lv[j].gpitch = (SX[j].n() + 3) & ~3;
        g.lv[j].opitch = (OX[j].n() + 3) & ~3;
    }
    return g;
}

// ---- row sharding (multi-GPU) -----------------------------------------------------------------------
// The frame's rows are split into contiguous bands, one per rank.  Rank r owns input/output rows
// [band.lo, band.hi]; at level j+1 it owns the rows whose first source row (2y) it owns:
// A_{j+1} = ceil(A_j / 2).  To produce its rows of level j+1 it needs, besides its own rows of level j,
// ONE row above and TWO rows below (the 1-3-3-1 taps 2y-1 .. 2y+2); to produce its rows of
// outGPyramid[j] it needs one row above and one below of outGPyramid[j+1] and gPyramid[j+1]
// (taps (y-1)/2, (y+1)/2).  Those halo rows are what neighbours exchange once per level.
// Rows beyond the first/last rank's band belong to that rank (the constant border of ll_geom).
struct BandLevel {
    Span own, stored;      // Gaussian side: rows computed here / rows held here (own + halo)
    Span own_o, stored_o;  // outGPyramid side
};

inline void compute_band_y(const Geom &whole, Span band, bool first, bool last, BandLevel *bl) {
    int a_lo = band.lo, a_hi = band.hi + 1;  // [a_lo, a_hi)
    for (int j = 0; j < whole.J; j++) {
        if (j > 0) {
            a_lo = (a_lo + 1) >> 1;
            a_hi = (a_hi + 1) >> 1;
        }
        const Span S = whole.lv[j].sy, O = whole.lv[j].oy;
        BandLevel &b = bl[j];
        b.own.lo = first ? S.lo : (a_lo > S.lo ? a_lo : S.lo);
        b.own.hi = last ? S.hi : (a_hi - 1 < S.hi ? a_hi - 1 : S.hi);
        b.stored.lo = first ? b.own.lo : (b.own.lo - 1 > S.lo ? b.own.lo - 1 : S.lo);
        b.stored.hi = last ? b.own.hi : (b.own.hi + 2 < S.hi ? b.own.hi + 2 : S.hi);
        b.own_o.lo = first ? O.lo : (a_lo > O.lo ? a_lo : O.lo);
        b.own_o.hi = last ? O.hi : (a_hi - 1 < O.hi ? a_hi - 1 : O.hi);
        b.stored_o.lo = first ? b.own_o.lo : (b.own_o.lo - 1 > O.lo ? b.own_o.lo - 1 : O.lo);
        b.stored_o.hi = last ? b.own_o.hi : (b.own_o.hi + 1 < O.hi ? b.own_o.hi + 1 : O.hi);
    }
}

// Geometry of one rank: the whole frame's geometry with the y spans replaced by the band's.
inline Geom make_band_geom(const Geom &whole, const BandLevel *bl) {
    Geom g = whole;
    for (int j = 0; j < whole.J; j++) {
        g.lv[j].gy = whole.lv[j].sy; g.lv[j].goy = whole.lv[j].oy;
        g.lv[j].sy = bl[j].stored; g.lv[j].cy = bl[j].own;
        g.lv[j].oy = bl[j].stored_o; g.lv[j].coy = bl[j].own_o;
    }
    return g;
}

}  // namespace ll
