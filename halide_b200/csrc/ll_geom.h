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
    Span ox, oy;   // region of outGPyramid[j]
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
        g.lv[j].gpitch = (SX[j].n() + 3) & ~3;
        g.lv[j].opitch = (OX[j].n() + 3) & ~3;
    }
    return g;
}

}  // namespace ll
