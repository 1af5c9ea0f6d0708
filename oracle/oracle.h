// oracle/oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry points of the CPU oracle: a restatement of the seven Halide app pipelines in plain
// scalar C++ following SURVEY.md Appendix A/B, evaluated the way Halide's bounds inference
// defines them (every Func is a pure function on Z^n, computed on whatever enlarged region its
// consumers touch; only the pipeline input is edge-clamped).
//
// PARITY PINNING: blur is pinned against the reference's in-tree C implementation
// (apps/blur/test.cpp:18-33, compiled from where it lies into oracle/_ref by oracle/Makefile).
// The other six pipelines have no golden outputs in the reference (its app tests only check
// for "Success!") and libHalide cannot be built in this image (needs LLVM), so for them this
// oracle is "parity unpinned": it is pinned only at the primitive level (tests/test_oracle_*.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may use it.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// Image view: `base` addresses the element at coordinates (min[0], min[1], min[2]); strides in elements.
typedef struct {
    void *base;
    int32_t min[4], extent[4], stride[4];
} oracle_image_t;

// apps/blur/halide_blur_generator.cpp:39-40.  in must cover out grown by +2 in x and y.
int oracle_blur(const oracle_image_t *in, const oracle_image_t *out);

// apps/local_laplacian/local_laplacian_generator.cpp:19-87,266-282 with pyramid_levels J (8 in the app).
int oracle_local_laplacian(const oracle_image_t *in, int levels, float alpha, float beta,
                           const oracle_image_t *out, int pyramid_levels);

// apps/bilateral_grid/bilateral_grid_generator.cpp:17-67 (s_sigma is the GeneratorParam, 8 in the app).
int oracle_bilateral_grid(const oracle_image_t *in, float r_sigma, const oracle_image_t *out, int s_sigma);

// apps/nl_means/nl_means_generator.cpp:24-63.  out has 3 channels.
int oracle_nl_means(const oracle_image_t *in, int patch_size, int search_area, float sigma, const oracle_image_t *out);

// apps/stencil_chain/stencil_chain_generator.cpp:16-34 (`stencils` is the GeneratorParam, 32 in the app).
int oracle_stencil_chain(const oracle_image_t *in, const oracle_image_t *out, int stencils);

// apps/camera_pipe/camera_pipe_generator.cpp (whole pipeline).  in: u16 2-D raw; m3200/m7000: f32 (4 x 3); out: u8 3-D.
// Returns -4 when the input does not cover the stencil footprint.
int oracle_camera_pipe(const oracle_image_t *in, const oracle_image_t *m3200, const oracle_image_t *m7000, float color_temp,
                       float gamma, float contrast, float sharpen_strength, int blackLevel, int whiteLevel,
                       const oracle_image_t *out);
// The float-derived tables of camera_pipe (Q8.8 matrix [y*4+x], 1024-entry tone curve, sharpen strength x32).
void oracle_camera_pipe_tables(const float *m3200, const float *m7000, float color_temp, float gamma, float contrast,
                               float sharpen_strength, int blackLevel, int whiteLevel, int16_t *matrix12, uint8_t *curve1024,
                               uint8_t *s32_out);

// apps/conv_layer/conv_layer_generator.cpp:17-27; dense arrays in the generator's fixed layouts.
int oracle_conv_layer(const float *input, const float *filter, const float *bias, float *out, int N, int CI, int CO, int W, int H);

// Primitive probes so the tests can pin the math helpers against known values.
float oracle_halide_exp(float x);
float oracle_halide_log(float x);
float oracle_halide_pow(float x, float y);
float oracle_fast_exp(float x);
int oracle_div_floor(int a, int b);
int oracle_mod_floor(int a, int b);
// remap LUT entry i in [-256(levels-1), 256(levels-1)] (local_laplacian_generator.cpp:24-25)
float oracle_ll_remap(int i, float alpha);

int oracle_num_threads(void);
// torchrun exports OMP_NUM_THREADS=1; the CPU baseline legs of bench.py restore all host cores through this.
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
