// oracle/ref_blur_wrap.cpp — TEST INFRASTRUCTURE.  Compiles the REFERENCE's own blur
// implementations (apps/blur/test.cpp: `blur` :18-33, `blur_fast` :35-132) into
// oracle/_ref/libref_blur.so by including the reference source where it lies.  No reference
// source is copied into this repository; without /root/reference this file does not build and
// the prebuilt .so (shipped to the GPU box by gpurun) is used.
#include <cstring>
#define main ref_blur_test_main
#include "apps_blur_test_include.h"
#undef main

// test.cpp calls the AOT filter through blur_halide(); the reference-only library never calls it,
// but the symbols must resolve.
extern "C" int halide_blur(struct halide_buffer_t *, struct halide_buffer_t *) { return -20; }
extern "C" const struct halide_filter_metadata_t *halide_blur_metadata() {
    static const halide_filter_metadata_t md = {1, 0, nullptr, "host", "halide_blur"};
    return &md;
}

// in: (w x h) dense u16; out: (w-8) x (h-2) dense u16, exactly the shapes test.cpp:19-20 uses.
extern "C" int ref_blur(const uint16_t *in, int w, int h, uint16_t *out, int fast) {
    Buffer<uint16_t, 2> input(w, h);
    memcpy(input.data(), in, (size_t)w * h * 2);
    Buffer<uint16_t, 2> result = fast ? blur_fast(input) : blur(input);
    for (int y = 0; y < h - 2; y++) memcpy(out + (size_t)y * (w - 8), &result(0, y), (size_t)(w - 8) * 2);
    return 0;
}
