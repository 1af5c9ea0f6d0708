// ref_image_io_tool.cpp — TEST INFRASTRUCTURE.  A command-line wrapper around the reference's own image I/O header,
// #included from where it lies (tools/halide_image_io.h, never copied), built into oracle/_ref/ref_image_io by
// oracle/Makefile.  It is the pin for halide_b200/image_io.py: element conversions
// (tools/halide_image_io.h:80-710), the format readers / writers (:1030-2400) and load_and_convert_image /
// convert_and_save_image (:2760-2815) are exercised through it and compared with the Python implementation.
//
//   ref_image_io convert  SRC DST in.bin out.bin      element-wise Internal::convert<DST, SRC> of a raw array
//   ref_image_io load     FILE out.dump               load<Buffer<>>: dump = text header line + planar payload
//   ref_image_io loadconv FILE DST out.dump           load_and_convert_image into Buffer<DST>
//   ref_image_io save     in.dump FILE                save_image (the format must hold the type exactly)
//   ref_image_io autosave in.dump FILE                convert_and_save_image
// Type names: u8 u16 u32 u64 i8 i16 i32 i64 f32 f64.  Dump header: "<type> <ndims> <extent0> ... \n" (Halide order: x first).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "HalideBuffer.h"
#include "halide_image_io.h"

using Halide::Runtime::Buffer;
namespace HT = Halide::Tools;

namespace {

const char *kNames[10] = {"u8", "u16", "u32", "u64", "i8", "i16", "i32", "i64", "f32", "f64"};

int type_index(const std::string &s) {
    for (int i = 0; i < 10; i++) {
        if (s == kNames[i]) return i;
    }
    fprintf(stderr, "unknown type %s\n", s.c_str());
    exit(2);
}

halide_type_t halide_type(int t) {
    switch (t) {
    case 0: return halide_type_of<uint8_t>();
    case 1: return halide_type_of<uint16_t>();
    case 2: return halide_type_of<uint32_t>();
    case 3: return halide_type_of<uint64_t>();
    case 4: return halide_type_of<int8_t>();
    case 5: return halide_type_of<int16_t>();
    case 6: return halide_type_of<int32_t>();
    case 7: return halide_type_of<int64_t>();
    case 8: return halide_type_of<float>();
    default: return halide_type_of<double>();
    }
}

int index_of(halide_type_t t) {
    for (int i = 0; i < 10; i++) {
        if (halide_type(i) == t) return i;
    }
    fprintf(stderr, "unsupported halide type\n");
    exit(2);
}

template<typename F>
auto dispatch(int t, F f) {
    switch (t) {
    case 0: return f(uint8_t());
    case 1: return f(uint16_t());
    case 2: return f(uint32_t());
    case 3: return f(uint64_t());
    case 4: return f(int8_t());
    case 5: return f(int16_t());
    case 6: return f(int32_t());
    case 7: return f(int64_t());
    case 8: return f(float());
    default: return f(double());
    }
}

std::vector<uint8_t> read_file(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    std::vector<uint8_t> v;
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) v.insert(v.end(), buf, buf + n);
    fclose(f);
    return v;
}

void write_dump(Buffer<> &im, const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) { perror(path); exit(2); }
    fprintf(f, "%s %d", kNames[index_of(im.type())], im.dimensions());
    for (int d = 0; d < im.dimensions(); d++) fprintf(f, " %d", im.dim(d).extent());
    fprintf(f, "\n");
    // planar payload, dimension 0 fastest
    std::vector<int> ext;
    for (int d = 0; d < im.dimensions(); d++) ext.push_back(im.dim(d).extent());
    Buffer<> copy(im.type(), ext);
    copy.copy_from(im);
    fwrite(copy.data(), 1, copy.size_in_bytes(), f);
    fclose(f);
}

Buffer<> read_dump(const char *path) {
    std::vector<uint8_t> v = read_file(path);
    size_t nl = 0;
    while (nl < v.size() && v[nl] != '\n') nl++;
    std::string hdr((const char *)v.data(), nl);
    char tname[16];
    int nd = 0, consumed = 0;
    if (sscanf(hdr.c_str(), "%15s %d%n", tname, &nd, &consumed) != 2) { fprintf(stderr, "bad dump header\n"); exit(2); }
    std::vector<int> ext;
    const char *p = hdr.c_str() + consumed;
    for (int d = 0; d < nd; d++) {
        int e = 0, c = 0;
        sscanf(p, "%d%n", &e, &c);
        ext.push_back(e);
        p += c;
    }
    Buffer<> im(halide_type(type_index(tname)), ext);
    if (v.size() - nl - 1 != im.size_in_bytes()) { fprintf(stderr, "dump payload size mismatch\n"); exit(2); }
    memcpy(im.data(), v.data() + nl + 1, im.size_in_bytes());
    return im;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const std::string cmd = argv[1];
    if (cmd == "convert" && argc == 6) {
        const int src = type_index(argv[2]), dst = type_index(argv[3]);
        std::vector<uint8_t> in = read_file(argv[4]);
        FILE *out = fopen(argv[5], "wb");
        dispatch(src, [&](auto s) {
            using S = decltype(s);
            const size_t n = in.size() / sizeof(S);
            const S *sp = (const S *)in.data();
            dispatch(dst, [&](auto d) {
                using D = decltype(d);
                std::vector<D> o(n);
                for (size_t i = 0; i < n; i++) o[i] = HT::Internal::convert<D, S>(sp[i]);
                fwrite(o.data(), sizeof(D), n, out);
                return 0;
            });
            return 0;
        });
        fclose(out);
        return 0;
    }
    if (cmd == "load" && argc == 4) {
        Buffer<> im;
        if (!HT::load<Buffer<>, HT::Internal::CheckReturn>(argv[2], &im)) { fprintf(stderr, "load failed\n"); return 1; }
        write_dump(im, argv[3]);
        return 0;
    }
    if (cmd == "loadconv" && argc == 5) {
        const int dst = type_index(argv[3]);
        return dispatch(dst, [&](auto d) {
            using D = decltype(d);
            Buffer<D> im = HT::load_and_convert_image(argv[2]);
            Buffer<> dyn = im;
            write_dump(dyn, argv[4]);
            return 0;
        });
    }
    if ((cmd == "save" || cmd == "autosave") && argc == 4) {
        Buffer<> im = read_dump(argv[2]);
        const bool autosave = cmd == "autosave";
        return dispatch(index_of(im.type()), [&](auto d) {
            using D = decltype(d);
            Buffer<D> typed = im.as<D>();
            if (autosave) {
                HT::convert_and_save_image<Buffer<D>, HT::Internal::CheckReturn>(typed, argv[3]);
            } else {
                if (!HT::save<Buffer<D>, HT::Internal::CheckReturn>(typed, argv[3])) { fprintf(stderr, "save failed\n"); return 1; }
            }
            return 0;
        });
    }
    fprintf(stderr, "usage: see the header comment\n");
    return 2;
}
