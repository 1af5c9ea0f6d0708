// local_laplacian.cu — host side of local_laplacian(input, levels, alpha, beta, output) for sm_100a:
// the C-ABI entry points (reference: apps/local_laplacian/local_laplacian_generator.cpp:12-16,287 and
// the harness apps/local_laplacian/process.cpp:31), argument validation, pyramid geometry, scratch
// allocation, the launch sequence, and the row-sharded multi-GPU variant.  Kernels: ll_kernels.cuh.
//
// Data layout in HBM (all f32, callee-owned scratch):
//   lut      [2*256*(levels-1)+1]            remap(i), i in [-256(levels-1), 256(levels-1)]
//   gp[j]    [sy_j][gpitch_j][K]  j=1..J-1   gPyramid[j], the K=levels planes interleaved per
//                                            pixel (K=8 -> one 32-byte sector per pixel, so the
//                                            data-dependent (li, li+1) plane pick costs one sector)
//   ing[j]   [sy_j][gpitch_j]     j=1..J-1   inGPyramid[j]
//   outg[j]  [oy_j][opitch_j]     j=1..J-1   outGPyramid[j]
// gray / gPyramid[0] / lPyramid / outLPyramid / outGPyramid[0] are never materialised: they are
// recomputed from the uint16 input where needed (8 f32 planes at full resolution would be
// 32 B/px of traffic against 12 B/px of compulsory I/O).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "hb_dist.h"
#include "ll_kernels.cuh"

namespace {

using namespace llk;

int g_shard_coarse_level = 0;  // halide_b200_ll_shard_coarse_level: 0 auto, -1 exchange level by level, n >= 2 gather level n
int g_force_naive = 0;  // test hook bitmask (halide_b200_ll_force_generic): 1 = generic down kernels, 2 = generic up, 4 = generic final, 8 = no fused coarse launch, 16 = general-layout final kernel, 32 = pair-column level-1 kernel

const hb::ArgSpec kIn = {"input", halide_type_uint, 16, 3, false};
const hb::ArgSpec kOut = {"output", halide_type_uint, 16, 3, true};

int64_t est_i[3][2] = {{0, 1536}, {0, 2560}, {0, 3}};
const int64_t *const est_ptrs[6] = {&est_i[0][0], &est_i[0][1], &est_i[1][0], &est_i[1][1], &est_i[2][0], &est_i[2][1]};
halide_scalar_value_t sv_levels, sv_alpha, sv_beta;
struct InitScalars {
    InitScalars() {
        sv_levels.u.i64 = 0; sv_levels.u.i32 = 8;
        sv_alpha.u.i64 = 0; sv_alpha.u.f32 = 1.0f;
        sv_beta.u.i64 = 0; sv_beta.u.f32 = 1.0f;
    }
} init_scalars;
// Argument records as the generator declares them (generator :12-16, estimates :92-99).
const halide_filter_argument_t kArgs[5] = {
    {"input", halide_argument_kind_input_buffer, 3, {halide_type_uint, 16, 0}, nullptr, nullptr, nullptr, nullptr, est_ptrs},
    {"levels", halide_argument_kind_input_scalar, 0, {halide_type_int, 32, 0}, nullptr, nullptr, nullptr, &sv_levels, nullptr},
    {"alpha", halide_argument_kind_input_scalar, 0, {halide_type_float, 32, 0}, nullptr, nullptr, nullptr, &sv_alpha, nullptr},
    {"beta", halide_argument_kind_input_scalar, 0, {halide_type_float, 32, 0}, nullptr, nullptr, nullptr, &sv_beta, nullptr},
    {"output", halide_argument_kind_output_buffer, 3, {halide_type_uint, 16, 0}, nullptr, nullptr, nullptr, nullptr, est_ptrs},
};
const halide_filter_metadata_t kMeta = {1, 5, kArgs, "x86-64-linux-cuda-cuda_capability_100-b200_native", "local_laplacian"};
const halide_filter_metadata_t kMetaAuto = {1, 5, kArgs, "x86-64-linux-cuda-cuda_capability_100-b200_native",
                                            "local_laplacian_auto_schedule"};


// Everything one call needs on the device: frame description + per-level buffers.
struct Plan {
    bool interior = false;  // row-sharded: this launch covers only rows whose taps stay inside the band (no halo rows, no clamp at the band edge)
    LLFrame f;
    LevelSet ls;
    ll::Geom geom;
    int J, K;
    float alpha;
    float *lut;
};

int check_and_query(halide_buffer_t *input, int levels, halide_buffer_t *output, bool *done) {
    int r;
    *done = true;
    if ((r = hb::check_arg(input, kIn))) return r;
    if ((r = hb::check_arg(output, kOut))) return r;
    // Bounds query: the only access that bypasses repeat_edge is input(x,y,c) in `color`
    // (generator :84), so the input must cover exactly the output region.
    bool query = false;
    {
        int mins[3] = {output->dim[0].min, output->dim[1].min, output->dim[2].min};
        int ext[3] = {output->dim[0].extent, output->dim[1].extent, output->dim[2].extent};
        if (hb::is_bounds_query(input)) {
            hb::propose_shape(input, mins, ext);
            query = true;
        }
        if (hb::is_bounds_query(output)) {
            hb::propose_shape(output, mins, ext);
            query = true;
        }
    }
    if (query) return 0;
    if ((r = hb::check_shape(input, kIn))) return r;
    if ((r = hb::check_shape(output, kOut))) return r;
    for (int d = 0; d < 3; d++) {
        if ((r = hb::check_covers(input, kIn, d, output->dim[d].min, output->dim[d].extent))) return r;
    }
    if (levels < 2 || levels > 32) {
        // 1/(levels-1) (generator :41) is meaningless below 2; the reference does not check, we do.
        return hb::fail(levels < 2 ? halide_error_code_param_too_small : halide_error_code_param_too_large,
                        "Parameter levels is %d but must be in [2, 32]", levels);
    }
    if (output->dim[0].extent <= 0 || output->dim[1].extent <= 0 || output->dim[2].extent <= 0) return 0;
    *done = false;
    return 0;
}

void fill_frame(Plan &p, halide_buffer_t *input, halide_buffer_t *output, void *din, void *dout, int levels, float alpha,
                float beta) {
    LLFrame &f = p.f;
    memset(&f.io, 0, sizeof(f.io));
    f.in = (const uint16_t *)din;
    f.in_sy = input->dim[1].stride; f.in_sc = input->dim[2].stride;
    f.in_x0 = input->dim[0].min; f.in_y0 = input->dim[1].min; f.in_c0 = input->dim[2].min;
    f.in_w = input->dim[0].extent; f.in_h = input->dim[1].extent; f.in_c = input->dim[2].extent;
    f.clamp_y0 = f.in_y0; f.clamp_h = f.in_h;
    f.halo_top = f.halo_bot = nullptr;
    f.halo_top_rows = f.halo_bot_rows = f.halo_pitch = 0;
    f.out = (uint16_t *)dout;
    f.out_sy = output->dim[1].stride; f.out_sc = output->dim[2].stride;
    f.out_x0 = output->dim[0].min; f.out_y0 = output->dim[1].min; f.out_c0 = output->dim[2].min;
    f.W = output->dim[0].extent; f.H = output->dim[1].extent; f.C = output->dim[2].extent;
    f.row0 = f.out_y0; f.nrows = f.H;
    f.levels = levels;
    f.beta = beta;
    f.flm1 = (float)(levels - 1);
    f.inv_lm1 = 1.0f / (float)(levels - 1);
    f.lut_half = 256 * (levels - 1);
    p.K = levels;
    p.J = ll::kMaxJ;
    p.alpha = alpha;
}

int alloc_levels(Plan &p, hb::Scratch &scratch) {
    p.lut = scratch.get<float>(2 * p.f.lut_half + 1);
    if (!p.lut) return hb::fail(halide_error_code_device_malloc_failed, "local_laplacian: scratch allocation failed");
    p.f.lut = p.lut;
    for (int j = 1; j < p.J; j++) {
        const ll::Level &lv = p.geom.lv[j];
        LevelBuf &b = p.ls.lv[j];
        b.sx = lv.sx; b.sy = lv.sy; b.ox = lv.ox; b.oy = lv.oy;
        b.cy = lv.cy; b.coy = lv.coy; b.gy = lv.gy;
        b.gpitch = lv.gpitch; b.opitch = lv.opitch;
        size_t gpix = (size_t)lv.sy.n() * lv.gpitch;
        b.gp = scratch.get<float>(gpix * p.K);
        b.ing = scratch.get<float>(gpix);
        b.outg = scratch.get<float>((size_t)lv.oy.n() * lv.opitch);
        if (!b.gp || !b.ing || !b.outg) {
            return hb::fail(halide_error_code_device_malloc_failed, "local_laplacian: scratch allocation failed");
        }
    }
    p.ls.lv[0] = p.ls.lv[1];  // level 0 is never stored; keep the slot initialised
    return 0;
}

// ---- launch helpers ------------------------------------------------------------------------------------
const dim3 kBlk(32, 8);
dim3 grid_for(int w, int h) { return dim3((w + 31) / 32, (h + 7) / 8); }

// Strip kernels run with one block per resident slot and a balanced static partition of the level's rows
// (ll_down_strip_kernel), so there is no tail wave.
template<typename Kern>
int strip_slots(Kern kern, size_t smem) {
    int dev = 0, sms = 148, per_sm = 8;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 4;
    }
    return sms * per_sm;
}
int strip_xblocks(const LevelBuf &d) {
    return ((d.sx.n() + kStripCols - 1) / kStripCols + 3) / 4;
}
int strip_grid(const LevelBuf &d, int slots) {
    // never more blocks than block-rows of work; at least 2 destination rows per block where possible
    long long work = (long long)strip_xblocks(d) * d.cy.n();
    long long g = work / 2 < 1 ? 1 : work / 2;
    return (int)(g < slots ? g : slots);
}

void launch_lut(Plan &p, cudaStream_t s) {
    HB_LAUNCH("ll_lut", ll_lut_kernel, (2 * p.f.lut_half + 1 + 255) / 256, 256, 0, s, p.lut, p.f.lut_half, p.alpha);
}

void launch_down(Plan &p, int j, cudaStream_t s) {  // produce level j (j >= 1) rows cy
    LevelBuf *lb = p.ls.lv;
    if (lb[j].cy.n() <= 0) return;
    const bool fast = (p.K == 8) && !(g_force_naive & 1);
    if (j == 1) {
        if (fast) {
            size_t smem = (size_t)(2 * p.f.lut_half + 1) * sizeof(float);
            const int xb = strip_xblocks(lb[1]);
            const bool sharded_rows = p.f.halo_top_rows || p.f.halo_bot_rows || p.f.clamp_y0 != p.f.in_y0 || p.f.clamp_h != p.f.in_h;
            if ((g_force_naive & 32) && !sharded_rows && p.f.in_sy > 0 && p.f.in_sc > 0 &&
                (int64_t)p.f.in_h * p.f.in_sy + 3 * p.f.in_sc < (1ll << 31)) {
                // alternative two-columns-per-lane variant (see ll_level1_pair_kernel: bit-exact, measured slightly slower)
                const LLFrame &f = p.f;
                const int wide = ((uintptr_t)f.in & 3) == 0 && (f.in_sy & 1) == 0 && (f.in_sc & 1) == 0 && (f.in_x0 & 1) == 0;
                const int xbp = ((lb[1].sx.n() + kPairCols - 1) / kPairCols + 3) / 4;
                static int slots = strip_slots(ll_level1_pair_kernel<8>, 16 * 1024);
                long long work = (long long)xbp * lb[1].cy.n();
                long long g = work / 2 < 1 ? 1 : work / 2;
                HB_LAUNCH("ll_level1_pair", (ll_level1_pair_kernel<8>), (int)(g < slots ? g : slots), 128, smem, s, p.f, lb[1], xbp, wide);
            } else if (!p.interior && sharded_rows) {
                static int slots = strip_slots(ll_down_strip_kernel<8, true, true>, 16 * 1024);
                HB_LAUNCH("ll_level1_strip", (ll_down_strip_kernel<8, true, true>), strip_grid(lb[1], slots), 128, smem, s, p.f,
                          lb[1], lb[1], xb);
            } else {
                static int slots = strip_slots(ll_down_strip_kernel<8, true, false>, 16 * 1024);
                HB_LAUNCH("ll_level1_strip", (ll_down_strip_kernel<8, true, false>), strip_grid(lb[1], slots), 128, smem, s, p.f,
                          lb[1], lb[1], xb);
            }
        } else {
            HB_LAUNCH("ll_level1", ll_level1_naive_kernel, grid_for(lb[1].sx.n(), lb[1].cy.n()), kBlk, 0, s, p.f, lb[1]);
        }
    } else if (fast) {
        const bool peer = p.f.io.up_flag || p.f.io.dn_flag || p.f.io.wait_up[0] || p.f.io.wait_dn[0];
        if (peer) {
            static int slots = strip_slots(ll_down_strip_kernel<8, false, true>, 0);
            HB_LAUNCH("ll_down_strip", (ll_down_strip_kernel<8, false, true>), strip_grid(lb[j], slots), 128, 0, s, p.f, lb[j - 1],
                      lb[j], strip_xblocks(lb[j]));
        } else {
            static int slots = strip_slots(ll_down_strip_kernel<8, false, false>, 0);
            HB_LAUNCH("ll_down_strip", (ll_down_strip_kernel<8, false, false>), strip_grid(lb[j], slots), 128, 0, s, p.f, lb[j - 1],
                      lb[j], strip_xblocks(lb[j]));
        }
    } else {
        HB_LAUNCH("ll_down", ll_down_naive_kernel, grid_for(lb[j].sx.n(), lb[j].cy.n()), kBlk, 0, s, lb[j - 1], lb[j], p.K);
    }
}

void launch_up(Plan &p, int j, cudaStream_t s) {  // produce outGPyramid[j] (1 <= j <= J-1) rows coy
    LevelBuf *lb = p.ls.lv;
    if (lb[j].coy.n() <= 0) return;
    const bool fast_up = (p.K == 8) && !(g_force_naive & 2);
    if (fast_up && j < p.J - 1) {
        dim3 g((lb[j].ox.n() + kUpTW - 1) / kUpTW, (lb[j].coy.n() + kUpTH - 1) / kUpTH);
        if (p.f.io.up_flag || p.f.io.dn_flag || p.f.io.wait_up[0] || p.f.io.wait_dn[0]) {
            HB_LAUNCH("ll_up_tile", (ll_up_tile_kernel<false, true>), g, 256, 0, s, p.f, lb[j], lb[j + 1]);
        } else {
            HB_LAUNCH("ll_up_tile", (ll_up_tile_kernel<false, false>), g, 256, 0, s, p.f, lb[j], lb[j + 1]);
        }
    } else {
        HB_LAUNCH("ll_up", ll_up_naive_kernel, grid_for(lb[j].ox.n(), lb[j].coy.n()), kBlk, 0, s, lb[j],
                  lb[j == p.J - 1 ? j : j + 1], p.K, p.f.flm1, p.f.levels, j == p.J - 1 ? 1 : 0, p.f.io);
    }
}

void launch_final(Plan &p, cudaStream_t s) {
    LevelBuf *lb = p.ls.lv;
    if (p.f.nrows <= 0) return;
    if (p.J > 1 && p.K == 8 && !(g_force_naive & 4) && p.f.C <= 3) {
        dim3 g((p.f.W + kUpTW - 1) / kUpTW, (p.f.nrows + kUpTH - 1) / kUpTH);
        size_t smem = 513 * sizeof(float);
        // the common layout takes the kernel's SIMPLE path (32-bit addressing, one aligned word per thread and channel)
        const LLFrame &f = p.f;
        const int64_t in_span = (int64_t)f.in_h * f.in_sy + 3 * f.in_sc, out_span = (int64_t)f.H * f.out_sy + 3 * f.out_sc;
        const bool simple = f.C == 3 && f.in_c0 == 0 && f.out_c0 == 0 && f.in_c >= 3 && (f.W & 1) == 0 &&
                            ((f.out_x0 - f.in_x0) & 1) == 0 && ((uintptr_t)f.in & 3) == 0 && ((uintptr_t)f.out & 3) == 0 &&
                            (f.in_sy & 1) == 0 && (f.in_sc & 1) == 0 && (f.out_sy & 1) == 0 && (f.out_sc & 1) == 0 &&
                            f.in_sy > 0 && f.in_sc > 0 && f.out_sy > 0 && f.out_sc > 0 && in_span < (1ll << 31) &&
                            out_span < (1ll << 31) && !(g_force_naive & 16);
        const bool peer = p.f.io.wait_up[0] || p.f.io.wait_dn[0];
        if (peer && simple) {
            HB_LAUNCH("ll_final_tile", (ll_up_tile_kernel<true, true, true>), g, 256, smem, s, p.f, lb[1], lb[1]);
        } else if (peer) {
            HB_LAUNCH("ll_final_tile", (ll_up_tile_kernel<true, true, false>), g, 256, smem, s, p.f, lb[1], lb[1]);
        } else if (simple) {
            HB_LAUNCH("ll_final_tile", (ll_up_tile_kernel<true, false, true>), g, 256, smem, s, p.f, lb[1], lb[1]);
        } else {
            HB_LAUNCH("ll_final_tile", (ll_up_tile_kernel<true, false, false>), g, 256, smem, s, p.f, lb[1], lb[1]);
        }
    } else {
        HB_LAUNCH("ll_final", ll_final_naive_kernel, grid_for(p.f.W, p.f.nrows), kBlk, 0, s, p.f, lb[1], p.J > 1 ? 1 : 0);
    }
}

// Coarse tail in one cooperative launch: levels j0+1 .. J-1 down and up (see ll_coarse_fused_kernel).
bool launch_coarse_fused(Plan &p, int j0, cudaStream_t s) {
    static int max_blocks = -1;
    if (max_blocks < 0) {
        int dev = 0, sms = 0, coop = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ll_coarse_fused_kernel, 256, 0);
        max_blocks = coop ? sms * (per_sm > 2 ? 2 : per_sm) : 0;
    }
    if (max_blocks <= 0) return false;
    int J = p.J, K = p.K, levels = p.f.levels;
    float flm1 = p.f.flm1;
    void *args[] = {&p.ls, &J, &j0, &K, &flm1, &levels};
    hb::count_launch("ll_coarse_fused", s);
    cudaError_t e = cudaLaunchCooperativeKernel((void *)ll_coarse_fused_kernel, dim3(max_blocks), dim3(256), args, 0, s);
    hb::after_launch(s);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return true;
}

int run_local_laplacian(halide_buffer_t *input, int levels, float alpha, float beta, halide_buffer_t *output) {
    int r;
    bool done;
    if ((r = check_and_query(input, levels, output, &done)) || done) return r;
    void *din = nullptr, *dout = nullptr;
    if ((r = hb::acquire_input(input, kIn, &din))) return r;
    if ((r = hb::acquire_output(output, kOut, &dout))) return r;

    Plan p;
    fill_frame(p, input, output, din, dout, levels, alpha, beta);
    const int W = p.f.W, H = p.f.H;
    Span outx = {p.f.out_x0, p.f.out_x0 + W - 1}, outy = {p.f.out_y0, p.f.out_y0 + H - 1};
    Span inx = {p.f.in_x0, p.f.in_x0 + p.f.in_w - 1}, iny = {p.f.in_y0, p.f.in_y0 + p.f.in_h - 1};
    p.geom = ll::make_geom(outx, outy, inx, iny, p.J);
    hb::Scratch scratch;
    if ((r = alloc_levels(p, scratch))) return r;

    cudaStream_t s = hb::stream();
    {
        hb::CallTimer timer(s);
        launch_lut(p, s);
        // Levels whose pixel count is small are launch-latency bound: they run in one cooperative kernel.
        int j0 = p.J - 1;
        if (!(g_force_naive & 8)) {
            while (j0 > 1 && (int64_t)p.ls.lv[j0].sx.n() * p.ls.lv[j0].sy.n() <= 40 * 1024) j0--;
        }
        // j0 = last level produced by its own launch; levels j0+1.. are fused (if any)
        bool fused = false;
        for (int j = 1; j <= j0; j++) launch_down(p, j, s);
        if (j0 < p.J - 1) fused = launch_coarse_fused(p, j0, s);
        if (!fused) {
            for (int j = j0 + 1; j < p.J; j++) launch_down(p, j, s);
            for (int j = p.J - 1; j > j0; j--) launch_up(p, j, s);
        }
        for (int j = j0; j >= 1; j--) launch_up(p, j, s);
        launch_final(p, s);
    }
    if ((r = hb::check_cuda(cudaGetLastError(), "local_laplacian launch", halide_error_code_device_run_failed))) return r;
    hb::mark_output_written(output);
    return 0;
}

// ---- peer-memory plan for the row-sharded variant ---------------------------------------------------------
// All level buffers of a rank live in one cudaMalloc'ed slab whose CUDA-IPC handle and internal layout are
// all-gathered once per geometry; afterwards every halo exchange is one kernel that stores this rank's boundary rows
// straight into the neighbours' slabs over NVLink and handshakes through flags (hb_dist.h: PeerXchg).
struct SlabLayout {  // POD, exchanged between ranks
    unsigned long long gp[ll::kMaxJ], ing[ll::kMaxJ], outg[ll::kMaxJ], halo_top, halo_bot, flags, lut, total;
    int sy_lo[ll::kMaxJ], oy_lo[ll::kMaxJ];
    int halo_top_rows, halo_bot_rows;
    cudaIpcMemHandle_t handle;
};
constexpr int kMaxPeers = 8;  // flag slots per rank (one NVSwitch domain)
struct ShardPlan {
    int key[12];
    bool valid = false;
    char *slab = nullptr;
    SlabLayout mine, up, dn;
    char *up_base = nullptr, *dn_base = nullptr;
    std::vector<SlabLayout> all;  // every rank's layout, rank order
    std::vector<char *> base;     // every rank's slab mapped here (null for this rank and for unmapped ranks)
    unsigned epoch = 0;
    unsigned *host_error = nullptr, *dev_error = nullptr;  // mapped pinned: set by a timed-out wait
};
ShardPlan g_shard;

void destroy_shard_plan() {
    if (!g_shard.valid) return;
    cudaDeviceSynchronize();
    for (char *b : g_shard.base) {
        if (b) cudaIpcCloseMemHandle(b);
    }
    if (g_shard.slab) cudaFree(g_shard.slab);
    if (g_shard.host_error) cudaFreeHost(g_shard.host_error);
    g_shard = ShardPlan();
}

// Collective: every rank must call it with its own geometry at the same point of the program.
int build_shard_plan(const Plan &p, const int *key, bool first, bool last, bool map_all) {
    destroy_shard_plan();
    ShardPlan &sp = g_shard;
    SlabLayout &L = sp.mine;
    memset(&L, 0, sizeof(L));
    unsigned long long off = 0;
    auto take = [&](unsigned long long bytes) {
        unsigned long long o = off;
        off += (bytes + 255) & ~255ull;
        return o;
    };
    // words 0..31: [step*2 + dir] halo epochs; 32..39: gather epochs by source rank; 40..47: ready epochs by source
    // rank; 48: the done counter
    L.flags = take(256);
    L.lut = take((2ull * p.f.lut_half + 1) * sizeof(float));
    L.halo_top_rows = first ? 0 : 1;
    L.halo_bot_rows = last ? 0 : 2;
    L.halo_top = take((unsigned long long)p.f.in_c * 1 * p.f.in_w * sizeof(uint16_t));
    L.halo_bot = take((unsigned long long)p.f.in_c * 2 * p.f.in_w * sizeof(uint16_t));
    for (int j = 1; j < p.J; j++) {
        const ll::Level &lv = p.geom.lv[j];
        unsigned long long gpix = (unsigned long long)lv.sy.n() * lv.gpitch;
        L.gp[j] = take(gpix * p.K * sizeof(float));
        L.ing[j] = take(gpix * sizeof(float));
        L.outg[j] = take((unsigned long long)lv.oy.n() * lv.opitch * sizeof(float));
        L.sy_lo[j] = lv.sy.lo;
        L.oy_lo[j] = lv.oy.lo;
    }
    L.total = off;
    if (cudaMalloc((void **)&sp.slab, L.total) != cudaSuccess) {
        cudaGetLastError();
        return hb::fail(halide_error_code_device_malloc_failed, "local_laplacian_sharded: slab allocation of %llu bytes failed", L.total);
    }
    cudaMemset(sp.slab + L.flags, 0, 256);
    if (cudaIpcGetMemHandle(&L.handle, sp.slab) != cudaSuccess) {
        cudaGetLastError();
        return hb::fail(halide_error_code_generic_error, "local_laplacian_sharded: cudaIpcGetMemHandle failed");
    }
    if (cudaHostAlloc((void **)&sp.host_error, sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&sp.dev_error, sp.host_error, 0) != cudaSuccess) {
        cudaGetLastError();
        return hb::fail(halide_error_code_generic_error, "local_laplacian_sharded: mapped error flag allocation failed");
    }
    *sp.host_error = 0;
    cudaDeviceSynchronize();
    const int n = hbdist::size(), me = hbdist::rank();
    sp.all.assign(n, SlabLayout());
    int r = hbdist::allgather_bytes(&L, sp.all.data(), sizeof(SlabLayout));
    if (r) return r;
    sp.base.assign(n, nullptr);
    for (int q = 0; q < n; q++) {
        // neighbours always; everybody when a level is gathered all-to-all (map_all)
        if (q == me || !(map_all || q == me - 1 || q == me + 1)) continue;
        if (cudaIpcOpenMemHandle((void **)&sp.base[q], sp.all[q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            sp.base[q] = nullptr;
            return hb::fail(halide_error_code_generic_error, "local_laplacian_sharded: cannot map rank %d's slab (no peer access?)", q);
        }
    }
    if (!first) {
        sp.up = sp.all[me - 1];
        sp.up_base = sp.base[me - 1];
    }
    if (!last) {
        sp.dn = sp.all[me + 1];
        sp.dn_base = sp.base[me + 1];
    }
    memcpy(sp.key, key, sizeof(sp.key));
    sp.valid = true;
    return 0;
}

void bind_slab(Plan &p) {  // point the plan's buffers into the slab
    const ShardPlan &sp = g_shard;
    p.lut = (float *)(sp.slab + sp.mine.lut);
    p.f.lut = p.lut;
    for (int j = 1; j < p.J; j++) {
        const ll::Level &lv = p.geom.lv[j];
        LevelBuf &b = p.ls.lv[j];
        b.sx = lv.sx; b.sy = lv.sy; b.ox = lv.ox; b.oy = lv.oy;
        b.cy = lv.cy; b.coy = lv.coy; b.gy = lv.gy;
        b.gpitch = lv.gpitch; b.opitch = lv.opitch;
        b.gp = (float *)(sp.slab + sp.mine.gp[j]);
        b.ing = (float *)(sp.slab + sp.mine.ing[j]);
        b.outg = (float *)(sp.slab + sp.mine.outg[j]);
    }
    p.ls.lv[0] = p.ls.lv[1];
}

// Row-sharded path: the level gathered all-to-all (levels above it are replicated on every rank), or J for none.
// Depends only on the whole frame's geometry and the rank count, so every rank computes the same value.
int choose_coarse_level(const ll::Geom &whole, int nranks, int J, int K) {
    if (nranks < 2 || nranks > kMaxPeers || g_shard_coarse_level < 0 || J < 4) return J;
    if (g_shard_coarse_level > 0) return g_shard_coarse_level < 2 ? 2 : (g_shard_coarse_level > J - 1 ? J - 1 : g_shard_coarse_level);
    for (int j = 2; j < J - 1; j++) {
        const unsigned long long rows = (unsigned long long)(whole.lv[j].sy.n() + nranks - 1) / nranks;
        // measured at N=2 (4K band per GPU): gathering level 3 (4.7 MB per peer) 312-326 us, level 4 (1.2 MB) 315-316 us,
        // level 5 359 us: one more level in the sharded sweeps costs ~40 us, the bytes hardly matter at this size
        if (rows * whole.lv[j].gpitch * (K + 1) * sizeof(float) <= (5ull << 18)) return j;  // <= 1.25 MiB per peer
    }
    return J - 1;
}

// ---- row-sharded variant (one process per GPU) -----------------------------------------------------------
// `input`/`output` describe this rank's band: all columns and channels of the frame, rows
// [dim[1].min, dim[1].min + extent) in the frame's coordinates.  frame_y_min/extent give the rows of the whole
// frame; ranks must be ordered top to bottom (rank r-1 holds the rows directly above rank r's).
// One halo exchange per pyramid level in each sweep plus one for the input rows.
int run_local_laplacian_sharded(halide_buffer_t *input, int levels, float alpha, float beta, halide_buffer_t *output,
                                int frame_y_min, int frame_y_extent) {
    int r;
    bool done;
    if ((r = check_and_query(input, levels, output, &done)) || done) return r;
    if (!hbdist::active()) {
        return hb::fail(halide_error_code_generic_error, "local_laplacian_sharded: call halide_b200_dist_init first");
    }
    if (levels != 8) {
        return hb::fail(halide_error_code_unimplemented, "local_laplacian_sharded: only levels == 8 is built for sharding");
    }
    const int rank = hbdist::rank(), nranks = hbdist::size();
    Span band = {output->dim[1].min, output->dim[1].min + output->dim[1].extent - 1};
    if (input->dim[1].min != band.lo || input->dim[1].extent != band.n()) {
        return hb::fail(halide_error_code_constraint_violated,
                        "local_laplacian_sharded: input rows [%d,%d) must equal the output band [%d,%d)", input->dim[1].min,
                        input->dim[1].min + input->dim[1].extent, band.lo, band.hi + 1);
    }
    const bool first = rank == 0, last = rank == nranks - 1;
    Span frame_y = {frame_y_min, frame_y_min + frame_y_extent - 1};
    if ((first && band.lo != frame_y.lo) || (last && band.hi != frame_y.hi) || band.lo < frame_y.lo || band.hi > frame_y.hi) {
        return hb::fail(halide_error_code_constraint_violated, "local_laplacian_sharded: band [%d,%d] inconsistent with frame rows [%d,%d] on rank %d/%d",
                        band.lo, band.hi, frame_y.lo, frame_y.hi, rank, nranks);
    }
    void *din = nullptr, *dout = nullptr;
    if ((r = hb::acquire_input(input, kIn, &din))) return r;
    if ((r = hb::acquire_output(output, kOut, &dout))) return r;

    Plan p;
    fill_frame(p, input, output, din, dout, levels, alpha, beta);
    const int W = p.f.W;
    Span outx = {p.f.out_x0, p.f.out_x0 + W - 1};
    Span inx = {p.f.in_x0, p.f.in_x0 + p.f.in_w - 1};
    ll::Geom whole = ll::make_geom(outx, frame_y, inx, frame_y, p.J);
    ll::BandLevel bl[ll::kMaxJ];
    ll::compute_band_y(whole, band, first, last, bl);
    static const bool use_peer = [] {
        const char *e = getenv("HALIDE_B200_HALO");
        return !(e && strcmp(e, "nccl") == 0);
    }();
    // Coarse replication: levels jr.. are tiny, so exchanging their halos level by level costs one NVLink flag round
    // trip per level and sweep with nothing to hide it behind.  Instead every rank's band of level jr is gathered
    // all-to-all once, and levels jr+1.. (down) and ..jr (up) are computed redundantly for the whole frame on every
    // rank.  jr = the first level whose band is at most 1.25 MiB (what each peer receives); it depends only on the frame and the rank
    // count, so all ranks agree.  jr == J: no replication (level-by-level exchange).
    const int jr = use_peer ? choose_coarse_level(whole, nranks, p.J, p.K) : p.J;
    for (int j = 1; j < p.J && j <= jr; j++) {
        if (bl[j].own.n() < 2 || bl[j].own_o.n() < 1) {
            return hb::fail(halide_error_code_constraint_violated,
                            "local_laplacian_sharded: band of %d rows is too small for %d pyramid levels (level %d owns %d rows)",
                            band.n(), p.J, j, bl[j].own.n());
        }
    }
    p.geom = ll::make_band_geom(whole, bl);
    if (jr < p.J) {
        // levels jr.. are held for the whole frame; level jr's rows are produced band by band (cy) and gathered
        for (int j = jr; j < p.J; j++) p.geom.lv[j] = whole.lv[j];
        p.geom.lv[jr].cy = bl[jr].own;
    }
    p.f.clamp_y0 = frame_y.lo;
    p.f.clamp_h = frame_y.n();
    cudaStream_t s = hb::stream();
    const int up = rank - 1, dn = rank + 1;
    const int C = p.f.in_c;
    if (use_peer) {
        // ---- peer-memory path: slab + IPC plan (built collectively on first use / geometry change) ----
        const int key[12] = {W, p.f.in_w, band.lo, band.hi, frame_y.lo, frame_y.hi, rank, nranks, C, (int)(p.f.in_sy & 0x7fffffff), jr, 0};
        if (!g_shard.valid || memcmp(g_shard.key, key, sizeof(key)) != 0) {
            if ((r = build_shard_plan(p, key, first, last, jr < p.J))) return r;
        }
        ShardPlan &sp = g_shard;
        if (*sp.host_error) {
            return hb::fail(halide_error_code_device_run_failed, "local_laplacian_sharded: a previous halo wait timed out (neighbour rank stalled?)");
        }
        bind_slab(p);
        p.f.halo_pitch = p.f.in_w;
        p.f.halo_top = (const uint16_t *)(sp.slab + sp.mine.halo_top);
        p.f.halo_bot = (const uint16_t *)(sp.slab + sp.mine.halo_bot);
        p.f.halo_top_rows = sp.mine.halo_top_rows;
        p.f.halo_bot_rows = sp.mine.halo_bot_rows;
        sp.epoch++;
        unsigned *flags = (unsigned *)(sp.slab + sp.mine.flags);
        int step = 0;
        hbdist::PeerXchg x;
        auto begin_step = [&]() {
            memset(&x, 0, sizeof(x));
            x.epoch = sp.epoch;
            x.done_counter = flags + 48;
            x.error_flag = sp.dev_error;
            // I am the DOWN neighbour of rank-1 (its slot dir 1) and the UP neighbour of rank+1 (its slot dir 0)
            x.peer_flag[0] = first ? nullptr : (unsigned *)(sp.up_base + sp.up.flags) + step * 2 + 1;
            x.peer_flag[1] = last ? nullptr : (unsigned *)(sp.dn_base + sp.dn.flags) + step * 2 + 0;
            x.my_flag[0] = first ? nullptr : flags + step * 2 + 0;
            x.my_flag[1] = last ? nullptr : flags + step * 2 + 1;
        };
        auto add_seg = [&](const void *src, void *dst, size_t bytes, unsigned elem) {
            x.seg[x.nseg].src = src; x.seg[x.nseg].dst = dst;
            x.seg[x.nseg].bytes = (unsigned)bytes; x.seg[x.nseg].elem = elem;
            x.nseg++;
        };
        auto end_step = [&]() {
            hbdist::launch_peer_exchange(x, s);
            step++;
        };
        {
            hb::CallTimer timer(s);
            // step 0: input rows (per channel; rows may be strided in the caller's buffer -> one segment per row)
            begin_step();
            const size_t rb = (size_t)p.f.in_w * sizeof(uint16_t);
            for (int c = 0; c < C; c++) {
                const uint16_t *plane = (const uint16_t *)din + (int64_t)c * p.f.in_sc;
                if (!first) {  // my rows 0,1 are the up neighbour's two bottom-halo rows
                    char *dst = sp.up_base + sp.up.halo_bot + (size_t)c * 2 * rb;
                    add_seg(plane, dst, rb, 2);
                    add_seg(plane + p.f.in_sy, dst + rb, rb, 2);
                }
                if (!last) {  // my last row is the down neighbour's top-halo row
                    add_seg(plane + (int64_t)(p.f.in_h - 1) * p.f.in_sy, sp.dn_base + sp.dn.halo_top + (size_t)c * rb, rb, 2);
                }
            }
            // push only: the level-1 kernel's boundary blocks acquire the step-0 flags themselves (PeerIO)
            x.my_flag[0] = x.my_flag[1] = nullptr;
            if (jr < p.J) {
                // this kernel runs after everything of the previous call: tell every rank its level-jr rows may be overwritten
                for (int q = 0; q < nranks; q++) {
                    if (q != rank) x.ready_flag[x.nready++] = (unsigned *)(sp.base[q] + sp.all[q].flags) + 40 + rank;
                }
            }
            end_step();
            launch_lut(p, s);
            LevelBuf *lb = p.ls.lv;
            // From here on there are no exchange kernels: producers mirror their boundary rows into the neighbours'
            // slabs and release the step's flag, consumers acquire the flags of the halo rows they read (PeerIO).
            // Flag slots: down-sweep level j = step j; up-sweep level j = step 15 - j.
            // A sharded sweep step is one launch of the sharded kernel variant over the band's rows: the blocks that hold the
            // band's first / last rows acquire the neighbour's flag of the previous step before reading its halo rows, mirror the
            // boundary rows they produce into the neighbour's slab, and the launch's last block releases the neighbour's flag of
            // this step; interior blocks never wait.  (HALIDE_B200_SHARD_SPLIT=1 issues the interior rows with the single-GPU
            // kernel variant and the edge rows as separate small launches instead — measured slower: the edge launches
            // serialise behind the interior one.)
            static const bool split = [] {
                const char *e = getenv("HALIDE_B200_SHARD_SPLIT");
                return e && e[0] == '1';
            }();
            auto io_begin = [&]() {
                memset(&p.f.io, 0, sizeof(p.f.io));
                p.f.io.epoch = sp.epoch;
                p.f.io.error_flag = sp.dev_error;
                p.f.io.done_counter = flags + 48;
            };
            auto io_clear = [&]() { memset(&p.f.io, 0, sizeof(p.f.io)); };
            // down-sweep level j: rows cy of lb[j] from level j-1 (step j-1 halos); `produce`: mirror to the neighbours (step j)
            auto down_level = [&](int j, bool produce) {
                const LevelBuf full = lb[j];
                const Span cy = full.cy;
                const size_t rb_a = (size_t)full.gpitch * p.K * sizeof(float), rb_b = (size_t)full.gpitch * sizeof(float);
                auto io_top = [&]() {
                    p.f.io.wait_up[0] = flags + (j - 1) * 2 + 0;
                    if (produce) {
                        p.f.io.up_a = sp.up_base + sp.up.gp[j] + (size_t)(cy.lo - sp.up.sy_lo[j]) * rb_a;
                        p.f.io.up_b = sp.up_base + sp.up.ing[j] + (size_t)(cy.lo - sp.up.sy_lo[j]) * rb_b;
                        p.f.io.up_flag = (unsigned *)(sp.up_base + sp.up.flags) + j * 2 + 1;
                    }
                };
                auto io_bot = [&]() {
                    p.f.io.wait_dn[0] = flags + (j - 1) * 2 + 1;
                    if (produce) {
                        p.f.io.dn_a = sp.dn_base + sp.dn.gp[j] + (size_t)(cy.hi - sp.dn.sy_lo[j]) * rb_a;
                        p.f.io.dn_b = sp.dn_base + sp.dn.ing[j] + (size_t)(cy.hi - sp.dn.sy_lo[j]) * rb_b;
                        p.f.io.dn_flag = (unsigned *)(sp.dn_base + sp.dn.flags) + j * 2 + 0;
                    }
                };
                if (!split) {
                    io_begin();
                    if (!first) io_top();
                    if (!last) io_bot();
                    launch_down(p, j, s);
                    io_clear();
                    return;
                }
                const int top_n = first ? 0 : (cy.n() < 2 ? cy.n() : 2);    // rows cy.lo, cy.lo+1: read 1 halo row, mirrored up
                const int bot_n = last ? 0 : (cy.n() - top_n < 1 ? 0 : 1);  // row cy.hi: reads 2 halo rows, mirrored down
                io_clear();
                p.interior = true;
                lb[j].cy = {cy.lo + top_n, cy.hi - bot_n};
                launch_down(p, j, s);
                p.interior = false;
                if (top_n) {
                    io_begin();
                    io_top();
                    lb[j].cy = {cy.lo, cy.lo + top_n - 1};
                    launch_down(p, j, s);
                }
                if (bot_n) {
                    io_begin();
                    io_bot();
                    lb[j].cy = {cy.hi, cy.hi};
                    launch_down(p, j, s);
                }
                lb[j] = full;
                io_clear();
            };
            // up-sweep level j (j >= 1) or the final kernel (j == 0): `wait`: the coarse level's halo rows come from the
            // neighbours (steps j+1 and 15-(j+1)); level j >= 1 mirrors its first / last row of outGPyramid[j] (step 15-j)
            auto up_level = [&](int j, bool wait) {
                const LevelBuf full = lb[j];
                const Span rows = j ? full.coy : Span{p.f.out_y0, p.f.out_y0 + p.f.H - 1};
                const size_t rb = (size_t)full.opitch * sizeof(float);
                auto launch_rows = [&](Span r) {
                    if (j) {
                        lb[j].coy = r;
                        launch_up(p, j, s);
                    } else {
                        p.f.row0 = r.lo;
                        p.f.nrows = r.n();
                        launch_final(p, s);
                    }
                };
                auto io_top = [&]() {
                    if (wait) {
                        p.f.io.wait_up[0] = flags + (j + 1) * 2 + 0;
                        p.f.io.wait_up[1] = flags + (15 - (j + 1)) * 2 + 0;
                    }
                    if (j) {
                        p.f.io.up_a = sp.up_base + sp.up.outg[j] + (size_t)(rows.lo - sp.up.oy_lo[j]) * rb;
                        p.f.io.up_flag = (unsigned *)(sp.up_base + sp.up.flags) + (15 - j) * 2 + 1;
                    }
                };
                auto io_bot = [&]() {
                    if (wait) {
                        p.f.io.wait_dn[0] = flags + (j + 1) * 2 + 1;
                        p.f.io.wait_dn[1] = flags + (15 - (j + 1)) * 2 + 1;
                    }
                    if (j) {
                        p.f.io.dn_a = sp.dn_base + sp.dn.outg[j] + (size_t)(rows.hi - sp.dn.oy_lo[j]) * rb;
                        p.f.io.dn_flag = (unsigned *)(sp.dn_base + sp.dn.flags) + (15 - j) * 2 + 0;
                    }
                };
                if (!split) {
                    io_begin();
                    if (!first) io_top();
                    if (!last) io_bot();
                    launch_rows(rows);
                } else {
                    const int top_n = first ? 0 : (rows.n() < kUpTH ? rows.n() : kUpTH);
                    const int bot_n = last ? 0 : (rows.n() - top_n < kUpTH ? rows.n() - top_n : kUpTH);
                    io_clear();
                    launch_rows({rows.lo + top_n, rows.hi - bot_n});
                    if (top_n) {
                        io_begin();
                        io_top();
                        launch_rows({rows.lo, rows.lo + top_n - 1});
                    }
                    if (bot_n) {
                        io_begin();
                        io_bot();
                        launch_rows({rows.hi - bot_n + 1, rows.hi});
                    }
                }
                lb[j] = full;
                p.f.row0 = p.f.out_y0;
                p.f.nrows = p.f.H;
                io_clear();
            };
            const int jd = jr < p.J ? jr : p.J - 1;  // last level of the sharded down sweep
            for (int j = 1; j <= jd; j++) down_level(j, j < jr);  // (level jr is gathered all-to-all instead of mirrored)
            int ju = p.J - 1;  // first level of the sharded up sweep
            if (jr < p.J) {
                // ---- gather level jr, then the coarse tail for the whole frame on every rank (no communication) ----
                hbdist::PeerGather g;
                memset(&g, 0, sizeof(g));
                g.epoch = sp.epoch;
                g.done_counter = flags + 48;
                g.error_flag = sp.dev_error;
                const size_t row_off = (size_t)(lb[jr].cy.lo - lb[jr].sy.lo) * lb[jr].gpitch;  // same rows, same pitch on every rank
                const size_t npx = (size_t)lb[jr].cy.n() * lb[jr].gpitch;
                for (int q = 0; q < nranks; q++) {
                    if (q == rank) continue;
                    char *qb = sp.base[q];
                    g.seg[g.nseg++] = {lb[jr].gp + row_off * p.K, qb + sp.all[q].gp[jr] + row_off * p.K * sizeof(float),
                                       (unsigned)(npx * p.K * sizeof(float)), 16};
                    g.seg[g.nseg++] = {lb[jr].ing + row_off, qb + sp.all[q].ing[jr] + row_off * sizeof(float),
                                       (unsigned)(npx * sizeof(float)), 16};
                    g.peer_flag[g.npeer] = (unsigned *)(qb + sp.all[q].flags) + 32 + rank;
                    g.my_flag[g.npeer] = flags + 32 + q;
                    g.ready[g.npeer] = flags + 40 + q;
                    g.npeer++;
                }
                hbdist::launch_peer_gather(g, s);
                memset(&p.f.io, 0, sizeof(p.f.io));
                lb[jr].cy = lb[jr].sy;  // from here on level jr is complete
                int j0 = p.J - 1;
                if (!(g_force_naive & 8)) {
                    while (j0 > jr && (int64_t)lb[j0].sx.n() * lb[j0].sy.n() <= 40 * 1024) j0--;
                }
                bool fused = false;
                for (int j = jr + 1; j <= j0; j++) launch_down(p, j, s);
                if (j0 < p.J - 1) fused = launch_coarse_fused(p, j0, s);
                if (!fused) {
                    for (int j = j0 + 1; j < p.J; j++) launch_down(p, j, s);
                    for (int j = p.J - 1; j > j0; j--) launch_up(p, j, s);
                }
                for (int j = j0; j >= jr; j--) launch_up(p, j, s);
                ju = jr - 1;
            }
            // the coarse level's halo rows are local when it is the gathered level; the final kernel (j == 0) reads level 1's
            for (int j = ju; j >= 0; j--) up_level(j, j < p.J - 1 && j + 1 < jr);
            memset(&p.f.io, 0, sizeof(p.f.io));
        }
    } else {
    hb::Scratch scratch;
    if ((r = alloc_levels(p, scratch))) return r;
    // input halo: 1 row above, 2 rows below (the 1-3-3-1 taps of level 1), per channel, x relative to in_x0
    const int ht = first ? 0 : 1, hbn = last ? 0 : 2;
    uint16_t *halo_top = nullptr, *halo_bot = nullptr;
    p.f.halo_pitch = p.f.in_w;
    if (ht) halo_top = scratch.get<uint16_t>((size_t)C * ht * p.f.in_w);
    if (hbn) halo_bot = scratch.get<uint16_t>((size_t)C * hbn * p.f.in_w);
    if ((ht && !halo_top) || (hbn && !halo_bot)) {
        return hb::fail(halide_error_code_device_malloc_failed, "local_laplacian_sharded: scratch allocation failed");
    }
    p.f.halo_top = halo_top; p.f.halo_bot = halo_bot;
    p.f.halo_top_rows = ht; p.f.halo_bot_rows = hbn;

    hbdist::Msg msgs[64];
    // queue the halo messages of one row-major f32 array: my first n_up owned rows go up and my last owned row goes
    // down; one row arrives above and n_dn_recv rows arrive below
    int nq = 0;
    auto queue_rows_f32 = [&](float *base, size_t row_elems, Span stored, Span own, int n_up, int n_dn_recv) {
        const size_t rb = row_elems * sizeof(float);
        if (!first) {
            msgs[nq++] = {base + (size_t)(own.lo - stored.lo) * row_elems, (size_t)n_up * rb, up, true};
            msgs[nq++] = {base + (size_t)(own.lo - 1 - stored.lo) * row_elems, rb, up, false};
        }
        if (!last) {
            msgs[nq++] = {base + (size_t)(own.hi - stored.lo) * row_elems, rb, dn, true};
            msgs[nq++] = {base + (size_t)(own.hi + 1 - stored.lo) * row_elems, (size_t)n_dn_recv * rb, dn, false};
        }
    };
    auto flush = [&]() -> int {
        int rr = hbdist::exchange(msgs, nq, s);
        nq = 0;
        return rr;
    };
    {
        hb::CallTimer timer(s);
        // input rows: first two owned rows go up, last owned row goes down (per channel; rows may be strided)
        {
            int n = 0;
            const size_t rb = (size_t)p.f.in_w * sizeof(uint16_t);
            for (int c = 0; c < C; c++) {
                uint16_t *plane = (uint16_t *)din + (int64_t)c * p.f.in_sc;
                if (!first) {
                    msgs[n++] = {plane, rb, up, true};
                    msgs[n++] = {plane + p.f.in_sy, rb, up, true};
                    msgs[n++] = {halo_top + (size_t)c * ht * p.f.in_w, rb, up, false};
                }
                if (!last) {
                    msgs[n++] = {plane + (int64_t)(p.f.in_h - 1) * p.f.in_sy, rb, dn, true};
                    // two receives: the neighbour sends its two (possibly strided) rows as two messages
                    msgs[n++] = {halo_bot + (size_t)c * hbn * p.f.in_w, rb, dn, false};
                    msgs[n++] = {halo_bot + ((size_t)c * hbn + 1) * p.f.in_w, rb, dn, false};
                }
            }
            if ((r = hbdist::exchange(msgs, n, s))) return r;
        }
        launch_lut(p, s);
        LevelBuf *lb = p.ls.lv;
        for (int j = 1; j < p.J; j++) {
            launch_down(p, j, s);
            // gPyramid[j] + inGPyramid[j] halo: 2 rows up, 1 row down (receive 1 above, 2 below)
            queue_rows_f32(lb[j].gp, (size_t)lb[j].gpitch * p.K, lb[j].sy, lb[j].cy, 2, 2);
            queue_rows_f32(lb[j].ing, (size_t)lb[j].gpitch, lb[j].sy, lb[j].cy, 2, 2);
            if ((r = flush())) return r;  // one ncclGroup per level
        }
        for (int j = p.J - 1; j >= 1; j--) {
            launch_up(p, j, s);
            // outGPyramid[j] halo: 1 row each way
            queue_rows_f32(lb[j].outg, (size_t)lb[j].opitch, lb[j].oy, lb[j].coy, 1, 1);
            if ((r = flush())) return r;
        }
        launch_final(p, s);
    }
    }  // NCCL path
    if ((r = hb::check_cuda(cudaGetLastError(), "local_laplacian_sharded launch", halide_error_code_device_run_failed))) return r;
    hb::mark_output_written(output);
    return 0;
}

}  // namespace

extern "C" int local_laplacian(halide_buffer_t *input, int32_t levels, float alpha, float beta, halide_buffer_t *output) {
    return run_local_laplacian(input, levels, alpha, beta, output);
}
extern "C" int local_laplacian_argv(void **args) {
    return run_local_laplacian((halide_buffer_t *)args[0], *(int32_t *)args[1], *(float *)args[2], *(float *)args[3],
                               (halide_buffer_t *)args[4]);
}
extern "C" const halide_filter_metadata_t *local_laplacian_metadata(void) {
    return &kMeta;
}
// The harness's second AOT variant (apps/local_laplacian/process.cpp:44-48): same implementation.
extern "C" int local_laplacian_auto_schedule(halide_buffer_t *input, int32_t levels, float alpha, float beta,
                                             halide_buffer_t *output) {
    return run_local_laplacian(input, levels, alpha, beta, output);
}
extern "C" int local_laplacian_auto_schedule_argv(void **args) {
    return local_laplacian_argv(args);
}
extern "C" const halide_filter_metadata_t *local_laplacian_auto_schedule_metadata(void) {
    return &kMetaAuto;
}

// Test hook: route K == 8 calls through the generic (any `levels`) kernels so both paths stay covered.
extern "C" void halide_b200_ll_force_generic(int enable) {
    g_force_naive = enable;
}

// Row-sharded path: which pyramid level is gathered all-to-all (see run_local_laplacian_sharded).  Collective
// setting: every rank must use the same value.  0 = choose by size (default), -1 = never (exchange halos level by
// level), n >= 2 = level n.
extern "C" void halide_b200_ll_shard_coarse_level(int level) {
    g_shard_coarse_level = level;
}

// Probe for the CPU-side tests (no CUDA calls): the level halide_b200_local_laplacian_sharded would gather for a
// frame of frame_w x frame_h split over nranks ranks, or 8 when halos are exchanged level by level.
extern "C" int halide_b200_ll_shard_plan_level(int32_t frame_w, int32_t frame_h, int32_t nranks) {
    Span fx = {0, frame_w - 1}, fy = {0, frame_h - 1};
    return choose_coarse_level(ll::make_geom(fx, fy, fx, fy, ll::kMaxJ), nranks, ll::kMaxJ, 8);
}

// Row-sharded entry point (B200 extension; see run_local_laplacian_sharded).
extern "C" int halide_b200_local_laplacian_sharded(halide_buffer_t *input, int32_t levels, float alpha, float beta,
                                                   halide_buffer_t *output, int32_t frame_y_min, int32_t frame_y_extent) {
    return run_local_laplacian_sharded(input, levels, alpha, beta, output, frame_y_min, frame_y_extent);
}

// Band geometry probe for the CPU-side tests of the sharding logic (no CUDA calls): fills
// out[j*8 .. j*8+7] = {own.lo, own.hi, stored.lo, stored.hi, own_o.lo, own_o.hi, stored_o.lo, stored_o.hi}.
extern "C" int halide_b200_ll_band_geometry(int32_t frame_w, int32_t frame_h, int32_t band_lo, int32_t band_hi, int32_t first,
                                            int32_t last, int32_t *out) {
    Span fx = {0, frame_w - 1}, fy = {0, frame_h - 1};
    ll::Geom whole = ll::make_geom(fx, fy, fx, fy, ll::kMaxJ);
    ll::BandLevel bl[ll::kMaxJ];
    ll::compute_band_y(whole, Span{band_lo, band_hi}, first != 0, last != 0, bl);
    for (int j = 0; j < ll::kMaxJ; j++) {
        int32_t *o = out + j * 8;
        o[0] = bl[j].own.lo; o[1] = bl[j].own.hi; o[2] = bl[j].stored.lo; o[3] = bl[j].stored.hi;
        o[4] = bl[j].own_o.lo; o[5] = bl[j].own_o.hi; o[6] = bl[j].stored_o.lo; o[7] = bl[j].stored_o.hi;
    }
    return 0;
}

// Device self-test of the arithmetic shortcuts used by the fast kernels (shared-reciprocal division, magic-number
// conversions): returns the number of mismatches against div.rn / cvt over n pseudo-random operand sets, or -1.
extern "C" long long halide_b200_selftest_arith(unsigned long long n, unsigned long long seed) {
    unsigned long long *bad = nullptr, host = 0;
    if (cudaMalloc(&bad, sizeof(*bad)) != cudaSuccess) return -1;
    cudaMemset(bad, 0, sizeof(*bad));
    cudaStream_t s = hb::stream();
    HB_LAUNCH("ll_selftest", ll_selftest_kernel, 148 * 8, 256, 0, s, n, seed, bad);
    if (cudaMemcpyAsync(&host, bad, sizeof(host), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
        cudaFree(bad);
        return -1;
    }
    cudaFree(bad);
    return (long long)host;
}
