// b2r_static_sizes.h -- the transform sizes whose schedules are compiled ahead of time.
//
// The reference JIT-compiles one GLSL shader per axis at plan time (VkFFTPlanAxis,
// vkFFT.h:6041-7540), so every size gets constants baked in.  Here the sizes of the BASELINE configs,
// a few common power-of-two sizes and the usual 16:9 video sizes at 2x (360p/540p/720p/1440p sources)
// are instantiated ahead of time; any other 2^a 3^b 5^c 7^d size gets the same templates compiled at
// plan time (b2r_jit.cpp, NVRTC) or, without NVRTC, runs through the dynamic kernels (b2r_dynamic.cu) at
// about half the speed -- add a line here and rebuild to promote a size.
//
// Row lists: X(N, PPB, T, radices...)   N-point complex transform of one row pair, T threads per
//            pair, PPB pairs per CTA; one list for K1 (N = W), one for K7 (N = upW).
// Column list: X(H, UPH, CC, FWD, INV)  fused column kernel, CC columns per CTA, FWD / INV are
//            StaticFft aliases with the same thread count.
#pragma once

#include "b2r_fft.cuh"

// Thread counts / pairs per CTA come from sweeps on B200 through the plan-time JIT (scripts/jit_sweep.py:
// B2R_FORCE_JIT=1 + B2R_TUNE_*, no rebuild per variant): K1 is fastest with one row pair per CTA and one
// butterfly per thread in the widest stage; K7 (persistent, bulk-copy fed) is fastest with two butterflies
// per thread once N >= 2560 (4096: 39.1 -> 35.6 us, 3840: 42.7 -> 37.0 us, 5120: 92.3 -> 82.3 us).
#define B2R_STATIC_R2C_ROWS(X)          \
    X(256, 8, 16, 16, 16)               \
    X(512, 4, 32, 16, 8, 4)             \
    X(1024, 4, 64, 16, 16, 4)           \
    X(2048, 1, 128, 16, 16, 8)          \
    X(4096, 1, 256, 16, 16, 16)         \
    X(1920, 1, 128, 16, 15, 8)          \
    X(3840, 1, 256, 16, 16, 15)         \
    X(640, 4, 64, 16, 8, 5)             \
    X(960, 1, 64, 16, 15, 4)            \
    X(1280, 1, 96, 16, 16, 5)           \
    X(2560, 1, 160, 16, 16, 10)

// (four parts: b2r_static_c2r.cu is compiled once per part so that the build uses more cores)
#define B2R_STATIC_C2R_ROWS_0(X)        \
    X(256, 8, 16, 16, 16)               \
    X(512, 4, 32, 16, 8, 4)             \
    X(1024, 4, 64, 16, 16, 4)           \
    X(2048, 2, 128, 16, 16, 8)
#define B2R_STATIC_C2R_ROWS_1(X)        \
    X(4096, 1, 128, 16, 16, 16)         \
    X(1920, 2, 128, 16, 15, 8)          \
    X(3840, 1, 128, 16, 16, 15)
#define B2R_STATIC_C2R_ROWS_2(X)        \
    X(7680, 1, 384, 16, 20, 24)         \
    X(640, 4, 64, 16, 8, 5)             \
    X(960, 4, 64, 16, 15, 4)            \
    X(1280, 2, 96, 16, 16, 5)
#define B2R_STATIC_C2R_ROWS_3(X)        \
    X(2560, 1, 128, 16, 16, 10)         \
    X(5120, 1, 160, 20, 16, 16)         \
    X(4320, 1, 288, 18, 16, 15)
#define B2R_STATIC_C2R_ROWS(X)          \
    B2R_STATIC_C2R_ROWS_0(X)            \
    B2R_STATIC_C2R_ROWS_1(X)            \
    B2R_STATIC_C2R_ROWS_2(X)            \
    B2R_STATIC_C2R_ROWS_3(X)

// every row schedule once (tests of the bare transforms)
#define B2R_STATIC_ROWS(X)              \
    B2R_STATIC_R2C_ROWS(X)              \
    X(7680, 1, 384, 16, 20, 24)         \
    X(5120, 1, 160, 20, 16, 16)         \
    X(4320, 1, 288, 18, 16, 15)

namespace b2r {
using ColF128 = StaticFft<128, 16, 16, 8>;
using ColI256 = StaticFft<256, 16, 16, 16>;
using ColF512 = StaticFft<512, 64, 16, 8, 4>;
using ColI1024 = StaticFft<1024, 64, 16, 16, 4>;
using ColF1024 = StaticFft<1024, 128, 16, 4, 16>;   // radix order from the sweep: 42.7 -> 41.7 us
using ColI2048 = StaticFft<2048, 128, 16, 16, 8>;
// 8-column tile (64-byte rows), two butterflies per thread: 47.1 us against 54.1 us for <.., 180, ..> x 4 columns
using ColF1080 = StaticFft<1080, 90, 15, 12, 6>;
using ColI2160 = StaticFft<2160, 90, 12, 12, 15>;
using ColF1080n = StaticFft<1080, 180, 15, 12, 6>;
using ColI2160n = StaticFft<2160, 180, 15, 12, 12>;
using ColF360 = StaticFft<360, 48, 15, 8, 3>;
using ColI720 = StaticFft<720, 48, 16, 15, 3>;
using ColF540 = StaticFft<540, 90, 15, 12, 3>;
using ColI1080 = StaticFft<1080, 90, 15, 12, 6>;
using ColF720 = StaticFft<720, 120, 16, 15, 3>;
using ColI1440 = StaticFft<1440, 120, 16, 15, 6>;
using ColF1440 = StaticFft<1440, 240, 16, 15, 6>;
using ColI2880 = StaticFft<2880, 240, 16, 15, 12>;
// 4-column tile (32 B = one DRAM sector per spectrum row), two butterflies per thread: 193 us against
// 312 us for <4320, 288, ...> with 2 columns per CTA (one CTA per SM either way)
using ColF2160 = StaticFft<2160, 144, 16, 9, 15>;    // radix-16 stage first: 199 -> 178 us (jit_sweep)
using ColI4320 = StaticFft<4320, 144, 16, 15, 18>;
using ColF1024h = StaticFft<1024, 64, 16, 16, 4>;
using ColI2048h = StaticFft<2048, 64, 16, 16, 8>;
using ColF1440h = StaticFft<1440, 96, 16, 15, 6>;
using ColI2880h = StaticFft<2880, 96, 16, 15, 12>;
using ColF2160n = StaticFft<2160, 288, 15, 12, 12>;
using ColI4320n = StaticFft<4320, 288, 18, 16, 15>;
}  // namespace b2r

#define B2R_STATIC_COLS(X)                     \
    X(128, 256, 8, ColF128, ColI256)           \
    X(512, 1024, 4, ColF512, ColI1024)         \
    X(1024, 2048, 4, ColF1024, ColI2048)       \
    X(1080, 2160, 8, ColF1080, ColI2160)       \
    X(2160, 4320, 4, ColF2160, ColI4320)       \
    X(360, 720, 8, ColF360, ColI720)           \
    X(540, 1080, 4, ColF540, ColI1080)         \
    X(720, 1440, 4, ColF720, ColI1440)         \
    X(1440, 2880, 4, ColF1440, ColI2880)

// extra tile widths of the c2 column kernel, selectable with B2R_COLS_CC for tuning runs
#define B2R_STATIC_COLS_TUNING(X)              \
    X(1024, 2048, 2, ColF1024, ColI2048)       \
    X(1024, 2048, 8, ColF1024h, ColI2048h)     \
    X(2160, 4320, 2, ColF2160n, ColI4320n)     \
    X(1080, 2160, 4, ColF1080n, ColI2160n)     \
    X(1440, 2880, 8, ColF1440h, ColI2880h)
