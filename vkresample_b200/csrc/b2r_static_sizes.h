// b2r_static_sizes.h -- the transform sizes whose schedules are compiled ahead of time.
//
// The reference JIT-compiles one GLSL shader per axis at plan time (VkFFTPlanAxis,
// vkFFT.h:6041-7540), so every size gets constants baked in.  The image has no NVRTC, so the same
// effect is obtained by instantiating the kernel templates here for the sizes of the BASELINE
// configs (and a few common power-of-two sizes); any other 2^a 3^b 5^c 7^d size runs through the
// dynamic kernels (b2r_dynamic.cu).
//
// Row list:  X(N, PPB, T, radices...)   N-point complex transform of one row pair, T threads per
//            pair, PPB pairs per CTA.  Used for both K1 (N = W) and K7 (N = upW).
// Column list: X(H, UPH, CC, FWD, INV)  fused column kernel, CC columns per CTA, FWD / INV are
//            StaticFft aliases with the same thread count.
#pragma once

#include "b2r_fft.cuh"

#define B2R_STATIC_ROWS(X)              \
    X(256, 8, 16, 16, 16)               \
    X(512, 4, 32, 16, 8, 4)             \
    X(1024, 4, 64, 16, 16, 4)           \
    X(2048, 2, 128, 16, 16, 8)          \
    X(4096, 1, 256, 16, 16, 16)         \
    X(1920, 2, 128, 16, 15, 8)          \
    X(3840, 1, 256, 16, 16, 15)         \
    X(7680, 1, 512, 16, 16, 15, 2)

namespace b2r {
using ColF128 = StaticFft<128, 16, 16, 8>;
using ColI256 = StaticFft<256, 16, 16, 16>;
using ColF1024 = StaticFft<1024, 128, 16, 16, 4>;
using ColI2048 = StaticFft<2048, 128, 16, 16, 8>;
using ColF1080 = StaticFft<1080, 180, 15, 12, 6>;
using ColI2160 = StaticFft<2160, 180, 15, 12, 12>;
using ColF2160 = StaticFft<2160, 360, 15, 12, 12>;
using ColI4320 = StaticFft<4320, 360, 16, 15, 6, 3>;
}  // namespace b2r

#define B2R_STATIC_COLS(X)                     \
    X(128, 256, 8, ColF128, ColI256)           \
    X(1024, 2048, 4, ColF1024, ColI2048)       \
    X(1080, 2160, 4, ColF1080, ColI2160)       \
    X(2160, 4320, 2, ColF2160, ColI4320)

// extra tile widths of the c2 column kernel, selectable with B2R_COLS_CC for tuning runs
#define B2R_STATIC_COLS_TUNING(X)              \
    X(1024, 2048, 2, ColF1024, ColI2048)       \
    X(1024, 2048, 8, ColF1024, ColI2048)
