// b2r_jit.h -- plan-time JIT (NVRTC) of statically scheduled kernels; see b2r_jit.cpp.
#pragma once

#include <string>

#include "b2r_launch.h"

namespace b2r {

struct JitModule;   // one loaded module per plan (kernels are per CUDA context)

struct JitRequest {
    Schedule w, h, uh, uw;     // W (R2C rows), H / upH (columns, same thread count), upW (C2R rows)
    bool want_r2c = false, want_cols = false, want_c2r = false;   // which kernels lack an ahead-of-time build
    int precision = 0;         // 0 fp32, 1 fp64 (whole pipeline compiled with -DB2R_REAL_IS_DOUBLE), 2 fp16 storage
    bool want_pixels = false;  // also build the sharpen and u8 pixel kernels (fp64: they have no ahead-of-time build)
    int up_w = 0;              // picks the sharpen variant
    bool up2 = false;          // upW == 2*W: static first-stage operand pattern in the C2R kernel
    bool c2c = false;          // also build k_c2c_rows (B2R_FLAG_C2C_PARITY)
    int cc = 4;                // column tile width
    int ppb_w = 0;             // row pairs per CTA of the R2C kernel (0: 256 threads' worth)
    int nx = 0;                // W/2 + 1
    bool cache_only = false;   // only use a cubin already in the disk cache, never compile
};

// true when libcuda + libnvrtc could be loaded and B2R_JIT != 0
bool jit_available(std::string* why);
// Compiles (or fetches from the disk cache) and loads the requested kernels on the current device and
// fills the launcher structs.  On failure returns false with a message; the caller falls back to the
// dynamic kernels.
bool jit_build(const JitRequest& rq, JitModule** out_mod, RowImpl* r2c, ColImpl* cols, RowImpl* c2r, std::string* err);
void jit_destroy(JitModule* m);
// launchers of the JIT-built sharpen / pixel kernels (fp64 plans)
cudaError_t jit_launch_sharpen(const JitModule* m, cudaStream_t s, const SharpenArgs& a);
cudaError_t jit_launch_u8_to_planar(const JitModule* m, cudaStream_t s, const unsigned char* src, void* dst, const FrameDims& dm);
cudaError_t jit_launch_planar_to_u8(const JitModule* m, cudaStream_t s, const void* src, unsigned char* dst, const FrameDims& dm);

}  // namespace b2r
