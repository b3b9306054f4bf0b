// b2r_launch.h -- host-side launch interface between the plan runtime (b2r_api.cu) and the kernel
// translation units.  Two families implement it:
//   * static  (b2r_static_rows.cu / b2r_static_cols.cu): schedules instantiated at build time for
//     the sizes in b2r_static_sizes.h -- every stride / count is an immediate in the SASS;
//   * dynamic (b2r_dynamic.cu): one kernel per kind that reads its schedule from a device-resident
//     FftDesc, for any 2^a 3^b 5^c 7^d size (what the reference gets from JIT-compiling GLSL,
//     vkFFT.h:4495-4642, we get from a runtime radix dispatch).
#pragma once

#include <cuda_runtime.h>

#include "b2r_fft.cuh"
#include "b2r_kernels.cuh"

namespace b2r {

struct R2cArgs {
    const void* in; float2* spec; const float2* tw; const FftDesc* dfd; FrameDims dm; int precision;
};
struct ColsArgs {
    const float2* in; float2* out; const float2 *tw_f, *tw_i; const FftDesc *dfd_f, *dfd_i; FrameDims dm; float scale;
    float2* nyq = nullptr;   // C2C parity mode: receives F[H/2][x] of the forward column transform
    const float2* ramp = nullptr;   // exact-2x column kernel: half-sample phase ramp, H entries (nullptr: k_cols)
};
struct C2rArgs {
    const float2* spec; void* pre; const float2* tw; const FftDesc* dfd; FrameDims dm; int precision; float scale;
    const float2* nyq = nullptr;   // C2C parity mode (c2c launcher only)
};
struct SharpenArgs {
    const void* pre; void* out; FrameDims dm; int precision;
    bool approx = false;   // exact kernels only: the round-1 approximate-division variant (tuning / comparison)
    bool exact = false;    // B2R_FLAG_EXACT_SHARPEN: bit-exact kernels instead of the tolerance-bound default
    int ry = 0;            // rows per thread of the fast kernels (0: default)
    int reverse = -1;      // fast kernels: walk planes / strips against K7's write order (-1: default)
};

struct FusedArgs {     // K7 + K8 in one kernel (b2r_fused.cuh) followed by the boundary-row fix-up
    const float2* spec; void* out; void* pre; const float2* tw; FrameDims dm; int precision; float scale;
    int nsp;               // strips per colour plane (one CTA per strip)
    const int* fix_list;   // device: rows / last pixels k_sharpen_fix finishes (per plane)
    int n_fix;
};

struct Schedule {      // radix list + cooperating threads of one transform
    int n = 0, nst = 0, threads = 0;
    int radices[kMaxStages] = {};
};

struct RowImpl {       // K1 or K7 resolved for one size
    const char* name = nullptr;
    bool is_static = false;
    Schedule sched;
    int ppb = 1;               // row pairs per CTA
    size_t smem = 0;
    const void* ctx = nullptr;  // launcher-private (JIT: the loaded kernels); passed back to every call below
    bool is_jit = false;
    cudaError_t (*prepare)(size_t smem, const void* ctx) = nullptr;   // per-device function attributes
    cudaError_t (*r2c)(cudaStream_t, const R2cArgs&, int threads, size_t smem, const void* ctx) = nullptr;
    cudaError_t (*c2r)(cudaStream_t, const C2rArgs&, int threads, size_t smem, const void* ctx) = nullptr;
    cudaError_t (*c2c)(cudaStream_t, const C2rArgs&, int threads, size_t smem, const void* ctx) = nullptr;   // k_c2c_rows
    cudaError_t (*prepare_c2c)(size_t smem, const void* ctx) = nullptr;
    int ppb_c2c = 1;
    size_t smem_c2c = 0;
    // fused C2R + sharpen (static schedules, fp32 / fp16): nullptr when not instantiated
    cudaError_t (*prepare_fused)(int precision, int nx) = nullptr;
    int (*fused_blocks_per_sm)(int precision, int nx) = nullptr;   // resident CTAs per SM (occupancy API)
    cudaError_t (*fused)(cudaStream_t, const FusedArgs&) = nullptr;
};

struct ColImpl {       // fused column kernel resolved for one (H, upH) pair
    const char* name = nullptr;
    bool is_static = false;
    Schedule fwd, inv;         // same thread count
    int cc = 4;                // spectrum columns per CTA
    size_t smem = 0;
    const void* ctx = nullptr;
    bool is_jit = false;
    cudaError_t (*prepare)(size_t smem, const void* ctx) = nullptr;
    cudaError_t (*launch)(cudaStream_t, const ColsArgs&, int threads, size_t smem, const void* ctx) = nullptr;
    // exact-2x form (k_cols2x): two H-point transforms, even rows copied; nullptr when not instantiated
    size_t smem2x = 0;
    cudaError_t (*prepare2x)(size_t smem, const void* ctx) = nullptr;
    cudaError_t (*launch2x)(cudaStream_t, const ColsArgs&, int threads, size_t smem, const void* ctx) = nullptr;
};

// static registries: return false when the size was not instantiated at build time
bool find_static_r2c(int n, RowImpl* out);
bool find_static_c2r(int n, RowImpl* out);
bool find_static_c2r_part0(int n, RowImpl* out);   // b2r_static_c2r.cu, one translation unit per part
bool find_static_c2r_part1(int n, RowImpl* out);
bool find_static_c2r_part2(int n, RowImpl* out);
bool find_static_c2r_part3(int n, RowImpl* out);
bool find_static_cols(int h, int up_h, ColImpl* out);
// dynamic fallbacks (always succeed for schedulable sizes); cc in {2,4,8}
void get_dynamic_r2c(RowImpl* out);
void get_dynamic_c2r(RowImpl* out);
void get_dynamic_cols(int cc, ColImpl* out);
void get_dynamic_cols_cc1(ColImpl* out);           // b2r_dynamic_cols.cu, one translation unit per tile width
void get_dynamic_cols_cc2(ColImpl* out);
void get_dynamic_cols_cc4(ColImpl* out);
void get_dynamic_cols_cc8(ColImpl* out);

cudaError_t launch_sharpen_fix(cudaStream_t s, const FusedArgs& a);   // boundary rows of the fused kernel
cudaError_t launch_sharpen_kernel(cudaStream_t s, const SharpenArgs& a);
bool sharpen_fast_applies(const SharpenArgs& a);   // true: the tolerance-bound kernels (b2r_cas.cuh) will run
cudaError_t launch_u8_to_planar(cudaStream_t s, const unsigned char* src, void* dst, const FrameDims& dm, int precision);
cudaError_t launch_planar_to_u8(cudaStream_t s, const void* src, unsigned char* dst, const FrameDims& dm, int precision);

}  // namespace b2r
