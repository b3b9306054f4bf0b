// b2r_common.cuh -- compiler glue shared by every device header.
//
// Two build modes:
//   * nvcc, -gencode arch=compute_100a,code=sm_100a  : the product (libb2resample.so).
//   * g++ with -DB2R_HOST_EMU                         : tests/emu only.  The same kernel source is
//     run by OS threads (one per CUDA thread of a CTA, std::barrier for __syncthreads) so that the
//     index arithmetic of every kernel can be checked against the oracle in the GPU-less authoring
//     container.  It is test infrastructure; the product never links it and has no CPU fallback.
#pragma once

#if defined(__CUDACC_RTC__)
// NVRTC (plan-time JIT of schedules for sizes without an ahead-of-time instantiation, b2r_jit.cpp):
// no host standard library -- take what the kernels need from libcu++ and the compiler built-ins.
#include <cuda_fp16.h>
#include <cuda/std/type_traits>
#include <cuda/std/cstdint>
namespace std { using ::cuda::std::integral_constant; }
#ifndef INFINITY
#define INFINITY __int_as_float(0x7f800000)
#endif
#ifndef NAN
#define NAN __int_as_float(0x7fffffff)
#endif
#else
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <type_traits>
#endif

#if defined(B2R_HOST_EMU)
#include <cmath>
namespace b2r_emu {
struct Ctx {
    unsigned tid_x, tid_y, bid_x, bid_y, bid_z, bdim_x, bdim_y, gdim_x, gdim_y, gdim_z;
    unsigned char* smem;
    void (*sync)(void*);
    void* sync_arg;
    void (*sync_group)(void*, int id, int count);   // named barrier `id` over `count` threads
};
extern thread_local Ctx g_ctx;
}  // namespace b2r_emu
#define B2R_HD inline
#define B2R_DEV inline
#define B2R_KERNEL inline void
#define B2R_TID_X (b2r_emu::g_ctx.tid_x)
#define B2R_TID_Y (b2r_emu::g_ctx.tid_y)
#define B2R_BID_X (b2r_emu::g_ctx.bid_x)
#define B2R_BID_Y (b2r_emu::g_ctx.bid_y)
#define B2R_BID_Z (b2r_emu::g_ctx.bid_z)
#define B2R_BDIM_X (b2r_emu::g_ctx.bdim_x)
#define B2R_BDIM_Y (b2r_emu::g_ctx.bdim_y)
#define B2R_GDIM_X (b2r_emu::g_ctx.gdim_x)
#define B2R_SYNC() (b2r_emu::g_ctx.sync(b2r_emu::g_ctx.sync_arg))
#define B2R_SYNC_GROUP(id, count) (b2r_emu::g_ctx.sync_group(b2r_emu::g_ctx.sync_arg, (id), (count)))
#define B2R_SMEM(T) (reinterpret_cast<T*>(b2r_emu::g_ctx.smem))
#define B2R_LDG(p) (*(p))
#define B2R_LAUNCH_BOUNDS(t, b)
inline float sinpif(float x) { return (float)std::sin(3.14159265358979323846 * (double)x); }
#else
#define B2R_HD __host__ __device__ __forceinline__
#define B2R_DEV __device__ __forceinline__
#define B2R_KERNEL __global__ void
#define B2R_TID_X (threadIdx.x)
#define B2R_TID_Y (threadIdx.y)
#define B2R_BID_X (blockIdx.x)
#define B2R_BID_Y (blockIdx.y)
#define B2R_BID_Z (blockIdx.z)
#define B2R_BDIM_X (blockDim.x)
#define B2R_BDIM_Y (blockDim.y)
#define B2R_GDIM_X (gridDim.x)
#define B2R_SYNC() __syncthreads()
// named barrier: only the `count` threads (a multiple of 32) that use the same id wait for each other
#define B2R_SYNC_GROUP(id, count) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory")
#define B2R_SMEM(T) (reinterpret_cast<T*>(b2r_dyn_smem))
#if defined(__CUDA_ARCH__)
#define B2R_LDG(p) __ldg(p)
#else
#define B2R_LDG(p) (*(p))
#endif
#define B2R_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
extern __shared__ __align__(16) unsigned char b2r_dyn_smem[];

// ---- mbarrier + bulk async copy (the copy engine behind TMA; SASS: UBLKCP / SYNCS) ----------------
__device__ __forceinline__ unsigned b2r_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void b2r_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b2r_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void b2r_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void b2r_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b2r_smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16; completes on `bar`
__device__ __forceinline__ void b2r_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     b2r_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(b2r_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void b2r_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(b2r_smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
#endif

// ---- arithmetic type of the FFT pipeline ---------------------------------------------------------
// float for -p 0 / -p 2 (half is a storage type only, vkFFT.h:7280-7293); the whole header set is
// compiled a second time with -DB2R_REAL_IS_DOUBLE (plan-time JIT only) for -p 1.
namespace b2r {
#if defined(B2R_REAL_IS_DOUBLE)
typedef double real;
typedef double2 real2;
B2R_HD real2 make_real2(real x, real y) { return make_double2(x, y); }
B2R_HD real rfma(real a, real b, real c) { return fma(a, b, c); }
B2R_HD real real_sqrt(real a) { return sqrt(a); }
#if defined(B2R_HOST_EMU)
B2R_HD real real_sinpi(real a) { return std::sin(3.14159265358979323846 * a); }
#else
B2R_HD real real_sinpi(real a) { return sinpi(a); }
#endif
#else
typedef float real;
typedef float2 real2;
B2R_HD real2 make_real2(real x, real y) { return make_float2(x, y); }
B2R_HD real rfma(real a, real b, real c) { return fmaf(a, b, c); }
B2R_HD real real_sqrt(real a) { return sqrtf(a); }
B2R_HD real real_sinpi(real a) { return sinpif(a); }
#endif
}  // namespace b2r
