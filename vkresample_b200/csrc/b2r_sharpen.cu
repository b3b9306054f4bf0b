// b2r_sharpen.cu -- launchers of K8 (CAS-style sharpen) and of the u8 pixel-format kernels (b2r_kernels.cuh).
#include "b2r_launch.h"
#include <type_traits>

#include <cstdlib>

namespace b2r {
namespace {
int env_or(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}
}  // namespace

// Tolerance-bound kernels (b2r_cas.cuh) whenever they apply: 0 <= s <= 0.24 (denominator provably in
// [0.04, 1]), fp32 / fp16, output width a multiple of 4 (fp32) or 8 (fp16); everything else -- and
// B2R_FLAG_EXACT_SHARPEN -- runs the bit-exact kernels below.
bool sharpen_fast_applies(const SharpenArgs& a) {
    if (a.exact || a.precision == 1) return false;
    if (!(a.dm.sharpen >= 0.0f && a.dm.sharpen <= kCasFastMaxSharpen)) return false;
    // 16-byte vector accesses: rows and planes must start on 16-byte boundaries
    if (a.precision == 2) return a.dm.up_w % 8 == 0 && a.dm.pre_plane % 8 == 0 && a.dm.out_plane % 8 == 0;
    return a.dm.up_w % 4 == 0 && a.dm.pre_plane % 4 == 0 && a.dm.out_plane % 4 == 0;
}

static cudaError_t launch_sharpen_fast(cudaStream_t s, const SharpenArgs& a) {
    static const int env_ry = env_or("B2R_SHARPEN_RY", 0), env_rev = env_or("B2R_SHARPEN_REVERSE", -1);
    int ry = a.ry > 0 ? a.ry : (env_ry > 0 ? env_ry : kCasFastRows);
    ry = (ry + 5) / 6 * 6;
    const int reverse = a.reverse >= 0 ? a.reverse : (env_rev >= 0 ? env_rev : 1);
    const int np = (a.precision == 2 || a.dm.up_w % 8 == 0) ? 8 : 4;
    static const int env_bx = env_or("B2R_SHARPEN_BX", 0);   // tuning aid (a multiple of 32)
    const int vecs = a.dm.up_w / np, bx = env_bx > 0 ? env_bx : cas_fast_block(vecs);
    dim3 block(bx), grid((vecs + bx - 1) / bx, (a.dm.up_h + ry - 1) / ry, 3);
    if (a.precision == 2) {
        static const int minb = env_or("B2R_SHARPEN_MINB", 3);   // 4: 64-register build, four resident CTAs (tuning aid)
        if (minb == 4) k_sharpen_fast_f16<8, 4><<<grid, block, 0, s>>>((const __half*)a.pre, (__half*)a.out, a.dm, ry, reverse);
        else k_sharpen_fast_f16<8><<<grid, block, 0, s>>>((const __half*)a.pre, (__half*)a.out, a.dm, ry, reverse);
    } else if (np == 8)
        k_sharpen_fast_f32<2><<<grid, block, 0, s>>>((const float*)a.pre, (float*)a.out, a.dm, ry, reverse);
    else
        k_sharpen_fast_f32<1><<<grid, block, 0, s>>>((const float*)a.pre, (float*)a.out, a.dm, ry, reverse);
    return cudaGetLastError();
}

cudaError_t launch_sharpen_fix(cudaStream_t s, const FusedArgs& a) {
    if (a.n_fix <= 0) return cudaSuccess;
    if (a.precision != 0 && a.precision != 2) return cudaErrorNotSupported;
    const int groups = a.dm.up_w / 8, bx = 128;   // 8 pixels per thread in both precisions
    dim3 block(bx), grid((groups + bx - 1) / bx, a.n_fix, 3);
    if (a.precision == 2) k_sharpen_fix_f16<0><<<grid, block, 0, s>>>((const __half*)a.pre, (__half*)a.out, a.dm, a.fix_list);
    else k_sharpen_fix_f32<0><<<grid, block, 0, s>>>((const float*)a.pre, (float*)a.out, a.dm, a.fix_list);
    return cudaGetLastError();
}

cudaError_t launch_sharpen_kernel(cudaStream_t s, const SharpenArgs& a) {
    if (sharpen_fast_applies(a)) return launch_sharpen_fast(s, a);
    const int bx = sharpen_rows_block(a.dm.up_w);
    if (bx > 0) {   // vectorised rolling-window kernel
        constexpr int RY = kSharpenRowsPerThread;
        dim3 block(bx), grid((a.dm.up_w / 4 + bx - 1) / bx, (a.dm.up_h + RY - 1) / RY, 3);
        const bool ragged = sharpen_rows_ragged(a.dm.up_w, bx);
        auto go = [&](auto tp, auto rg, auto ap) {
            using TP = decltype(tp);
            k_sharpen_rows<TP, RY, decltype(rg)::value, decltype(ap)::value><<<grid, block, 0, s>>>((const TP*)a.pre, (TP*)a.out, a.dm);
        };
        auto pick = [&](auto tp) {
            using T = std::true_type; using F = std::false_type;
            if (ragged) { if (a.approx) go(tp, T{}, T{}); else go(tp, T{}, F{}); }
            else        { if (a.approx) go(tp, F{}, T{}); else go(tp, F{}, F{}); }
        };
        if (a.precision == 2) pick(__half{}); else pick(float{});
        return cudaGetLastError();
    }
    constexpr int PX = 4;   // any-width fallback
    dim3 block(256), grid((a.dm.up_w + PX * 256 - 1) / (PX * 256), a.dm.up_h, 3);
    if (a.precision == 2)
        k_sharpen<__half, PX><<<grid, block, 0, s>>>((const __half*)a.pre, (__half*)a.out, a.dm);
    else
        k_sharpen<float, PX><<<grid, block, 0, s>>>((const float*)a.pre, (float*)a.out, a.dm);
    return cudaGetLastError();
}
cudaError_t launch_u8_to_planar(cudaStream_t s, const unsigned char* src, void* dst, const FrameDims& dm, int precision) {
    const size_t n4 = ((size_t)dm.w * dm.h + 3) / 4;
    dim3 block(256), grid((unsigned)((n4 + 255) / 256));
    if (precision == 2) k_u8_to_planar<__half><<<grid, block, 0, s>>>(src, (__half*)dst, dm);
    else k_u8_to_planar<float><<<grid, block, 0, s>>>(src, (float*)dst, dm);
    return cudaGetLastError();
}
cudaError_t launch_planar_to_u8(cudaStream_t s, const void* src, unsigned char* dst, const FrameDims& dm, int precision) {
    const size_t n4 = ((size_t)dm.up_w * dm.up_h + 3) / 4;
    dim3 block(256), grid((unsigned)((n4 + 255) / 256));
    if (precision == 2) k_planar_to_u8<__half><<<grid, block, 0, s>>>((const __half*)src, dst, dm);
    else k_planar_to_u8<float><<<grid, block, 0, s>>>((const float*)src, dst, dm);
    return cudaGetLastError();
}
}  // namespace b2r
