// b2r_static_c2r.cu -- K7 (inverse C2R rows) instantiated for the sizes in b2r_static_sizes.h.
#include "b2r_launch.h"
#include "b2r_static_sizes.h"

#include <cstdlib>

namespace b2r {
namespace {
template <class P, int PPB> cudaError_t prep(size_t smem, const void*) {
    if (smem <= 48 * 1024) return cudaSuccess;
    cudaError_t e;
    const int n = (int)smem;
    if ((e = cudaFuncSetAttribute(k_c2r_rows<P, float, PPB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
    if ((e = cudaFuncSetAttribute(k_c2r_rows<P, float, PPB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
    if ((e = cudaFuncSetAttribute(k_c2r_rows<P, __half, PPB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
    return cudaFuncSetAttribute(k_c2r_rows<P, __half, PPB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n);
}
template <class P, int PPB> cudaError_t run(cudaStream_t s, const C2rArgs& a, int, size_t smem, const void*) {
    const int pairs = 3 * a.dm.up_h / 2;
    dim3 block(P::kT, PPB), grid((pairs + PPB - 1) / PPB);
    const bool up2 = (a.dm.up_w == 2 * a.dm.w);   // exact 2x: first-stage operand pattern is static
    if (a.precision == 2) {
        if (up2) k_c2r_rows<P, __half, PPB, true><<<grid, block, smem, s>>>(a.spec, (__half*)a.pre, a.tw, P{}, a.dm, pairs, a.scale);
        else k_c2r_rows<P, __half, PPB, false><<<grid, block, smem, s>>>(a.spec, (__half*)a.pre, a.tw, P{}, a.dm, pairs, a.scale);
    } else {
        if (up2) k_c2r_rows<P, float, PPB, true><<<grid, block, smem, s>>>(a.spec, (float*)a.pre, a.tw, P{}, a.dm, pairs, a.scale);
        else k_c2r_rows<P, float, PPB, false><<<grid, block, smem, s>>>(a.spec, (float*)a.pre, a.tw, P{}, a.dm, pairs, a.scale);
    }
    return cudaGetLastError();
}
// ---- bulk-copy (persistent, mbarrier-prefetched) variant ------------------------------------------
// grid = resident CTAs of the device (see DESIGN.md section 4 for the measured comparison with the
// direct-load kernel).
constexpr size_t kBulkSmemMax = 16 + 4 * 4100 * sizeof(float2);   // staging for nx <= 4100 (W <= 8198)
template <class P> cudaError_t prep_bulk(size_t, const void*) {
    const int n = (int)(kBulkSmemMax + smem_padded_len(P::kN) * sizeof(float2));
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_c2r_rows_bulk<P, float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
    if ((e = cudaFuncSetAttribute(k_c2r_rows_bulk<P, float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
    if ((e = cudaFuncSetAttribute(k_c2r_rows_bulk<P, __half, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
    return cudaFuncSetAttribute(k_c2r_rows_bulk<P, __half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n);
}
// single staging buffer (refilled after the first stage): B2R_C2R_BULK=2; measured in profiles/README.md
template <class P, class TOut, bool UP2> cudaError_t run_bulk1_t(cudaStream_t s, const C2rArgs& a) {
    const int pairs = 3 * a.dm.up_h / 2;
    const size_t smem = c2r_bulk1_smem_bytes(P::kN, a.dm.nx);
    static thread_local int dev_cached = -1, sms = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != dev_cached) { cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); dev_cached = dev; }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_c2r_rows_bulk<P, TOut, UP2, true>, P::kT, smem);
    if (e != cudaSuccess) return e;
    int grid = sms * (per_sm > 0 ? per_sm : 1);
    if (grid > pairs) grid = pairs;
    k_c2r_rows_bulk<P, TOut, UP2, true><<<grid, P::kT, smem, s>>>(a.spec, (TOut*)a.pre, a.tw, P{}, a.dm, pairs, a.scale);
    return cudaGetLastError();
}
template <class P, int PPB> cudaError_t run_bulk1(cudaStream_t s, const C2rArgs& a, int, size_t, const void*) {
    const bool up2 = (a.dm.up_w == 2 * a.dm.w);
    if (a.precision == 2) return up2 ? run_bulk1_t<P, __half, true>(s, a) : run_bulk1_t<P, __half, false>(s, a);
    return up2 ? run_bulk1_t<P, float, true>(s, a) : run_bulk1_t<P, float, false>(s, a);
}
template <class P> cudaError_t prep_bulk1(size_t, const void*) {
    const int n = (int)(16 + 2 * 4100 * sizeof(float2) + smem_padded_len(P::kN) * sizeof(float2));
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_c2r_rows_bulk<P, float, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
    if ((e = cudaFuncSetAttribute(k_c2r_rows_bulk<P, float, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
    if ((e = cudaFuncSetAttribute(k_c2r_rows_bulk<P, __half, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
    return cudaFuncSetAttribute(k_c2r_rows_bulk<P, __half, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n);
}

template <class P, class TOut, bool UP2> cudaError_t run_bulk_t(cudaStream_t s, const C2rArgs& a) {
    const int pairs = 3 * a.dm.up_h / 2;
    const size_t smem = c2r_bulk_smem_bytes(P::kN, a.dm.nx);
    static thread_local int dev_cached = -1, sms = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != dev_cached) { cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); dev_cached = dev; }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_c2r_rows_bulk<P, TOut, UP2>, P::kT, smem);
    if (e != cudaSuccess) return e;
    int grid = sms * (per_sm > 0 ? per_sm : 1);
    if (grid > pairs) grid = pairs;
    k_c2r_rows_bulk<P, TOut, UP2><<<grid, P::kT, smem, s>>>(a.spec, (TOut*)a.pre, a.tw, P{}, a.dm, pairs, a.scale);
    return cudaGetLastError();
}
template <class P, int PPB> cudaError_t run_bulk(cudaStream_t s, const C2rArgs& a, int, size_t, const void*) {
    const bool up2 = (a.dm.up_w == 2 * a.dm.w);
    if (a.precision == 2) return up2 ? run_bulk_t<P, __half, true>(s, a) : run_bulk_t<P, __half, false>(s, a);
    return up2 ? run_bulk_t<P, float, true>(s, a) : run_bulk_t<P, float, false>(s, a);
}

// ---- C2C parity mode rows (k_c2c_rows): same schedule, one complex transform per output row
template <class P, int PPB> cudaError_t prep_c2c(size_t smem, const void*) {
    if (smem <= 48 * 1024) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_c2c_rows<P, float, PPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_c2c_rows<P, __half, PPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <class P, int PPB> cudaError_t run_c2c(cudaStream_t s, const C2rArgs& a, int, size_t smem, const void*) {
    const int rows = 3 * a.dm.up_h;
    dim3 block(P::kT, PPB), grid((rows + PPB - 1) / PPB);
    if (a.precision == 2)
        k_c2c_rows<P, __half, PPB><<<grid, block, smem, s>>>(a.spec, a.nyq, (__half*)a.pre, a.tw, P{}, a.dm, rows, a.scale);
    else
        k_c2c_rows<P, float, PPB><<<grid, block, smem, s>>>(a.spec, a.nyq, (float*)a.pre, a.tw, P{}, a.dm, rows, a.scale);
    return cudaGetLastError();
}

// ---- fused C2R + sharpen (b2r_fused.cuh): one CTA per strip of row pairs -----------------------------
template <class P> cudaError_t prep_fused(int precision, int nx) {
    if (precision != 0 && precision != 2) return cudaErrorNotSupported;
    const int n = (int)fused_smem_bytes(P::kN, nx);
    cudaError_t e;
    if (precision == 2) {
        if ((e = cudaFuncSetAttribute(k_c2r_sharpen_f16<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
        return cudaFuncSetAttribute(k_c2r_sharpen_f16<P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n);
    }
    if ((e = cudaFuncSetAttribute(k_c2r_sharpen_f32<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
    return cudaFuncSetAttribute(k_c2r_sharpen_f32<P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n);
}
template <class P> int fused_per_sm(int precision, int nx) {
    if (precision != 0 && precision != 2) return 0;
    int per_sm = 0;
    const size_t smem = fused_smem_bytes(P::kN, nx);
    cudaError_t e = precision == 2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_c2r_sharpen_f16<P, true>, P::kT, smem)
                                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_c2r_sharpen_f32<P, true>, P::kT, smem);
    return e == cudaSuccess ? per_sm : 0;
}
template <class P> cudaError_t run_fused(cudaStream_t s, const FusedArgs& a) {
    if (a.precision != 0 && a.precision != 2) return cudaErrorNotSupported;
    const bool up2 = (a.dm.up_w == 2 * a.dm.w);
    const size_t smem = fused_smem_bytes(P::kN, a.dm.nx);
    const int grid = 3 * a.nsp;
    if (a.precision == 2) {
        if (up2) k_c2r_sharpen_f16<P, true><<<grid, P::kT, smem, s>>>(a.spec, (__half*)a.out, (__half*)a.pre, a.tw, P{}, a.dm, a.scale, a.nsp);
        else k_c2r_sharpen_f16<P, false><<<grid, P::kT, smem, s>>>(a.spec, (__half*)a.out, (__half*)a.pre, a.tw, P{}, a.dm, a.scale, a.nsp);
    } else {
        if (up2) k_c2r_sharpen_f32<P, true><<<grid, P::kT, smem, s>>>(a.spec, (float*)a.out, (float*)a.pre, a.tw, P{}, a.dm, a.scale, a.nsp);
        else k_c2r_sharpen_f32<P, false><<<grid, P::kT, smem, s>>>(a.spec, (float*)a.out, (float*)a.pre, a.tw, P{}, a.dm, a.scale, a.nsp);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_sharpen_fix(s, a);
}

template <class P, int PPB> void fill(RowImpl* o, const char* name) {
    *o = RowImpl{};
    o->name = name; o->is_static = true;
    o->sched.n = P::kN; o->sched.nst = P::kStages; o->sched.threads = P::kT;
    for (int s = 0; s < P::kStages; ++s) o->sched.radices[s] = P::radix(s);
    o->ppb = PPB;
    o->smem = (size_t)PPB * smem_padded_len(P::kN) * sizeof(float2);
    o->prepare = &prep<P, PPB>;
    o->c2r = &run<P, PPB>;
    o->c2c = &run_c2c<P, PPB>;
    o->prepare_c2c = &prep_c2c<P, PPB>;
    o->ppb_c2c = PPB;
    o->smem_c2c = o->smem;
    if constexpr (P::kStages >= 2 && P::kN % 8 == 0) {
        using PF = typename FusedSchedule<P>::type;   // same radix list (same twiddle table), its own thread count
        o->prepare_fused = &prep_fused<PF>;
        o->fused_blocks_per_sm = &fused_per_sm<PF>;
        o->fused = &run_fused<PF>;
    }
    // default: the bulk-copy (mbarrier-prefetched, persistent) kernel; B2R_C2R_BULK=0 selects the
    // direct-load kernel.  Measured on B200, c2: 44.4 -> 39.1 us stand-alone (profiles/README.md).
    const char* e = getenv("B2R_C2R_BULK");
    if (!(e && atoi(e) == 0)) {
        o->name = "c2r_rows_bulk";
        o->ppb = 1;
        o->prepare = &prep_bulk<P>;
        o->c2r = &run_bulk<P, PPB>;
        if (e && atoi(e) == 2) {
            o->name = "c2r_rows_bulk1";
            o->prepare = &prep_bulk1<P>;
            o->c2r = &run_bulk1<P, PPB>;
        }
    }
}
}  // namespace

// This file is compiled once per part of the size list (-DB2R_C2R_PART=0..3, see the Makefile).
#ifndef B2R_C2R_PART
#error "compile with -DB2R_C2R_PART=0..3"
#endif
#define B2R_CAT2(a, b) a##b
#define B2R_CAT(a, b) B2R_CAT2(a, b)
bool B2R_CAT(find_static_c2r_part, B2R_C2R_PART)(int n, RowImpl* out) {
#define X(N, PPB, T, ...) \
    if (n == N) { fill<StaticFft<N, T, __VA_ARGS__>, PPB>(out, "c2r_rows<" #N ">"); return true; }
    B2R_CAT(B2R_STATIC_C2R_ROWS_, B2R_C2R_PART)(X)
#undef X
    return false;
}
#if B2R_C2R_PART == 0
bool find_static_c2r(int n, RowImpl* out) {
    return find_static_c2r_part0(n, out) || find_static_c2r_part1(n, out) || find_static_c2r_part2(n, out) ||
           find_static_c2r_part3(n, out);
}
#endif
}  // namespace b2r
