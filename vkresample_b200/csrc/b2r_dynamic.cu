// b2r_dynamic.cu -- any-size fallback: the same kernel templates driven by a device-resident
// FftDesc (runtime radix dispatch).  Slower to compile and to run than the static schedules; used
// only for sizes that are not listed in b2r_static_sizes.h.
#include "b2r_launch.h"

namespace b2r {
namespace {
constexpr int kDynPPB = 1;

cudaError_t prep_r2c(size_t smem, const void*) {
    if (smem <= 48 * 1024) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_r2c_rows<DynFft, float, kDynPPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_r2c_rows<DynFft, __half, kDynPPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
cudaError_t run_r2c(cudaStream_t s, const R2cArgs& a, int threads, size_t smem, const void*) {
    const int pairs = 3 * a.dm.h / 2;
    dim3 block(threads, kDynPPB), grid(pairs);
    if (a.precision == 2)
        k_r2c_rows<DynFft, __half, kDynPPB><<<grid, block, smem, s>>>((const __half*)a.in, a.spec, a.tw, DynFft{a.dfd}, a.dm, pairs);
    else
        k_r2c_rows<DynFft, float, kDynPPB><<<grid, block, smem, s>>>((const float*)a.in, a.spec, a.tw, DynFft{a.dfd}, a.dm, pairs);
    return cudaGetLastError();
}
cudaError_t prep_c2r(size_t smem, const void*) {
    if (smem <= 48 * 1024) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_c2r_rows<DynFft, float, kDynPPB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_c2r_rows<DynFft, __half, kDynPPB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
cudaError_t run_c2r(cudaStream_t s, const C2rArgs& a, int threads, size_t smem, const void*) {
    const int pairs = 3 * a.dm.up_h / 2;
    dim3 block(threads, kDynPPB), grid(pairs);
    if (a.precision == 2)
        k_c2r_rows<DynFft, __half, kDynPPB, false><<<grid, block, smem, s>>>(a.spec, (__half*)a.pre, a.tw, DynFft{a.dfd}, a.dm, pairs, a.scale);
    else
        k_c2r_rows<DynFft, float, kDynPPB, false><<<grid, block, smem, s>>>(a.spec, (float*)a.pre, a.tw, DynFft{a.dfd}, a.dm, pairs, a.scale);
    return cudaGetLastError();
}
cudaError_t prep_c2c(size_t smem, const void*) {
    if (smem <= 48 * 1024) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_c2c_rows<DynFft, float, kDynPPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_c2c_rows<DynFft, __half, kDynPPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
cudaError_t run_c2c(cudaStream_t s, const C2rArgs& a, int threads, size_t smem, const void*) {
    const int rows = 3 * a.dm.up_h;
    dim3 block(threads, kDynPPB), grid(rows);
    if (a.precision == 2)
        k_c2c_rows<DynFft, __half, kDynPPB><<<grid, block, smem, s>>>(a.spec, a.nyq, (__half*)a.pre, a.tw, DynFft{a.dfd}, a.dm, rows, a.scale);
    else
        k_c2c_rows<DynFft, float, kDynPPB><<<grid, block, smem, s>>>(a.spec, a.nyq, (float*)a.pre, a.tw, DynFft{a.dfd}, a.dm, rows, a.scale);
    return cudaGetLastError();
}
}  // namespace

void get_dynamic_r2c(RowImpl* o) {
    *o = RowImpl{};
    o->name = "r2c_rows<dynamic>"; o->ppb = kDynPPB; o->prepare = &prep_r2c; o->r2c = &run_r2c;
}
void get_dynamic_c2r(RowImpl* o) {
    *o = RowImpl{};
    o->name = "c2r_rows<dynamic>"; o->ppb = kDynPPB; o->prepare = &prep_c2r; o->c2r = &run_c2r;
    o->c2c = &run_c2c; o->prepare_c2c = &prep_c2c; o->ppb_c2c = kDynPPB;
}
void get_dynamic_cols(int cc, ColImpl* o) {
    if (cc == 8) get_dynamic_cols_cc8(o);
    else if (cc == 4) get_dynamic_cols_cc4(o);
    else if (cc == 1) get_dynamic_cols_cc1(o);
    else get_dynamic_cols_cc2(o);
}
}  // namespace b2r
