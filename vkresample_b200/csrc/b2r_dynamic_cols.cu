// b2r_dynamic_cols.cu -- any-size fused column kernel (runtime radix dispatch), one translation unit per
// column-tile width: compiled with -DB2R_DYN_CC=1|2|4|8 (see the Makefile) so that the build uses more cores.
#include "b2r_launch.h"

#ifndef B2R_DYN_CC
#error "compile with -DB2R_DYN_CC=1|2|4|8"
#endif
#define B2R_CAT2(a, b) a##b
#define B2R_CAT(a, b) B2R_CAT2(a, b)

namespace b2r {
namespace {
constexpr int CC = B2R_DYN_CC;
cudaError_t prep_cols(size_t smem, const void*) {
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(k_cols<DynFft, DynFft, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
cudaError_t run_cols(cudaStream_t s, const ColsArgs& a, int threads, size_t smem, const void*) {
    dim3 block(threads * CC), grid((a.dm.nx + CC - 1) / CC, 3);
    k_cols<DynFft, DynFft, CC><<<grid, block, smem, s>>>(a.in, a.out, a.tw_f, a.tw_i, DynFft{a.dfd_f}, DynFft{a.dfd_i}, a.dm, a.scale, a.nyq);
    return cudaGetLastError();
}
}  // namespace

void B2R_CAT(get_dynamic_cols_cc, B2R_DYN_CC)(ColImpl* o) {
    *o = ColImpl{};
    o->name = "cols<dynamic>"; o->cc = CC; o->prepare = &prep_cols; o->launch = &run_cols;
}
}  // namespace b2r
