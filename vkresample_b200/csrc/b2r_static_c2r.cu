// b2r_static_c2r.cu -- K7 (inverse C2R rows) instantiated for the sizes in b2r_static_sizes.h.
#include "b2r_launch.h"
#include "b2r_static_sizes.h"

namespace b2r {
namespace {
template <class P, int PPB> cudaError_t prep(size_t smem) {
    if (smem <= 48 * 1024) return cudaSuccess;
    cudaError_t e;
    const int n = (int)smem;
    if ((e = cudaFuncSetAttribute(k_c2r_rows<P, float, PPB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
    if ((e = cudaFuncSetAttribute(k_c2r_rows<P, float, PPB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
    if ((e = cudaFuncSetAttribute(k_c2r_rows<P, __half, PPB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, n))) return e;
    return cudaFuncSetAttribute(k_c2r_rows<P, __half, PPB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, n);
}
template <class P, int PPB> cudaError_t run(cudaStream_t s, const C2rArgs& a, int, size_t smem) {
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
template <class P, int PPB> void fill(RowImpl* o, const char* name) {
    *o = RowImpl{};
    o->name = name; o->is_static = true;
    o->sched.n = P::kN; o->sched.nst = P::kStages; o->sched.threads = P::kT;
    for (int s = 0; s < P::kStages; ++s) o->sched.radices[s] = P::radix(s);
    o->ppb = PPB;
    o->smem = (size_t)PPB * smem_padded_len(P::kN) * sizeof(float2);
    o->prepare = &prep<P, PPB>;
    o->c2r = &run<P, PPB>;
}
}  // namespace

bool find_static_c2r(int n, RowImpl* out) {
#define X(N, PPB, T, ...) \
    if (n == N) { fill<StaticFft<N, T, __VA_ARGS__>, PPB>(out, "c2r_rows<" #N ">"); return true; }
    B2R_STATIC_ROWS(X)
#undef X
    return false;
}
}  // namespace b2r
