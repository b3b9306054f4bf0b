// b2r_static_r2c.cu -- K1 (forward R2C rows) instantiated for the sizes in b2r_static_sizes.h.
#include "b2r_launch.h"
#include "b2r_static_sizes.h"

namespace b2r {
namespace {
template <class P, int PPB> cudaError_t prep(size_t smem, const void*) {
    if (smem <= 48 * 1024) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_r2c_rows<P, float, PPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_r2c_rows<P, __half, PPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <class P, int PPB> cudaError_t run(cudaStream_t s, const R2cArgs& a, int, size_t smem, const void*) {
    const int pairs = 3 * a.dm.h / 2;
    dim3 block(P::kT, PPB), grid((pairs + PPB - 1) / PPB);
    if (a.precision == 2)
        k_r2c_rows<P, __half, PPB><<<grid, block, smem, s>>>((const __half*)a.in, a.spec, a.tw, P{}, a.dm, pairs);
    else
        k_r2c_rows<P, float, PPB><<<grid, block, smem, s>>>((const float*)a.in, a.spec, a.tw, P{}, a.dm, pairs);
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
    o->r2c = &run<P, PPB>;
}
}  // namespace

bool find_static_r2c(int n, RowImpl* out) {
#define X(N, PPB, T, ...) \
    if (n == N) { fill<StaticFft<N, T, __VA_ARGS__>, PPB>(out, "r2c_rows<" #N ">"); return true; }
    B2R_STATIC_R2C_ROWS(X)
#undef X
    return false;
}
}  // namespace b2r
