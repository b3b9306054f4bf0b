// b2r_static_r2c.cu -- K1 (forward R2C rows) instantiated for the sizes in b2r_static_sizes.h.
#include "b2r_launch.h"
#include "b2r_static_sizes.h"

#include <cstdlib>

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
// ---- bulk-copy (persistent, mbarrier-prefetched) variant: default; B2R_R2C_BULK=0 selects the direct-load kernel
template <class P> cudaError_t prep_bulk(size_t, const void*) {
    cudaError_t e = cudaFuncSetAttribute(k_r2c_rows_bulk<P, float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)r2c_bulk_smem_bytes(P::kN, 4));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_r2c_rows_bulk<P, __half>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)r2c_bulk_smem_bytes(P::kN, 2));
}
template <class P, class TIn> cudaError_t run_bulk_t(cudaStream_t s, const R2cArgs& a, int pairs) {
    const size_t smem = r2c_bulk_smem_bytes(P::kN, sizeof(TIn));
    static thread_local int dev_cached = -1, sms = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != dev_cached) { cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); dev_cached = dev; }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_r2c_rows_bulk<P, TIn>, P::kT, smem);
    if (e != cudaSuccess) return e;
    const int slots = sms * (per_sm > 0 ? per_sm : 1);
    static const int min_trips = [] { const char* e = getenv("B2R_R2C_TRIPS"); return e ? atoi(e) : 0; }();   // tuning aid
    int trips = (pairs + slots - 1) / slots;                 // every CTA runs `trips` (or trips-1) pairs
    if (min_trips > trips) trips = min_trips;
    const int grid = (pairs + trips - 1) / trips;
    k_r2c_rows_bulk<P, TIn><<<grid, P::kT, smem, s>>>((const TIn*)a.in, a.spec, a.tw, P{}, a.dm, pairs);
    return cudaGetLastError();
}
template <class P, int PPB> cudaError_t run_bulk(cudaStream_t s, const R2cArgs& a, int t, size_t smem, const void* ctx) {
    const int pairs = 3 * a.dm.h / 2;
    const size_t elem = a.precision == 2 ? 2 : 4;
    // the bulk copy needs 16-byte aligned row pairs: plane stride and row-pair size multiples of 16 bytes
    if ((a.dm.in_plane * elem) % 16 != 0 || (2 * (size_t)a.dm.w * elem) % 16 != 0) return run<P, PPB>(s, a, t, smem, ctx);
    return a.precision == 2 ? run_bulk_t<P, __half>(s, a, pairs) : run_bulk_t<P, float>(s, a, pairs);
}
template <class P, int PPB> cudaError_t prep_both(size_t smem, const void* ctx) {
    cudaError_t e = prep<P, PPB>(smem, ctx);
    if (e != cudaSuccess) return e;
    return prep_bulk<P>(smem, ctx);
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
    const char* e = getenv("B2R_R2C_BULK");
    if (PPB == 1 && !(e && atoi(e) == 0) && r2c_bulk_smem_bytes(P::kN, 4) <= 200 * 1024) {   // sizes tuned for one pair per CTA
        o->name = "r2c_rows_bulk";
        o->prepare = &prep_both<P, PPB>;
        o->r2c = &run_bulk<P, PPB>;
    }
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
