// b2r_static_cols.cu -- fused column kernel (K2..K6) instantiated for the (H, upH) pairs in
// b2r_static_sizes.h.
#include "b2r_launch.h"
#include "b2r_static_sizes.h"

#include <cstdlib>

namespace b2r {
namespace {
template <class PF, class PI, int CC> cudaError_t prep(size_t smem, const void*) {
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(k_cols<PF, PI, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <class PF, class PI, int CC> cudaError_t run(cudaStream_t s, const ColsArgs& a, int, size_t smem, const void*) {
    dim3 block(PI::kT * CC), grid((a.dm.nx + CC - 1) / CC, 3);
    k_cols<PF, PI, CC><<<grid, block, smem, s>>>(a.in, a.out, a.tw_f, a.tw_i, PF{}, PI{}, a.dm, a.scale, a.nyq);
    return cudaGetLastError();
}
// ---- staged persistent variant (k_cols_staged): opt-in with B2R_COLS_STAGED=1 (measured: profiles/README.md)
template <class PF, class PI, int CC> cudaError_t prep_staged(size_t, const void*) {
    const size_t smem = cols_staged_smem_bytes(PF::kN, PI::kN, CC, sizeof(float2));
    return cudaFuncSetAttribute(k_cols_staged<PF, PI, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <class PF, class PI, int CC> cudaError_t run_staged(cudaStream_t s, const ColsArgs& a, int, size_t, const void*) {
    const size_t smem = cols_staged_smem_bytes(PF::kN, PI::kN, CC, sizeof(float2));
    static thread_local int dev_cached = -1, sms = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != dev_cached) { cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); dev_cached = dev; }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cols_staged<PF, PI, CC>, PI::kT * CC, smem);
    if (e != cudaSuccess) return e;
    const int tiles_per_ch = (a.dm.nx + CC - 1) / CC, tiles = 3 * tiles_per_ch;
    const int slots = sms * (per_sm > 0 ? per_sm : 1);
    const int trips = (tiles + slots - 1) / slots;
    const int grid = (tiles + trips - 1) / trips;             // equal trips for every CTA
    k_cols_staged<PF, PI, CC><<<grid, PI::kT * CC, smem, s>>>(a.in, a.out, a.tw_f, a.tw_i, PF{}, PI{}, a.dm, a.scale, a.nyq, tiles_per_ch);
    return cudaGetLastError();
}

template <class PF, class PI, int CC> constexpr bool grouped_ok() {
    return PI::kT % 32 == 0 && PF::kStages >= 2 && PI::kStages >= 3 && CC <= 15;
}
template <class PF, class PI, int CC> cudaError_t prep_grouped(size_t smem, const void*) {
    if constexpr (grouped_ok<PF, PI, CC>()) {
        if (smem <= 48 * 1024) return cudaSuccess;
        return cudaFuncSetAttribute(k_cols_grouped<PF, PI, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    } else {
        return cudaErrorInvalidValue;
    }
}
template <class PF, class PI, int CC> cudaError_t run_grouped(cudaStream_t s, const ColsArgs& a, int, size_t smem, const void*) {
    if constexpr (grouped_ok<PF, PI, CC>()) {
        dim3 block(PI::kT * CC), grid((a.dm.nx + CC - 1) / CC, 3);
        k_cols_grouped<PF, PI, CC><<<grid, block, smem, s>>>(a.in, a.out, a.tw_f, a.tw_i, PF{}, PI{}, a.dm, a.scale, a.nyq);
        return cudaGetLastError();
    } else {
        return cudaErrorInvalidValue;
    }
}
template <class P> void sched_of(Schedule* sc) {
    sc->n = P::kN; sc->nst = P::kStages; sc->threads = P::kT;
    for (int s = 0; s < P::kStages; ++s) sc->radices[s] = P::radix(s);
}
template <class PF, class PI, int CC> void fill(ColImpl* o, const char* name) {
    static_assert(PF::kT == PI::kT, "forward and inverse column schedules must share the thread count");
    *o = ColImpl{};
    o->name = name; o->is_static = true; o->cc = CC;
    sched_of<PF>(&o->fwd); sched_of<PI>(&o->inv);
    o->smem = (size_t)smem_padded_len(PI::kN * CC) * sizeof(float2);
    o->prepare = &prep<PF, PI, CC>;
    o->launch = &run<PF, PI, CC>;
    // Named-barrier variant (one thread group per column), opt-in with B2R_COLS_GROUPED=1.  Measured on
    // B200 it is NOT faster than the CTA-barrier kernel (c2: 42.7 vs 41.4 us; c5: 335 vs 321 us): the
    // barrier stalls ncu attributes to k_cols are warps waiting for shared-memory traffic of their
    // peers, which a narrower barrier does not remove.  Kept for that record and for the tests.
    const char* es = getenv("B2R_COLS_STAGED");
    if (es && atoi(es) != 0 && cols_staged_smem_bytes(PF::kN, PI::kN, CC, sizeof(float2)) <= 227 * 1024) {
        o->name = "cols_staged";
        o->smem = cols_staged_smem_bytes(PF::kN, PI::kN, CC, sizeof(float2));
        o->prepare = &prep_staged<PF, PI, CC>;
        o->launch = &run_staged<PF, PI, CC>;
        return;
    }
    const char* e = getenv("B2R_COLS_GROUPED");
    if (grouped_ok<PF, PI, CC>() && e && atoi(e) != 0) {
        o->smem = (size_t)CC * cols_group_stride(PI::kN) * sizeof(float2);
        o->prepare = &prep_grouped<PF, PI, CC>;
        o->launch = &run_grouped<PF, PI, CC>;
    }
}
}  // namespace

bool find_static_cols(int h, int up_h, ColImpl* out) {
    // tuning aid: B2R_COLS_CC=2|4|8 picks another column-tile width where it is instantiated
    const char* e = getenv("B2R_COLS_CC");
    const int want = e ? atoi(e) : 0;
#define X(H, UPH, CC, PF, PI) \
    if (h == H && up_h == UPH && want == CC) { fill<PF, PI, CC>(out, "cols<" #H "->" #UPH "," #CC ">"); return true; }
    B2R_STATIC_COLS_TUNING(X)
#undef X
#define X(H, UPH, CC, PF, PI) \
    if (h == H && up_h == UPH) { fill<PF, PI, CC>(out, "cols<" #H "->" #UPH ">"); return true; }
    B2R_STATIC_COLS(X)
#undef X
    return false;
}
}  // namespace b2r
