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
