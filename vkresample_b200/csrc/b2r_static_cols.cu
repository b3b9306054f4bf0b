// b2r_static_cols.cu -- fused column kernel (K2..K6) instantiated for the (H, upH) pairs in
// b2r_static_sizes.h.
#include "b2r_launch.h"
#include "b2r_static_sizes.h"

#include <cstdlib>

namespace b2r {
namespace {
template <class P> void sched_of_fwd(Schedule* sc);
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

// ---- exact-2x column kernel (k_cols2x): default for the 2x pairs; B2R_COLS_2X=0 keeps k_cols
template <class PF, int CC, int MINB = 0> cudaError_t prep2x(size_t smem, const void*) {
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(k_cols2x<PF, CC, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <class PF, int CC, int MINB = 0> cudaError_t run2x(cudaStream_t s, const ColsArgs& a, int, size_t smem, const void*) {
    dim3 block(PF::kT * CC), grid((a.dm.nx + CC - 1) / CC, 3);
    k_cols2x<PF, CC, MINB><<<grid, block, smem, s>>>(a.in, a.out, a.tw_f, a.ramp, PF{}, a.dm, a.scale, a.nyq);
    return cudaGetLastError();
}
// tuning variants of the exact-2x kernel (B2R_COLS_2X_VARIANT=1..): other thread counts / tile widths / register targets
template <class PF, int CC, int MINB> void use2x(ColImpl* o) {
    o->cc = CC;
    sched_of_fwd<PF>(&o->fwd);
    o->smem2x = (size_t)smem_padded_len(PF::kN * CC) * sizeof(float2);
    o->prepare2x = &prep2x<PF, CC, MINB>;
    o->launch2x = &run2x<PF, CC, MINB>;
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
template <class P> void sched_of_fwd(Schedule* sc) {
    sc->n = P::kN; sc->nst = P::kStages; sc->threads = P::kT;
    for (int s = 0; s < P::kStages; ++s) sc->radices[s] = P::radix(s);
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
    if constexpr (PI::kN == 2 * PF::kN) {
        // exact-2x pairs run k_cols2x unless B2R_COLS_2X=0.  Decided HERE because the tuned 2x variants may use another
        // H-point schedule than the generic kernel's forward transform: o->fwd (from which the plan builds the
        // twiddle table) and o->cc must describe the kernel that actually runs.  Tuned on B200 (profiles/README.md):
        // 1024: 64 threads per column x 8 columns (28.7 us against 35.0 for 128 x 4); 2160: 144 x 4 with a
        // two-CTA register target (116 against 133 us); 1080: 90 x 8.
        const char* e2 = getenv("B2R_COLS_2X");
        if (!(e2 && atoi(e2) == 0)) {
            const char* ev = getenv("B2R_COLS_2X_VARIANT");
            const int var = ev ? atoi(ev) : 0;
            use2x<PF, CC, 0>(o);
            if constexpr (PF::kN == 1024) {
                use2x<ColF1024h, 8, 0>(o);
                if (var == 1) use2x<ColF1024h, 4, 0>(o);        // 64 threads per column, 256-thread CTAs
                if (var == 2) use2x<ColF1024, 4, 0>(o);         // the generic kernel's shape
                if (var == 3) use2x<ColF1024, 8, 0>(o);         // 8-column tiles, 1024 threads
            }
            if constexpr (PF::kN == 2160) {
                use2x<ColF2160, 4, 2>(o);
                if (var == 1) use2x<ColF2160, 4, 0>(o);         // one resident CTA (86 registers)
            }
            if constexpr (PF::kN == 1080) {
                if (var == 1) use2x<ColF1080, 4, 0>(o);         // 360-thread CTAs
                if (var == 2) use2x<ColF1080n, 4, 0>(o);
            }
            if constexpr (PF::kN == 512) {
                if (var == 1) use2x<StaticFft<512, 32, 16, 8, 4>, 8, 0>(o);
            }
            return;
        }
    }
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
