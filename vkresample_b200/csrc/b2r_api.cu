// b2r_api.cu -- C-ABI (include/b2resample.h) and host runtime of the frame pipeline:
// device buffers, stream, events, CUDA graph of the per-frame kernel sequence.
//
// Replaces, call for call, the Vulkan side of launchResample (VkResample.cpp:1280-1780):
//   plan create   <- configuration + allocateFFTBuffer x3 + initializeVulkanFFT x2 + createShiftApp
//                    + createSharpenApp                       (:1409-1617)
//   upload        <- transferDataFromCPU                      (:1688, def :385-429)
//   execute       <- performVulkanUpscale                     (:1692, def :1249-1279)
//   download      <- transferDataToCPU                        (:1697-1700, def :430-473)
//   destroy       <- deleteVulkanFFT / deleteShiftApp / vkDestroy*   (:1762-1778)
// The reference records `numIter` copies of its 23 dispatches into one command buffer and times
// submit -> fence; here one frame is a CUDA graph of 4 kernels, replayed numIter times between two
// events on the plan's stream.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/b2resample.h"
#include "b2r_jit.h"
#include "b2r_launch.h"
#include "b2r_plan.h"

using namespace b2r;

namespace {
thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(B2R_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Entry points run on the plan's device and leave the caller's current device as they found it.
struct DeviceGuard {
    int prev = -1, dev_ = -1;
    bool good = true;
    explicit DeviceGuard(int dev) : dev_(dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        // always: cudaSetDevice is also what creates the primary context and makes it current on this thread,
        // which the driver-API calls of the plan-time JIT (cuModuleLoadData, ...) rely on
        good = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0 && prev != dev_) cudaSetDevice(prev); }
    bool ok() const { return good; }
};
#define ON_PLAN_DEVICE(p)                                                                           \
    DeviceGuard b2r_guard_((p)->device);                                                            \
    if (!b2r_guard_.ok()) return fail(B2R_ERR_CUDA, "cudaSetDevice(%d) failed", (p)->device)

float literal_f(float v) {  // value of the "%f" text the reference pastes into its GLSL (VkResample.cpp:893-920)
    char t[64];
    snprintf(t, sizeof t, "%f", (double)v);
    return strtof(t, nullptr);
}
double literal_d(float v) {  // the same text read as a double literal (-p 1 shader)
    char t[64];
    snprintf(t, sizeof t, "%f", (double)v);
    return strtod(t, nullptr);
}
int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}
}  // namespace

// One set of per-frame working buffers + a stream.  Lane 0 is the plan's primary set; extra lanes
// (b2r_plan_set_lanes) let independent frames of a stream overlap on the GPU and let the PCIe copies
// of one frame overlap the kernels of another -- what running the reference with -numthreads N does
// with N private VkFFT applications on one device (VkResample.cpp:1959-1969).
struct Lane {
    cudaStream_t stream = nullptr;
    void* d_in = nullptr;
    void* d_pre = nullptr;
    void* d_out = nullptr;
    float2* d_spec1 = nullptr;
    float2* d_spec2 = nullptr;
    cudaEvent_t done = nullptr;
    float2* d_nyq = nullptr;           // C2C parity mode: y-Nyquist row of the forward column transform
    unsigned char* u8_in = nullptr;    // interleaved u8 staging, allocated on first use of the u8 API
    unsigned char* u8_out = nullptr;
};

constexpr int kTicketDepth = 4;        // completion events kept per lane (b2r_wait_ticket)

struct b2r_plan {
    int device = 0;
    uint32_t flags = 0;
    Geometry g;
    HostFft fw, fh, fuh, fuw;  // W (R2C rows), H (fwd cols), upH (inv cols), upW (C2R rows)
    RowImpl k_r2c, k_c2r;      // resolved kernels (static schedule or dynamic fallback)
    ColImpl k_cols;
    FrameDims dm{};
    void* d_in = nullptr;
    void* d_pre = nullptr;
    void* d_out = nullptr;
    float2* d_spec1 = nullptr;
    float2* d_spec2 = nullptr;
    float2* d_tw = nullptr;
    float2* d_nyq = nullptr;   // lane 0, C2C parity mode only
    bool c2c = false;
    FftDesc* d_fd = nullptr;   // 4 descriptors for the dynamic kernels: W, H, upH, upW
    const float2 *tw_w = nullptr, *tw_h = nullptr, *tw_uh = nullptr, *tw_uw = nullptr;
    size_t device_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    uint64_t launches = 0;
    int kernels_per_frame = 4;
    unsigned char* u8_in0 = nullptr;   // lane 0's u8 staging
    unsigned char* u8_out0 = nullptr;
    float2* d_ramp = nullptr;  // exact-2x column kernel: half-sample phase ramp (H entries); nullptr = k_cols
    // fused C2R + sharpen (b2r_fused.cuh): strips per plane, fix-up list (device), 0 = separate kernels
    int fused_nsp = 0;
    int* d_fix = nullptr;
    int n_fix = 0;
    JitModule* jit = nullptr;   // kernels compiled at plan time for sizes without an ahead-of-time build
    std::string jit_note;       // why JIT was not used (if it was not)
    std::vector<Lane> extra;   // lanes 1..n-1
    uint32_t next_lane = 0;
    // completion tickets (b2r_wait_ticket): per lane a ring of kTicketDepth events; tick_log remembers which
    // event the last kTicketLog tickets were stamped with.  An event that was re-recorded since belongs to a
    // LATER frame of the same in-order stream, so waiting on it still implies the older frame has finished.
    std::vector<std::vector<cudaEvent_t>> tick_ev;
    std::vector<uint32_t> tick_pos;
    struct TickEntry { uint64_t ticket = 0; cudaEvent_t ev = nullptr; };
    TickEntry tick_log[64];
    std::mutex tick_mu;        // b2r_wait_ticket may run on another thread than the enqueuing one
    uint64_t tickets = 0;
    Lane lane(uint32_t i) const {
        if (i == 0) { Lane l; l.stream = stream; l.d_in = d_in; l.d_pre = d_pre; l.d_out = d_out; l.d_spec1 = d_spec1; l.d_spec2 = d_spec2; l.done = ev1; l.d_nyq = d_nyq; return l; }
        return extra[i - 1];
    }
    uint32_t num_lanes() const { return 1 + (uint32_t)extra.size(); }
};

namespace {

int launch_sharpen(b2r_plan* p, cudaStream_t s, void* d_out = nullptr, const Lane* ln = nullptr) {
    SharpenArgs a{ln ? ln->d_pre : p->d_pre, d_out ? d_out : (ln ? ln->d_out : p->d_out), p->dm, p->g.precision};
    a.exact = (p->flags & B2R_FLAG_EXACT_SHARPEN) != 0;   // default: tolerance-bound kernels (b2r_cas.cuh)
    a.approx = (p->flags & B2R_FLAG_FAST_SHARPEN) != 0;   // exact kernels only: round-1 approximate-division variant
    if (p->g.precision == 1) CU(jit_launch_sharpen(p->jit, s, a));
    else CU(launch_sharpen_kernel(s, a));
    return B2R_SUCCESS;
}

// The per-frame sequence: K1 -> fused columns -> K7 -> K8 (performVulkanUpscale body, :1260-1269)
// ev (optional, 5 events): recorded before K1 and after each kernel for per-kernel timing.
int launch_frame(b2r_plan* p, cudaStream_t s, const void* d_in = nullptr, void* d_out = nullptr,
                 cudaEvent_t* ev = nullptr, const Lane* ln = nullptr) {
    const Geometry& g = p->g;
    float2* spec1 = ln ? ln->d_spec1 : p->d_spec1;
    float2* spec2 = ln ? ln->d_spec2 : p->d_spec2;
    void* pre = ln ? ln->d_pre : p->d_pre;
    float2* nyq = p->c2c ? (ln ? ln->d_nyq : p->d_nyq) : nullptr;
    if (ev) CU(cudaEventRecord(ev[0], s));
    R2cArgs a1{d_in ? d_in : (ln ? ln->d_in : p->d_in), spec1, p->tw_w, p->d_fd + 0, p->dm, g.precision};
    CU(p->k_r2c.r2c(s, a1, p->k_r2c.sched.threads, p->k_r2c.smem, p->k_r2c.ctx));
    if (ev) CU(cudaEventRecord(ev[1], s));
    ColsArgs a2{spec1, spec2, p->tw_h, p->tw_uh, p->d_fd + 1, p->d_fd + 2, p->dm, 1.0f / (float)g.up_h, nyq};
    if (p->d_ramp) {   // upH == 2H: even rows copied, odd rows by an H-point inverse (k_cols2x)
        a2.ramp = p->d_ramp;
        CU(p->k_cols.launch2x(s, a2, p->k_cols.fwd.threads, p->k_cols.smem2x, p->k_cols.ctx));
    } else {
        CU(p->k_cols.launch(s, a2, p->k_cols.inv.threads, p->k_cols.smem, p->k_cols.ctx));
    }
    if (ev) CU(cudaEventRecord(ev[2], s));
    if (p->fused_nsp > 0) {   // K7 + K8 in one kernel; the boundary rows go through the pre-sharpen buffer
        FusedArgs af{spec2, d_out ? d_out : (ln ? ln->d_out : p->d_out), pre, p->tw_uw, p->dm, g.precision,
                     1.0f / (float)g.up_w, p->fused_nsp, p->d_fix, p->n_fix};
        CU(p->k_c2r.fused(s, af));
        if (ev) { CU(cudaEventRecord(ev[3], s)); CU(cudaEventRecord(ev[4], s)); }
        return B2R_SUCCESS;
    }
    C2rArgs a3{spec2, pre, p->tw_uw, p->d_fd + 3, p->dm, g.precision, 1.0f / (float)g.up_w, nyq};
    if (p->c2c) CU(p->k_c2r.c2c(s, a3, p->k_c2r.sched.threads, p->k_c2r.smem_c2c, p->k_c2r.ctx));
    else CU(p->k_c2r.c2r(s, a3, p->k_c2r.sched.threads, p->k_c2r.smem, p->k_c2r.ctx));
    if (ev) CU(cudaEventRecord(ev[3], s));
    int rc = launch_sharpen(p, s, d_out, ln);
    if (rc) return rc;
    if (ev) CU(cudaEventRecord(ev[4], s));
    return B2R_SUCCESS;
}

int run_frame(b2r_plan* p) {
    if (p->graph_exec) {
        CU(cudaGraphLaunch(p->graph_exec, p->stream));
    } else {
        int rc = launch_frame(p, p->stream);
        if (rc) return rc;
    }
    p->launches += p->kernels_per_frame;
    return B2R_SUCCESS;
}

void sched_from(const HostFft& f, Schedule* sc) {
    sc->n = f.desc.n; sc->nst = f.desc.nstages; sc->threads = f.desc.threads;
    for (int s = 0; s < f.desc.nstages; ++s) sc->radices[s] = f.desc.st[s].radix;
}

int build(b2r_plan* p) {
    Geometry& g = p->g;
    std::string err;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, p->device));
    const size_t smem_max = prop.sharedMemPerBlockOptin;
    const bool dbl = g.precision == 1;
    // fp64: nothing is built ahead of time -- the whole kernel set is compiled at plan time; the static
    // table still supplies the schedule for the sizes it lists
    const bool force_dyn = env_int("B2R_FORCE_DYNAMIC", 0) != 0 || dbl;
    // tuning aid: B2R_FORCE_JIT=1 compiles the schedules at plan time even for sizes with an ahead-of-time
    // build, so that B2R_TUNE_* (thread counts, tile widths, radix orders) can be swept without rebuilding
    const bool force_jit = !force_dyn && env_int("B2R_FORCE_JIT", 0) != 0;
    const bool seed = dbl || force_jit;        // take the static table's schedule as the starting point
    const size_t cb = g.cplx_bytes();
    RowImpl st_r2c, st_c2r; ColImpl st_cols;
    const bool has_r2c = find_static_r2c(g.w, &st_r2c), has_c2r = find_static_c2r(g.up_w, &st_c2r);
    const bool has_cols = find_static_cols(g.h, g.up_h, &st_cols);

    // ---- resolve kernels: static schedule when the size was instantiated, else dynamic
    if (!force_dyn && !force_jit && find_static_r2c(g.w, &p->k_r2c)) {
        build_fft(g.w, p->k_r2c.sched.radices, p->k_r2c.sched.nst, p->k_r2c.sched.threads, &p->fw);
    } else {
        if (seed && has_r2c) build_fft(g.w, st_r2c.sched.radices, st_r2c.sched.nst, st_r2c.sched.threads, &p->fw);
        else if (!schedule_fft(g.w, &p->fw, &err)) return fail(B2R_ERR_UNSUPPORTED, "%s", err.c_str());
        get_dynamic_r2c(&p->k_r2c);
        sched_from(p->fw, &p->k_r2c.sched);
        p->k_r2c.smem = (size_t)p->k_r2c.ppb * smem_padded_len(g.w) * sizeof(float2);
    }
    if (!force_dyn && !force_jit && find_static_c2r(g.up_w, &p->k_c2r)) {
        build_fft(g.up_w, p->k_c2r.sched.radices, p->k_c2r.sched.nst, p->k_c2r.sched.threads, &p->fuw);
    } else {
        if (seed && has_c2r) build_fft(g.up_w, st_c2r.sched.radices, st_c2r.sched.nst, st_c2r.sched.threads, &p->fuw);
        else if (!schedule_fft(g.up_w, &p->fuw, &err)) return fail(B2R_ERR_UNSUPPORTED, "%s", err.c_str());
        get_dynamic_c2r(&p->k_c2r);
        sched_from(p->fuw, &p->k_c2r.sched);
        p->k_c2r.smem = (size_t)p->k_c2r.ppb * smem_padded_len(g.up_w) * sizeof(float2);
    }
    if (!force_dyn && !force_jit && find_static_cols(g.h, g.up_h, &p->k_cols)) {
        build_fft(g.h, p->k_cols.fwd.radices, p->k_cols.fwd.nst, p->k_cols.fwd.threads, &p->fh);
        build_fft(g.up_h, p->k_cols.inv.radices, p->k_cols.inv.nst, p->k_cols.inv.threads, &p->fuh);
    } else {
        if (seed && has_cols) {
            build_fft(g.h, st_cols.fwd.radices, st_cols.fwd.nst, st_cols.fwd.threads, &p->fh);
            build_fft(g.up_h, st_cols.inv.radices, st_cols.inv.nst, st_cols.inv.threads, &p->fuh);
        } else {
            if (!schedule_fft(g.h, &p->fh, &err) || !schedule_fft(g.up_h, &p->fuh, &err))
                return fail(B2R_ERR_UNSUPPORTED, "%s", err.c_str());
            const int tc0 = std::max(p->fh.desc.threads, p->fuh.desc.threads);
            schedule_fft(g.h, &p->fh, &err, tc0);
            schedule_fft(g.up_h, &p->fuh, &err, tc0);
        }
        const int tc = p->fuh.desc.threads;
        // column tile: widest of {8,4,2} with <= 512 threads and room for >= 2 CTAs per SM
        int cc = env_int("B2R_COLS_CC", 0);
        if (cc != 1 && cc != 2 && cc != 4 && cc != 8) {
            cc = (2 * tc <= kDynMaxThreads) ? 2 : 1;   // 1: last resort for very long columns
            for (int cand : {8, 4})
                if (cand * tc <= kDynMaxThreads && smem_padded_len(g.up_h * cand) * sizeof(float2) <= 100 * 1024) { cc = cand; break; }
        }
        get_dynamic_cols(cc, &p->k_cols);
        sched_from(p->fh, &p->k_cols.fwd);
        sched_from(p->fuh, &p->k_cols.inv);
        p->k_cols.smem = (size_t)smem_padded_len(g.up_h * p->k_cols.cc) * sizeof(float2);
    }
    // ---- plan-time JIT for whatever has no ahead-of-time schedule (the reference JIT-compiles every plan)
    if ((dbl || (!force_dyn && !(p->flags & B2R_FLAG_NO_JIT))) &&
        (!p->k_r2c.is_static || !p->k_cols.is_static || !p->k_c2r.is_static)) {
        std::string why;
        if (jit_available(&why)) {
            JitRequest rq;
            sched_from(p->fw, &rq.w); sched_from(p->fh, &rq.h); sched_from(p->fuh, &rq.uh); sched_from(p->fuw, &rq.uw);
            rq.want_r2c = !p->k_r2c.is_static; rq.want_cols = !p->k_cols.is_static; rq.want_c2r = !p->k_c2r.is_static;
            rq.precision = g.precision; rq.up2 = (g.up_w == 2 * g.w); rq.c2c = p->c2c; rq.nx = g.nx;
            rq.cache_only = false;
            rq.want_pixels = dbl; rq.up_w = g.up_w;
            const int tc = rq.uh.threads;
            rq.cc = (tc <= 32) ? 8 : ((4 * tc <= 1024 && smem_padded_len(g.up_h * 4) * cb <= (dbl ? 200u : 110u) * 1024) ? 4 : 2);
            if (dbl && has_cols && (size_t)smem_padded_len(g.up_h * st_cols.cc) * cb <= smem_max) rq.cc = st_cols.cc;
            // long columns: a 2-column tile moves 16 B per spectrum row (half a DRAM sector; measured 312 vs
            // 193 us at 2160 -> 4320).  Keep 4 columns per CTA and let every thread run several butterflies.
            if (!dbl && rq.want_cols && rq.cc == 2 && (size_t)smem_padded_len(g.up_h * 4) * cb <= smem_max) {
                int t2 = tc;
                while (4 * t2 > 640) t2 = (((t2 + 1) / 2) + 7) & ~7;
                rq.cc = 4; rq.h.threads = rq.uh.threads = t2;
            }
            // row kernels (sweeps on B200, scripts/jit_sweep.py): K1 with one row pair per CTA from W = 960 up;
            // K7 with two butterflies per thread from upW = 2560 up while a thread holds <= 16 values
            if (!dbl && !force_jit) {
                if (rq.want_r2c && g.w >= 960) rq.ppb_w = 1;
                if (rq.want_c2r && g.up_w >= 2560) {
                    int rmax = 0;
                    for (int i = 0; i < rq.uw.nst; ++i) rmax = std::max(rmax, rq.uw.radices[i]);
                    const int t2 = (((rq.uw.threads + 1) / 2) + 7) & ~7;
                    int rmin = 99;
                    for (int i = 0; i < rq.uw.nst; ++i) rmin = std::min(rmin, rq.uw.radices[i]);
                    if (rmax <= 16 && rmin >= 8 && t2 >= 96) rq.uw.threads = t2;
                }
            }
            if (force_jit) {
                if (has_cols) rq.cc = st_cols.cc;
                auto tune = [&](Schedule& sc, const char* axis) {   // B2R_TUNE_T<axis>=threads  B2R_TUNE_R<axis>=r0,r1,...
                    const std::string a(axis);
                    sc.threads = env_int(("B2R_TUNE_T" + a).c_str(), sc.threads);
                    if (const char* e = getenv(("B2R_TUNE_R" + a).c_str())) {
                        int prod = 1, n = 0, rad[kMaxStages];
                        for (const char* q = e; *q && n < kMaxStages; ) {
                            rad[n] = atoi(q); prod *= rad[n] > 0 ? rad[n] : 0; ++n;
                            while (*q && *q != ',') ++q;
                            if (*q == ',') ++q;
                        }
                        if (prod == sc.n) { sc.nst = n; for (int i = 0; i < n; ++i) sc.radices[i] = rad[i]; }
                    }
                };
                tune(rq.w, "W"); tune(rq.h, "H"); tune(rq.uh, "UH"); tune(rq.uw, "UW");
                rq.uh.threads = rq.h.threads = std::max(rq.h.threads, rq.uh.threads);
                rq.cc = env_int("B2R_TUNE_CC", rq.cc);
                rq.ppb_w = env_int("B2R_TUNE_PPBW", has_r2c ? st_r2c.ppb : 0);
            }
            RowImpl jr, jc; ColImpl jcol;
            if (jit_build(rq, &p->jit, &jr, &jcol, &jc, &why)) {
                if (rq.want_r2c) p->k_r2c = jr;
                if (rq.want_cols) p->k_cols = jcol;
                if (rq.want_c2r) p->k_c2r = jc;
                if (force_jit) {   // twiddle tables follow the (possibly re-ordered) radix lists
                    build_fft(g.w, rq.w.radices, rq.w.nst, rq.w.threads, &p->fw);
                    build_fft(g.h, rq.h.radices, rq.h.nst, rq.h.threads, &p->fh);
                    build_fft(g.up_h, rq.uh.radices, rq.uh.nst, rq.uh.threads, &p->fuh);
                    build_fft(g.up_w, rq.uw.radices, rq.uw.nst, rq.uw.threads, &p->fuw);
                }
            } else {
                p->jit_note = why;
            }
        } else {
            p->jit_note = why;
        }
        // B2R_FORCE_JIT seeds the descriptors with the ahead-of-time schedules (two butterflies per thread, ...),
        // which the any-size kernels cannot run: without the JIT build there is nothing valid to fall back to
        if (force_jit && !(p->k_r2c.is_jit && p->k_cols.is_jit && p->k_c2r.is_jit))
            return fail(B2R_ERR_UNSUPPORTED, "B2R_FORCE_JIT: plan-time JIT failed (%s)", p->jit_note.c_str());
    }
    if (dbl && !(p->k_r2c.is_jit && p->k_cols.is_jit && p->k_c2r.is_jit))
        return fail(B2R_ERR_UNSUPPORTED, "precision 1 (double) needs the plan-time JIT: %s", p->jit_note.c_str());
    const int lim_c = p->k_cols.is_static ? 1024 : kDynMaxThreads;
    const int lim_1 = p->k_r2c.is_static ? 1024 : kDynMaxThreads, lim_7 = p->k_c2r.is_static ? 1024 : kDynMaxThreads;
    if (p->k_cols.cc * p->k_cols.inv.threads > lim_c || p->k_cols.smem > smem_max)
        return fail(B2R_ERR_UNSUPPORTED, "column transform %d x %d does not fit one CTA (%zu B shared, %d threads)",
                    g.up_h, p->k_cols.cc, p->k_cols.smem, p->k_cols.cc * p->k_cols.inv.threads);
    if (p->k_r2c.sched.threads * p->k_r2c.ppb > lim_1 || p->k_c2r.sched.threads * p->k_c2r.ppb > lim_7 ||
        p->k_r2c.smem > smem_max || p->k_c2r.smem > smem_max)
        return fail(B2R_ERR_UNSUPPORTED, "row transform %d / %d does not fit one CTA", g.w, g.up_w);
    CU(p->k_r2c.prepare(p->k_r2c.smem, p->k_r2c.ctx));
    CU(p->k_c2r.prepare(p->k_c2r.smem, p->k_c2r.ctx));
    if (p->c2c) {
        if (!p->k_c2r.is_static) p->k_c2r.smem_c2c = (size_t)p->k_c2r.ppb_c2c * smem_padded_len(g.up_w) * cb;
        if (p->k_c2r.sched.threads * p->k_c2r.ppb_c2c > lim_7 || p->k_c2r.smem_c2c > smem_max)
            return fail(B2R_ERR_UNSUPPORTED, "row transform %d does not fit one CTA", g.up_w);
        CU(p->k_c2r.prepare_c2c(p->k_c2r.smem_c2c, p->k_c2r.ctx));
    }
    CU(p->k_cols.prepare(p->k_cols.smem, p->k_cols.ctx));

    // kernel-side dimensions
    FrameDims& d = p->dm;
    d.w = g.w; d.h = g.h; d.up_w = g.up_w; d.up_h = g.up_h; d.nx = g.nx; d.spec_stride = g.spec_stride;
    d.zp_lo = g.zp_lo; d.zp_hi = g.zp_hi; d.neg_shift = g.neg_shift;
    d.in_plane = g.in_plane; d.pre_plane = g.pre_plane; d.out_plane = g.out_plane;
    const bool raw = p->flags & B2R_FLAG_NO_SHARPEN_LITERAL_ROUNDING;
    d.up2 = raw ? g.up2 : literal_f(g.up2);
    d.sharpen = raw ? g.sharpen : literal_f(g.sharpen);
    { const CasK k = cas_k(d.sharpen); d.cas_a = k.a; d.cas_b = k.b; }
    d.up2_d = raw ? (double)g.up2 : literal_d(g.up2);
    d.sharpen_d = raw ? (double)g.sharpen : literal_d(g.sharpen);

    // ---- fused C2R + sharpen: whenever the tolerance-bound sharpen applies and the row schedule is built in
    // (B2R_FUSED=0 / B2R_FLAG_SEPARATE_SHARPEN keep the two kernels; B2R_FUSED_NSP overrides the strip count)
    {
        SharpenArgs sa{nullptr, nullptr, p->dm, g.precision};
        sa.exact = (p->flags & B2R_FLAG_EXACT_SHARPEN) != 0;
        const int ppp = g.up_h / 2;
        const bool want = !(p->flags & B2R_FLAG_SEPARATE_SHARPEN) && env_int("B2R_FUSED", 1) != 0;
        bool ok = want && !p->c2c && (g.precision == 0 || g.precision == 2) && sharpen_fast_applies(sa) && p->k_c2r.fused && !p->k_c2r.is_jit &&
                  ppp >= 6 && fused_smem_bytes(g.up_w, g.nx) <= smem_max;
        if (ok && p->k_c2r.prepare_fused(g.precision, g.nx) != cudaSuccess) { (void)cudaGetLastError(); ok = false; }
        if (ok) {
            // measured on B200 (profiles/README.md): the fused kernel wins for fp32 while two strip CTAs fit one SM
            // (4096- and 3840-wide rows: +5 % sustained, half the HBM traffic, no power-cap throttling); with a single
            // resident CTA (7680-wide rows: 349 vs 307 us) and for fp16 storage (whose stand-alone sharpen moves half
            // the bytes: c4 80 vs 64 us) the separate kernels are faster.  B2R_FUSED=2 forces the fused kernel.
            const int per_sm = p->k_c2r.fused_blocks_per_sm(g.precision, g.nx);
            const bool forced = env_int("B2R_FUSED", 1) == 2;
            if ((per_sm >= 2 && g.precision == 0) || (per_sm >= 1 && forced)) {
                const int slots = prop.multiProcessorCount * per_sm;
                int nsp = env_int("B2R_FUSED_NSP", 0);
                if (nsp <= 0) nsp = std::max(1, slots / 3);          // one wave of strip CTAs over the three planes
                nsp = std::min(nsp, ppp / 3);                         // at least 3 pairs per strip
                p->fused_nsp = std::max(1, nsp);
            }
        }
    }

    // device memory
    const size_t eb = g.elem_bytes();
    const size_t b_in = g.input_bytes(), b_pre = g.pre_elems * eb, b_out = g.output_bytes();
    const size_t b_s1 = g.spec_in_elems() * cb, b_s2 = g.spec_out_elems() * cb;
    const size_t n_tw = p->fw.twiddles.size() + p->fh.twiddles.size() + p->fuh.twiddles.size() + p->fuw.twiddles.size();
    CU(cudaMalloc(&p->d_in, b_in));
    CU(cudaMalloc(&p->d_pre, b_pre));
    CU(cudaMalloc(&p->d_out, b_out));
    CU(cudaMalloc((void**)&p->d_spec1, b_s1));
    CU(cudaMalloc((void**)&p->d_spec2, b_s2));
    CU(cudaMalloc((void**)&p->d_tw, (n_tw + 1) * cb));
    CU(cudaMalloc((void**)&p->d_fd, 4 * sizeof(FftDesc)));
    if (p->c2c) {
        CU(cudaMalloc((void**)&p->d_nyq, 3 * (size_t)g.spec_stride * cb));
        CU(cudaMemset(p->d_nyq, 0, 3 * (size_t)g.spec_stride * cb));
    }
    p->device_bytes = b_in + b_pre + b_out + b_s1 + b_s2 + (n_tw + 1) * cb + 4 * sizeof(FftDesc);
    CU(cudaMemset(p->d_in, 0, b_in));
    CU(cudaMemset(p->d_pre, 0, b_pre));  // the plane pad regions stay zero for the plan's lifetime
    CU(cudaMemset(p->d_spec1, 0, b_s1));
    CU(cudaMemset(p->d_spec2, 0, b_s2));
    const FftDesc descs[4] = {p->fw.desc, p->fh.desc, p->fuh.desc, p->fuw.desc};
    CU(cudaMemcpy(p->d_fd, descs, sizeof descs, cudaMemcpyHostToDevice));
    size_t off = 0;
    auto put = [&](const HostFft& f, const float2** slot) -> int {   // fp64: the table holds double2 (byte offsets scale with cb)
        char* base = reinterpret_cast<char*>(p->d_tw) + off * cb;
        *slot = reinterpret_cast<const float2*>(base);
        if (!f.twiddles.empty()) {
            if (dbl) CU(cudaMemcpy(base, f.twiddles_d.data(), f.twiddles_d.size() * sizeof(double2), cudaMemcpyHostToDevice));
            else CU(cudaMemcpy(base, f.twiddles.data(), f.twiddles.size() * sizeof(float2), cudaMemcpyHostToDevice));
        }
        off += f.twiddles.size();
        return B2R_SUCCESS;
    };
    int rc;
    if ((rc = put(p->fw, &p->tw_w)) || (rc = put(p->fh, &p->tw_h)) || (rc = put(p->fuh, &p->tw_uh)) ||
        (rc = put(p->fuw, &p->tw_uw)))
        return rc;

    // exact-2x column kernel: built-in schedule pairs with upH == 2H, single precision (B2R_COLS_2X=0 keeps k_cols)
    if (p->k_cols.launch2x) {   // (set by the static registry only: upH == 2H, single precision, B2R_COLS_2X != 0)
        if (p->k_cols.smem2x > smem_max) return fail(B2R_ERR_UNSUPPORTED, "column tile does not fit one CTA");
        CU(p->k_cols.prepare2x(p->k_cols.smem2x, p->k_cols.ctx));
        std::vector<float2> ramp((size_t)g.h);
        for (int k = 0; k < g.h; ++k) {
            const int ks = (k < g.h / 2) ? k : k - g.h;       // signed frequency after the reference's shift
            const double a = 3.14159265358979323846 * (double)ks / (double)g.h;
            ramp[(size_t)k] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
        CU(cudaMalloc((void**)&p->d_ramp, ramp.size() * sizeof(float2)));
        CU(cudaMemcpy(p->d_ramp, ramp.data(), ramp.size() * sizeof(float2), cudaMemcpyHostToDevice));
    }
    if (p->fused_nsp > 0) {   // rows k_sharpen_fix finishes, per plane (b2r_fused.cuh)
        std::vector<int> fix;
        const int ppp = g.up_h / 2, nsp = p->fused_nsp;
        for (int q = 1; q < nsp; ++q) fix.push_back(2 * fused_strip_begin(q, nsp, ppp));   // first row of strip q
        fix.push_back(g.up_h);                                       // plane end: the row below is the pad region
        p->n_fix = (int)fix.size();
        CU(cudaMalloc((void**)&p->d_fix, fix.size() * sizeof(int)));
        CU(cudaMemcpy(p->d_fix, fix.data(), fix.size() * sizeof(int), cudaMemcpyHostToDevice));
    }

    CU(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&p->ev0));
    CU(cudaEventCreate(&p->ev1));

    if (!(p->flags & B2R_FLAG_NO_GRAPH)) {
        CU(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
        rc = launch_frame(p, p->stream);
        cudaError_t e = cudaStreamEndCapture(p->stream, &p->graph);
        if (rc) return rc;
        CU(e);
        CU(cudaGraphInstantiate(&p->graph_exec, p->graph, 0));
    }
    CU(cudaStreamSynchronize(p->stream));
    return B2R_SUCCESS;
}

}  // namespace

extern "C" {

const char* b2r_last_error(void) { return g_err.c_str(); }
const char* b2r_version(void) { return "b2resample 0.2 (sm_100a)"; }

int b2r_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { fail(B2R_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e)); return 0; }
    return n;
}

int b2r_device_name(int device, char* buf, size_t buf_len) {
    if (!buf || !buf_len) return fail(B2R_ERR_INVALID_ARG, "null buffer");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    snprintf(buf, buf_len, "%s", prop.name);
    return B2R_SUCCESS;
}

int b2r_plan_create(b2r_plan** out, int device, uint32_t w, uint32_t h, float upscale, uint32_t precision,
                    float sharpen, uint32_t flags) {
    if (!out) return fail(B2R_ERR_INVALID_ARG, "out is null");
    *out = nullptr;
    Geometry g;
    std::string err;
    const bool c2c = (flags & B2R_FLAG_C2C_PARITY) != 0;
    if (!make_geometry((int)w, (int)h, upscale, (int)precision, sharpen, &g, &err, c2c))
        return fail(err.find("not of the form") != std::string::npos ? B2R_ERR_UNSUPPORTED : B2R_ERR_INVALID_ARG, "%s", err.c_str());
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(B2R_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return fail(B2R_ERR_INVALID_ARG, "device %d out of range [0,%d)", device, ndev);
    DeviceGuard guard(device);   // the caller's current device is restored on return
    if (!guard.ok()) return fail(B2R_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    b2r_plan* p = new (std::nothrow) b2r_plan();
    if (!p) return fail(B2R_ERR_NOMEM, "out of host memory");
    p->device = device; p->flags = flags; p->g = g; p->c2c = c2c;
    int rc = build(p);
    if (rc) { std::string keep = g_err; b2r_plan_destroy(p); g_err = keep; return rc; }
    *out = p;
    return B2R_SUCCESS;
}

void b2r_plan_destroy(b2r_plan* p) {
    if (!p) return;
    DeviceGuard guard(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (Lane& l : p->extra) {
        if (l.stream) { cudaStreamSynchronize(l.stream); cudaStreamDestroy(l.stream); }
        if (l.done) cudaEventDestroy(l.done);
        cudaFree(l.d_in); cudaFree(l.d_pre); cudaFree(l.d_out); cudaFree(l.d_spec1); cudaFree(l.d_spec2);
        cudaFree(l.u8_in); cudaFree(l.u8_out); cudaFree(l.d_nyq);
    }
    cudaFree(p->u8_in0); cudaFree(p->u8_out0);
    for (auto& ring : p->tick_ev) for (auto e : ring) if (e) cudaEventDestroy(e);
    p->tick_ev.clear();
    p->extra.clear();
    if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
    if (p->graph) cudaGraphDestroy(p->graph);
    jit_destroy(p->jit);
    p->jit = nullptr;
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->stream) cudaStreamDestroy(p->stream);
    cudaFree(p->d_in); cudaFree(p->d_pre); cudaFree(p->d_out);
    cudaFree(p->d_spec1); cudaFree(p->d_spec2); cudaFree(p->d_tw); cudaFree(p->d_fd); cudaFree(p->d_nyq);
    cudaFree(p->d_fix); cudaFree(p->d_ramp);
    delete p;
}

size_t b2r_plan_input_bytes(const b2r_plan* p) { return p ? p->g.input_bytes() : 0; }
size_t b2r_plan_output_bytes(const b2r_plan* p) { return p ? p->g.output_bytes() : 0; }
size_t b2r_plan_pre_sharpen_bytes(const b2r_plan* p) { return p ? 3 * p->g.pre_plane * p->g.elem_bytes() : 0; }

int b2r_plan_get_info(const b2r_plan* p, b2r_plan_info* info) {
    if (!p || !info) return fail(B2R_ERR_INVALID_ARG, "null argument");
    memset(info, 0, sizeof *info);
    const Geometry& g = p->g;
    info->w = g.w; info->h = g.h; info->up_w = g.up_w; info->up_h = g.up_h;
    info->precision = g.precision; info->upscale = g.upscale; info->sharpen = g.sharpen;
    info->zeropad_lo_y = g.zp_lo; info->zeropad_hi_y = g.zp_hi;
    info->spectrum_row_stride = g.spec_stride;
    info->input_bytes = g.input_bytes(); info->output_bytes = g.output_bytes();
    info->device_bytes = p->device_bytes;
    const HostFft* f[4] = {&p->fw, &p->fh, &p->fuh, &p->fuw};
    // thread counts of the kernels that actually launch (the JIT chooser may halve the seed's counts)
    const int launched[4] = {p->k_r2c.sched.threads, p->k_cols.fwd.threads, p->k_cols.inv.threads, p->k_c2r.sched.threads};
    for (int i = 0; i < 4; ++i) {
        info->n_stages[i] = f[i]->desc.nstages;
        info->threads[i] = launched[i] > 0 ? launched[i] : f[i]->desc.threads;
        for (int s = 0; s < f[i]->desc.nstages; ++s) info->radices[i][s] = f[i]->desc.st[s].radix;
    }
    info->c2c_mode = p->c2c ? 1u : 0u;
    info->pre_sharpen_plane_stride = g.pre_plane;
    info->column_tile = p->k_cols.cc;
    info->static_kernels = (p->k_r2c.is_static ? 1u : 0u) | (p->k_cols.is_static ? 2u : 0u) | (p->k_c2r.is_static ? 4u : 0u);
    info->jit_kernels = (p->k_r2c.is_jit ? 1u : 0u) | (p->k_cols.is_jit ? 2u : 0u) | (p->k_c2r.is_jit ? 4u : 0u);
    snprintf(info->jit_note, sizeof info->jit_note, "%s", p->jit_note.c_str());
    info->kernels_per_frame = p->kernels_per_frame;
    info->fused_strips_per_plane = (uint32_t)p->fused_nsp;
    {
        SharpenArgs sa{nullptr, nullptr, p->dm, g.precision};
        sa.exact = (p->flags & B2R_FLAG_EXACT_SHARPEN) != 0;
        info->sharpen_mode = p->c2c ? (sa.exact ? 0u : (sharpen_fast_applies(sa) ? 1u : 0u)) : (sharpen_fast_applies(sa) ? 1u : 0u);
    }
    return B2R_SUCCESS;
}

int b2r_upload(b2r_plan* p, const void* host_in) {
    if (!p || !host_in) return fail(B2R_ERR_INVALID_ARG, "null argument");
    ON_PLAN_DEVICE(p);
    CU(cudaMemcpyAsync(p->d_in, host_in, p->g.input_bytes(), cudaMemcpyHostToDevice, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    return B2R_SUCCESS;
}

int b2r_execute(b2r_plan* p, uint32_t num_iter, double* ms_per_iter) {
    if (!p || num_iter == 0) return fail(B2R_ERR_INVALID_ARG, "null plan or num_iter == 0");
    ON_PLAN_DEVICE(p);
    CU(cudaEventRecord(p->ev0, p->stream));
    for (uint32_t i = 0; i < num_iter; ++i) {
        int rc = run_frame(p);
        if (rc) return rc;
    }
    CU(cudaEventRecord(p->ev1, p->stream));
    CU(cudaEventSynchronize(p->ev1));
    CU(cudaGetLastError());
    if (ms_per_iter) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
        *ms_per_iter = (double)ms / (double)num_iter;
    }
    return B2R_SUCCESS;
}

int b2r_download(b2r_plan* p, void* host_out) {
    if (!p || !host_out) return fail(B2R_ERR_INVALID_ARG, "null argument");
    ON_PLAN_DEVICE(p);
    CU(cudaMemcpyAsync(host_out, p->d_out, p->g.output_bytes(), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    return B2R_SUCCESS;
}

int b2r_upscale_host(b2r_plan* p, const void* host_in, void* host_out, double* ms_total) {
    if (!p || !host_in || !host_out) return fail(B2R_ERR_INVALID_ARG, "null argument");
    ON_PLAN_DEVICE(p);
    CU(cudaEventRecord(p->ev0, p->stream));
    CU(cudaMemcpyAsync(p->d_in, host_in, p->g.input_bytes(), cudaMemcpyHostToDevice, p->stream));
    int rc = run_frame(p);
    if (rc) return rc;
    CU(cudaMemcpyAsync(host_out, p->d_out, p->g.output_bytes(), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaEventRecord(p->ev1, p->stream));
    CU(cudaEventSynchronize(p->ev1));
    if (ms_total) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
        *ms_total = ms;
    }
    return B2R_SUCCESS;
}

void* b2r_device_input(b2r_plan* p) { return p ? p->d_in : nullptr; }
void* b2r_device_output(b2r_plan* p) { return p ? p->d_out : nullptr; }
void* b2r_plan_stream(b2r_plan* p) { return p ? (void*)p->stream : nullptr; }
uint64_t b2r_plan_launch_count(const b2r_plan* p) { return p ? p->launches : 0; }

int b2r_download_pre_sharpen(b2r_plan* p, void* host_out) {
    if (!p || !host_out) return fail(B2R_ERR_INVALID_ARG, "null argument");
    ON_PLAN_DEVICE(p);
    if (p->fused_nsp > 0) {
        // a fused plan never holds the whole plane: rebuild it with the stand-alone C2R kernel from the column
        // spectra of the last frame, which are still resident (same arithmetic, same values)
        C2rArgs a3{p->d_spec2, p->d_pre, p->tw_uw, p->d_fd + 3, p->dm, p->g.precision, 1.0f / (float)p->g.up_w, nullptr};
        CU(p->k_c2r.c2r(p->stream, a3, p->k_c2r.sched.threads, p->k_c2r.smem, p->k_c2r.ctx));
        p->launches += 1;
    }
    CU(cudaMemcpyAsync(host_out, p->d_pre, b2r_plan_pre_sharpen_bytes(p), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    return B2R_SUCCESS;
}

int b2r_sharpen_host(b2r_plan* p, const void* host_pre, void* host_out) {
    if (!p || !host_pre || !host_out) return fail(B2R_ERR_INVALID_ARG, "null argument");
    ON_PLAN_DEVICE(p);
    CU(cudaMemcpyAsync(p->d_pre, host_pre, b2r_plan_pre_sharpen_bytes(p), cudaMemcpyHostToDevice, p->stream));
    int rc = launch_sharpen(p, p->stream);
    if (rc) return rc;
    p->launches += 1;
    CU(cudaMemcpyAsync(host_out, p->d_out, p->g.output_bytes(), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    return B2R_SUCCESS;
}

namespace {
void free_lane(Lane& l) {
    if (l.stream) { cudaStreamSynchronize(l.stream); cudaStreamDestroy(l.stream); }
    if (l.done) cudaEventDestroy(l.done);
    cudaFree(l.d_in); cudaFree(l.d_pre); cudaFree(l.d_out); cudaFree(l.d_spec1); cudaFree(l.d_spec2);
    cudaFree(l.u8_in); cudaFree(l.u8_out); cudaFree(l.d_nyq);
    l = Lane{};
}
size_t lane_bytes(const Geometry& g) {
    return g.input_bytes() + g.pre_elems * g.elem_bytes() + g.output_bytes() + (g.spec_in_elems() + g.spec_out_elems()) * g.cplx_bytes();
}
// allocates every resource of one extra lane; on failure whatever was created is released again
int make_lane(b2r_plan* p, Lane* out) {
    const Geometry& g = p->g;
    const size_t eb = g.elem_bytes(), cb = g.cplx_bytes();
    Lane l;
    auto body = [&]() -> int {
        CU(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
        CU(cudaMalloc(&l.d_in, g.input_bytes()));
        CU(cudaMalloc(&l.d_pre, g.pre_elems * eb));
        CU(cudaMalloc(&l.d_out, g.output_bytes()));
        CU(cudaMalloc((void**)&l.d_spec1, g.spec_in_elems() * cb));
        CU(cudaMalloc((void**)&l.d_spec2, g.spec_out_elems() * cb));
        CU(cudaMemset(l.d_in, 0, g.input_bytes()));
        CU(cudaMemset(l.d_pre, 0, g.pre_elems * eb));
        CU(cudaMemset(l.d_spec1, 0, g.spec_in_elems() * cb));
        CU(cudaMemset(l.d_spec2, 0, g.spec_out_elems() * cb));
        if (p->c2c) {
            CU(cudaMalloc((void**)&l.d_nyq, 3 * (size_t)g.spec_stride * cb));
            CU(cudaMemset(l.d_nyq, 0, 3 * (size_t)g.spec_stride * cb));
        }
        return B2R_SUCCESS;
    };
    const int rc = body();
    if (rc) { free_lane(l); return rc; }
    *out = l;
    return B2R_SUCCESS;
}
}  // namespace

int b2r_plan_set_lanes(b2r_plan* p, uint32_t lanes) {
    if (!p || lanes < 1 || lanes > 8) return fail(B2R_ERR_INVALID_ARG, "lanes must be in [1, 8]");
    DeviceGuard guard(p->device);
    if (!guard.ok()) return fail(B2R_ERR_CUDA, "cudaSetDevice(%d) failed", p->device);
    int rc = b2r_synchronize(p);   // nothing may be in flight while lanes come and go
    if (rc) return rc;
    while (p->num_lanes() < lanes) {
        Lane l;
        if ((rc = make_lane(p, &l))) return rc;
        p->device_bytes += lane_bytes(p->g);
        p->extra.push_back(l);
    }
    while (p->num_lanes() > lanes) {   // shrink: release the highest lanes and their staging buffers
        Lane& l = p->extra.back();
        if (l.u8_in) p->device_bytes -= b2r_plan_input_u8_bytes(p) + b2r_plan_output_u8_bytes(p);
        free_lane(l);
        p->device_bytes -= lane_bytes(p->g);
        p->extra.pop_back();
    }
    {
        std::lock_guard<std::mutex> g(p->tick_mu);   // tickets refer to lanes: start afresh
        for (auto& ring : p->tick_ev) for (auto e : ring) if (e) cudaEventDestroy(e);
        p->tick_ev.clear(); p->tick_pos.clear();
        for (auto& e : p->tick_log) e = b2r_plan::TickEntry{};
    }
    p->next_lane = 0;
    CU(cudaDeviceSynchronize());
    return B2R_SUCCESS;
}

uint32_t b2r_plan_lanes(const b2r_plan* p) { return p ? p->num_lanes() : 0; }

namespace {
// records the completion event of the frame just enqueued on lane li and advances the ticket counter
int stamp_ticket(b2r_plan* p, uint32_t li, cudaStream_t s) {
    const uint32_t nl = p->num_lanes();
    std::lock_guard<std::mutex> g(p->tick_mu);
    if (p->tick_ev.size() < nl) { p->tick_ev.resize(nl); p->tick_pos.resize(nl, 0); }
    auto& ring = p->tick_ev[li];
    if (ring.empty()) {
        ring.resize(kTicketDepth, nullptr);
        for (auto& e : ring) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaEvent_t ev = ring[p->tick_pos[li]++ % kTicketDepth];
    CU(cudaEventRecord(ev, s));
    const uint64_t t = ++p->tickets;
    p->tick_log[t % 64] = b2r_plan::TickEntry{t, ev};
    return B2R_SUCCESS;
}
}  // namespace

int b2r_enqueue_device(b2r_plan* p, const void* d_in, void* d_out) {
    if (!p || !d_in || !d_out) return fail(B2R_ERR_INVALID_ARG, "null argument");
    ON_PLAN_DEVICE(p);
    const uint32_t li = p->next_lane++ % p->num_lanes();
    const Lane l = p->lane(li);
    int rc = launch_frame(p, l.stream, d_in, d_out, nullptr, li ? &l : nullptr);
    if (rc) return rc;
    p->launches += p->kernels_per_frame;
    return B2R_SUCCESS;
}

int b2r_enqueue_host(b2r_plan* p, const void* host_in, void* host_out) {
    if (!p || !host_in || !host_out) return fail(B2R_ERR_INVALID_ARG, "null argument");
    ON_PLAN_DEVICE(p);
    const uint32_t li = p->next_lane++ % p->num_lanes();
    const Lane l = p->lane(li);
    CU(cudaMemcpyAsync(l.d_in, host_in, p->g.input_bytes(), cudaMemcpyHostToDevice, l.stream));
    int rc = launch_frame(p, l.stream, l.d_in, l.d_out, nullptr, li ? &l : nullptr);
    if (rc) return rc;
    p->launches += p->kernels_per_frame;
    CU(cudaMemcpyAsync(host_out, l.d_out, p->g.output_bytes(), cudaMemcpyDeviceToHost, l.stream));
    return stamp_ticket(p, li, l.stream);
}

size_t b2r_plan_input_u8_bytes(const b2r_plan* p) { return p ? 3ull * p->g.w * p->g.h : 0; }
size_t b2r_plan_output_u8_bytes(const b2r_plan* p) { return p ? 3ull * p->g.up_w * p->g.up_h : 0; }

namespace {
int u8_buffers(b2r_plan* p, uint32_t li, unsigned char** in, unsigned char** out) {
    unsigned char** pin = li ? &p->extra[li - 1].u8_in : &p->u8_in0;
    unsigned char** pout = li ? &p->extra[li - 1].u8_out : &p->u8_out0;
    if (!*pin) {   // both buffers exist before either pointer is published
        unsigned char *a = nullptr, *b = nullptr;
        CU(cudaMalloc((void**)&a, b2r_plan_input_u8_bytes(p)));
        cudaError_t e2 = cudaMalloc((void**)&b, b2r_plan_output_u8_bytes(p));
        if (e2 != cudaSuccess) { cudaFree(a); return fail(B2R_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e2)); }
        *pin = a; *pout = b;
        p->device_bytes += b2r_plan_input_u8_bytes(p) + b2r_plan_output_u8_bytes(p);
    }
    *in = *pin; *out = *pout;
    return B2R_SUCCESS;
}
}  // namespace

int b2r_upload_u8(b2r_plan* p, const unsigned char* host_hwc) {
    if (!p || !host_hwc) return fail(B2R_ERR_INVALID_ARG, "null argument");
    ON_PLAN_DEVICE(p);
    unsigned char *in, *out;
    int rc = u8_buffers(p, 0, &in, &out);
    if (rc) return rc;
    CU(cudaMemcpyAsync(in, host_hwc, b2r_plan_input_u8_bytes(p), cudaMemcpyHostToDevice, p->stream));
    if (p->g.precision == 1) CU(jit_launch_u8_to_planar(p->jit, p->stream, in, p->d_in, p->dm));
    else CU(launch_u8_to_planar(p->stream, in, p->d_in, p->dm, p->g.precision));
    p->launches += 1;
    CU(cudaStreamSynchronize(p->stream));
    return B2R_SUCCESS;
}

int b2r_download_u8(b2r_plan* p, unsigned char* host_hwc) {
    if (!p || !host_hwc) return fail(B2R_ERR_INVALID_ARG, "null argument");
    ON_PLAN_DEVICE(p);
    unsigned char *in, *out;
    int rc = u8_buffers(p, 0, &in, &out);
    if (rc) return rc;
    if (p->g.precision == 1) CU(jit_launch_planar_to_u8(p->jit, p->stream, p->d_out, out, p->dm));
    else CU(launch_planar_to_u8(p->stream, p->d_out, out, p->dm, p->g.precision));
    p->launches += 1;
    CU(cudaMemcpyAsync(host_hwc, out, b2r_plan_output_u8_bytes(p), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    return B2R_SUCCESS;
}

int b2r_enqueue_host_u8(b2r_plan* p, const unsigned char* host_in, unsigned char* host_out) {
    if (!p || !host_in || !host_out) return fail(B2R_ERR_INVALID_ARG, "null argument");
    ON_PLAN_DEVICE(p);
    const uint32_t li = p->next_lane++ % p->num_lanes();
    const Lane l = p->lane(li);
    unsigned char *in, *out;
    int rc = u8_buffers(p, li, &in, &out);
    if (rc) return rc;
    CU(cudaMemcpyAsync(in, host_in, b2r_plan_input_u8_bytes(p), cudaMemcpyHostToDevice, l.stream));
    if (p->g.precision == 1) CU(jit_launch_u8_to_planar(p->jit, l.stream, in, l.d_in, p->dm));
    else CU(launch_u8_to_planar(l.stream, in, l.d_in, p->dm, p->g.precision));
    rc = launch_frame(p, l.stream, l.d_in, l.d_out, nullptr, li ? &l : nullptr);
    if (rc) return rc;
    if (p->g.precision == 1) CU(jit_launch_planar_to_u8(p->jit, l.stream, l.d_out, out, p->dm));
    else CU(launch_planar_to_u8(l.stream, l.d_out, out, p->dm, p->g.precision));
    p->launches += p->kernels_per_frame + 2;
    CU(cudaMemcpyAsync(host_out, out, b2r_plan_output_u8_bytes(p), cudaMemcpyDeviceToHost, l.stream));
    return stamp_ticket(p, li, l.stream);
}

uint64_t b2r_plan_last_ticket(const b2r_plan* p) { return p ? p->tickets : 0; }

int b2r_wait_ticket(b2r_plan* p, uint64_t ticket) {
    if (!p || ticket == 0) return fail(B2R_ERR_INVALID_ARG, "no such ticket");
    cudaEvent_t ev = nullptr;
    {
        std::lock_guard<std::mutex> g(p->tick_mu);
        if (ticket > p->tickets) return fail(B2R_ERR_INVALID_ARG, "no such ticket");
        if (p->tick_log[ticket % 64].ticket == ticket) ev = p->tick_log[ticket % 64].ev;
    }
    if (ev) { CU(cudaEventSynchronize(ev)); return B2R_SUCCESS; }
    return b2r_synchronize(p);   // older than the log: everything enqueued so far covers it
}

void* b2r_host_alloc(size_t bytes) {
    void* ptr = nullptr;
    if (cudaHostAlloc(&ptr, bytes, cudaHostAllocPortable) != cudaSuccess) {
        fail(B2R_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return ptr;
}
void b2r_host_free(void* ptr) { if (ptr) cudaFreeHost(ptr); }

int b2r_timer_start(b2r_plan* p) {
    if (!p) return fail(B2R_ERR_INVALID_ARG, "null plan");
    ON_PLAN_DEVICE(p);
    CU(cudaEventRecord(p->ev0, p->stream));
    for (Lane& l : p->extra) CU(cudaStreamWaitEvent(l.stream, p->ev0, 0));   // every lane starts after t0
    return B2R_SUCCESS;
}

int b2r_timer_stop(b2r_plan* p, double* ms) {
    if (!p) return fail(B2R_ERR_INVALID_ARG, "null plan");
    ON_PLAN_DEVICE(p);
    for (Lane& l : p->extra) {                                                // t1 is after every lane's tail
        CU(cudaEventRecord(l.done, l.stream));
        CU(cudaStreamWaitEvent(p->stream, l.done, 0));
    }
    CU(cudaEventRecord(p->ev1, p->stream));
    CU(cudaEventSynchronize(p->ev1));
    CU(cudaGetLastError());
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, p->ev0, p->ev1));
    if (ms) *ms = t;
    return B2R_SUCCESS;
}

int b2r_profile_kernels(b2r_plan* p, uint32_t num_iter, double* ms_per_kernel) {
    if (!p || !num_iter || !ms_per_kernel) return fail(B2R_ERR_INVALID_ARG, "null argument");
    ON_PLAN_DEVICE(p);
    struct Events {   // destroyed on every return path
        std::vector<cudaEvent_t> v;
        ~Events() { for (auto e : v) if (e) cudaEventDestroy(e); }
    } evs;
    evs.v.assign(5 * (size_t)num_iter, nullptr);
    std::vector<cudaEvent_t>& ev = evs.v;
    for (auto& e : ev) CU(cudaEventCreate(&e));
    for (uint32_t i = 0; i < num_iter; ++i) {
        int rc = launch_frame(p, p->stream, nullptr, nullptr, &ev[5 * (size_t)i]);
        if (rc) return rc;
        p->launches += p->kernels_per_frame;
    }
    CU(cudaStreamSynchronize(p->stream));
    for (int k = 0; k < 4; ++k) ms_per_kernel[k] = 0.0;
    for (uint32_t i = 0; i < num_iter; ++i)
        for (int k = 0; k < 4; ++k) {
            float t = 0.f;
            CU(cudaEventElapsedTime(&t, ev[5 * (size_t)i + k], ev[5 * (size_t)i + k + 1]));
            ms_per_kernel[k] += (double)t / num_iter;
        }
    return B2R_SUCCESS;
}

int b2r_synchronize(b2r_plan* p) {
    if (!p) return fail(B2R_ERR_INVALID_ARG, "null plan");
    ON_PLAN_DEVICE(p);
    for (Lane& l : p->extra) CU(cudaStreamSynchronize(l.stream));
    CU(cudaStreamSynchronize(p->stream));
    return B2R_SUCCESS;
}

}  // extern "C"
