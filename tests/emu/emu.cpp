// tests/emu/emu.cpp -- CPU thread emulator for the CUDA kernels.  TEST INFRASTRUCTURE ONLY.
//
// The authoring container has no GPU.  To check the index arithmetic of every kernel before
// spending GPU time, this file compiles the *same* kernel source (b2r_kernels.cuh) as plain C++
// (-DB2R_HOST_EMU): each CUDA thread of a CTA becomes an OS thread, __syncthreads() a
// std::barrier, dynamic shared memory a heap block.  CTAs run one after another.  It is built
// into tests/emu/libb2r_emu.so by tests/emu/build.sh and driven by tests/test_emu_kernels.py.
// The product library (libb2resample.so) never links or calls any of this.
#include <barrier>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../vkresample_b200/csrc/b2r_kernels.cuh"
#include "../../vkresample_b200/csrc/b2r_plan.h"
#include "../../vkresample_b200/csrc/b2r_static_sizes.h"

namespace b2r_emu {
thread_local Ctx g_ctx;

struct Dim3 { unsigned x = 1, y = 1, z = 1; };

struct Barriers {
    std::barrier<> all;
    std::mutex mu;
    std::map<int, std::unique_ptr<std::barrier<>>> named;   // created on first use with that id's count
    explicit Barriers(std::ptrdiff_t n) : all(n) {}
    std::barrier<>* get(int id, int count) {
        std::lock_guard<std::mutex> g(mu);
        auto& b = named[id];
        if (!b) b = std::make_unique<std::barrier<>>(count);
        return b.get();
    }
};

static void launch(Dim3 grid, Dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    const unsigned nthreads = block.x * block.y;
    Barriers bar((std::ptrdiff_t)nthreads);
    std::vector<unsigned char> smem(smem_bytes + 64, 0);
    auto sync_fn = [](void* p) { static_cast<Barriers*>(p)->all.arrive_and_wait(); };
    auto sync_group_fn = [](void* p, int id, int count) { static_cast<Barriers*>(p)->get(id, count)->arrive_and_wait(); };
    auto worker = [&](unsigned t) {
        for (unsigned bz = 0; bz < grid.z; ++bz)
            for (unsigned by = 0; by < grid.y; ++by)
                for (unsigned bx = 0; bx < grid.x; ++bx) {
                    Ctx& c = g_ctx;
                    c.tid_x = t % block.x; c.tid_y = t / block.x;
                    c.bid_x = bx; c.bid_y = by; c.bid_z = bz;
                    c.bdim_x = block.x; c.bdim_y = block.y;
                    c.gdim_x = grid.x; c.gdim_y = grid.y; c.gdim_z = grid.z;
                    c.smem = smem.data();
                    c.sync = sync_fn; c.sync_arg = &bar; c.sync_group = sync_group_fn;
                    body();
                    bar.all.arrive_and_wait();  // CTA boundary: shared memory is reused
                }
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
    for (auto& x : th) x.join();
}
}  // namespace b2r_emu

using namespace b2r;

// generic single-sequence complex FFT through the stage engine (tests every radix / size)
template <int DIR, class P>
static void k_fft_test(const float2* in, float2* out, const float2* tw, const P plan) {
    const int T = (int)B2R_BDIM_X, tid = (int)B2R_TID_X;
    float2* sm = B2R_SMEM(float2);
    for (int i = tid; i < plan.n(); i += T) sm[smem_pad(i)] = in[i];
    B2R_SYNC();
    plan.template for_stages<0, 0>([&](auto st, int) {
        using St = decltype(st);
        float2 v[St::NB][St::R];
        stage_load_compute<DIR, 1>(st, sm, tw, T, tid, 0, v);
        B2R_SYNC();
        stage_store<1>(st, sm, T, tid, 0, v);
        B2R_SYNC();
    });
    for (int i = tid; i < plan.n(); i += T) out[i] = sm[smem_pad(i)];
}

static FrameDims dims_of(const Geometry& g) {
    FrameDims d{};
    d.w = g.w; d.h = g.h; d.up_w = g.up_w; d.up_h = g.up_h; d.nx = g.nx; d.spec_stride = g.spec_stride;
    d.zp_lo = g.zp_lo; d.zp_hi = g.zp_hi; d.neg_shift = g.neg_shift;
    d.in_plane = g.in_plane; d.pre_plane = g.pre_plane; d.out_plane = g.out_plane;
    d.up2 = g.up2; d.sharpen = g.sharpen;
    { const CasK k = cas_k(d.sharpen); d.cas_a = k.a; d.cas_b = k.b; }
    return d;
}

template <class P> static void host_fft_of(HostFft* hf) {
    int r[kMaxStages];
    for (int s = 0; s < P::kStages; ++s) r[s] = P::radix(s);
    build_fft(P::kN, r, P::kStages, P::kT, hf);
}

using b2r_emu::Dim3;

struct FrameCtx {
    std::vector<float2> nyq;
    Geometry g; FrameDims dm; int precision;
    const void* in; void* out;
    std::vector<float2> spec1, spec2;
    std::vector<unsigned char> pre;
};

static int g_r2c_bulk = 0;   // emulate the bulk-copy (persistent) K1 instead of the direct-load one

template <class P, int PPB> static void emu_r2c(FrameCtx& c, const P plan, const HostFft& hf) {
    int pairs = 3 * c.g.h / 2;
    if (g_r2c_bulk) {
        if constexpr (P::kStatic) {
            Dim3 grid, block; block.x = hf.desc.threads; grid.x = 7;   // few persistent CTAs, many trips
            const float2* tw = hf.twiddles.data();
            const size_t elem = c.precision == 2 ? 2 : 4;
            b2r_emu::launch(grid, block, r2c_bulk_smem_bytes(P::kN, elem), [&] {
                if (c.precision == 2) k_r2c_rows_bulk<P, __half>((const __half*)c.in, c.spec1.data(), tw, plan, c.dm, pairs);
                else k_r2c_rows_bulk<P, float>((const float*)c.in, c.spec1.data(), tw, plan, c.dm, pairs);
            });
            return;
        }
    }
    Dim3 grid, block; block.x = hf.desc.threads; block.y = PPB; grid.x = (pairs + PPB - 1) / PPB;
    const float2* tw = hf.twiddles.data();
    b2r_emu::launch(grid, block, PPB * smem_padded_len(c.g.w) * sizeof(float2), [&] {
        if (c.precision == 2) k_r2c_rows<P, __half, PPB>((const __half*)c.in, c.spec1.data(), tw, plan, c.dm, pairs);
        else k_r2c_rows<P, float, PPB>((const float*)c.in, c.spec1.data(), tw, plan, c.dm, pairs);
    });
}
static int g_c2c = 0;        // C2C parity mode (B2R_FLAG_C2C_PARITY)
static int g_c2r_bulk = 0;   // emulate the bulk-copy (persistent) C2R kernel instead of the direct one

template <class P, int PPB> static void emu_c2r(FrameCtx& c, const P plan, const HostFft& hf) {
    int pairs = 3 * c.g.up_h / 2;
    if (g_c2c) {
        const int rows = 3 * c.g.up_h;
        Dim3 grid, block; block.x = hf.desc.threads; block.y = PPB; grid.x = (rows + PPB - 1) / PPB;
        const float2* tw = hf.twiddles.data();
        const float scale = 1.0f / (float)c.g.up_w;
        b2r_emu::launch(grid, block, PPB * smem_padded_len(c.g.up_w) * sizeof(float2), [&] {
            if (c.precision == 2) k_c2c_rows<P, __half, PPB>(c.spec2.data(), c.nyq.data(), (__half*)c.pre.data(), tw, plan, c.dm, rows, scale);
            else k_c2c_rows<P, float, PPB>(c.spec2.data(), c.nyq.data(), (float*)c.pre.data(), tw, plan, c.dm, rows, scale);
        });
        return;
    }
    if (g_c2r_bulk) {
        if constexpr (P::kStatic) {
            Dim3 grid, block; block.x = hf.desc.threads; grid.x = 5;   // few persistent CTAs, many trips
            const float2* tw = hf.twiddles.data();
            const float scale = 1.0f / (float)c.g.up_w;
            const bool up2 = (c.g.up_w == 2 * c.g.w);
            b2r_emu::launch(grid, block, c2r_bulk_smem_bytes(P::kN, c.g.nx), [&] {
                if (c.precision == 2) {
                    if (up2) k_c2r_rows_bulk<P, __half, true>(c.spec2.data(), (__half*)c.pre.data(), tw, plan, c.dm, pairs, scale);
                    else k_c2r_rows_bulk<P, __half, false>(c.spec2.data(), (__half*)c.pre.data(), tw, plan, c.dm, pairs, scale);
                } else {
                    if (up2) k_c2r_rows_bulk<P, float, true>(c.spec2.data(), (float*)c.pre.data(), tw, plan, c.dm, pairs, scale);
                    else k_c2r_rows_bulk<P, float, false>(c.spec2.data(), (float*)c.pre.data(), tw, plan, c.dm, pairs, scale);
                }
            });
            return;
        }
    }
    Dim3 grid, block; block.x = hf.desc.threads; block.y = PPB; grid.x = (pairs + PPB - 1) / PPB;
    const float2* tw = hf.twiddles.data();
    const float scale = 1.0f / (float)c.g.up_w;
    b2r_emu::launch(grid, block, PPB * smem_padded_len(c.g.up_w) * sizeof(float2), [&] {
        const bool up2 = P::kStatic && (c.g.up_w == 2 * c.g.w);
        if (c.precision == 2) {
            if (up2) k_c2r_rows<P, __half, PPB, true>(c.spec2.data(), (__half*)c.pre.data(), tw, plan, c.dm, pairs, scale);
            else k_c2r_rows<P, __half, PPB, false>(c.spec2.data(), (__half*)c.pre.data(), tw, plan, c.dm, pairs, scale);
        } else {
            if (up2) k_c2r_rows<P, float, PPB, true>(c.spec2.data(), (float*)c.pre.data(), tw, plan, c.dm, pairs, scale);
            else k_c2r_rows<P, float, PPB, false>(c.spec2.data(), (float*)c.pre.data(), tw, plan, c.dm, pairs, scale);
        }
    });
}
static int g_cols_grouped = 0;   // opt-in like the product (B2R_COLS_GROUPED=1)
static int g_cols_staged = 0;    // the persistent, asynchronously staged column kernel (B2R_COLS_STAGED=1)
static int g_cols_2x = 0;        // the exact-2x column kernel (k_cols2x), as the product uses for upH == 2H
static int g_fused_nsp = 0;      // > 0: K7 + K8 through the fused strip kernel (b2r_fused.cuh) with this many strips per plane

// fused C2R + sharpen + boundary fix-up, as launch_frame / run_fused do it (fp32 / fp16, static schedules)
template <class P> static void emu_fused(FrameCtx& c, const P plan, const HostFft& hf, void* out) {
    const bool half = c.precision == 2;
    const int nsp = g_fused_nsp, ppp = c.g.up_h / 2;
    const float2* tw = hf.twiddles.data();
    const float scale = 1.0f / (float)c.g.up_w;
    const bool up2 = (c.g.up_w == 2 * c.g.w);
    Dim3 grid, block; block.x = P::kT; grid.x = 3 * nsp;
    b2r_emu::launch(grid, block, fused_smem_bytes(P::kN, c.g.nx), [&] {
        if (half) {
            if (up2) k_c2r_sharpen_f16<P, true>(c.spec2.data(), (__half*)out, (__half*)c.pre.data(), tw, plan, c.dm, scale, nsp);
            else k_c2r_sharpen_f16<P, false>(c.spec2.data(), (__half*)out, (__half*)c.pre.data(), tw, plan, c.dm, scale, nsp);
        } else {
            if (up2) k_c2r_sharpen_f32<P, true>(c.spec2.data(), (float*)out, (float*)c.pre.data(), tw, plan, c.dm, scale, nsp);
            else k_c2r_sharpen_f32<P, false>(c.spec2.data(), (float*)out, (float*)c.pre.data(), tw, plan, c.dm, scale, nsp);
        }
    });
    std::vector<int> fix;
    for (int q = 1; q < nsp; ++q) fix.push_back(2 * fused_strip_begin(q, nsp, ppp));
    fix.push_back(c.g.up_h);
    Dim3 g2, b2; b2.x = 32; g2.x = (c.g.up_w / 8 + 31) / 32; g2.y = (unsigned)fix.size(); g2.z = 3;
    b2r_emu::launch(g2, b2, 0, [&] {
        if (half) k_sharpen_fix_f16<0>((const __half*)c.pre.data(), (__half*)out, c.dm, fix.data());
        else k_sharpen_fix_f32<0>((const float*)c.pre.data(), (float*)out, c.dm, fix.data());
    });
}

template <class PF, class PI, int CC>
static void emu_cols(FrameCtx& c, const PF pf, const PI pi, const HostFft& hf, const HostFft& hi) {
    Dim3 grid, block; block.x = CC * hi.desc.threads; grid.x = (c.g.nx + CC - 1) / CC; grid.y = 3;
    const float2 *twf = hf.twiddles.data(), *twi = hi.twiddles.data();
    const float scale = 1.0f / (float)c.g.up_h;
    if constexpr (PI::kStatic) {
        if constexpr (PI::kN == 2 * PF::kN) {
            if (g_cols_2x) {
                std::vector<float2> ramp((size_t)c.g.h);
                for (int k = 0; k < c.g.h; ++k) {
                    const int ks = (k < c.g.h / 2) ? k : k - c.g.h;
                    const double a = 3.14159265358979323846 * (double)ks / (double)c.g.h;
                    ramp[(size_t)k] = make_float2((float)std::cos(a), (float)std::sin(a));
                }
                Dim3 g2, b2; b2.x = CC * hf.desc.threads; g2.x = (c.g.nx + CC - 1) / CC; g2.y = 3;
                b2r_emu::launch(g2, b2, smem_padded_len(c.g.h * CC) * sizeof(float2), [&] {
                    k_cols2x<PF, CC>(c.spec1.data(), c.spec2.data(), twf, ramp.data(), pf, c.dm, scale, g_c2c ? c.nyq.data() : nullptr);
                });
                return;
            }
        }
        if (g_cols_staged) {
            Dim3 g2, b2; b2.x = CC * hi.desc.threads; g2.x = 5;   // few persistent CTAs, several tiles each
            const int tiles_per_ch = (c.g.nx + CC - 1) / CC;
            b2r_emu::launch(g2, b2, cols_staged_smem_bytes(c.g.h, c.g.up_h, CC, sizeof(float2)), [&] {
                k_cols_staged<PF, PI, CC>(c.spec1.data(), c.spec2.data(), twf, twi, pf, pi, c.dm, scale, g_c2c ? c.nyq.data() : nullptr, tiles_per_ch);
            });
            return;
        }
        if constexpr (PI::kT % 32 == 0 && PF::kStages >= 2 && PI::kStages >= 3) {
            if (g_cols_grouped) {
                b2r_emu::launch(grid, block, (size_t)CC * cols_group_stride(c.g.up_h) * sizeof(float2), [&] {
                    k_cols_grouped<PF, PI, CC>(c.spec1.data(), c.spec2.data(), twf, twi, pf, pi, c.dm, scale, g_c2c ? c.nyq.data() : nullptr);
                });
                return;
            }
        }
    }
    b2r_emu::launch(grid, block, smem_padded_len(c.g.up_h * CC) * sizeof(float2), [&] {
        k_cols<PF, PI, CC>(c.spec1.data(), c.spec2.data(), twf, twi, pf, pi, c.dm, scale, g_c2c ? c.nyq.data() : nullptr);
    });
}

// K8 exactly as launch_sharpen_kernel chooses it (b2r_sharpen.cu)
static void emu_sharpen(const FrameDims dm, int precision, const void* pre, void* out) {
    Dim3 grid, block;
    const int bx = sharpen_rows_block(dm.up_w);
    if (bx > 0) {
        constexpr int RY = kSharpenRowsPerThread;
        block.x = bx; grid.x = (dm.up_w / 4 + bx - 1) / bx; grid.y = (dm.up_h + RY - 1) / RY; grid.z = 3;
        b2r_emu::launch(grid, block, 0, [&] {
            const bool ragged = sharpen_rows_ragged(dm.up_w, bx);
            if (precision == 2) {
                if (ragged) k_sharpen_rows<__half, RY, true>((const __half*)pre, (__half*)out, dm);
                else k_sharpen_rows<__half, RY, false>((const __half*)pre, (__half*)out, dm);
            } else {
                if (ragged) k_sharpen_rows<float, RY, true>((const float*)pre, (float*)out, dm);
                else k_sharpen_rows<float, RY, false>((const float*)pre, (float*)out, dm);
            }
        });
        return;
    }
    constexpr int PX = 4;
    block.x = 64; grid.x = (dm.up_w + PX * 64 - 1) / (PX * 64); grid.y = dm.up_h; grid.z = 3;
    b2r_emu::launch(grid, block, 0, [&] {
        if (precision == 2) k_sharpen<__half, PX>((const __half*)pre, (__half*)out, dm);
        else k_sharpen<float, PX>((const float*)pre, (float*)out, dm);
    });
}

static int g_sharpen_fast = 0;
static void emu_sharpen_fast(const FrameDims dm, int precision, const void* pre, void* out) {
    const int np = (precision == 2 || dm.up_w % 8 == 0) ? 8 : 4, ry = 12;
    const int vecs = dm.up_w / np, bx = cas_fast_block(vecs);
    Dim3 grid, block;
    block.x = bx; grid.x = (vecs + bx - 1) / bx; grid.y = (dm.up_h + ry - 1) / ry; grid.z = 3;
    b2r_emu::launch(grid, block, 0, [&] {
        if (precision == 2) k_sharpen_fast_f16<8>((const __half*)pre, (__half*)out, dm, ry, 1);
        else if (np == 8) k_sharpen_fast_f32<2>((const float*)pre, (float*)out, dm, ry, 1);
        else k_sharpen_fast_f32<1>((const float*)pre, (float*)out, dm, ry, 1);
    });
}

template <class P> static int emu_fft_static(int n, int dir, const float* in, float* out) {
    HostFft hf;
    Dim3 grid, block;
    host_fft_of<P>(&hf); block.x = P::kT;
    const float2* tw = hf.twiddles.data();
    b2r_emu::launch(grid, block, smem_padded_len(n) * sizeof(float2), [&] {
        if (dir < 0) k_fft_test<-1>((const float2*)in, (float2*)out, tw, P{});
        else k_fft_test<+1>((const float2*)in, (float2*)out, tw, P{});
    });
    return 1;
}

extern "C" {

void b2r_emu_set_c2r_bulk(int on) { g_c2r_bulk = on; }
void b2r_emu_set_r2c_bulk(int on) { g_r2c_bulk = on; }
void b2r_emu_set_c2c(int on) { g_c2c = on; }
void b2r_emu_set_cols_grouped(int on) { g_cols_grouped = on; }
void b2r_emu_set_cols_staged(int on) { g_cols_staged = on; }
void b2r_emu_set_cols_2x(int on) { g_cols_2x = on; }
void b2r_emu_set_fused(int nsp) { g_fused_nsp = nsp; }
void b2r_emu_set_sharpen_fast(int on) { g_sharpen_fast = on; }

// returns number of stages (>0) or -1; radices[] receives the schedule
int b2r_emu_schedule(int n, int* radices, int* threads) {
    HostFft hf; std::string err;
    if (!schedule_fft(n, &hf, &err)) return -1;
    for (int s = 0; s < hf.desc.nstages; ++s) radices[s] = hf.desc.st[s].radix;
    *threads = hf.desc.threads;
    return hf.desc.nstages;
}

// use_static != 0: run an ahead-of-time schedule for n if one exists (returns 1 if it did):
//   1 = the K1 row list (+ the sizes only K7 has), 2 = the K7 row list (two butterflies per thread for the long
//   rows), 3 / 4 = the forward / inverse schedules of the fused column kernels
int b2r_emu_fft(int n, int dir, int use_static, const float* in, float* out) {
    HostFft hf; std::string err;
    Dim3 grid, block;
#define X(N, PPB, T, ...) if (n == N) return emu_fft_static<StaticFft<N, T, __VA_ARGS__>>(n, dir, in, out);
    if (use_static == 1) { B2R_STATIC_ROWS(X) }
    if (use_static == 2) { B2R_STATIC_C2R_ROWS(X) }
#undef X
#define X(H, UPH, CC, PF, PI) if (n == H) return emu_fft_static<PF>(n, dir, in, out);
    if (use_static == 3) { B2R_STATIC_COLS(X) }
#undef X
#define X(H, UPH, CC, PF, PI) if (n == UPH) return emu_fft_static<PI>(n, dir, in, out);
    if (use_static == 4) { B2R_STATIC_COLS(X) }
#undef X
    if (!schedule_fft(n, &hf, &err)) return -1;
    block.x = (unsigned)hf.desc.threads;
    const float2* tw = hf.twiddles.data();
    const DynFft plan{&hf.desc};
    b2r_emu::launch(grid, block, smem_padded_len(n) * sizeof(float2), [&] {
        if (dir < 0) k_fft_test<-1>((const float2*)in, (float2*)out, tw, plan);
        else k_fft_test<+1>((const float2*)in, (float2*)out, tw, plan);
    });
    return 0;
}

// Runs the 4-kernel frame on the CPU emulator.  Buffers as in the C-ABI (input: reference layout
// with (W+2)*H plane stride; output compact).  Optional dumps: spec1 [3][H][stride] complex,
// spec2 [3][upH][stride] complex, pre [3*(upW+2)*upH + slack] elements.
// use_static: prefer the ahead-of-time schedules (as the product does); cc only affects the
// dynamic column kernel.  *used_static gets bit0 K1, bit1 columns, bit2 K7.
int b2r_emu_frame(int w, int h, float upscale, int precision, float sharpen_const, float up2_lit, int cc,
                  int use_static, const void* in, void* out, float* spec1_dump, float* spec2_dump,
                  void* pre_dump, int* spec_stride_out, int* used_static) {
    FrameCtx c; std::string err;
    Geometry& g = c.g;
    if (!make_geometry(w, h, upscale, precision, sharpen_const, &g, &err, g_c2c != 0)) { fprintf(stderr, "%s\n", err.c_str()); return -1; }
    g.up2 = up2_lit;
    c.dm = dims_of(g); c.precision = precision; c.in = in; c.out = out;
    c.nyq.assign(3 * (size_t)g.spec_stride, make_float2(0, 0));
    if (spec_stride_out) *spec_stride_out = g.spec_stride;
    c.spec1.assign(g.spec_in_elems(), make_float2(0, 0));
    c.spec2.assign(g.spec_out_elems(), make_float2(0, 0));
    c.pre.assign(g.pre_elems * g.elem_bytes(), 0);
    int used = 0;
    HostFft hf, hi;
    {   // K1
        bool done = false;
        if (use_static) {
#define X(N, PPB, T, ...) \
            if (!done && g.w == N) { using P = StaticFft<N, T, __VA_ARGS__>; host_fft_of<P>(&hf); emu_r2c<P, PPB>(c, P{}, hf); done = true; used |= 1; }
            B2R_STATIC_R2C_ROWS(X)
#undef X
        }
        if (!done) {
            if (!schedule_fft(g.w, &hf, &err)) return -2;
            emu_r2c<DynFft, 1>(c, DynFft{&hf.desc}, hf);
        }
    }
    {   // K2..K6
        bool done = false;
        if (use_static) {
#define X(H, UPH, CC, PF, PI) \
            if (!done && g.h == H && g.up_h == UPH) { host_fft_of<PF>(&hf); host_fft_of<PI>(&hi); emu_cols<PF, PI, CC>(c, PF{}, PI{}, hf, hi); done = true; used |= 2; }
            B2R_STATIC_COLS(X)
#undef X
        }
        if (!done) {
            if (!schedule_fft(g.h, &hf, &err) || !schedule_fft(g.up_h, &hi, &err)) return -2;
            int tc = std::max(hf.desc.threads, hi.desc.threads);
            schedule_fft(g.h, &hf, &err, tc); schedule_fft(g.up_h, &hi, &err, tc);
            const DynFft pf{&hf.desc}, pi{&hi.desc};
            if (cc == 8) emu_cols<DynFft, DynFft, 8>(c, pf, pi, hf, hi);
            else if (cc == 4) emu_cols<DynFft, DynFft, 4>(c, pf, pi, hf, hi);
            else emu_cols<DynFft, DynFft, 2>(c, pf, pi, hf, hi);
        }
    }
    bool fused_done = false;
    if (g_fused_nsp > 0 && (precision == 0 || precision == 2) && use_static && !g_c2c) {   // K7 + K8 fused
#define X(N, PPB, T, ...) \
        if (!fused_done && g.up_w == N) { using P = StaticFft<N, T, __VA_ARGS__>; host_fft_of<P>(&hf); emu_fused<P>(c, P{}, hf, out); fused_done = true; used |= 4; }
        B2R_STATIC_C2R_ROWS(X)
#undef X
    }
    if (!fused_done) {   // K7
        bool done = false;
        if (use_static) {
#define X(N, PPB, T, ...) \
            if (!done && g.up_w == N) { using P = StaticFft<N, T, __VA_ARGS__>; host_fft_of<P>(&hf); emu_c2r<P, PPB>(c, P{}, hf); done = true; used |= 4; }
            B2R_STATIC_C2R_ROWS(X)
#undef X
        }
        if (!done) {
            if (!schedule_fft(g.up_w, &hf, &err)) return -2;
            emu_c2r<DynFft, 1>(c, DynFft{&hf.desc}, hf);
        }
    }
    if (!fused_done) {   // K8 (g_sharpen_fast: the tolerance-bound kernels, as the product's default)
        if (g_sharpen_fast) emu_sharpen_fast(c.dm, precision, c.pre.data(), out);
        else emu_sharpen(c.dm, precision, c.pre.data(), out);
    }
    if (used_static) *used_static = used;
    if (spec1_dump) memcpy(spec1_dump, c.spec1.data(), c.spec1.size() * sizeof(float2));
    if (spec2_dump) memcpy(spec2_dump, c.spec2.data(), c.spec2.size() * sizeof(float2));
    if (pre_dump) memcpy(pre_dump, c.pre.data(), c.pre.size());
    return 0;
}

// the u8 pixel-format kernels (precision 0/2) on a w x h frame / its upscaled output
int b2r_emu_u8_to_planar(int w, int h, int precision, const unsigned char* rgb, void* planar) {
    Geometry g; std::string err;
    if (!make_geometry(w, h, 1.0f, precision, 0.2f, &g, &err)) return -1;
    const FrameDims dm = dims_of(g);
    Dim3 grid, block; block.x = 64; grid.x = (unsigned)(((size_t)w * h / 4 + 63) / 64);
    b2r_emu::launch(grid, block, 0, [&] {
        if (precision == 2) k_u8_to_planar<__half>(rgb, (__half*)planar, dm);
        else k_u8_to_planar<float>(rgb, (float*)planar, dm);
    });
    return 0;
}
int b2r_emu_planar_to_u8(int w, int h, int precision, const void* planar, unsigned char* rgb) {
    Geometry g; std::string err;
    if (!make_geometry(w, h, 1.0f, precision, 0.2f, &g, &err)) return -1;   // up = 1: out dims == in dims
    const FrameDims dm = dims_of(g);
    Dim3 grid, block; block.x = 64; grid.x = (unsigned)(((size_t)w * h / 4 + 63) / 64);
    b2r_emu::launch(grid, block, 0, [&] {
        if (precision == 2) k_planar_to_u8<__half>((const __half*)planar, rgb, dm);
        else k_planar_to_u8<float>((const float*)planar, rgb, dm);
    });
    return 0;
}

// the tolerance-bound sharpen kernels (b2r_cas.cuh), launched as launch_sharpen_fast does (b2r_sharpen.cu)
int b2r_emu_sharpen_fast(int w, int h, float upscale, int precision, float sharpen_const, float up2_lit, int ry,
                         int reverse, int block_x, const void* pre, void* out) {
    Geometry g; std::string err;
    if (!make_geometry(w, h, upscale, precision, sharpen_const, &g, &err)) return -1;
    g.up2 = up2_lit;
    const FrameDims dm = dims_of(g);
    const int np = (precision == 2 || dm.up_w % 8 == 0) ? 8 : 4;
    if (dm.up_w % np) return -2;
    const int vecs = dm.up_w / np, bx = block_x > 0 ? block_x : cas_fast_block(vecs);
    Dim3 grid, block;
    block.x = bx; grid.x = (vecs + bx - 1) / bx; grid.y = (dm.up_h + ry - 1) / ry; grid.z = 3;
    b2r_emu::launch(grid, block, 0, [&] {
        if (precision == 2) k_sharpen_fast_f16<8>((const __half*)pre, (__half*)out, dm, ry, reverse);
        else if (np == 8) k_sharpen_fast_f32<2>((const float*)pre, (float*)out, dm, ry, reverse);
        else k_sharpen_fast_f32<1>((const float*)pre, (float*)out, dm, ry, reverse);
    });
    return 0;
}

// sharpen alone on a caller-provided padded plane buffer (bit-exactness test of K8)
int b2r_emu_sharpen(int w, int h, float upscale, int precision, float sharpen_const, float up2_lit,
                    const void* pre, void* out) {
    Geometry g; std::string err;
    if (!make_geometry(w, h, upscale, precision, sharpen_const, &g, &err)) return -1;
    g.up2 = up2_lit;
    const FrameDims dm = dims_of(g);
    emu_sharpen(dm, precision, pre, out);
    return 0;
}

}  // extern "C"
