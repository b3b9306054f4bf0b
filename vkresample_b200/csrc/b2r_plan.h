// b2r_plan.h -- host-side frame geometry and FFT scheduling (no CUDA runtime calls).
//
// Replaces the plan-time half of the reference:
//   VkResampleConfiguration / VkFFTConfiguration fill   VkResample.cpp:45-59, :1409-1503
//   VkFFTScheduler (radix split of an axis)              vkFFT.h:4707-5189
// The reference factors an axis into <=3 "uploads" of radix-2..8 stages sized to 32-48 KB of
// shared memory and JIT-compiles GLSL per axis; here one CTA always holds a whole sequence (B200:
// 227 KB shared memory per CTA), so scheduling reduces to choosing a radix list (radices up to 16)
// and the number of threads that cooperate on one sequence.
#pragma once

#include <string>
#include <vector>

#include "b2r_fft.cuh"

namespace b2r {

struct HostFft {
    FftDesc desc{};
    std::vector<float2> twiddles;  // concatenated per-stage rows, forward sign
    std::vector<double2> twiddles_d;  // the same in double (for -p 1)
};

// Factor n (= 2^a 3^b 5^c 7^d, the set the reference accepts: vkFFT.h:4719-4726) into radices
// <= 16, fewest stages first, and derive the per-stage constants.  min_threads lets a caller force
// a common thread count on two transforms that share a CTA (column kernel).
bool schedule_fft(int n, HostFft* out, std::string* err, int force_threads = 0);
// Same tables for a caller-chosen radix list (the statically instantiated schedules).
void build_fft(int n, const int* radices, int nst, int threads, HostFft* out);
int fft_min_threads(const std::vector<int>& radices, int n);
bool factor_radices(int n, std::vector<int>* radices);

// Geometry of one frame plan.  Element counts, not bytes.
struct Geometry {
    int w = 0, h = 0;          // input size (both even)
    float upscale = 1.f;
    int up_w = 0, up_h = 0;    // truncated float products, VkResample.cpp:1417-1418
    int nx = 0;                // kept spectrum bins per row: W/2+1 (kx = 0..W/2)
    int spec_stride = 0;       // row stride (complex elements) of both spectrum buffers
    int zp_lo = 0, zp_hi = 0;  // inverse reads rows [zp_lo, zp_hi) as zero, VkResample.cpp:1494-1495
    int neg_shift = 0;         // rows >= up_h - h/2 come from source row (m - neg_shift)
    int precision = 0;         // 0 fp32, 1 fp64, 2 fp16 storage
    float up2 = 1.f;           // appSharpen.upscale, VkResample.cpp:1615
    float sharpen = 0.2f;
    // strides in elements (float or half)
    size_t in_row = 0, in_plane = 0;    // W, (W+2)*H                      VkResample.cpp:1644
    size_t pre_row = 0, pre_plane = 0;  // upW, (upW+2)*upH                VkResample.cpp:1593-1596
    size_t out_row = 0, out_plane = 0;  // upW, upW*upH (compact)          VkResample.cpp:1599-1601
    size_t pre_elems = 0;               // allocation incl. zero slack after the last plane
    size_t elem_bytes() const { return precision == 2 ? 2 : (precision == 1 ? 8 : 4); }
    size_t cplx_bytes() const { return precision == 1 ? 16 : 8; }   // one spectrum / workspace element
    size_t input_bytes() const { return 3 * in_plane * elem_bytes(); }      // == 3*cs*(W/2+1)*H
    size_t output_bytes() const { return 3 * out_plane * elem_bytes(); }    // VkResample.cpp:1698
    size_t spec_in_elems() const { return 3ull * h * spec_stride; }
    size_t spec_out_elems() const { return 3ull * up_h * spec_stride; }
};

// c2c_layout: the C2R/C2C-rows result is stored on a COMPACT plane (stride upW*upH, no pad rows) like the
// reference's C2C branch (VkResample.cpp:1598) instead of the R2C branch's (upW+2)*upH.
bool make_geometry(int w, int h, float upscale, int precision, float sharpen, Geometry* g, std::string* err,
                   bool c2c_layout = false);

}  // namespace b2r
