// b2r_cas.cuh -- K8 (CAS-style sharpen, shaderGenSharpen r2c branch, VkResample.cpp:849-923) in its
// TOLERANCE-BOUND form: the default since round 2.  Same inputs, same neighbour rule, same formula as
// k_sharpen_rows (b2r_kernels.cuh), but the arithmetic is arranged for the B200 issue pipes instead of
// for bit-identity with the numpy oracle (the reference's own GLSL `/` and sqrt are not correctly rounded
// either; north_star's bar is max-abs 1e-5 fp32 / 1e-2 fp16, tests/test_gpu_parity.py checks it against
// oracle.sharpen on the identical plane).  B2R_FLAG_EXACT_SHARPEN selects the bit-exact kernels.
//
// What is different from the exact form (measured pipe rates in profiles/r2_pipes.md: FMNMX / integer
// ops issue at half rate, MUFU at 1/8, 3-input FMNMX3 costs the same as a 2-input one):
//   * scale = min(a, b),  a = mn/(1-mn),  b = (1-mx)/mx  with  mn = (mn0+mn1)/2, mx = (mx0+mx1)/2.
//     With sA = mn0+mn1 and u = 2-(mx0+mx1):  a = f(sA), b = f(u), f(z) = z/(2-z) increasing on [0,2),
//     so min(a, b) = f(min(sA, u)) -- one quotient instead of two, and since mn <= mx gives
//     sA + u <= 2 the selected m = min(sA, u) is <= 1: the divisor 2-m lies in [1, 2] (no special cases).
//   * -s*sqrt(m/(2-m)) = -m * rsqrt(m*(2-m)/s^2 + 1e-30): one MUFU instead of two, and the sharpen constant
//     rides in the FFMA that forms 2-m ((2-m)/s^2 = m*(-1/s^2) + 2/s^2, CasK); the 1e-30 (folded into the
//     FFMA that forms the product) makes m = 0 give exactly 0 instead of 0*inf.
//   * the tap clamp min(|up2*x|, 1) is the saturating multiply up2*|x| (one FMUL.SAT / HMUL2.SAT).
//   * the final quotient is num * rcp(den); den = 1 + 4*scale*(-s) stays in [0.04, 1] for 0 <= s <= 0.24
//     (other constants keep the exact kernel).
//   * vertical 3-min / 3-max per column are shared by the three pixels that use the column; every
//     min/max is a single FMNMX3.
// fp32: 19 arithmetic instructions per pixel + 4 per column (clamp + vertical extrema), 2 MUFU per pixel
// (exact form: 62.5 instructions and 4 MUFU-class sequences per pixel).
// fp16 (precision 2): the same arrangement in native half2 arithmetic, two pixels per instruction, the
// quotient / rsqrt evaluated in float on the half operands (the reference's shader computes in float16_t,
// VkResample.cpp:823-827; tolerance 1e-2).
#pragma once

namespace b2r {

constexpr int kCasFastRows = 12;   // rows per thread of the fast kernels (B2R_SHARPEN_RY overrides; sweep in profiles/)

B2R_HD constexpr int cas_fast_block(int vecs) {   // threads per block for `vecs` thread-columns per row
    for (int b = 256; b >= 128; b -= 32)
        if (vecs % b == 0) return b;
    if (vecs <= 256) return (vecs + 31) / 32 * 32;
    return 256;
}

#if defined(__CUDA_ARCH__)
B2R_DEV float cas_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
B2R_DEV float cas_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
B2R_DEV float cas_rcp(float x) { return 1.0f / x; }
B2R_DEV float cas_rsqrt(float x) { return 1.0f / sqrtf(x); }
#endif

// clamped magnitude of one tap: min(|up2 * x|, 1)   (VkResample.cpp:893-907; the < 0 clamp cannot fire):
// up2 > 0, so |up2 * x| = up2 * |x| exactly and the clamp is the multiply's saturation modifier
#if defined(__CUDA_ARCH__)
B2R_DEV float cas_tap(float up2, float x) { return __saturatef(up2 * fabsf(x)); }
#else
B2R_DEV float cas_tap(float up2, float x) { return fminf(fmaxf(up2 * fabsf(x), 0.0f), 1.0f); }
#endif

// sharpen constant folded for the rsqrt argument: (2 - m) / s^2 = fma(m, a, b).  s = 0 takes 1/s^2 = 1e30:
// the scale then comes out as -1e-15 * sqrt(m/(2-m)), i.e. the pixel passes through to 1e-15.
// Evaluated once per plan on the host (FrameDims::cas_a / cas_b).
struct CasK { float a, b; };
inline CasK cas_k(float sharpen) {
    const float inv = (sharpen != 0.0f) ? 1.0f / (sharpen * sharpen) : 1e30f;
    CasK k; k.a = -inv; k.b = 2.0f * inv;
    return k;
}

// One output row segment.  up / mid / dn: clamped magnitudes of columns x0-1 .. x0+NP of rows y-1, y, y+1;
// o[i]: sharpened pixel x0+i.  ks = cas_k(sharpen).
template <int NP>
B2R_DEV void cas_row_f32(const float (&up)[NP + 2], const float (&mid)[NP + 2], const float (&dn)[NP + 2],
                         const CasK ks, float (&o)[NP]) {
    float vmn[NP + 2], vmx[NP + 2];
#pragma unroll
    for (int c = 0; c < NP + 2; ++c) {
        vmn[c] = fminf(up[c], fminf(mid[c], dn[c]));
        vmx[c] = fmaxf(up[c], fmaxf(mid[c], dn[c]));
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const float mn0 = fminf(vmn[i + 1], fminf(mid[i], mid[i + 2]));      // cross: up, left, centre, right, down
        const float mn1 = fminf(vmn[i], fminf(vmn[i + 1], vmn[i + 2]));      // all nine
        const float mx0 = fmaxf(vmx[i + 1], fmaxf(mid[i], mid[i + 2]));
        const float mx1 = fmaxf(vmx[i], fmaxf(vmx[i + 1], vmx[i + 2]));
        const float sA = mn0 + mn1;
        const float u = 2.0f - (mx0 + mx1);
        const float m = fminf(sA, u);
        const float sc = -m * cas_rsqrt(fmaf(m, fmaf(m, ks.a, ks.b), 1e-30f));   // -s * sqrt(m / (2 - m))
        const float cross = (up[i + 1] + dn[i + 1]) + (mid[i] + mid[i + 2]);
        o[i] = fmaf(sc, cross, mid[i + 1]) * cas_rcp(fmaf(sc, 4.0f, 1.0f));
    }
}

// ---- half2 form -------------------------------------------------------------------------------------
// Two horizontally adjacent pixels per instruction.  Column c of a row is held twice: as lane `c & 1` of
// the aligned pair P[c >> 1] and -- for the odd-aligned accesses (left / right neighbours) -- in the
// shifted pairs built once per row.
#if defined(__CUDA_ARCH__)
B2R_DEV __half2 h2min3(__half2 a, __half2 b, __half2 c) { return __hmin2(a, __hmin2(b, c)); }
B2R_DEV __half2 h2max3(__half2 a, __half2 b, __half2 c) { return __hmax2(a, __hmax2(b, c)); }
#else
B2R_DEV __half2 h2min3(__half2 a, __half2 b, __half2 c) {
    auto mn = [](__half x, __half y) { return __half2float(x) < __half2float(y) ? x : y; };
    return __halves2half2(mn(__low2half(a), mn(__low2half(b), __low2half(c))), mn(__high2half(a), mn(__high2half(b), __high2half(c))));
}
B2R_DEV __half2 h2max3(__half2 a, __half2 b, __half2 c) {
    auto mx = [](__half x, __half y) { return __half2float(x) > __half2float(y) ? x : y; };
    return __halves2half2(mx(__low2half(a), mx(__low2half(b), __low2half(c))), mx(__high2half(a), mx(__high2half(b), __high2half(c))));
}
#endif
#if defined(__CUDA_ARCH__)
B2R_DEV __half2 cas_hfma2(__half2 a, __half2 b, __half2 c) { return __hfma2(a, b, c); }
#else
B2R_DEV __half2 cas_hfma2(__half2 a, __half2 b, __half2 c) {   // exact product + sum in double, one rounding
    const float2 x = __half22float2(a), y = __half22float2(b), z = __half22float2(c);
    return __halves2half2(__double2half((double)x.x * y.x + z.x), __double2half((double)x.y * y.y + z.y));
}
#endif
#if defined(__CUDA_ARCH__)
B2R_DEV __half2 cas_tap2(__half2 up2, __half2 x) { return __hmul2_sat(up2, __habs2(x)); }
#else
B2R_DEV __half2 cas_tap2(__half2 up2, __half2 x) {
    const __half2 one = __float2half2_rn(1.0f);
    return h2min3(__habs2(__hmul2(up2, x)), one, one);
}
#endif

// a[k] = columns (x0-1+2k, x0+2k), k = 0 .. NP/2     (pairs starting at the odd column x0-1; a[NP/2] ends at x0+NP)
// b[k] = columns (x0+2k, x0+2k+1),   k = 0 .. NP/2-1   (the thread's own aligned pairs)
template <int NP> struct CasRowH {
    __half2 a[NP / 2 + 1], b[NP / 2];
    // b filled, halo = (column x0-1, column x0+NP): derive the odd-aligned pairs
    B2R_DEV void link(__half left, __half right) {
        a[0] = __halves2half2(left, __low2half(b[0]));
#pragma unroll
        for (int k = 1; k < NP / 2; ++k) a[k] = __halves2half2(__high2half(b[k - 1]), __low2half(b[k]));
        a[NP / 2] = __halves2half2(__high2half(b[NP / 2 - 1]), right);
    }
};

// o[k] = sharpened pixels (x0+2k, x0+2k+1)
template <int NP>
B2R_DEV void cas_row_f16(const CasRowH<NP>& up, const CasRowH<NP>& mid, const CasRowH<NP>& dn, const CasK ks,
                         __half2 (&o)[NP / 2]) {
    constexpr int H = NP / 2;
    __half2 vna[H + 1], vxa[H + 1], vnb[H], vxb[H];
#pragma unroll
    for (int k = 0; k <= H; ++k) { vna[k] = h2min3(up.a[k], mid.a[k], dn.a[k]); vxa[k] = h2max3(up.a[k], mid.a[k], dn.a[k]); }
#pragma unroll
    for (int k = 0; k < H; ++k) { vnb[k] = h2min3(up.b[k], mid.b[k], dn.b[k]); vxb[k] = h2max3(up.b[k], mid.b[k], dn.b[k]); }
    const __half2 two = __float2half2_rn(2.0f);
#pragma unroll
    for (int k = 0; k < H; ++k) {
        // pixel pair (x0+2k, x0+2k+1): centre = b[k]; left = a[k]; right = a[k+1]
        const __half2 mn0 = h2min3(vnb[k], mid.a[k], mid.a[k + 1]);
        const __half2 mn1 = h2min3(vna[k], vnb[k], vna[k + 1]);
        const __half2 mx0 = h2max3(vxb[k], mid.a[k], mid.a[k + 1]);
        const __half2 mx1 = h2max3(vxa[k], vxb[k], vxa[k + 1]);
        const __half2 sA = __hadd2(mn0, mn1);
        const __half2 u = __hsub2(two, __hadd2(mx0, mx1));
        const __half2 m = __hmin2(sA, u);
        // the two MUFU steps in float on the half operand: scale = -s*sqrt(m/(2-m)) as in cas_row_f32, its
        // half rounding feeds the numerator, the float value the reciprocal of the denominator
        const float2 mf = __half22float2(m);
        const float sx = -mf.x * cas_rsqrt(fmaf(mf.x, fmaf(mf.x, ks.a, ks.b), 1e-30f));
        const float sy = -mf.y * cas_rsqrt(fmaf(mf.y, fmaf(mf.y, ks.a, ks.b), 1e-30f));
        const __half2 sc = __floats2half2_rn(sx, sy);
        const __half2 cross = __hadd2(__hadd2(up.b[k], dn.b[k]), __hadd2(mid.a[k], mid.a[k + 1]));
        const __half2 num = cas_hfma2(sc, cross, mid.b[k]);
        o[k] = __hmul2(num, __floats2half2_rn(cas_rcp(fmaf(sx, 4.0f, 1.0f)), cas_rcp(fmaf(sy, 4.0f, 1.0f))));
    }
}


// =================================================================================================
// Kernels.  One thread = NP consecutive pixels x `ry` rows (rolling three-row window: every input row is
// loaded once per thread, turned into clamped magnitudes once and used by three output rows); the halo
// columns x0-1 / x0+NP come from the neighbouring lanes by warp shuffle, only the first / last lane of a
// warp (and the last pixel group of a row) load them.  Flat neighbour rule as in k_sharpen_rows: left / up
// clamp at 0, right / down do not (the pixel after a row's last one is the next row's first; row upH is the
// plane's zero pad region).  grid = (ceil(upW/NP/blockDim.x), ceil(upH/ry), 3); ry a multiple of 6.
// reverse != 0 walks the planes and row strips in the opposite order of K7's writes, so that the first
// CTAs find the rows K7 wrote last still in L2.
// =================================================================================================
#if !defined(B2R_GDIM_Y)
#if defined(B2R_HOST_EMU)
#define B2R_GDIM_Y (b2r_emu::g_ctx.gdim_y)
#define B2R_GDIM_Z (b2r_emu::g_ctx.gdim_z)
#else
#define B2R_GDIM_Y (gridDim.y)
#define B2R_GDIM_Z (gridDim.z)
#endif
#endif

template <int NV>
B2R_KERNEL B2R_LAUNCH_BOUNDS(256, 2)
k_sharpen_fast_f32(const float* __restrict__ pre, float* __restrict__ out, const FrameDims dm, const int ry, const int reverse) {
    constexpr int NP = 4 * NV;
    const int x0 = (int)(B2R_BID_X * B2R_BDIM_X + B2R_TID_X) * NP;
    int by = (int)B2R_BID_Y, ch = (int)B2R_BID_Z;
    if (reverse) { by = (int)B2R_GDIM_Y - 1 - by; ch = (int)B2R_GDIM_Z - 1 - ch; }
    const int y_begin = by * ry;
    const float up2 = dm.up2;
    const CasK ks = {dm.cas_a, dm.cas_b};
    const float* plane = pre + (size_t)ch * dm.pre_plane;
    float* oplane = out + (size_t)ch * dm.out_plane;
    const bool in_row = x0 < dm.up_w;            // lanes past the row end only take part in the shuffles
    const bool first_in_row = (x0 == 0);
    const bool last_in_row = (x0 + NP == dm.up_w);
#if !defined(B2R_HOST_EMU)
    const int lane = (int)B2R_TID_X & 31;
    const bool need_e0 = (lane == 0) && !first_in_row && in_row;
    const bool need_e1 = ((lane == 31) || last_in_row) && in_row;
#else
    const bool need_e0 = !first_in_row && in_row, need_e1 = in_row;
#endif
    struct Raw { float v[NP]; float e0, e1; };
    auto fetch_row = [&](int y, Raw& q) {
        const float* p = plane + (size_t)y * dm.up_w + x0;
        q.e0 = 0.f; q.e1 = 0.f;
#pragma unroll
        for (int i = 0; i < NP; ++i) q.v[i] = 0.f;
        if (in_row) {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const float4 t = *reinterpret_cast<const float4*>(p + 4 * k);
                q.v[4 * k] = t.x; q.v[4 * k + 1] = t.y; q.v[4 * k + 2] = t.z; q.v[4 * k + 3] = t.w;
            }
        }
        if (need_e0) q.e0 = p[-1];
        if (need_e1) q.e1 = p[NP];   // flat +1: the next row's first pixel at the row end
    };
    auto finish_row = [&](const Raw& q, float (&t)[NP + 2]) {
#pragma unroll
        for (int i = 0; i < NP; ++i) t[i + 1] = cas_tap(up2, q.v[i]);
#if defined(B2R_HOST_EMU)
        t[0] = first_in_row ? t[1] : cas_tap(up2, q.e0);
        t[NP + 1] = cas_tap(up2, q.e1);
#else
        float l = __shfl_up_sync(0xffffffffu, t[NP], 1);
        float r = __shfl_down_sync(0xffffffffu, t[1], 1);
        if (lane == 0) l = first_in_row ? t[1] : cas_tap(up2, q.e0);
        if (lane == 31 || last_in_row) r = cas_tap(up2, q.e1);
        t[0] = l; t[NP + 1] = r;
#endif
    };
    float ta[NP + 2], tb[NP + 2], tc[NP + 2];
    Raw q0, q1;
    fetch_row(y_begin > 0 ? y_begin - 1 : 0, q0);
    fetch_row(y_begin, q1);
    finish_row(q0, ta);
    finish_row(q1, tb);
    // rows are fetched TWO output rows ahead (q0 / q1 alternate): 2 x 16*NV bytes per thread in flight,
    // which is what the HBM latency-bandwidth product needs at 3 CTAs per SM
    fetch_row(y_begin + 1, q0);
    if (y_begin + 1 < dm.up_h) fetch_row(y_begin + 2, q1); else fetch_row(y_begin + 1, q1);
    auto do_row = [&](int y, bool more, Raw& q, float (&up)[NP + 2], float (&mid)[NP + 2], float (&dn)[NP + 2]) {
        finish_row(q, dn);                                   // row y+1
        if (more && y + 2 < dm.up_h) fetch_row(y + 3, q);    // needed by output row y+2
        float o[NP];
        cas_row_f32<NP>(up, mid, dn, ks, o);
        if (in_row) {
            float* dst = oplane + (size_t)y * dm.up_w + x0;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const float4 v = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
#if defined(__CUDA_ARCH__)
                __stcs(reinterpret_cast<float4*>(dst + 4 * k), v);
#else
                *reinterpret_cast<float4*>(dst + 4 * k) = v;
#endif
            }
        }
    };
    // six rows per trip: the three-row window rotates by renaming (period 3), the two raw buffers alternate (period 2)
    for (int r = 0; r < ry; r += 6) {
        const int y = y_begin + r;
        const bool more = r + 6 < ry;
        if (y >= dm.up_h) break;
        do_row(y, true, q0, ta, tb, tc);
        if (y + 1 >= dm.up_h) break;
        do_row(y + 1, true, q1, tb, tc, ta);
        if (y + 2 >= dm.up_h) break;
        do_row(y + 2, true, q0, tc, ta, tb);
        if (y + 3 >= dm.up_h) break;
        do_row(y + 3, true, q1, ta, tb, tc);
        if (y + 4 >= dm.up_h) break;
        do_row(y + 4, more, q0, tb, tc, ta);
        if (y + 5 >= dm.up_h) break;
        do_row(y + 5, more, q1, tc, ta, tb);
    }
}

// fp16: one thread = 8 pixels (one 16-byte vector) per row
template <int NP, int MINB = 3>
B2R_KERNEL B2R_LAUNCH_BOUNDS(256, MINB)
k_sharpen_fast_f16(const __half* __restrict__ pre, __half* __restrict__ out, const FrameDims dm, const int ry, const int reverse) {
    static_assert(NP == 8, "one 16-byte vector per thread and row");
    const int x0 = (int)(B2R_BID_X * B2R_BDIM_X + B2R_TID_X) * NP;
    int by = (int)B2R_BID_Y, ch = (int)B2R_BID_Z;
    if (reverse) { by = (int)B2R_GDIM_Y - 1 - by; ch = (int)B2R_GDIM_Z - 1 - ch; }
    const int y_begin = by * ry;
    const __half2 up2 = __float2half2_rn(dm.up2);
    const CasK ks = {dm.cas_a, dm.cas_b};
    const __half* plane = pre + (size_t)ch * dm.pre_plane;
    __half* oplane = out + (size_t)ch * dm.out_plane;
    const bool in_row = x0 < dm.up_w;
    const bool first_in_row = (x0 == 0);
    const bool last_in_row = (x0 + NP == dm.up_w);
#if !defined(B2R_HOST_EMU)
    const int lane = (int)B2R_TID_X & 31;
    const bool need_e0 = (lane == 0) && !first_in_row && in_row;
    const bool need_e1 = ((lane == 31) || last_in_row) && in_row;
#else
    const bool need_e0 = !first_in_row && in_row, need_e1 = in_row;
#endif
    struct Raw { uint4 v; __half e0, e1; };
    const __half hzero = __float2half_rn(0.f);
    auto fetch_row = [&](int y, Raw& q) {
        const __half* p = plane + (size_t)y * dm.up_w + x0;
        q.e0 = hzero; q.e1 = hzero;
        q.v = make_uint4(0u, 0u, 0u, 0u);
        if (in_row) q.v = *reinterpret_cast<const uint4*>(p);
        if (need_e0) q.e0 = p[-1];
        if (need_e1) q.e1 = p[NP];
    };
    auto as_h2 = [](unsigned w) { __half2 h; *reinterpret_cast<unsigned*>(&h) = w; return h; };
    auto as_u32 = [](__half2 h) { return *reinterpret_cast<unsigned*>(&h); };
    auto finish_row = [&](const Raw& q, CasRowH<NP>& t) {
        t.b[0] = cas_tap2(up2, as_h2(q.v.x)); t.b[1] = cas_tap2(up2, as_h2(q.v.y));
        t.b[2] = cas_tap2(up2, as_h2(q.v.z)); t.b[3] = cas_tap2(up2, as_h2(q.v.w));
        const __half2 e = cas_tap2(up2, __halves2half2(q.e0, q.e1));
#if defined(B2R_HOST_EMU)
        t.link(first_in_row ? __low2half(t.b[0]) : __low2half(e), __high2half(e));
#else
        __half l = __high2half(as_h2(__shfl_up_sync(0xffffffffu, as_u32(t.b[3]), 1)));
        __half r = __low2half(as_h2(__shfl_down_sync(0xffffffffu, as_u32(t.b[0]), 1)));
        if (lane == 0) l = first_in_row ? __low2half(t.b[0]) : __low2half(e);
        if (lane == 31 || last_in_row) r = __high2half(e);
        t.link(l, r);
#endif
    };
    CasRowH<NP> ta, tb, tc;
    Raw q0, q1;
    fetch_row(y_begin > 0 ? y_begin - 1 : 0, q0);
    fetch_row(y_begin, q1);
    finish_row(q0, ta);
    finish_row(q1, tb);
    fetch_row(y_begin + 1, q0);   // two output rows ahead, like the fp32 kernel
    if (y_begin + 1 < dm.up_h) fetch_row(y_begin + 2, q1); else fetch_row(y_begin + 1, q1);
    auto do_row = [&](int y, bool more, Raw& q, CasRowH<NP>& up, CasRowH<NP>& mid, CasRowH<NP>& dn) {
        finish_row(q, dn);
        if (more && y + 2 < dm.up_h) fetch_row(y + 3, q);
        __half2 o[NP / 2];
        cas_row_f16<NP>(up, mid, dn, ks, o);
        if (in_row) {
            const uint4 v = make_uint4(as_u32(o[0]), as_u32(o[1]), as_u32(o[2]), as_u32(o[3]));
            uint4* dst = reinterpret_cast<uint4*>(oplane + (size_t)y * dm.up_w + x0);
#if defined(__CUDA_ARCH__)
            __stcs(dst, v);
#else
            *dst = v;
#endif
        }
    };
    for (int r = 0; r < ry; r += 6) {
        const int y = y_begin + r;
        const bool more = r + 6 < ry;
        if (y >= dm.up_h) break;
        do_row(y, true, q0, ta, tb, tc);
        if (y + 1 >= dm.up_h) break;
        do_row(y + 1, true, q1, tb, tc, ta);
        if (y + 2 >= dm.up_h) break;
        do_row(y + 2, true, q0, tc, ta, tb);
        if (y + 3 >= dm.up_h) break;
        do_row(y + 3, true, q1, ta, tb, tc);
        if (y + 4 >= dm.up_h) break;
        do_row(y + 4, more, q0, tb, tc, ta);
        if (y + 5 >= dm.up_h) break;
        do_row(y + 5, more, q1, tc, ta, tb);
    }
}

}  // namespace b2r
