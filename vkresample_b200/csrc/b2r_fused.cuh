// b2r_fused.cuh -- K7 + K8 in ONE kernel: inverse C2R rows whose result never reaches HBM as a plane.
//
// Replaces the reference's inverse axis-0 C2R dispatches (vkFFT.h:8246-8288, read :2059-2201, write
// :4378-4491) AND the sharpen dispatch (shaderGenSharpen r2c branch, VkResample.cpp:849-923), which
// communicate through `tempBuffer` (3*(upW+2)*upH elements written, then read nine times per pixel).
// Here a CTA owns a STRIP of consecutive row pairs of one colour plane:
//   * two shared-memory slots, each the FFT workspace of one row pair; the last FFT stage leaves the
//     pair IN PLACE as two real rows of clamped magnitudes t = min(|up2*v|, 1) -- all the sharpen reads;
//   * after pair j (rows 2j, 2j+1) the slots hold rows 2j-2 .. 2j+1 and the CTA sharpens rows 2j-1 and 2j
//     straight from shared memory with the tolerance-bound arithmetic of b2r_cas.cuh (cas_row_f32 /
//     cas_row_f16: identical expressions, so the output equals the separate kernels' bit for bit);
//   * the flat neighbour rule (right neighbour of a row's last pixel = first pixel of the NEXT row,
//     VkResample.cpp:888-892) makes pixel (upW-1, 2j) need row 2j+2: it is finished one pair later by the
//     thread that owns the row's last pixel group (two taps of row 2j-1 kept in registers);
//   * rows that need a neighbouring strip -- the last row of a strip, the first row of the next one, the
//     last pixel of the row before, and each plane's last row (whose lower neighbours are the plane's pad
//     region) -- are NOT produced here: the CTA stores the raw C2R values of its first and last pair (plus
//     three single elements) into the ordinary pre-sharpen buffer and k_sharpen_fix finishes those rows
//     from there (2 of every 2*S rows; S = pairs per strip).
// HBM traffic per frame at c2: spectrum in 50 MB + output 101 MB + ~10 % boundary rows, instead of
// 50 + 101 (pre write) + 101 (pre read) + 101 (out).
#pragma once

namespace b2r {

// ---- strip geometry (host and device agree through these) -----------------------------------------
B2R_HD int fused_strip_begin(int q, int nsp, int pairs_per_plane) { return (int)(((long long)q * pairs_per_plane) / nsp); }
// shared memory of one CTA: [2 mbarriers | staging 0 | staging 1 | workspace 0 | workspace 1].  A staging
// buffer holds the two spectrum rows of one pair (bulk-copied one pair ahead, like k_c2r_rows_bulk); the
// workspaces are multiples of 16 bytes (the rows they end up holding are read as 16-byte vectors).
B2R_HD constexpr int fused_ws_len(int n) { return (smem_padded_len(n) + 1) & ~1; }
B2R_HD constexpr size_t fused_smem_bytes(int n, int nx) {
    return 16 + 4 * (size_t)c2r_stage_row_elems(nx) * sizeof(real2) + 2 * (size_t)fused_ws_len(n) * sizeof(real2);
}
// fix-up list: one entry per strip boundary of a plane = the first row of the lower strip; the plane's end is
// the entry upH (b2r_api.cu builds it; the same list serves the three planes)
// register budget per thread: 128 with one butterfly per thread in every stage (two 256-thread CTAs per SM),
// 168 when a thread holds more than 16 complex values
template <class P> constexpr int fused_min_blocks() {
    const int regs = (P::max_elems() > 16) ? 168 : 128;
    const int b = 65536 / (regs * P::kT);
    return b < 1 ? 1 : (b > 6 ? 6 : b);
}

#if !defined(B2R_REAL_IS_DOUBLE)

// Clamped magnitudes of columns x0-1 .. x0+3 of one row held in shared memory, given the row's four own values
// `v` (loaded ahead of time).  right_end: the element that follows the row's last one in flat order (used when
// x0 + 4 == n).  SHFL: the halo columns come from the neighbouring lanes -- a scalar shared load with a 16-byte
// lane stride is a 4-way bank conflict -- which needs every lane of a full warp to take part; otherwise (CTAs
// that are not a multiple of 32 threads, and the CPU emulator) they are loaded.
template <bool SHFL>
B2R_DEV void fused_taps4(const float* row, const float4 v, int x0, int n, float right_end, float (&t)[6]) {
    t[1] = v.x; t[2] = v.y; t[3] = v.z; t[4] = v.w;
#if !defined(B2R_HOST_EMU)
    if constexpr (SHFL) {
        const int lane = (int)B2R_TID_X & 31;
        float l = __shfl_up_sync(0xffffffffu, v.w, 1);
        float r = __shfl_down_sync(0xffffffffu, v.x, 1);
        if (lane == 0) l = row[x0 > 0 ? x0 - 1 : 0];             // left clamps at 0
        if (lane == 31) r = row[x0 + 4 < n ? x0 + 4 : x0];
        t[0] = l;
        t[5] = (x0 + 4 < n) ? r : right_end;                     // right does not clamp: flat +1
        return;
    }
#endif
    t[0] = row[x0 > 0 ? x0 - 1 : 0];
    t[5] = (x0 + 4 < n) ? row[x0 + 4] : right_end;
}

// thread count of the fused kernel per row length: one butterfly per thread in the widest stage (twice the
// warps of the stand-alone C2R schedules, which run two butterflies per thread: the sharpen phase needs
// the extra warps to cover its MUFU / shared-memory latencies; measured in profiles/)
template <class P> struct FusedSchedule { using type = P; };
template <> struct FusedSchedule<StaticFft<4096, 128, 16, 16, 16>> { using type = StaticFft<4096, 256, 16, 16, 16>; };
template <> struct FusedSchedule<StaticFft<3840, 128, 16, 16, 15>> { using type = StaticFft<3840, 256, 16, 16, 15>; };
template <> struct FusedSchedule<StaticFft<7680, 384, 16, 20, 24>> { using type = StaticFft<7680, 512, 16, 20, 24>; };
template <> struct FusedSchedule<StaticFft<5120, 160, 20, 16, 16>> { using type = StaticFft<5120, 320, 20, 16, 16>; };
template <> struct FusedSchedule<StaticFft<2560, 128, 16, 16, 10>> { using type = StaticFft<2560, 256, 16, 16, 10>; };

template <class P, bool UP2>
B2R_KERNEL B2R_LAUNCH_BOUNDS((P::kT), (fused_min_blocks<P>()))
k_c2r_sharpen_f32(const real2* __restrict__ spec, float* __restrict__ out, float* __restrict__ pre,
                  const real2* __restrict__ tw, const P plan, const FrameDims dm, const real scale, const int nsp) {
    constexpr int NP = 4;
    const int T = plan.threads(), tid = (int)B2R_TID_X, n = plan.n();
    const int ppp = dm.up_h >> 1;
    const int c = (int)B2R_BID_X / nsp, q = (int)B2R_BID_X - c * nsp;
    const int j0 = fused_strip_begin(q, nsp, ppp), j1 = fused_strip_begin(q + 1, nsp, ppp);
    const int S = j1 - j0;
    const bool top = (j0 == 0);
    const int row_elems = c2r_stage_row_elems(dm.nx);
    unsigned char* const smem_base = B2R_SMEM(unsigned char);
    unsigned long long* const bar = reinterpret_cast<unsigned long long*>(smem_base);
    real2* const stg = reinterpret_cast<real2*>(smem_base + 16);
    real2* const ws0 = stg + 4 * (size_t)row_elems;
    const int ws_len = fused_ws_len(n);
    const float up2 = dm.up2;
    const CasK ks = {dm.cas_a, dm.cas_b};
    const real2* sp = spec + (size_t)c * dm.up_h * dm.spec_stride;
    float* oplane = out + (size_t)c * dm.out_plane;
    float* pplane = pre + (size_t)c * dm.pre_plane;
    const int G = n / NP;                       // pixel groups per row
    const int g_last = G - 1;
    float corner_up0 = 0.f, corner_up1 = 0.f;   // taps (n-2, n-1) of the row above the pending corner pixel
    bool corner_pending = false;                // CTA-uniform

    auto store4 = [&](float* dst, const float (&o)[NP]) {
        const float4 v = make_float4(o[0], o[1], o[2], o[3]);
#if defined(__CUDA_ARCH__)
        __stcs(reinterpret_cast<float4*>(dst), v);
#else
        *reinterpret_cast<float4*>(dst) = v;
#endif
    };

    // producer side (thread 0): both spectrum rows of pair `pj` -> staging buffer `buf` (cp.async.bulk + mbarrier)
    auto issue = [&](int pj, int buf) {
        const real2* src = sp + (size_t)(2 * pj) * dm.spec_stride;
        real2* dst = stg + (size_t)buf * 2 * row_elems;
#if defined(B2R_HOST_EMU)
        for (int e = 0; e < row_elems; ++e) { dst[e] = src[e]; dst[row_elems + e] = src[dm.spec_stride + e]; }
#else
        const unsigned bytes = (unsigned)(row_elems * sizeof(real2));
        b2r_mbar_expect_tx(&bar[buf], 2 * bytes);
        b2r_bulk_g2s(dst, src, bytes, &bar[buf]);
        b2r_bulk_g2s(dst + row_elems, src + dm.spec_stride, bytes, &bar[buf]);
#endif
    };
#if !defined(B2R_HOST_EMU)
    if (tid == 0) { b2r_mbar_init(&bar[0], 1); b2r_mbar_init(&bar[1], 1); b2r_mbar_fence_init(); }
    B2R_SYNC();
#endif
    if (tid == 0) issue(j0, 0);

    for (int i = 0; i < S; ++i) {
        const int j = j0 + i;
        // the next pair's rows travel while this pair is transformed and sharpened; its staging buffer was last
        // read by the first FFT stage of pair i-1, two CTA barriers ago
        if (tid == 0 && i + 1 < S) issue(j + 1, (i + 1) & 1);
#if defined(B2R_HOST_EMU)
        B2R_SYNC();
#else
        b2r_mbar_wait(&bar[i & 1], (unsigned)((i >> 1) & 1));
#endif
        const real2* const sa = stg + (size_t)(i & 1) * 2 * row_elems;
        real2* const wsc = ws0 + (i & 1) * ws_len;
        float* cur = reinterpret_cast<float*>(wsc);                 // rows 2j (cur[0..n)) and 2j+1 (cur[n..2n))
        const float* prv = reinterpret_cast<const float*>(ws0 + ((i & 1) ^ 1) * ws_len);   // rows 2j-2, 2j-1
        // which raw values go to the pre-sharpen buffer for k_sharpen_fix: 1 = both rows, 2 = single elements
        const int pre_mode = (i == 0 || i == S - 1) ? 1 : ((i == 1 || i == S - 2) ? 2 : 0);
        const bool head = (i == 1), tail = (i == S - 2);
        float* p0 = pplane + (size_t)(2 * j) * n;
        float* p1 = p0 + n;
        c2r_pair_emit<P, UP2, true, true>(plan, sa, sa + row_elems, wsc, tw, dm, tid, true, [&](int idx, real2 z) {
            const float v0 = z.x * scale, v1 = z.y * scale;
            cur[idx] = cas_tap(up2, v0);
            cur[n + idx] = cas_tap(up2, v1);
            if (pre_mode == 1) { p0[idx] = v0; p1[idx] = v1; }
            else if (pre_mode == 2) {
                if (head && idx == 0) p0[0] = v0;
                if (tail && idx >= n - 2) p1[idx] = v1;
            }
        });
        B2R_SYNC();
        constexpr bool kShfl = (P::kT % 32 == 0);
        constexpr int K = (P::kN / NP + P::kT - 1) / P::kT;      // pixel groups per thread and row
        const float* rowa = prv;          // row 2j-2
        const float* rowb = prv + n;      // row 2j-1
        const float* rowc = cur;          // row 2j
        const float* rowd = cur + n;      // row 2j+1
        auto ld4 = [&](const float* row, int x0) { return *reinterpret_cast<const float4*>(row + x0); };
        auto x_of = [&](int k) { const int g = k * T + tid; return (g < G ? g : g_last) * NP; };   // lanes past the row end re-read the last group
        if (i == 0) {
            if (top) {   // row 0 of the plane: the row above clamps to row 0 itself
#pragma unroll 1
                for (int k = 0; k < K; ++k) {     // every lane runs every trip (warp shuffles inside)
                    const bool valid = k * T + tid < G;
                    const int x0 = x_of(k);
                    float tc[NP + 2], td[NP + 2], o[NP];
                    fused_taps4<kShfl>(rowc, ld4(rowc, x0), x0, n, rowc[n], tc);   // row 0; after its end comes row 1
                    fused_taps4<kShfl>(rowd, ld4(rowd, x0), x0, n, 0.f, td);       // row 1; its right end (row 2) is not here yet
                    cas_row_f32<NP>(tc, tc, td, ks, o);
                    if (valid) store4(oplane + x0, o);
                }
                corner_pending = true;
                if (tid == g_last % T) { corner_up0 = rowc[n - 2]; corner_up1 = rowc[n - 1]; }
            }
        } else {
            const int ya = 2 * j - 1, yb = 2 * j;
            const float end_a = rowa[n], end_b = rowc[0], end_c = rowc[n];   // the elements after each row's end (flat order)
            // the four rows of the NEXT group are loaded before the current group is computed
            float4 va = ld4(rowa, x_of(0)), vb = ld4(rowb, x_of(0)), vc = ld4(rowc, x_of(0)), vd = ld4(rowd, x_of(0));
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const bool valid = k * T + tid < G;
                const int x0 = x_of(k);
                float4 na = va, nb = vb, nc = vc, nd = vd;
                if (k + 1 < K) { const int x1 = x_of(k + 1); na = ld4(rowa, x1); nb = ld4(rowb, x1); nc = ld4(rowc, x1); nd = ld4(rowd, x1); }
                float ta[NP + 2], tb[NP + 2], tc[NP + 2], td[NP + 2], o[NP];
                fused_taps4<kShfl>(rowa, va, x0, n, end_a, ta);
                fused_taps4<kShfl>(rowb, vb, x0, n, end_b, tb);
                fused_taps4<kShfl>(rowc, vc, x0, n, end_c, tc);
                fused_taps4<kShfl>(rowd, vd, x0, n, 0.f, td);                // row 2j+2 comes with the next pair
                cas_row_f32<NP>(ta, tb, tc, ks, o);
                if (valid) store4(oplane + (size_t)ya * n + x0, o);
                cas_row_f32<NP>(tb, tc, td, ks, o);
                if (valid) store4(oplane + (size_t)yb * n + x0, o);          // its last pixel is redone one pair later
                va = na; vb = nb; vc = nc; vd = nd;
            }
            if (tid == g_last % T) {   // the thread that owns the rows' last pixel group
                if (corner_pending) {  // pixel (n-1, 2j-2): rows 2j-3 (registers), 2j-2, 2j-1, and row 2j's first tap
                    // the tap after the upper row's end: row 2j-2's first one -- or, for plane row 0 (whose
                    // upper row clamps to row 0 itself), row 1's
                    float up3[3] = {corner_up0, corner_up1, (yb - 2 == 0) ? rowa[n] : rowa[0]};
                    float mid3[3] = {rowa[n - 2], rowa[n - 1], rowa[n]};
                    float dn3[3] = {rowb[n - 2], rowb[n - 1], rowc[0]};
                    float o1[1];
                    cas_row_f32<1>(up3, mid3, dn3, ks, o1);
                    oplane[(size_t)(yb - 2) * n + n - 1] = o1[0];
                }
                corner_up0 = rowb[n - 2]; corner_up1 = rowb[n - 1];
            }
            corner_pending = true;
        }
        B2R_SYNC();   // the other slot becomes the next pair's workspace
    }
}

// Rows (and single last pixels) the fused kernel leaves out, finished from the raw values it stored in the
// pre-sharpen buffer.  One list entry = one strip boundary b (the first row of the lower strip; b == upH for the
// plane's end): a thread loads its 8 columns of rows b-2 .. b+1 once and produces rows b-1 and b (only b-1 at
// the plane's end); the thread that owns the rows' last pixels also redoes pixel (upW-1, b-2), whose lower-right
// tap is row b's first element.  grid = (ceil(upW/8/blockDim.x), boundaries per plane, 3).
B2R_DEV void sharpen_fix_f32_impl(const float* __restrict__ pre, float* __restrict__ out, const FrameDims& dm,
                                  const int* __restrict__ list) {
    constexpr int NP = 8;
    const int b = list[B2R_BID_Y], ch = (int)B2R_BID_Z, n = dm.up_w;
    const int x0 = (int)(B2R_BID_X * B2R_BDIM_X + B2R_TID_X) * NP;
    if (x0 >= n) return;
    const float* plane = pre + (size_t)ch * dm.pre_plane;
    float* oplane = out + (size_t)ch * dm.out_plane;
    const float up2 = dm.up2;
    const CasK ks = {dm.cas_a, dm.cas_b};
    const bool plane_end = (b >= dm.up_h);
    // clamped magnitudes of columns x0-1 .. x0+8 of row y (flat +1 on the right, clamp on the left)
    auto taps = [&](int y, float (&t)[NP + 2]) {
        const float* p = plane + (size_t)y * n + x0;
        const float4 v0 = *reinterpret_cast<const float4*>(p), v1 = *reinterpret_cast<const float4*>(p + 4);
        t[1] = cas_tap(up2, v0.x); t[2] = cas_tap(up2, v0.y); t[3] = cas_tap(up2, v0.z); t[4] = cas_tap(up2, v0.w);
        t[5] = cas_tap(up2, v1.x); t[6] = cas_tap(up2, v1.y); t[7] = cas_tap(up2, v1.z); t[8] = cas_tap(up2, v1.w);
        t[0] = (x0 > 0) ? cas_tap(up2, p[-1]) : t[1];
        t[NP + 1] = cas_tap(up2, p[NP]);
    };
    auto store8 = [&](int y, const float (&o)[NP]) {
        float* dst = oplane + (size_t)y * n + x0;
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
    };
    float ta[NP + 2], tb[NP + 2], tc[NP + 2], td[NP + 2], o[NP];
    taps(b - 2, ta);
    taps(b - 1, tb);
    taps(b, tc);                       // at the plane's end: the zero pad region below the plane
    cas_row_f32<NP>(ta, tb, tc, ks, o);
    store8(b - 1, o);
    if (!plane_end) {
        taps(b + 1, td);
        cas_row_f32<NP>(tb, tc, td, ks, o);
        store8(b, o);
    }
    if (x0 + NP == n) {                // pixel (n-1, b-2): rows b-3 (two elements were stored), b-2, b-1
        const float* pu = plane + (size_t)(b - 3) * n;
        float up3[3] = {cas_tap(up2, pu[n - 2]), cas_tap(up2, pu[n - 1]), cas_tap(up2, pu[n])};
        float mid3[3] = {ta[NP - 1], ta[NP], ta[NP + 1]};
        float dn3[3] = {tb[NP - 1], tb[NP], tb[NP + 1]};
        float o1[1];
        cas_row_f32<1>(up3, mid3, dn3, ks, o1);
        oplane[(size_t)(b - 2) * n + n - 1] = o1[0];
    }
}
template <int DUMMY>
B2R_KERNEL k_sharpen_fix_f32(const float* __restrict__ pre, float* __restrict__ out, const FrameDims dm,
                             const int* __restrict__ list) {
    sharpen_fix_f32_impl(pre, out, dm, list);
}

// ---- fp16 storage (precision 2) ---------------------------------------------------------------------
// Same strip structure; the rows are kept as HALF clamped magnitudes (the reference's sharpen shader computes
// in float16_t on the half C2R output, VkResample.cpp:823-827, vkFFT.h:3525-3527), eight pixels per group and
// two pixels per instruction (cas_row_f16).  t = min(|up2_h * half(v)|, 1) with individually rounded half
// operations, i.e. exactly what k_sharpen_fast_f16 derives from the stored plane.
#if defined(__CUDA_ARCH__)
B2R_DEV __half fused_tap_h(__half up2, __half v) { return __hmin(__habs(__hmul(up2, v)), __float2half_rn(1.0f)); }
#else
B2R_DEV __half fused_tap_h(__half up2, __half v) {
    const __half2 r = cas_tap2(__halves2half2(up2, up2), __halves2half2(v, v));
    return __low2half(r);
}
#endif

B2R_DEV __half2 fused_u2h(unsigned w) { __half2 h; *reinterpret_cast<unsigned*>(&h) = w; return h; }
B2R_DEV unsigned fused_h2u(__half2 h) { return *reinterpret_cast<unsigned*>(&h); }

// the thread's eight own values `v` of one row + halo columns -> CasRowH<8>
template <bool SHFL>
B2R_DEV void fused_taps8h(const __half* row, const uint4 v, int x0, int n, __half right_end, CasRowH<8>& t) {
    t.b[0] = fused_u2h(v.x); t.b[1] = fused_u2h(v.y); t.b[2] = fused_u2h(v.z); t.b[3] = fused_u2h(v.w);
    __half l, r;
#if !defined(B2R_HOST_EMU)
    if constexpr (SHFL) {
        const int lane = (int)B2R_TID_X & 31;
        l = __high2half(fused_u2h(__shfl_up_sync(0xffffffffu, v.w, 1)));
        r = __low2half(fused_u2h(__shfl_down_sync(0xffffffffu, v.x, 1)));
        if (lane == 0) l = row[x0 > 0 ? x0 - 1 : 0];
        if (lane == 31) r = row[x0 + 8 < n ? x0 + 8 : x0];
        if (x0 + 8 >= n) r = right_end;
        t.link(l, r);
        return;
    }
#endif
    l = row[x0 > 0 ? x0 - 1 : 0];
    r = (x0 + 8 < n) ? row[x0 + 8] : right_end;
    t.link(l, r);
}

template <class P, bool UP2>
B2R_KERNEL B2R_LAUNCH_BOUNDS((P::kT), (fused_min_blocks<P>()))
k_c2r_sharpen_f16(const real2* __restrict__ spec, __half* __restrict__ out, __half* __restrict__ pre,
                  const real2* __restrict__ tw, const P plan, const FrameDims dm, const real scale, const int nsp) {
    constexpr int NP = 8;
    const int T = plan.threads(), tid = (int)B2R_TID_X, n = plan.n();
    const int ppp = dm.up_h >> 1;
    const int c = (int)B2R_BID_X / nsp, q = (int)B2R_BID_X - c * nsp;
    const int j0 = fused_strip_begin(q, nsp, ppp), j1 = fused_strip_begin(q + 1, nsp, ppp);
    const int S = j1 - j0;
    const bool top = (j0 == 0);
    const int row_elems = c2r_stage_row_elems(dm.nx);
    unsigned char* const smem_base = B2R_SMEM(unsigned char);
    unsigned long long* const bar = reinterpret_cast<unsigned long long*>(smem_base);
    real2* const stg = reinterpret_cast<real2*>(smem_base + 16);
    real2* const ws0 = stg + 4 * (size_t)row_elems;
    const int ws_len = fused_ws_len(n);
    const __half up2 = __float2half_rn(dm.up2);
    const CasK ks = {dm.cas_a, dm.cas_b};
    const __half hzero = __float2half_rn(0.f);
    const real2* sp = spec + (size_t)c * dm.up_h * dm.spec_stride;
    __half* oplane = out + (size_t)c * dm.out_plane;
    __half* pplane = pre + (size_t)c * dm.pre_plane;
    const int G = n / NP, g_last = G - 1;
    __half corner_up[3] = {hzero, hzero, hzero};   // taps (n-3, n-2, n-1) of the row above the pending corner pixel
    bool corner_pending = false;

    auto store8 = [&](__half* dst, const __half2 (&o)[4]) {
        const uint4 v = make_uint4(fused_h2u(o[0]), fused_h2u(o[1]), fused_h2u(o[2]), fused_h2u(o[3]));
#if defined(__CUDA_ARCH__)
        __stcs(reinterpret_cast<uint4*>(dst), v);
#else
        *reinterpret_cast<uint4*>(dst) = v;
#endif
    };
    auto issue = [&](int pj, int buf) {
        const real2* src = sp + (size_t)(2 * pj) * dm.spec_stride;
        real2* dst = stg + (size_t)buf * 2 * row_elems;
#if defined(B2R_HOST_EMU)
        for (int e = 0; e < row_elems; ++e) { dst[e] = src[e]; dst[row_elems + e] = src[dm.spec_stride + e]; }
#else
        const unsigned bytes = (unsigned)(row_elems * sizeof(real2));
        b2r_mbar_expect_tx(&bar[buf], 2 * bytes);
        b2r_bulk_g2s(dst, src, bytes, &bar[buf]);
        b2r_bulk_g2s(dst + row_elems, src + dm.spec_stride, bytes, &bar[buf]);
#endif
    };
#if !defined(B2R_HOST_EMU)
    if (tid == 0) { b2r_mbar_init(&bar[0], 1); b2r_mbar_init(&bar[1], 1); b2r_mbar_fence_init(); }
    B2R_SYNC();
#endif
    if (tid == 0) issue(j0, 0);

    for (int i = 0; i < S; ++i) {
        const int j = j0 + i;
        if (tid == 0 && i + 1 < S) issue(j + 1, (i + 1) & 1);
#if defined(B2R_HOST_EMU)
        B2R_SYNC();
#else
        b2r_mbar_wait(&bar[i & 1], (unsigned)((i >> 1) & 1));
#endif
        const real2* const sa = stg + (size_t)(i & 1) * 2 * row_elems;
        real2* const wsc = ws0 + (i & 1) * ws_len;
        __half* cur = reinterpret_cast<__half*>(wsc);                 // rows 2j (cur[0..n)) and 2j+1 (cur[n..2n))
        const __half* prv = reinterpret_cast<const __half*>(ws0 + ((i & 1) ^ 1) * ws_len);
        const int pre_mode = (i == 0 || i == S - 1) ? 1 : ((i == 1 || i == S - 2) ? 2 : 0);
        const bool head = (i == 1), tail = (i == S - 2);
        __half* p0 = pplane + (size_t)(2 * j) * n;
        __half* p1 = p0 + n;
        c2r_pair_emit<P, UP2, true, true>(plan, sa, sa + row_elems, wsc, tw, dm, tid, true, [&](int idx, real2 z) {
            const __half v0 = __float2half_rn((float)(z.x * scale)), v1 = __float2half_rn((float)(z.y * scale));
            cur[idx] = fused_tap_h(up2, v0);
            cur[n + idx] = fused_tap_h(up2, v1);
            if (pre_mode == 1) { p0[idx] = v0; p1[idx] = v1; }
            else if (pre_mode == 2) {
                if (head && idx == 0) p0[0] = v0;
                if (tail && idx >= n - 3) p1[idx] = v1;
            }
        });
        B2R_SYNC();
        constexpr bool kShfl = (P::kT % 32 == 0);
        constexpr int K = (P::kN / NP + P::kT - 1) / P::kT;
        const __half* rowa = prv;
        const __half* rowb = prv + n;
        const __half* rowc = cur;
        const __half* rowd = cur + n;
        auto ld8 = [&](const __half* row, int x0) { return *reinterpret_cast<const uint4*>(row + x0); };
        auto x_of = [&](int k) { const int g = k * T + tid; return (g < G ? g : g_last) * NP; };
        // the pixel pair (n-2, n-1) of a row whose lower-right tap arrived one pair late
        auto corner_pixel = [&](const __half (&u)[3], __half u_end, const __half* mid, __half m_end, const __half* dn, __half d_end, int y) {
            CasRowH<2> tu, tm, td;
            tu.b[0] = __halves2half2(u[1], u[2]); tu.link(u[0], u_end);
            tm.b[0] = __halves2half2(mid[n - 2], mid[n - 1]); tm.link(mid[n - 3], m_end);
            td.b[0] = __halves2half2(dn[n - 2], dn[n - 1]); td.link(dn[n - 3], d_end);
            __half2 o1[1];
            cas_row_f16<2>(tu, tm, td, ks, o1);
            oplane[(size_t)y * n + n - 1] = __high2half(o1[0]);
        };
        if (i == 0) {
            if (top) {
#pragma unroll 1
                for (int k = 0; k < K; ++k) {
                    const bool valid = k * T + tid < G;
                    const int x0 = x_of(k);
                    CasRowH<NP> tc, td;
                    __half2 o[4];
                    fused_taps8h<kShfl>(rowc, ld8(rowc, x0), x0, n, rowc[n], tc);
                    fused_taps8h<kShfl>(rowd, ld8(rowd, x0), x0, n, hzero, td);
                    cas_row_f16<NP>(tc, tc, td, ks, o);
                    if (valid) store8(oplane + x0, o);
                }
                corner_pending = true;
                if (tid == g_last % T) { corner_up[0] = rowc[n - 3]; corner_up[1] = rowc[n - 2]; corner_up[2] = rowc[n - 1]; }
            }
        } else {
            const int ya = 2 * j - 1, yb = 2 * j;
            const __half end_a = rowa[n], end_b = rowc[0], end_c = rowc[n];
            uint4 va = ld8(rowa, x_of(0)), vb = ld8(rowb, x_of(0)), vc = ld8(rowc, x_of(0)), vd = ld8(rowd, x_of(0));
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const bool valid = k * T + tid < G;
                const int x0 = x_of(k);
                uint4 na = va, nb = vb, nc = vc, nd = vd;
                if (k + 1 < K) { const int x1 = x_of(k + 1); na = ld8(rowa, x1); nb = ld8(rowb, x1); nc = ld8(rowc, x1); nd = ld8(rowd, x1); }
                CasRowH<NP> ta, tb, tc, td;
                __half2 o[4];
                fused_taps8h<kShfl>(rowa, va, x0, n, end_a, ta);
                fused_taps8h<kShfl>(rowb, vb, x0, n, end_b, tb);
                fused_taps8h<kShfl>(rowc, vc, x0, n, end_c, tc);
                fused_taps8h<kShfl>(rowd, vd, x0, n, hzero, td);
                cas_row_f16<NP>(ta, tb, tc, ks, o);
                if (valid) store8(oplane + (size_t)ya * n + x0, o);
                cas_row_f16<NP>(tb, tc, td, ks, o);
                if (valid) store8(oplane + (size_t)yb * n + x0, o);
                va = na; vb = nb; vc = nc; vd = nd;
            }
            if (tid == g_last % T) {
                if (corner_pending)
                    corner_pixel(corner_up, (yb - 2 == 0) ? rowa[n] : rowa[0], rowa, rowa[n], rowb, rowc[0], yb - 2);
                corner_up[0] = rowb[n - 3]; corner_up[1] = rowb[n - 2]; corner_up[2] = rowb[n - 1];
            }
            corner_pending = true;
        }
        B2R_SYNC();
    }
}

B2R_DEV void sharpen_fix_f16_impl(const __half* __restrict__ pre, __half* __restrict__ out, const FrameDims& dm,
                                  const int* __restrict__ list) {
    constexpr int NP = 8;
    const int b = list[B2R_BID_Y], ch = (int)B2R_BID_Z, n = dm.up_w;
    const int x0 = (int)(B2R_BID_X * B2R_BDIM_X + B2R_TID_X) * NP;
    if (x0 >= n) return;
    const __half* plane = pre + (size_t)ch * dm.pre_plane;
    __half* oplane = out + (size_t)ch * dm.out_plane;
    const __half2 up2 = __float2half2_rn(dm.up2);
    const CasK ks = {dm.cas_a, dm.cas_b};
    const bool plane_end = (b >= dm.up_h);
    auto taps = [&](int y, CasRowH<NP>& t) {
        const __half* p = plane + (size_t)y * n + x0;
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        t.b[0] = cas_tap2(up2, fused_u2h(v.x)); t.b[1] = cas_tap2(up2, fused_u2h(v.y));
        t.b[2] = cas_tap2(up2, fused_u2h(v.z)); t.b[3] = cas_tap2(up2, fused_u2h(v.w));
        const __half2 ed = cas_tap2(up2, __halves2half2(p[x0 > 0 ? -1 : 0], p[NP]));   // flat +1 on the right
        t.link(__low2half(ed), __high2half(ed));
    };
    auto store8 = [&](int y, const __half2 (&o)[4]) {
        *reinterpret_cast<uint4*>(oplane + (size_t)y * n + x0) = make_uint4(fused_h2u(o[0]), fused_h2u(o[1]), fused_h2u(o[2]), fused_h2u(o[3]));
    };
    CasRowH<NP> ta, tb, tc, td;
    __half2 o[4];
    taps(b - 2, ta);
    taps(b - 1, tb);
    taps(b, tc);
    cas_row_f16<NP>(ta, tb, tc, ks, o);
    store8(b - 1, o);
    if (!plane_end) {
        taps(b + 1, td);
        cas_row_f16<NP>(tb, tc, td, ks, o);
        store8(b, o);
    }
    if (x0 + NP == n) {                // pixel (n-1, b-2)
        const __half* pu = plane + (size_t)(b - 3) * n;
        CasRowH<2> tu, tm, tdn;
        const __half2 u01 = cas_tap2(up2, __halves2half2(pu[n - 2], pu[n - 1])), u2e = cas_tap2(up2, __halves2half2(pu[n - 3], pu[n]));
        tu.b[0] = u01; tu.link(__low2half(u2e), __high2half(u2e));
        tm.b[0] = ta.b[3]; tm.link(__high2half(ta.b[2]), __high2half(ta.a[4]));
        tdn.b[0] = tb.b[3]; tdn.link(__high2half(tb.b[2]), __high2half(tb.a[4]));
        __half2 o1[1];
        cas_row_f16<2>(tu, tm, tdn, ks, o1);
        oplane[(size_t)(b - 2) * n + n - 1] = __high2half(o1[0]);
    }
}
template <int DUMMY>
B2R_KERNEL k_sharpen_fix_f16(const __half* __restrict__ pre, __half* __restrict__ out, const FrameDims dm,
                             const int* __restrict__ list) {
    sharpen_fix_f16_impl(pre, out, dm, list);
}

#endif  // !B2R_REAL_IS_DOUBLE

}  // namespace b2r
