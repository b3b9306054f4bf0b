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
// shared memory of one CTA: two workspaces, each a multiple of 16 bytes (the rows are read as 16-byte vectors)
B2R_HD constexpr int fused_ws_len(int n) { return (smem_padded_len(n) + 1) & ~1; }
B2R_HD constexpr size_t fused_smem_bytes(int n) { return 2 * (size_t)fused_ws_len(n) * sizeof(real2); }
// fix-up list entry: output row y, or only its last pixel
constexpr int kFixCornerBit = 1 << 30;
// register budget: 168 per thread (two radix-16 butterflies in flight in the FFT phase, the 4 x 6 tap window
// in the sharpen phase) -- three 128-thread CTAs per SM
constexpr int fused_min_blocks(int threads) {
    int b = 65536 / (168 * threads);
    return b < 1 ? 1 : b;
}

#if !defined(B2R_REAL_IS_DOUBLE)

// One group of NP pixels of rows `up`, `mid`, `dn` held in shared memory as clamped magnitudes.
// right_* : the element that follows the row's last one in flat order (used when x0 + NP == n).
template <int NP>
B2R_DEV void fused_taps_f32(const float* row, int x0, int n, float right_end, float (&t)[NP + 2]) {
#pragma unroll
    for (int k = 0; k < NP / 4; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(row + x0 + 4 * k);
        t[4 * k + 1] = v.x; t[4 * k + 2] = v.y; t[4 * k + 3] = v.z; t[4 * k + 4] = v.w;
    }
    t[0] = (x0 > 0) ? row[x0 - 1] : t[1];                       // left clamps at 0
    t[NP + 1] = (x0 + NP < n) ? row[x0 + NP] : right_end;       // right does not clamp: flat +1
}

template <class P, bool UP2>
B2R_KERNEL B2R_LAUNCH_BOUNDS((row_launch_bound<P, 1>()), (fused_min_blocks(row_launch_bound<P, 1>())))
k_c2r_sharpen_f32(const real2* __restrict__ spec, float* __restrict__ out, float* __restrict__ pre,
                  const real2* __restrict__ tw, const P plan, const FrameDims dm, const real scale, const int nsp) {
    constexpr int NP = 4;
    const int T = plan.threads(), tid = (int)B2R_TID_X, n = plan.n();
    const int ppp = dm.up_h >> 1;
    const int c = (int)B2R_BID_X / nsp, q = (int)B2R_BID_X - c * nsp;
    const int j0 = fused_strip_begin(q, nsp, ppp), j1 = fused_strip_begin(q + 1, nsp, ppp);
    const int S = j1 - j0;
    const bool top = (j0 == 0);
    real2* const ws0 = B2R_SMEM(real2);
    const int ws_len = fused_ws_len(n);
    const float up2 = dm.up2, neg_s = -dm.sharpen;
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

    for (int i = 0; i < S; ++i) {
        const int j = j0 + i;
        real2* const wsc = ws0 + (i & 1) * ws_len;
        float* cur = reinterpret_cast<float*>(wsc);                 // rows 2j (cur[0..n)) and 2j+1 (cur[n..2n))
        const float* prv = reinterpret_cast<const float*>(ws0 + ((i & 1) ^ 1) * ws_len);   // rows 2j-2, 2j-1
        // which raw values go to the pre-sharpen buffer for k_sharpen_fix: 1 = both rows, 2 = single elements
        const int pre_mode = (i == 0 || i == S - 1) ? 1 : ((i == 1 || i == S - 2) ? 2 : 0);
        const bool head = (i == 1), tail = (i == S - 2);
        float* p0 = pplane + (size_t)(2 * j) * n;
        float* p1 = p0 + n;
        const real2* a = sp + (size_t)(2 * j) * dm.spec_stride;
        c2r_pair_emit<P, UP2, false, true>(plan, a, a + dm.spec_stride, wsc, tw, dm, tid, true, [&](int idx, real2 z) {
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
        if (i == 0) {
            if (top) {   // row 0 of the plane: the row above clamps to row 0 itself
                for (int g = tid; g < G; g += T) {
                    const int x0 = g * NP;
                    float tc[NP + 2], td[NP + 2], o[NP];
                    fused_taps_f32<NP>(cur, x0, n, cur[n], tc);          // row 0; after its end comes row 1
                    fused_taps_f32<NP>(cur + n, x0, n, 0.f, td);         // row 1; its right end (row 2) is not here yet
                    cas_row_f32<NP>(tc, tc, td, neg_s, o);
                    store4(oplane + x0, o);
                    if (g == g_last) { corner_up0 = tc[NP - 1]; corner_up1 = tc[NP]; }
                }
                corner_pending = true;
            }
        } else {
            const int ya = 2 * j - 1, yb = 2 * j;
            for (int g = tid; g < G; g += T) {
                const int x0 = g * NP;
                float ta[NP + 2], tb[NP + 2], tc[NP + 2], td[NP + 2], o[NP];
                fused_taps_f32<NP>(prv, x0, n, prv[n], ta);              // row 2j-2, then row 2j-1
                fused_taps_f32<NP>(prv + n, x0, n, cur[0], tb);          // row 2j-1, then row 2j (other slot)
                fused_taps_f32<NP>(cur, x0, n, cur[n], tc);              // row 2j, then row 2j+1
                fused_taps_f32<NP>(cur + n, x0, n, 0.f, td);             // row 2j+1; row 2j+2 comes with the next pair
                cas_row_f32<NP>(ta, tb, tc, neg_s, o);
                store4(oplane + (size_t)ya * n + x0, o);
                cas_row_f32<NP>(tb, tc, td, neg_s, o);
                store4(oplane + (size_t)yb * n + x0, o);                 // its last pixel is redone one pair later
                if (g == g_last) {
                    if (corner_pending) {   // pixel (n-1, 2j-2): rows 2j-3 (registers), 2j-2, 2j-1, and row 2j's first tap
                        // the tap after the upper row's end: row 2j-2's first one -- or, for plane row 0 (whose
                        // upper row clamps to row 0 itself), row 1's
                        float up3[3] = {corner_up0, corner_up1, (yb - 2 == 0) ? prv[n] : prv[0]};
                        float mid3[3] = {ta[NP - 1], ta[NP], prv[n]};
                        float dn3[3] = {tb[NP - 1], tb[NP], cur[0]};
                        float o1[1];
                        cas_row_f32<1>(up3, mid3, dn3, neg_s, o1);
                        oplane[(size_t)(yb - 2) * n + n - 1] = o1[0];
                    }
                    corner_up0 = tb[NP - 1]; corner_up1 = tb[NP];
                }
            }
            corner_pending = true;
        }
        B2R_SYNC();   // the other slot becomes the next pair's workspace
    }
}

// Rows (or single last pixels) the fused kernel leaves out, finished from the raw values it stored in the
// pre-sharpen buffer.  grid = (ceil(upW/4/blockDim.x), entries per plane, 3); list: entries of one plane
// (the same for all three), y | kFixCornerBit for "last pixel only".
B2R_DEV void sharpen_fix_f32_impl(const float* __restrict__ pre, float* __restrict__ out, const FrameDims& dm,
                                  const int* __restrict__ list) {
    constexpr int NP = 4;
    const int e = list[B2R_BID_Y];
    const bool corner = (e & kFixCornerBit) != 0;
    const int y = e & (kFixCornerBit - 1), ch = (int)B2R_BID_Z, n = dm.up_w;
    const int x0 = (int)(B2R_BID_X * B2R_BDIM_X + B2R_TID_X) * NP;
    if (x0 >= n) return;
    if (corner && x0 + NP != n) return;
    const float* plane = pre + (size_t)ch * dm.pre_plane;
    const float up2 = dm.up2, neg_s = -dm.sharpen;
    const int yu = y > 0 ? y - 1 : 0;
    float t[3][NP + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float* p = plane + (size_t)(r == 0 ? yu : y + r - 1) * n + x0;
        const float4 v = *reinterpret_cast<const float4*>(p);
        t[r][1] = cas_tap(up2, v.x); t[r][2] = cas_tap(up2, v.y); t[r][3] = cas_tap(up2, v.z); t[r][4] = cas_tap(up2, v.w);
        t[r][0] = (x0 > 0) ? cas_tap(up2, p[-1]) : t[r][1];
        t[r][NP + 1] = cas_tap(up2, p[NP]);     // flat +1
    }
    float o[NP];
    cas_row_f32<NP>(t[0], t[1], t[2], neg_s, o);
    float* dst = out + (size_t)ch * dm.out_plane + (size_t)y * n + x0;
    if (corner) dst[NP - 1] = o[NP - 1];
    else *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
}
template <int DUMMY>
B2R_KERNEL k_sharpen_fix_f32(const float* __restrict__ pre, float* __restrict__ out, const FrameDims dm,
                             const int* __restrict__ list) {
    sharpen_fix_f32_impl(pre, out, dm, list);
}

#endif  // !B2R_REAL_IS_DOUBLE

}  // namespace b2r
