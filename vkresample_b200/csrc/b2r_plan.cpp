// b2r_plan.cpp -- frame geometry and FFT scheduling (host, no CUDA runtime calls).
// See b2r_plan.h for the reference regions this replaces.
#include "b2r_plan.h"

#include <algorithm>
#include <cmath>
#include <cstdio>

namespace b2r {

namespace {
const int kRadices[] = {16, 15, 14, 12, 10, 9, 8, 7, 6, 5, 4, 3, 2};

// depth-first search for the shortest non-increasing radix list with product n
void search(int n, int max_r, std::vector<int>& cur, std::vector<int>& best, int total_n) {
    if (n == 1) {
        bool better = best.empty() || cur.size() < best.size();
        if (!better && cur.size() == best.size()) {
            int tc = fft_min_threads(cur, total_n), tb = fft_min_threads(best, total_n);
            better = tc < tb || (tc == tb && cur > best);
        }
        if (better) best = cur;
        return;
    }
    if (!best.empty() && cur.size() + 1 > best.size()) return;
    for (int r : kRadices) {
        if (r > max_r || n % r) continue;
        cur.push_back(r);
        search(n / r, r, cur, best, total_n);
        cur.pop_back();
    }
}
}  // namespace

int fft_min_threads(const std::vector<int>& radices, int n) {
    int t = 1;
    for (int r : radices) {
        int nb = n / r, m = max_per_thread(r);
        t = std::max(t, (nb + m - 1) / m);
    }
    return t;
}

bool factor_radices(int n, std::vector<int>* radices) {
    int m = n;
    for (int p : {2, 3, 5, 7})
        while (m % p == 0) m /= p;
    if (n < 2 || m != 1) return false;
    std::vector<int> cur, best;
    search(n, 16, cur, best, n);
    if (best.empty() || (int)best.size() > kMaxStages) return false;
    *radices = best;
    return true;
}

void build_fft(int n, const int* radices, int nst, int threads, HostFft* out) {
    FftDesc& d = out->desc;
    d = FftDesc{};
    d.n = n;
    d.nstages = nst;
    d.threads = threads;
    out->twiddles.clear();
    out->twiddles_d.clear();
    int stride = 1;
    for (int s = 0; s < d.nstages; ++s) {
        StageDesc& sd = d.st[s];
        sd.radix = radices[s];
        sd.nb = n / sd.radix;
        sd.stride = stride;
        sd.divS.d = (unsigned)stride;
        sd.divS.magic = stride > 1 ? (unsigned)(((1ull << 32) + stride - 1) / stride) : 0u;
        sd.per_thread = (sd.nb + d.threads - 1) / d.threads;
        sd.tw_off = (int)out->twiddles.size();
        if (stride > 1) {
            const double m = (double)stride * sd.radix;
            for (int p = 0; p < stride; ++p) {
                double a = -2.0 * M_PI * (double)p / m;
                out->twiddles.push_back(make_float2((float)std::cos(a), (float)std::sin(a)));
                // double table: exact octant reduction (the same routine the in-kernel constants use)
                out->twiddles_d.push_back(make_double2(cx::cos2pi(-p, (long)m), cx::sin2pi(-p, (long)m)));
            }
        }
        stride *= sd.radix;
    }
}

bool schedule_fft(int n, HostFft* out, std::string* err, int force_threads) {
    std::vector<int> radices;
    if (n >= (1 << 16) || !factor_radices(n, &radices)) {
        if (err) *err = "FFT length " + std::to_string(n) + " is not of the form 2^a 3^b 5^c 7^d (< 65536)";
        return false;
    }
    build_fft(n, radices.data(), (int)radices.size(), std::max(fft_min_threads(radices, n), force_threads), out);
    return true;
}

bool make_geometry(int w, int h, float upscale, int precision, float sharpen, Geometry* g, std::string* err,
                   bool c2c_layout) {
    auto fail = [&](const std::string& m) { if (err) *err = m; return false; };
    if (w < 4 || h < 4 || (w & 1) || (h & 1)) return fail("input width and height must be even and >= 4");
    if (!(upscale >= 1.0f)) return fail("upscale factor must be >= 1");
    if (precision != 0 && precision != 1 && precision != 2)
        return fail("precision must be 0 (fp32), 1 (fp64) or 2 (fp16 storage)");
    Geometry r;
    r.w = w; r.h = h; r.upscale = upscale; r.precision = precision; r.sharpen = sharpen;
    // float products truncated on assignment to uint32_t, as the reference does
    r.up_w = (int)(uint32_t)(upscale * (float)w);
    r.up_h = (int)(uint32_t)(upscale * (float)h);
    if ((r.up_w & 1) || (r.up_h & 1)) return fail("upscaled width and height must be even");
    r.nx = w / 2 + 1;
    r.spec_stride = (r.nx + 15) / 16 * 16;
    r.zp_lo = (int)(uint32_t)((float)r.up_h / (2.0f * upscale));
    r.zp_hi = (int)(uint32_t)(((2.0f * upscale - 1.0f) * (float)r.up_h) / (2.0f * upscale));
    r.neg_shift = r.up_h - h;
    r.up2 = upscale * upscale;
    r.in_row = (size_t)w;            r.in_plane = (size_t)(w + 2) * h;
    r.pre_row = (size_t)r.up_w;      r.pre_plane = c2c_layout ? (size_t)r.up_w * r.up_h : (size_t)(r.up_w + 2) * r.up_h;
    r.out_row = (size_t)r.up_w;      r.out_plane = (size_t)r.up_w * r.up_h;
    r.pre_elems = 3 * r.pre_plane + (size_t)r.up_w + 8;
    std::vector<int> tmp;
    for (int n : {w, h, r.up_w, r.up_h})
        if (n >= (1 << 16) || !factor_radices(n, &tmp))
            return fail("size " + std::to_string(n) + " is not of the form 2^a 3^b 5^c 7^d (< 65536)");
    *g = r;
    return true;
}

}  // namespace b2r
