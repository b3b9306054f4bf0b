// b2r_fft.cuh -- in-register DFT butterflies and the shared-memory Stockham stage engine.
//
// Replaces (semantics only) the GLSL the reference generates at plan time:
//   inlineRadixKernelVkFFT   vkFFT.h:731-1182   (radix-2/3/4/5/7/8 butterflies)
//   appendRadixStage*        vkFFT.h:2390-2705  (shared -> registers, twiddle, butterfly)
//   appendRadixShuffle*      vkFFT.h:2917-3156  (registers -> shared, Stockham index)
// Nothing here is derived from that text: the butterflies are templates over a compile-time
// radix (prime radices written out, composite radices 6..16 built as R1 x R2 in registers with
// compile-time inner twiddles), the outer twiddle of a stage is ONE table load per butterfly
// (w = exp(-2*pi*i*p/(S*R))) whose powers are formed by a log-depth multiplication tree, and the
// stage engine is "read all -> barrier -> write all" in place on a bank-padded shared array.
//
// Sign convention: DIR = -1 forward (e^{-i..}), DIR = +1 inverse (e^{+i..}); no normalisation.
// The file compiles for sm_100a with nvcc and, with B2R_HOST_EMU defined, as plain C++ for the
// CPU thread emulator under tests/emu (test infrastructure; never part of the product path).
#pragma once

#include "b2r_common.cuh"

#if !defined(__CUDACC_RTC__)
#include <cmath>
#endif

namespace b2r {

// ------------------------------------------------------------------ complex helpers
B2R_DEV real2 cadd(real2 a, real2 b) { return make_real2(a.x + b.x, a.y + b.y); }
B2R_DEV real2 csub(real2 a, real2 b) { return make_real2(a.x - b.x, a.y - b.y); }
B2R_DEV real2 cmul(real2 a, real2 b) {
    return make_real2(rfma(-a.y, b.y, a.x * b.x), rfma(a.y, b.x, a.x * b.y));
}
B2R_DEV real2 cscale(real2 a, real s) { return make_real2(a.x * s, a.y * s); }
B2R_DEV real2 cconj(real2 a) { return make_real2(a.x, -a.y); }
// multiply by DIR * i  (a quarter turn in the transform's direction)
template <int DIR> B2R_DEV real2 rotq(real2 a) {
    if constexpr (DIR < 0) return make_real2(a.y, -a.x);
    else return make_real2(-a.y, a.x);
}

// ------------------------------------------------------------------ compile-time trig
namespace cx {
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double sin_taylor(double x) {  // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int k = 1; k < 14; ++k) { term *= -x2 / double((2 * k) * (2 * k + 1)); sum += term; }
    return sum;
}
constexpr double cos_taylor(double x) {  // |x| <= pi/4
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int k = 1; k < 14; ++k) { term *= -x2 / double((2 * k - 1) * (2 * k)); sum += term; }
    return sum;
}
// cos / sin of 2*pi*num/den with exact octant reduction (exact 0, +-1 at multiples of pi/2)
constexpr double cos2pi(long num, long den) {
    num %= den; if (num < 0) num += den;          // angle = 2 pi num/den in [0, 2 pi)
    long oct = (8 * num) / den;                   // octant 0..7
    long r8 = 8 * num - oct * den;                // remainder: angle = (oct + r8/den) * pi/4
    double t = (double(r8) / double(den)) * (kPi / 4.0);
    double u = kPi / 4.0 - t;
    switch (oct) {
        case 0: return cos_taylor(t);
        case 1: return sin_taylor(u);
        case 2: return -sin_taylor(t);
        case 3: return -cos_taylor(u);
        case 4: return -cos_taylor(t);
        case 5: return -sin_taylor(u);
        case 6: return sin_taylor(t);
        default: return cos_taylor(u);
    }
}
constexpr double sin2pi(long num, long den) { return cos2pi(4 * num - den, 4 * den); }
}  // namespace cx

// v * exp(DIR * 2*pi*i * NUM/DEN) with the trivial rotations folded at compile time
template <int NUM, int DEN, int DIR> B2R_DEV real2 cmul_root(real2 v) {
    constexpr int n = ((NUM % DEN) + DEN) % DEN;
    if constexpr (n == 0) return v;
    else if constexpr (4 * n == DEN) return rotq<DIR>(v);
    else if constexpr (2 * n == DEN) return make_real2(-v.x, -v.y);
    else if constexpr (4 * n == 3 * DEN) return rotq<-DIR>(v);
    else if constexpr ((8 * n) % DEN == 0) {
        constexpr real h = real(0.70710678118654752440);
        constexpr int o = (8 * n) / DEN;  // 1,3,5,7
        // exp(i*d*o*pi/4), d = DIR
        constexpr real cr = (o == 1 || o == 7) ? h : -h;
        constexpr real ci = ((o == 1 || o == 3) ? h : -h) * real(DIR);
        return make_real2(cr * v.x - ci * v.y, cr * v.y + ci * v.x);
    } else {
        constexpr real cr = real(cx::cos2pi(n, DEN));
        constexpr real ci = real(cx::sin2pi(n, DEN)) * real(DIR);
        return make_real2(rfma(-ci, v.y, cr * v.x), rfma(ci, v.x, cr * v.y));
    }
}

// ------------------------------------------------------------------ prime / radix-4 butterflies
// All operate in place on references and leave X[k] in the k-th argument (natural order).
template <int DIR> B2R_DEV void dft2(real2& a, real2& b) {
    real2 t = csub(a, b); a = cadd(a, b); b = t;
}
template <int DIR> B2R_DEV void dft3(real2& a, real2& b, real2& c) {
    constexpr real s3 = real(0.86602540378443864676);
    real2 t = cadd(b, c);
    real2 u = cscale(rotq<DIR>(csub(b, c)), s3);
    real2 m = make_real2(rfma(real(-0.5), t.x, a.x), rfma(real(-0.5), t.y, a.y));
    a = cadd(a, t); b = cadd(m, u); c = csub(m, u);
}
template <int DIR> B2R_DEV void dft4(real2& a, real2& b, real2& c, real2& d) {
    real2 s0 = cadd(a, c), d0 = csub(a, c), s1 = cadd(b, d), d1 = rotq<DIR>(csub(b, d));
    a = cadd(s0, s1); c = csub(s0, s1); b = cadd(d0, d1); d = csub(d0, d1);
}
template <int DIR> B2R_DEV void dft5(real2& a, real2& b, real2& c, real2& d, real2& e) {
    constexpr real c1 = real(cx::cos2pi(1, 5)), c2 = real(cx::cos2pi(2, 5));
    constexpr real s1 = real(cx::sin2pi(1, 5)), s2 = real(cx::sin2pi(2, 5));
    real2 t1 = cadd(b, e), t2 = cadd(c, d), t3 = csub(b, e), t4 = csub(c, d);
    real2 m1 = make_real2(rfma(c2, t2.x, rfma(c1, t1.x, a.x)), rfma(c2, t2.y, rfma(c1, t1.y, a.y)));
    real2 m2 = make_real2(rfma(c1, t2.x, rfma(c2, t1.x, a.x)), rfma(c1, t2.y, rfma(c2, t1.y, a.y)));
    real2 n1 = rotq<DIR>(make_real2(rfma(s2, t4.x, s1 * t3.x), rfma(s2, t4.y, s1 * t3.y)));
    real2 n2 = rotq<DIR>(make_real2(rfma(-s1, t4.x, s2 * t3.x), rfma(-s1, t4.y, s2 * t3.y)));
    a = cadd(a, cadd(t1, t2));
    b = cadd(m1, n1); e = csub(m1, n1); c = cadd(m2, n2); d = csub(m2, n2);
}
template <int DIR>
B2R_DEV void dft7(real2& a, real2& b, real2& c, real2& d, real2& e, real2& f, real2& g) {
    constexpr real c1 = real(cx::cos2pi(1, 7)), c2 = real(cx::cos2pi(2, 7)), c3 = real(cx::cos2pi(3, 7));
    constexpr real s1 = real(cx::sin2pi(1, 7)), s2 = real(cx::sin2pi(2, 7)), s3 = real(cx::sin2pi(3, 7));
    real2 t1 = cadd(b, g), t2 = cadd(c, f), t3 = cadd(d, e);
    real2 u1 = csub(b, g), u2 = csub(c, f), u3 = csub(d, e);
    real2 m1 = make_real2(rfma(c3, t3.x, rfma(c2, t2.x, rfma(c1, t1.x, a.x))),
                            rfma(c3, t3.y, rfma(c2, t2.y, rfma(c1, t1.y, a.y))));
    real2 m2 = make_real2(rfma(c1, t3.x, rfma(c3, t2.x, rfma(c2, t1.x, a.x))),
                            rfma(c1, t3.y, rfma(c3, t2.y, rfma(c2, t1.y, a.y))));
    real2 m3 = make_real2(rfma(c2, t3.x, rfma(c1, t2.x, rfma(c3, t1.x, a.x))),
                            rfma(c2, t3.y, rfma(c1, t2.y, rfma(c3, t1.y, a.y))));
    real2 n1 = rotq<DIR>(make_real2(rfma(s3, u3.x, rfma(s2, u2.x, s1 * u1.x)),
                                      rfma(s3, u3.y, rfma(s2, u2.y, s1 * u1.y))));
    real2 n2 = rotq<DIR>(make_real2(rfma(-s1, u3.x, rfma(-s3, u2.x, s2 * u1.x)),
                                      rfma(-s1, u3.y, rfma(-s3, u2.y, s2 * u1.y))));
    real2 n3 = rotq<DIR>(make_real2(rfma(s2, u3.x, rfma(-s1, u2.x, s3 * u1.x)),
                                      rfma(s2, u3.y, rfma(-s1, u2.y, s3 * u1.y))));
    a = cadd(a, cadd(t1, cadd(t2, t3)));
    b = cadd(m1, n1); g = csub(m1, n1);
    c = cadd(m2, n2); f = csub(m2, n2);
    d = cadd(m3, n3); e = csub(m3, n3);
}

// strided primitive dispatch on a register array (all indices fold after unrolling)
template <int R, int STRIDE, int DIR> B2R_DEV void dft_prim(real2* v) {
    if constexpr (R == 2) dft2<DIR>(v[0], v[STRIDE]);
    else if constexpr (R == 3) dft3<DIR>(v[0], v[STRIDE], v[2 * STRIDE]);
    else if constexpr (R == 4) dft4<DIR>(v[0], v[STRIDE], v[2 * STRIDE], v[3 * STRIDE]);
    else if constexpr (R == 5) dft5<DIR>(v[0], v[STRIDE], v[2 * STRIDE], v[3 * STRIDE], v[4 * STRIDE]);
    else if constexpr (R == 7)
        dft7<DIR>(v[0], v[STRIDE], v[2 * STRIDE], v[3 * STRIDE], v[4 * STRIDE], v[5 * STRIDE], v[6 * STRIDE]);
    else static_assert(R == 2, "unsupported primitive radix");
}

// ------------------------------------------------------------------ radix traits
// A radix is either primitive (2,3,4,5,7) or composite R = R1*R2 with primitive R1 and R2 primitive or
// itself composite (18 = 2 x (3x3), 24 = 3 x (2x4): used by the ahead-of-time / JIT schedules of long
// transforms to save a whole shared-memory stage; the dynamic dispatcher stops at 16).
template <int R> struct RadixTraits { static constexpr int r1 = R, r2 = 1; };
template <> struct RadixTraits<6>  { static constexpr int r1 = 2, r2 = 3; };
template <> struct RadixTraits<8>  { static constexpr int r1 = 2, r2 = 4; };
template <> struct RadixTraits<9>  { static constexpr int r1 = 3, r2 = 3; };
template <> struct RadixTraits<10> { static constexpr int r1 = 2, r2 = 5; };
template <> struct RadixTraits<12> { static constexpr int r1 = 3, r2 = 4; };
template <> struct RadixTraits<14> { static constexpr int r1 = 2, r2 = 7; };
template <> struct RadixTraits<15> { static constexpr int r1 = 3, r2 = 5; };
template <> struct RadixTraits<16> { static constexpr int r1 = 4, r2 = 4; };
template <> struct RadixTraits<18> { static constexpr int r1 = 2, r2 = 9; };
template <> struct RadixTraits<20> { static constexpr int r1 = 4, r2 = 5; };
template <> struct RadixTraits<24> { static constexpr int r1 = 3, r2 = 8; };

// register slot that holds output bin k after dft<R>
template <int R> B2R_HD constexpr int dft_slot(int k) {
    constexpr int r1 = RadixTraits<R>::r1, r2 = RadixTraits<R>::r2;
    if constexpr (r2 == 1) return k;
    else return (k % r1) * r2 + dft_slot<r2>(k / r1);   // X[k1 + r1*k2] sits in sub-array k1 at the slot of k2
}

template <int I, int N, class F> B2R_DEV void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(static_cast<F&&>(f));
    }
}

// In-register R-point DFT.  Input x[n] in v[n]; output X[k] in v[dft_slot<R>(k)].
template <int R, int DIR> B2R_DEV void dft(real2 (&v)[R]) {
    constexpr int r1 = RadixTraits<R>::r1, r2 = RadixTraits<R>::r2;
    if constexpr (r2 == 1) {
        dft_prim<R, 1, DIR>(v);
    } else {
        // n = n1*r2 + n2 ; step 1: r1-point DFTs over n1 (stride r2)
        static_for<0, r2>([&](auto n2) { dft_prim<r1, r2, DIR>(v + decltype(n2)::value); });
        // step 2: inner twiddles w_R^{n2*k1}
        static_for<1, r1>([&](auto k1) {
            static_for<1, r2>([&](auto n2) {
                constexpr int K1 = decltype(k1)::value, N2 = decltype(n2)::value;
                v[K1 * r2 + N2] = cmul_root<K1 * N2, R, DIR>(v[K1 * r2 + N2]);
            });
        });
        // step 3: r2-point DFTs over n2 (contiguous) -> X[k1 + r1*k2] in v[k1*r2 + dft_slot<r2>(k2)]
        static_for<0, r1>([&](auto k1) {
            dft<r2, DIR>(*reinterpret_cast<real2(*)[r2]>(v + decltype(k1)::value * r2));
        });
    }
}

// v[i] *= w^i, i = 1..R-1, powers by a balanced multiplication tree (depth <= log2 R)
template <int R> B2R_DEV void apply_twiddle_powers(real2 (&v)[R], real2 w) {
    real2 pw[R];
    pw[1] = w;
    static_for<2, R>([&](auto i) {
        constexpr int I = decltype(i)::value;
        pw[I] = cmul(pw[I / 2], pw[I - I / 2]);
    });
    static_for<1, R>([&](auto i) {
        constexpr int I = decltype(i)::value;
        v[I] = cmul(v[I], pw[I]);
    });
}

// ------------------------------------------------------------------ shared-memory layout
// One real2 of padding after every 16: keeps the stride-R writes of the early Stockham stages
// and the contiguous reads of all stages free of bank conflicts for power-of-two radices.
B2R_HD int smem_pad(int i) { return i + (i >> 4); }
B2R_HD constexpr int smem_padded_len(int n) { return n + (n >> 4) + 1; }
// Layout policy of a kernel: PAD = false addresses the workspace linearly.  Schedules whose first radix is odd
// want that in the column kernels: with CS columns side by side a half-warp touches 16/CS runs of CS consecutive
// elements whose distance (in sequence elements) is 1 for every read and every later-stage write and R0 for the
// first-stage writes -- odd distances spread the runs over the sixteen 8-byte bank pairs by themselves, and the
// extra element after every 16 then only misaligns them (scripts/smem_conflicts.py: 1080 = 15.12.6 x 8 columns,
// 23.5 % of the wavefronts are conflict replays with the padding, 1.3 % without).
template <bool PAD> B2R_HD int smem_at(int i) { return PAD ? i + (i >> 4) : i; }

// exact j / d for j < 2^16 via one mul.hi (magic = ceil(2^32 / d)), d >= 2
struct FastDiv {
    unsigned d, magic;
};
B2R_HD unsigned fd_div(unsigned j, FastDiv f) {
#if defined(__CUDA_ARCH__)
    return __umulhi(j, f.magic);
#else
    return (unsigned)(((unsigned long long)j * f.magic) >> 32);
#endif
}

// Per-stage constants of the DYNAMIC (any-size fallback) path, host-filled in b2r_plan.cpp
struct StageDesc {
    int radix;     // R
    int nb;        // number of butterflies per sequence = N / R
    int stride;    // S = product of the radices of the earlier stages
    FastDiv divS;  // for p = j mod S
    int tw_off;    // offset of this stage's twiddle row in the plan table (S entries, forward sign)
    int per_thread;  // ceil(nb / T): butterflies each thread runs in this stage
};

constexpr int kMaxStages = 8;
constexpr int kMaxElemsPerThread = 16;  // R * per_thread never exceeds this

struct FftDesc {
    int n;        // transform length
    int nstages;
    int threads;  // T: threads cooperating on one sequence
    StageDesc st[kMaxStages];
};

// ---- stage views ---------------------------------------------------------------------------------
// The stage engine is written against a "stage view": compile-time radix R and butterflies per
// thread NB, plus nb()/stride()/tw_off()/split().  StaticStage folds everything to immediates
// (the fast path: schedules instantiated ahead of time for the sizes in b2r_static_sizes.h);
// DynStage reads a StageDesc (any 2^a 3^b 5^c 7^d size).
template <int N_, int R_, int S_, int TWOFF_, int T_> struct StaticStage {
    static constexpr int R = R_;
    static constexpr int kNb = N_ / R_;
    static constexpr int NB = (kNb + T_ - 1) / T_;
    B2R_DEV constexpr int nb() const { return kNb; }
    B2R_DEV constexpr int stride() const { return S_; }
    B2R_DEV constexpr int tw_off() const { return TWOFF_; }
    B2R_DEV void split(int j, int& q, int& p) const { q = j / S_; p = j - q * S_; }
    // Padded-address shortcuts (smem_pad(i) = i + i/16): when the element stride between the R
    // operands of a butterfly is a multiple of 16 (or small enough never to carry into the next
    // group of 16) the R addresses are pad(first) + compile-time offsets -> LDS/STS immediates.
    template <int CS> static constexpr bool read_const() { return (kNb * CS) % 16 == 0; }
    template <int CS> static constexpr bool write_const() {
        return (S_ * CS) % 16 == 0 ||
               (S_ == 1 && 16 % CS == 0 && ((R_ * CS) % 16 == 0 || 16 % (R_ * CS) == 0));
    }
};
template <int R_, int NB_> struct DynStage {
    static constexpr int R = R_;
    static constexpr int NB = NB_;
    const StageDesc* sd;
    B2R_DEV int nb() const { return sd->nb; }
    B2R_DEV int stride() const { return sd->stride; }
    B2R_DEV int tw_off() const { return sd->tw_off; }
    B2R_DEV void split(int j, int& q, int& p) const {
        q = (sd->stride > 1) ? (int)fd_div((unsigned)j, sd->divS) : j;
        p = j - q * sd->stride;
    }
    template <int CS> static constexpr bool read_const() { return false; }
    template <int CS> static constexpr bool write_const() { return false; }
};

// ---- stage engine --------------------------------------------------------------------------------
// Element e of lane-column `c` lives at sm[smem_pad(e * cs + c)]  (cs = 1, c = 0 for row kernels;
// cs = columns per CTA for the column kernel, so that the batch index is the fastest dimension).

// shared -> registers, outer twiddle w = exp(DIR*2*pi*i*p/(S*R)) (one table load), butterfly
template <int DIR, int CS, bool PAD = true, class St>
B2R_DEV void stage_load_compute(const St st, const real2* sm, const real2* __restrict__ tw, int T, int tid,
                                int c, real2 (&v)[St::NB][St::R]) {
    constexpr int R = St::R;
#pragma unroll
    for (int b = 0; b < St::NB; ++b) {
        int j = tid + b * T;
        if (j < st.nb()) {
            if constexpr (!PAD) {
                const real2* base = sm + (j * CS + c);
#pragma unroll
                for (int i = 0; i < R; ++i) v[b][i] = base[i * st.nb() * CS];
            } else if constexpr (St::template read_const<CS>()) {
                const real2* base = sm + smem_pad(j * CS + c);
                const int step = st.nb() * CS + ((st.nb() * CS) >> 4);
#pragma unroll
                for (int i = 0; i < R; ++i) v[b][i] = base[i * step];
            } else {
#pragma unroll
                for (int i = 0; i < R; ++i) v[b][i] = sm[smem_pad((j + i * st.nb()) * CS + c)];
            }
            if (st.stride() > 1) {
                int q, p;
                st.split(j, q, p);
                real2 w = B2R_LDG(tw + st.tw_off() + p);
                if constexpr (DIR > 0) w.y = -w.y;
                apply_twiddle_powers<R>(v[b], w);
            }
            dft<R, DIR>(v[b]);
        }
    }
}

// butterfly on values already in registers (first stage fed from global memory; S = 1: no twiddle)
template <int DIR, class St>
B2R_DEV void stage_compute_first(const St st, int T, int tid, real2 (&v)[St::NB][St::R]) {
#pragma unroll
    for (int b = 0; b < St::NB; ++b) {
        int j = tid + b * T;
        if (j < st.nb()) dft<St::R, DIR>(v[b]);
    }
}

// registers -> shared at the Stockham output index  p + (j div S)*S*R + k*S
template <int CS, bool PAD = true, class St>
B2R_DEV void stage_store(const St st, real2* sm, int T, int tid, int c, real2 (&v)[St::NB][St::R]) {
    constexpr int R = St::R;
#pragma unroll
    for (int b = 0; b < St::NB; ++b) {
        int j = tid + b * T;
        if (j < st.nb()) {
            int q, p;
            st.split(j, q, p);
            int base = q * st.stride() * R + p;
            if constexpr (!PAD) {
                real2* dst = sm + (base * CS + c);
                static_for<0, R>([&](auto k) {
                    constexpr int K = decltype(k)::value;
                    dst[K * st.stride() * CS] = v[b][dft_slot<R>(K)];
                });
            } else if constexpr (St::template write_const<CS>()) {
                real2* dst = sm + smem_pad(base * CS + c);
                static_for<0, R>([&](auto k) {
                    constexpr int K = decltype(k)::value;
                    const int off = K * st.stride() * CS;
                    dst[off + (off >> 4)] = v[b][dft_slot<R>(K)];
                });
            } else {
                static_for<0, R>([&](auto k) {
                    constexpr int K = decltype(k)::value;
                    sm[smem_pad((base + K * st.stride()) * CS + c)] = v[b][dft_slot<R>(K)];
                });
            }
        }
    }
}

// ---- plan providers ------------------------------------------------------------------------------
// A provider exposes n(), nstages(), threads() and for_stage / for_stages, which call
// f(stage_view, s) with the view's R / NB as compile-time members.

template <int N_, int T_, int... Rs> struct StaticFft {
    static constexpr bool kStatic = true;
    static constexpr int kN = N_, kT = T_, kStages = (int)sizeof...(Rs);
    static constexpr int kR[sizeof...(Rs)] = {Rs...};
    static constexpr int radix(int s) { return kR[s]; }
    static constexpr int stride_of(int s) { int m = 1; for (int i = 0; i < s; ++i) m *= kR[i]; return m; }
    static constexpr int twoff_of(int s) { int o = 0; for (int i = 1; i < s; ++i) o += stride_of(i); return o; }
    static constexpr int product() { return stride_of(kStages); }
    static constexpr int max_elems() {
        int m = 0;
        for (int s = 0; s < kStages; ++s) { int nb = N_ / kR[s]; int e = ((nb + T_ - 1) / T_) * kR[s]; if (e > m) m = e; }
        return m;
    }
    static_assert(product() == N_, "radices must multiply to N");
    B2R_DEV constexpr int n() const { return N_; }
    B2R_DEV constexpr int nstages() const { return kStages; }
    B2R_DEV constexpr int threads() const { return T_; }
    template <int S> using Stage = StaticStage<N_, kR[S], stride_of(S), twoff_of(S), T_>;
    template <int S, class F> B2R_DEV void for_stage(F&& f) const { f(Stage<S>{}, S); }
    // stages [BEGIN, kStages - TAIL)
    template <int BEGIN, int TAIL, class F> B2R_DEV void for_stages(F&& f) const {
        static_for<BEGIN, kStages - TAIL>([&](auto s) { f(Stage<decltype(s)::value>{}, decltype(s)::value); });
    }
    template <class F> B2R_DEV void for_first(F&& f) const { f(Stage<0>{}, 0); }
    template <class F> B2R_DEV void for_last(F&& f) const { f(Stage<kStages - 1>{}, kStages - 1); }
};

// Dispatch a callable on the (radix, per-thread bound) pair of a dynamic stage.
template <class F> B2R_DEV void dispatch_stage(const StageDesc* sd, int s, F&& f) {
#define B2R_CASE(R_, NB_) case R_: f(DynStage<R_, NB_>{sd}, s); break;
    switch (sd->radix) {
        B2R_CASE(16, 1) B2R_CASE(15, 1) B2R_CASE(14, 1) B2R_CASE(12, 1) B2R_CASE(10, 1) B2R_CASE(9, 1)
        B2R_CASE(8, 2) B2R_CASE(7, 2) B2R_CASE(6, 2) B2R_CASE(5, 3) B2R_CASE(4, 4) B2R_CASE(3, 5) B2R_CASE(2, 8)
        default: break;
    }
#undef B2R_CASE
}

struct DynFft {
    static constexpr bool kStatic = false;
    const FftDesc* fd;  // device-resident descriptor
    B2R_DEV int n() const { return fd->n; }
    B2R_DEV int nstages() const { return fd->nstages; }
    B2R_DEV int threads() const { return fd->threads; }
    template <int BEGIN, int TAIL, class F> B2R_DEV void for_stages(F&& f) const {
        const int end = fd->nstages - TAIL;
        for (int s = BEGIN; s < end; ++s) dispatch_stage(&fd->st[s], s, f);
    }
    template <class F> B2R_DEV void for_first(F&& f) const { dispatch_stage(&fd->st[0], 0, f); }
    template <class F> B2R_DEV void for_last(F&& f) const {
        dispatch_stage(&fd->st[fd->nstages - 1], fd->nstages - 1, f);
    }
};

// max butterflies per thread the dynamic dispatcher serves for a radix (host scheduler uses this)
inline int max_per_thread(int radix) {
    switch (radix) {
        case 8: case 7: case 6: return 2;
        case 5: return 3;
        case 4: return 4;
        case 3: return 5;
        case 2: return 8;
        default: return 1;
    }
}

}  // namespace b2r
