// pipes2.cu -- second pass of the issue-rate microbenchmarks with rotating cross-chain dependencies so
// that ptxas cannot fold repeated operations.  Rates are computed from the event time of the whole grid:
// warp-instr/clk/SM = instructions / (ms * 1e-3 * clock * 148).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 1000
#define U 16
typedef unsigned long long u64;
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0,%1,%2,%3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0,%1,%2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmin2(float a, float b) { float r; asm volatile("min.f32 %0,%1,%2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmax2(float a, float b) { float r; asm volatile("max.f32 %0,%1,%2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm volatile("min.f32 %0,%1,%2,%3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float rsq(float a) { float r; asm volatile("rsqrt.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float rcp(float a) { float r; asm volatile("rcp.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ int iadd(int a, int b) { int r; asm volatile("add.s32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ int lop(int a, int b) { int r; asm volatile("xor.b32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// NF fp ops + NO other ops per inner step, per chain
template <int MODE> __global__ void __launch_bounds__(256, 4) k(float* out, float seed, int mask) {
    float a[U]; u64 p[U]; int n[U];
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = seed * i;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < U; ++i) { a[i] = seed + i + threadIdx.x; n[i] = threadIdx.x * i + 1; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); }
    const float b = seed * 1.0001f, c = seed * 0.5f;
    const u64 pb = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b), pc = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const int j = (i + 5) % U;
            if (MODE == 0) a[i] = ffma(a[i], b, a[j]);
            if (MODE == 1) p[i] = ffma2(p[i], pb, p[j]);
            if (MODE == 2) a[i] = fadd(a[i], a[j]);
            if (MODE == 3) p[i] = fadd2(p[i], p[j]);
            if (MODE == 4) a[i] = (i & 1) ? fmin2(a[i], a[j]) : fmax2(a[i], a[j]);
            if (MODE == 5) a[i] = fmin3(a[i], a[j], a[(i + 9) % U]);
            if (MODE == 6) { a[i] = ffma(a[i], b, a[j]); n[i] = iadd(n[i], n[j]); }                 // 1 FFMA + 1 IADD
            if (MODE == 7) { p[i] = ffma2(p[i], pb, p[j]); n[i] = iadd(n[i], n[j]); }               // 1 FFMA2 + 1 IADD
            if (MODE == 8) { p[i] = ffma2(p[i], pb, p[j]); n[i] = iadd(n[i], n[j]); n[i] = lop(n[i], n[(i + 3) % U]); }  // 1 FFMA2 + 2 INT
            if (MODE == 9) { a[i] = ffma(a[i], b, a[j]); n[i] = iadd(n[i], n[j]); n[i] = lop(n[i], n[(i + 3) % U]); }    // 1 FFMA + 2 INT
            if (MODE == 10) { a[i] = ffma(a[i], b, a[j]); a[i] = (i & 1) ? fmin2(a[i], a[(i + 7) % U]) : fmax2(a[i], a[(i + 7) % U]); }  // FFMA + FMNMX
            if (MODE == 11) { p[i] = ffma2(p[i], pb, p[j]); a[i] = (i & 1) ? fmin2(a[i], a[j]) : fmax2(a[i], a[j]); }  // FFMA2 + FMNMX
            if (MODE == 12) { p[i] = ffma2(p[i], pb, p[j]); a[i] = ffma(a[i], b, a[j]); }           // FFMA2 + FFMA
            if (MODE == 13) { a[i] = ffma(a[i], b, a[j]); if ((i & 3) == 0) a[i] = rsq(a[i]); }       // 4 FFMA + 1 MUFU
            if (MODE == 14) { a[i] = ffma(a[i], b, a[j]); if ((i & 7) == 0) a[i] = rsq(a[i]); }       // 8 FFMA + 1 MUFU
            if (MODE == 15) a[i] = rcp(fadd(a[i], a[j]));                                           // 1 FADD + 1 MUFU
            if (MODE == 16) { float2 v = *reinterpret_cast<float2*>(&sm[(n[i] & mask) * 2]); a[i] = fadd(a[i], v.x); n[i] = iadd(n[i], __float_as_int(v.y)); }  // LDS.64 + FADD + IADD
            if (MODE == 17) { float4 v = *reinterpret_cast<float4*>(&sm[(n[i] & (mask >> 1)) * 4]); a[i] = fadd(a[i], v.x); n[i] = iadd(n[i], __float_as_int(v.w)); }  // LDS.128 + ...
            if (MODE == 18) { a[i] = __shfl_down_sync(0xffffffffu, a[i], 1); a[i] = ffma(a[i], b, a[j]); a[j] = ffma(a[j], c, a[i]); a[i] = ffma(a[i], b, c); }  // 1 SHFL + 3 FFMA
            if (MODE == 19) { *reinterpret_cast<float2*>(&sm[((n[i] & mask) * 2)]) = make_float2(a[i], a[j]); n[i] = iadd(n[i], n[j]); }  // STS.64 + IADD
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < U; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + n[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + sm[threadIdx.x];
}

template <int MODE> void run(const char* name, double instr_per_step) {
    const int ctas = 148 * 4;
    float* out; cudaMalloc(&out, ctas * 256 * sizeof(float));
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<ctas, 256, 32768>>>(out, 1.0f, 4095);
    cudaEventRecord(e0);
    k<MODE><<<ctas, 256, 32768>>>(out, 1.0f, 4095);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double steps = (double)ctas * 8 * ITERS * U;
    const double per_clk_sm = steps / (ms * 1e-3 * 1.965e9 * 148);
    printf("%-34s %7.3f ms  steps/clk/SM %6.3f  instr/clk/SM %6.3f  (%s)\n", name, ms, per_clk_sm, per_clk_sm * instr_per_step, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    run<0>("FFMA", 1); run<1>("FFMA2", 1); run<2>("FADD", 1); run<3>("FADD2", 1); run<4>("FMNMX min/max alt", 1); run<5>("FMNMX3 (3 distinct)", 1);
    run<6>("FFMA + IADD", 2); run<7>("FFMA2 + IADD", 2); run<8>("FFMA2 + IADD + LOP", 3); run<9>("FFMA + IADD + LOP", 3);
    run<10>("FFMA + FMNMX", 2); run<11>("FFMA2 + FMNMX", 2); run<12>("FFMA2 + FFMA", 2);
    run<13>("4 FFMA + 1 MUFU", 1.25); run<14>("8 FFMA + 1 MUFU", 1.125); run<15>("FADD + MUFU.RCP", 2);
    run<16>("LDS.64 + FADD + IADD (+LOP)", 4); run<17>("LDS.128 + FADD + IADD (+LOP)", 4); run<18>("SHFL + 3 FFMA", 4); run<19>("STS.64 + IADD (+LOP)", 3);
    return 0;
}
