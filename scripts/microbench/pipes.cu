// pipes.cu -- issue-rate microbenchmarks for the instruction classes the b2resample kernels are made of
// (scalar vs packed fp32, 2- vs 3-input min/max, MUFU, shared-memory loads), run once per round on B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pipes scripts/microbench/pipes.cu && build/pipes
// Prints warp-instructions per clock per SM for each stream (each CTA: 256 threads, 8 CTAs per SM resident).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2000
#define U 16   // independent chains

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, long long* clk, float seed) {
    float a[U]; unsigned long long p[U];
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = seed * i;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < U; ++i) { a[i] = seed + i + threadIdx.x; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); }
    const float b = seed * 1.0001f, c = seed * 0.5f;
    unsigned long long pb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    unsigned long long pc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
            if (MODE == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
            if (MODE == 4) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (MODE == 5) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            if (MODE == 6) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 7) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 8) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 9) {   // FFMA + FMNMX interleaved (two pipes)
                if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                else asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            }
            if (MODE == 10) {  // FFMA2 + FMNMX interleaved
                if (i & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
                else asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            }
            if (MODE == 11) {  // FFMA2 + FFMA interleaved
                if (i & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            }
            if (MODE == 12) {  // LDS.64
                float2 v = *reinterpret_cast<float2*>(&sm[(idx * 2 + i * 64) & 4094]);
                a[i] += v.x + v.y;
            }
            if (MODE == 13) {  // LDS.128
                float4 v = *reinterpret_cast<float4*>(&sm[(idx * 4 + i * 128) & 4092]);
                a[i] += v.x + v.w;
            }
            if (MODE == 14) {  // FFMA2 + 3 x FMNMX per 4
                if ((i & 3) == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
                else asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            }
            if (MODE == 15) {  // mul.f32x2
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
            }
            if (MODE == 16) {  // FFMA + integer IADD interleaved
                if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                else asm volatile("add.s32 %0, %0, %1;" : "+r"(*(int*)&a[i]) : "r"(idx));
            }
            if (MODE == 17) {  // half2 fma
                asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(*(unsigned*)&a[i]) : "r"(__float_as_uint(b)), "r"(__float_as_uint(c)));
            }
            if (MODE == 18) {  // half2 min
                asm volatile("min.f16x2 %0, %0, %1;" : "+r"(*(unsigned*)&a[i]) : "r"(__float_as_uint(b)));
            }
            if (MODE == 20) {  // 1 MUFU : 3 FFMA
                if ((i & 3) == 0) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            }
            if (MODE == 21) {  // 1 MUFU : 7 FFMA
                if ((i & 7) == 0) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            }
            if (MODE == 19) {  // shfl
                a[i] = __shfl_down_sync(0xffffffffu, a[i], 1);
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < U; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int per_iter_extra = 0) {
    int sms = 148, ctas = sms * 8;
    float* out; long long* clk;
    cudaMalloc(&out, ctas * 256 * sizeof(float)); cudaMalloc(&clk, ctas * sizeof(long long));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<ctas, 256>>>(out, clk, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<ctas, 256>>>(out, clk, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[ctas];
    cudaMemcpy(h, clk, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
    // warp-instructions per clock per SM: 8 CTAs x 8 warps x ITERS x U / cycles (cycles of one CTA ~ all concurrent)
    double wi = 8.0 * 8 * ITERS * U;
    printf("%-28s %8.3f ms  cta_cycles %10.0f  warp-instr/clk/SM %6.3f  (%s)\n", name, ms, avg, wi / avg, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(clk); delete[] h;
}

int main() {
    run<0>("FFMA"); run<1>("FFMA2 (f32x2)"); run<2>("FADD"); run<3>("FADD2"); run<15>("FMUL2");
    run<4>("FMNMX"); run<5>("FMNMX3"); run<6>("MUFU.RCP"); run<7>("MUFU.RSQ"); run<8>("MUFU.SQRT");
    run<9>("FFMA+FMNMX 1:1"); run<10>("FFMA2+FMNMX 1:1"); run<11>("FFMA2+FFMA 1:1"); run<14>("FFMA2+3xFMNMX");
    run<16>("FFMA+IADD 1:1"); run<20>("MUFU+3xFFMA"); run<21>("MUFU+7xFFMA"); run<17>("HFMA2"); run<18>("HMNMX2"); run<19>("SHFL");
    run<12>("LDS.64 (+2 FADD)"); run<13>("LDS.128 (+2 FADD)");
    return 0;
}
