// Throughput of packed fp32 (FFMA2 / FADD2 / FMUL2, PTX fma/add/mul.rn.f32x2) against the scalar forms on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_bench f32x2_bench.cu && ./f32x2_bench
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float2 v) { return *reinterpret_cast<u64*>(&v); }
__device__ __forceinline__ float2 up(u64 v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return up(r); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return up(r); }

template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, int iters, float2 m, float2 c) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }      // 2 FFMA
            if (MODE == 1) a[i] = fma2(a[i], m, c);                                                   // 1 FFMA2
            if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }                            // 2 FADD
            if (MODE == 3) a[i] = add2(a[i], c);                                                      // 1 FADD2
            if (MODE == 4) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = a[i].y + c.y; }                  // FFMA + FADD
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int iters = 20000, ctas = sms * warps_per_sm / 8;
    float2* out; cudaMalloc(&out, sizeof(float2) * ctas * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<ctas, 256>>>(out, 100, make_float2(0.999f, 1.001f), make_float2(1e-3f, 2e-3f));
    cudaEventRecord(e0);
    k<MODE><<<ctas, 256>>>(out, iters, make_float2(0.999f, 1.001f), make_float2(1e-3f, 2e-3f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = (double)ctas * 256 * iters * 16;          // scalar fp32 operations (FMA = 1)
    printf("%-22s %2d warps/SM  %7.3f ms  %6.1f fp32 lane-ops / clk / SM (at %d MHz nominal)\n", name, warps_per_sm, ms,
           lane_ops / (ms * 1e-3) / ((double)khz * 1e3) / sms, khz / 1000);
    cudaFree(out);
}

int main() {
    for (int w : {16, 32, 64}) {
        run<0>("2 x FFMA", w); run<1>("1 x FFMA2", w); run<2>("2 x FADD", w); run<3>("1 x FADD2", w); run<4>("FFMA + FADD", w);
    }
    return 0;
}
