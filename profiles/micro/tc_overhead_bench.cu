// Would a tensor-core (tcgen05) radix-16 stage beat the register radix-16 pass of the 8192-point transform?
// A lower bound that needs no MMA: time only the CUDA-core work a tensor-core stage cannot avoid, with the MMA itself,
// the TMEM read-back (tcgen05.ld) and all of its synchronisation taken as FREE, and compare it with one complete
// register pass as k_render runs it.
//
//   MODE 0  register pass:  32 LDS.64 (exchange load) -> 2 x twiddled radix-16 (ss_core.cuh fft16_w, twiddles by
//           tw_get) -> 32 STS.64 (exchange store).  This is everything a pass costs today.
//   MODE 1  tensor-core pass, CUDA-core remainder only: per complex point the inter-pass twiddle product (the DFT-matrix
//           MMA has no per-column twiddles), the split of re / im into bf16 hi + bf16 lo operands (3 products
//           hi*hi + lo*hi + hi*lo keep ~16 mantissa bits; a third term would be needed for full fp32), and the store of
//           the packed operands into shared memory, 8 B per point.  Accumulators are assumed to appear in registers for free.
//
// If MODE 1 is not clearly cheaper than MODE 0, a tcgen05 stage cannot win: it still has to pay the MMA, TMEM traffic,
// the operand reads from shared memory by the tensor core (the same data pipe) and one commit / wait per pass.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../sonicsim_b200/csrc -o tc_overhead_bench tc_overhead_bench.cu
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ss_core.cuh"
using namespace ss;

__device__ float2 g_twB[kTabB];

template <int MODE>
__global__ void __launch_bounds__(256, 2) k_pass(float2* out, int iters) {
    extern __shared__ float2 s[];
    const int t = threadIdx.x;
    for (int i = t; i < kPadF; i += 256) s[i] = make_float2(1e-3f * (i & 255), -2e-3f * (i & 127));
    __syncthreads();
    float2 a[16], b[16], w[16];
    for (int it = 0; it < iters; ++it) {
        const float2* p = s + pad(t);
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 16; ++r) { a[r] = p[544 * r]; b[r] = p[544 * r + 272]; }
            __syncthreads();
            tw_get<true, 16>(g_twB + (t & 15), w);
            fft16_w<true>(a, w);
            fft16_w<true>(b, w);
            float2* d = s + (t >> 4) * 272 + (t & 15);
#pragma unroll
            for (int r = 0; r < 16; ++r) { d[17 * r] = a[out16(r)]; d[17 * r + 4352] = b[out16(r)]; }
            __syncthreads();
        } else {
            // accumulators "arrive" in registers (free); keep a data dependence on shared memory so nothing is hoisted
            const float2 seed = p[0];
#pragma unroll
            for (int r = 0; r < 16; ++r) { a[r] = make_float2(seed.x + r, seed.y - r); b[r] = make_float2(seed.y + r, seed.x - r); }
            __syncthreads();
            tw_get<true, 16>(g_twB + (t & 15), w);
            unsigned long long* d = reinterpret_cast<unsigned long long*>(s + (t >> 4) * 272 + (t & 15));
#pragma unroll
            for (int r = 0; r < 16; ++r) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 v = cmul(h ? b[r] : a[r], w[r]);                    // inter-pass twiddle
                    const __nv_bfloat162 hi = __floats2bfloat162_rn(v.x, v.y);       // operand split: hi
                    const float2 hf = __bfloat1622float2(hi);
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x - hf.x, v.y - hf.y);   // ... and lo
                    const unsigned long long pk = ((unsigned long long)*reinterpret_cast<const unsigned*>(&lo) << 32) |
                                                  *reinterpret_cast<const unsigned*>(&hi);
                    d[17 * r + (h ? 4352 : 0)] = pk;                                 // 8 B per point into the operand tile
                }
            }
            __syncthreads();
        }
    }
    out[blockIdx.x * 256 + t] = s[pad(t)];
}

template <int MODE>
void run(const char* name) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4000, ctas = sms * 2, smem = kPadF * (int)sizeof(float2);
    cudaFuncSetAttribute(k_pass<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float2* out; cudaMalloc(&out, sizeof(float2) * ctas * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_pass<MODE><<<ctas, 256, smem>>>(out, 50);
    cudaEventRecord(e0);
    k_pass<MODE><<<ctas, 256, smem>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-58s %8.3f ms  %7.1f ns per 8192-point pass and CTA (2 CTAs / SM)\n", name, ms, ms * 1e6 / iters);
    cudaFree(out);
}

int main() {
    float2 tb[kTabB];
    for (int r = 0; r < 16; ++r) for (int k = 0; k < 16; ++k) { double a = -2.0 * 3.14159265358979323846 * (k * r) / 256.0; tb[r * 16 + k] = make_float2((float)cos(a), (float)sin(a)); }
    cudaMemcpyToSymbol(g_twB, tb, sizeof(tb));
    run<0>("register radix-16 pass (load, 2 butterflies, store)");
    run<1>("tensor-core pass, CUDA-core remainder only (MMA, TMEM free)");
    return 0;
}
