// sonicsim_b200 :: ss_mix.cu - mixture assembly of the training dataloader
// (separation/look2hear/datas/movingdatamodule.py:29-32,105-124): active-energy (RMS dB) of the
// reference speaker, SIR gains of the interferers (clamped to +40 dB), sum, SNR gain of the summed
// noise, sum.  Three streaming passes over stems that are already in HBM + one tiny gain kernel;
// per-block partial sums are combined in a fixed order, so results are run-to-run deterministic.
#include <math.h>
#include <string.h>

#include "ss_internal.h"

namespace {

constexpr int kMixBlocks = 64;        // partial sums per mixture
constexpr int kMaxStems = 8;

struct MixItem {
    const float* spk;      // (S, E)
    const float* noise;    // (M, E)
    const float* sirs;     // (S - 1)
    float* mix;            // (E)
    float* spk_out;        // (S, E) scaled speakers (may alias spk)
    double* scratch;       // kMixBlocks * (S + 2) partial sums, then S + 1 gains at the end
    long long E;
    int S, M;
    float snr;
    int delay;             // overlap_audio shift of the summed noise in elements (0 = none)
};

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0;
    if (threadIdx.x == 0) for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[i];
    return r;          // valid in thread 0
}


// summed noise stems at element e; with delay D > 0 the enhancement dataloader's overlap_audio
// (enhancement/look2hear/datas/movingdatamodule.py:34-48) on top: (s[e-D] + s[e+D]) + s[e], zeros outside [0, E)
__device__ __forceinline__ float noise_sum(const MixItem& it, long long e) {
    float n = 0.f;
    for (int m = 0; m < it.M; ++m) n += it.noise[(long long)m * it.E + e];
    return n;
}
__device__ __forceinline__ float noise_at(const MixItem& it, long long e) {
    const float n = noise_sum(it, e);
    if (it.delay <= 0) return n;
    const float f = e >= it.delay ? noise_sum(it, e - it.delay) : 0.f;
    const float b = e + it.delay < it.E ? noise_sum(it, e + it.delay) : 0.f;
    return __fadd_rn(__fadd_rn(f, b), n);
}

// 16-byte path of the three streaming passes: element count a multiple of 4, every row 16-byte aligned, no overlap_audio
// shift (whose +-D neighbours need not be aligned)
__device__ __forceinline__ bool mix_vec_ok(const MixItem& it) {
    return it.delay <= 0 && (it.E & 3) == 0 &&
           ((((uintptr_t)it.spk) | ((uintptr_t)it.noise) | ((uintptr_t)it.mix) | ((uintptr_t)it.spk_out)) & 15) == 0;
}
__device__ __forceinline__ float4 noise_sum4(const MixItem& it, long long q) {
    float4 n = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = 0; m < it.M; ++m) {
        const float4 v = ((const float4*)(it.noise + (long long)m * it.E))[q];
        n.x += v.x; n.y += v.y; n.z += v.z; n.w += v.w;
    }
    return n;
}

// pass 1: sum of squares of every speaker stem and of the summed noise
__global__ void __launch_bounds__(256) k_mix_energy(const MixItem* __restrict__ items) {
    __shared__ double sh[8];
    const MixItem& it = items[blockIdx.y];
    double acc[kMaxStems + 1];
#pragma unroll
    for (int i = 0; i <= kMaxStems; ++i) acc[i] = 0;
    if (mix_vec_ok(it)) {
        for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < (it.E >> 2); q += (long long)gridDim.x * blockDim.x) {
#pragma unroll
            for (int s = 0; s < kMaxStems; ++s) if (s < it.S) {
                const float4 v = ((const float4*)(it.spk + (long long)s * it.E))[q];
                acc[s] += (double)(v.x * v.x) + (double)(v.y * v.y) + (double)(v.z * v.z) + (double)(v.w * v.w);
            }
            const float4 n = noise_sum4(it, q);
            acc[kMaxStems] += (double)(n.x * n.x) + (double)(n.y * n.y) + (double)(n.z * n.z) + (double)(n.w * n.w);
        }
    } else
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < it.E; e += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int s = 0; s < kMaxStems; ++s) if (s < it.S) { float v = it.spk[(long long)s * it.E + e]; acc[s] += (double)(v * v); }
        const float n = noise_at(it, e);
        acc[kMaxStems] += (double)(n * n);
    }
#pragma unroll
    for (int s = 0; s <= kMaxStems; ++s) {
        if (s < it.S || s == kMaxStems) {
            double r = block_sum(acc[s], sh);
            if (threadIdx.x == 0) it.scratch[(long long)blockIdx.x * (kMaxStems + 2) + s] = r;
        }
    }
}

__device__ double rms_db(double sum_sq, long long count) {
    double ms = sum_sq / (double)count;               // torch.mean(x ** 2)
    if (ms < 1e-20) ms = 1e-20;                       // max(1e-20, .)   movingdatamodule.py:31
    return 10.0 * log10(ms);
}

// gains of the interferers: gain_i = min(E(spk0) - E(spk_i) - sir_i, 40) dB   (:108-113)
__global__ void k_mix_gain(const MixItem* __restrict__ items, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const MixItem& it = items[i];
    double tot[kMaxStems + 1];
    for (int s = 0; s <= kMaxStems; ++s) tot[s] = 0;
    for (int b = 0; b < kMixBlocks; ++b)
        for (int s = 0; s <= kMaxStems; ++s) if (s < it.S || s == kMaxStems) tot[s] += it.scratch[(long long)b * (kMaxStems + 2) + s];
    double* g = it.scratch + (long long)kMixBlocks * (kMaxStems + 2);
    const double target = rms_db(tot[0], it.E);
    g[0] = 1.0;
    for (int s = 1; s < it.S; ++s) {
        double gain = target - rms_db(tot[s], it.E) - (double)it.sirs[s - 1];
        if (gain > 40.0) gain = 40.0;
        g[s] = pow(10.0, gain / 20.0);
    }
    g[kMaxStems] = rms_db(tot[kMaxStems], it.E);      // noise energy (dB), consumed by k_mix_write
}

// pass 2: energy of all_speech = sum_i g_i spk_i
__global__ void __launch_bounds__(256) k_mix_speech_energy(const MixItem* __restrict__ items) {
    __shared__ double sh[8];
    const MixItem& it = items[blockIdx.y];
    const double* g = it.scratch + (long long)kMixBlocks * (kMaxStems + 2);
    float gf[kMaxStems];
#pragma unroll
    for (int s = 0; s < kMaxStems; ++s) gf[s] = s < it.S ? (float)g[s] : 0.f;
    double acc = 0;
    if (mix_vec_ok(it)) {
        for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < (it.E >> 2); q += (long long)gridDim.x * blockDim.x) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int s = 0; s < kMaxStems; ++s) if (s < it.S) {
                const float4 a = ((const float4*)(it.spk + (long long)s * it.E))[q];
                v.x += a.x * gf[s]; v.y += a.y * gf[s]; v.z += a.z * gf[s]; v.w += a.w * gf[s];
            }
            acc += (double)(v.x * v.x) + (double)(v.y * v.y) + (double)(v.z * v.z) + (double)(v.w * v.w);
        }
    } else
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < it.E; e += (long long)gridDim.x * blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int s = 0; s < kMaxStems; ++s) if (s < it.S) v += it.spk[(long long)s * it.E + e] * gf[s];
        acc += (double)(v * v);
    }
    double r = block_sum(acc, sh);
    if (threadIdx.x == 0) it.scratch[(long long)blockIdx.x * (kMaxStems + 2) + kMaxStems + 1] = r;
}

// pass 3: noise gain (:118-121), mixture and scaled speakers (:123-124)
__global__ void __launch_bounds__(256) k_mix_write(const MixItem* __restrict__ items) {
    const MixItem& it = items[blockIdx.y];
    const double* g = it.scratch + (long long)kMixBlocks * (kMaxStems + 2);
    double tot = 0;
    for (int b = 0; b < kMixBlocks; ++b) tot += it.scratch[(long long)b * (kMaxStems + 2) + kMaxStems + 1];
    double gain = rms_db(tot, it.E) - g[kMaxStems] - (double)it.snr;
    if (gain > 40.0) gain = 40.0;
    const float gn = (float)pow(10.0, gain / 20.0);
    float gf[kMaxStems];
#pragma unroll
    for (int s = 0; s < kMaxStems; ++s) gf[s] = s < it.S ? (float)g[s] : 0.f;
    if (mix_vec_ok(it)) {
        for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < (it.E >> 2); q += (long long)gridDim.x * blockDim.x) {
            float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int s = 0; s < kMaxStems; ++s) if (s < it.S) {
                float4 v = ((const float4*)(it.spk + (long long)s * it.E))[q];
                v.x = __fmul_rn(v.x, gf[s]); v.y = __fmul_rn(v.y, gf[s]); v.z = __fmul_rn(v.z, gf[s]); v.w = __fmul_rn(v.w, gf[s]);
                if (it.spk_out) ((float4*)(it.spk_out + (long long)s * it.E))[q] = v;
                sp.x += v.x; sp.y += v.y; sp.z += v.z; sp.w += v.w;
            }
            const float4 n = noise_sum4(it, q);
            float4 o;
            o.x = __fadd_rn(sp.x, __fmul_rn(n.x, gn)); o.y = __fadd_rn(sp.y, __fmul_rn(n.y, gn));
            o.z = __fadd_rn(sp.z, __fmul_rn(n.z, gn)); o.w = __fadd_rn(sp.w, __fmul_rn(n.w, gn));
            ((float4*)it.mix)[q] = o;
        }
        return;
    }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < it.E; e += (long long)gridDim.x * blockDim.x) {
        float sp = 0.f;
#pragma unroll
        for (int s = 0; s < kMaxStems; ++s) if (s < it.S) {
            float v = __fmul_rn(it.spk[(long long)s * it.E + e], gf[s]);        // products and sums rounded separately, as torch does
            if (it.spk_out) it.spk_out[(long long)s * it.E + e] = v;
            sp += v;
        }
        it.mix[e] = __fadd_rn(sp, __fmul_rn(noise_at(it, e), gn));
    }
}

}  // namespace

extern "C" int ss_mix_dev(ss_ctx* c, const ss_mix_item* items, int n_items, void* stream_) {
    if (!c || (!items && n_items > 0) || n_items < 0) return SS_ERR_INVALID;
    if (n_items == 0) return SS_OK;
    CK(cudaSetDevice(c->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    for (int i = 0; i < n_items; ++i) {
        const ss_mix_item& a = items[i];
        if (!a.speakers || !a.noises || !a.mix || !a.scratch || a.E <= 0 || a.S < 1 || a.M < 1) return SS_ERR_INVALID;
        if (a.S > kMaxStems || a.M > kMaxStems) return SS_ERR_UNSUPPORTED;
        if (a.S > 1 && !a.sirs) return SS_ERR_INVALID;
        if (a.noise_delay < 0) return SS_ERR_INVALID;
    }
    const size_t bytes = align_up(sizeof(MixItem) * n_items, 16);
    int slot; char *hblk, *dblk;
    { int st = ring_acquire(c, bytes, &slot, &hblk, &dblk); if (st) return st; }
    MixItem* h = (MixItem*)c->h_desc[slot];
    for (int i = 0; i < n_items; ++i) {
        const ss_mix_item& a = items[i];
        MixItem m; memset(&m, 0, sizeof(m));
        m.spk = a.speakers; m.noise = a.noises; m.sirs = a.sirs; m.mix = a.mix; m.spk_out = a.speakers_out;
        m.scratch = a.scratch; m.E = a.E; m.S = a.S; m.M = a.M; m.snr = a.snr; m.delay = a.noise_delay;
        h[i] = m;
    }
    CK(cudaMemcpyAsync(c->d_desc[slot], c->h_desc[slot], bytes, cudaMemcpyHostToDevice, stream));
    CK(cudaEventRecord(c->desc_ev[slot], stream));
    const MixItem* d = (const MixItem*)c->d_desc[slot];
    dim3 grid(kMixBlocks, n_items);
    k_mix_energy<<<grid, 256, 0, stream>>>(d);
    k_mix_gain<<<(n_items + 63) / 64, 64, 0, stream>>>(d, n_items);
    k_mix_speech_energy<<<grid, 256, 0, stream>>>(d);
    k_mix_write<<<grid, 256, 0, stream>>>(d);
    CK(cudaGetLastError());
    c->launches += 4;
    return SS_OK;
}

extern "C" int64_t ss_mix_scratch_doubles(void) { return (int64_t)kMixBlocks * (kMaxStems + 2) + kMaxStems + 1; }

extern "C" int ss_mix_host_ex(ss_ctx* c, const float* speakers, const float* noises, const float* sirs, float snr,
                              float* mix, float* speakers_out, int32_t S, int32_t M, int64_t E, int32_t noise_delay) {
    if (!c || !speakers || !noises || !mix || S < 1 || M < 1 || E <= 0 || noise_delay < 0) return SS_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    const size_t o_spk = 0, o_noise = align_up(4 * (size_t)S * E, 256), o_mix = o_noise + align_up(4 * (size_t)M * E, 256);
    const size_t o_sir = o_mix + align_up(4 * (size_t)E, 256), o_scr = o_sir + 256;
    const size_t total = o_scr + align_up(8 * (size_t)ss_mix_scratch_doubles(), 256);
    ss_ctx::Slot& sl = c->slot[0];
    cudaStream_t st = c->s_cmp;
    CK(cudaStreamSynchronize(st));
    if (total > sl.in_cap) {
        CK(cudaDeviceSynchronize());
        if (sl.d_in) CK(cudaFree(sl.d_in));
        sl.d_in = nullptr; sl.in_cap = 0;
        CK(cudaMalloc((void**)&sl.d_in, align_up(total, 1 << 20)));
        sl.in_cap = align_up(total, 1 << 20);
    }
    char* b = sl.d_in;
    CK(cudaMemcpyAsync(b + o_spk, speakers, 4 * (size_t)S * E, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b + o_noise, noises, 4 * (size_t)M * E, cudaMemcpyHostToDevice, st));
    if (S > 1) CK(cudaMemcpyAsync(b + o_sir, sirs, 4 * (size_t)(S - 1), cudaMemcpyHostToDevice, st));
    ss_mix_item it; memset(&it, 0, sizeof(it));
    it.speakers = (const float*)(b + o_spk); it.noises = (const float*)(b + o_noise); it.sirs = (const float*)(b + o_sir);
    it.mix = (float*)(b + o_mix); it.speakers_out = speakers_out ? (float*)(b + o_spk) : nullptr;
    it.scratch = (double*)(b + o_scr); it.E = E; it.S = S; it.M = M; it.snr = snr; it.noise_delay = noise_delay;
    int rc = ss_mix_dev(c, &it, 1, (void*)st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(mix, b + o_mix, 4 * (size_t)E, cudaMemcpyDeviceToHost, st));
    if (speakers_out) CK(cudaMemcpyAsync(speakers_out, b + o_spk, 4 * (size_t)S * E, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SS_OK;
}

extern "C" int ss_mix_host(ss_ctx* c, const float* speakers, const float* noises, const float* sirs, float snr,
                           float* mix, float* speakers_out, int32_t S, int32_t M, int64_t E) {
    return ss_mix_host_ex(c, speakers, noises, sirs, snr, mix, speakers_out, S, M, E, 0);
}

// ---- overlap_audio (enhancement/look2hear/datas/movingdatamodule.py:34-48) as a stand-alone op
namespace {
__global__ void __launch_bounds__(256) k_overlap(const float* __restrict__ x, float* __restrict__ y, long long T, long long D) {
    const float* xr = x + (long long)blockIdx.y * T;
    float* yr = y + (long long)blockIdx.y * T;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < T; n += (long long)gridDim.x * blockDim.x) {
        const float f = n >= D ? xr[n - D] : 0.f;
        const float b = n + D < T ? xr[n + D] : 0.f;
        yr[n] = __fadd_rn(__fadd_rn(f, b), xr[n]);
    }
}
}  // namespace

extern "C" int ss_overlap_dev(ss_ctx* c, const float* x, float* y, int32_t rows, int64_t T, int64_t delay, void* stream_) {
    if (!c || !x || !y || x == y || rows < 0 || T < 0 || delay < 0) return SS_ERR_INVALID;
    if (rows == 0 || T == 0) return SS_OK;
    if (rows > 65535) return SS_ERR_UNSUPPORTED;
    CK(cudaSetDevice(c->device));
    long long nb = (T + 255) / 256;
    if (nb > 1184) nb = 1184;                       // 8 waves of 148 SMs, grid-stride beyond that
    k_overlap<<<dim3((unsigned)nb, (unsigned)rows), 256, 0, (cudaStream_t)stream_>>>(x, y, T, delay);
    CK(cudaGetLastError());
    c->launches += 1;
    return SS_OK;
}

extern "C" int ss_overlap_host(ss_ctx* c, const float* x, float* y, int32_t rows, int64_t T, int64_t delay) {
    if (!c || !x || !y || rows < 0 || T < 0 || delay < 0) return SS_ERR_INVALID;
    if (rows == 0 || T == 0) return SS_OK;
    CK(cudaSetDevice(c->device));
    const size_t bytes = 4 * (size_t)rows * (size_t)T, o_y = align_up(bytes, 256), total = o_y + align_up(bytes, 256);
    ss_ctx::Slot& sl = c->slot[0];
    cudaStream_t st = c->s_cmp;
    CK(cudaStreamSynchronize(st));
    if (total > sl.in_cap) {
        CK(cudaDeviceSynchronize());
        if (sl.d_in) CK(cudaFree(sl.d_in));
        sl.d_in = nullptr; sl.in_cap = 0;
        CK(cudaMalloc((void**)&sl.d_in, align_up(total, 1 << 20)));
        sl.in_cap = align_up(total, 1 << 20);
    }
    char* b = sl.d_in;
    CK(cudaMemcpyAsync(b, x, bytes, cudaMemcpyHostToDevice, st));
    int rc = ss_overlap_dev(c, (const float*)b, (float*)(b + o_y), rows, T, delay, (void*)st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(y, b + o_y, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SS_OK;
}
