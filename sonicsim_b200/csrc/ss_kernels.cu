// sonicsim_b200 :: ss_kernels.cu  -  sm_100a kernels + the C ABI of include/sonicsim_b200.h
//
// Path:  SonicSim_moving.convolve_moving_receiver (SonicSim_moving.py:63-96) and
// convolve_fixed_receiver (:47-61), restructured as a uniformly partitioned overlap-save
// convolution in the segment-local form
//      y[c,n] = sum_p hat_p(n) (x * h[p,c])[n],     hat_p = linear hat over segments p-1, p,
// so only the two positions a sample needs are ever convolved (2/P of the reference's work) and the
// (P, C, N) intermediate of :86 is never materialised.
//
// Launches per chunk of sources:
//   k_blocks  : block table (waypoint-aligned or 4096-grid) + dense numbering of the work items; skipped
//               when the trajectory bounds are on the host, which then builds the same tables itself
//   k_prepare : real rows (RIR partitions, dry windows) -> half spectra (two rows per complex FFT);
//               one warp per block fills that block's k_render work items
//   k_render  : persistent CTAs; per transform Z = sum_part X[b-part] (H[p] + i H[p+1]) from spectra the
//               bulk-copy engine staged in shared memory one transform ahead; one 8192-point inverse FFT
//               gives both positions' convolutions as Re / Im; the closing radix-2 is fused with the
//               per-sample lerp and the (C, N) store.  <LONG>: RIR partitions >= 1 (staged one after the other).
//   k_render_fast : the same for all-aligned chunks (one item = one transform), see its own header below
//   k_rir_absmax  : optional, SS_RIR_NORMALIZE: global abs-max of a source's RIR tensor for k_prepare's row loads
// Consecutive chunks rotate through three streams / scratch buffers so that they overlap; a bound batch (ss_plan_*)
// captures all of it into one CUDA graph.
// fp32 throughout (the reference is float32 end to end), no cuFFT.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <vector>

#include "ss_internal.h"
#include "ss_phases.cuh"

using namespace ss;

// Device-path trajectory check: the kernels cannot raise, so a source whose device-side trajectory breaks the contract
// of ss_render_dev (idx outside [0, P - 2]; bounds not ascending from 0 to N) sets a bit here; ss_device_errors reads
// and clears it.  bit 0: index out of range, bit 1: bounds not monotone / not ending at N.
__device__ unsigned g_dev_err;

// twiddle tables, filled from the host in double precision (ss_create)
__device__ float2 g_tw[kF];          // exp(-2 pi i m / 8192)
__device__ float2 g_twB[kTabB];      // [r][k] exp(-2 pi i k r / 256)
__device__ float2 g_twC[kTabC];      // [r][k] exp(-2 pi i k r / 4096)

// ----------------------------------------------------------------------------- k_blocks
// One small CTA per chunk: block table of every source (one warp per source), 1 / n_s table, then the
// exclusive scan that packs the render work items of all sources into one dense table.
__global__ void __launch_bounds__(256)
k_blocks(const Source* __restrict__ srcs, int n_src, int* __restrict__ total_items) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int si = warp; si < n_src; si += nwarps) {
        const Source& S = srcs[si];
        int nblk;
        if (S.aligned) {
            int carry = 0;
            for (int base = 0; base < S.P - 1; base += 32) {
                const int sg = base + lane;
                const int b0 = sg < S.P - 1 ? S.bounds[sg] : 0;
                const int n_s = sg < S.P - 1 ? S.bounds[sg + 1] - b0 : 0;
                const int cnt = seg_blocks(n_s);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                const int first = carry + incl - cnt;
                for (int q = 0; q < cnt; ++q) {
                    Block bk; bk.start = b0 + kB * q; bk.len = n_s - kB * q < kB ? n_s - kB * q : kB; bk.p_lo = sg; bk.p_hi = sg + 1;
                    S.blocks[first + q] = bk;
                }
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            nblk = carry;
        } else {
            nblk = S.nb;
            for (int bi = lane; bi < nblk; bi += 32) {
                Block bk; bk.start = bi * kB; bk.len = S.N - bi * kB < kB ? S.N - bi * kB : kB; bk.p_lo = 0; bk.p_hi = 0;
                S.blocks[bi] = bk;
            }
        }
        for (int bi = nblk + lane; bi < S.nblk_max; bi += 32) { Block z; z.start = 0; z.len = 0; z.p_lo = 0; z.p_hi = 0; S.blocks[bi] = z; }
        if (S.mode == MODE_MOVING_BOUNDS)
            for (int sg = lane; sg < S.P - 1; sg += 32) S.rstep[sg] = 1.0 / (double)(S.bounds[sg + 1] - S.bounds[sg]);
        if (lane == 0) S.counts[0] = nblk;
    }
    __syncthreads();
    if (warp == 0) {
        int carry = 0;
        for (int base = 0; base < n_src; base += 32) {
            const int si = base + lane;
            const int cnt = si < n_src ? srcs[si].counts[0] * items_per_block(srcs[si]) : 0;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (si < n_src) srcs[si].counts[1] = carry + incl - cnt;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) *total_items = carry;
    }
}

// ----------------------------------------------------------------------------- k_rir_absmax
// SS_RIR_NORMALIZE: partial maxima of |h| over a source's whole (P, C, L) tensor, kNormParts CTAs per source, each
// writing its own slot (no atomics, nothing to reset between runs); k_prepare's RIR rows take the maximum of the
// slots and divide their taps by it - generate_rir_combination's `ir_output /= ir_output.abs().max()`
// (SonicSim_audio.py:398) without a separate pass over the RIRs.
__global__ void __launch_bounds__(256)
k_rir_absmax(const Source* __restrict__ srcs) {
    const Source& S = srcs[blockIdx.x / kNormParts];
    if (!S.norm_part) return;
    const int part = blockIdx.x % kNormParts;
    const size_t total = (size_t)S.P * S.C * S.L;
    float m = 0.f;
    if ((((uintptr_t)S.rir) & 15) == 0) {               // 16-byte loads over the aligned bulk, scalar tail
        const float4* r4 = (const float4*)S.rir;
        const size_t n4 = total >> 2;
        for (size_t i = (size_t)part * blockDim.x + threadIdx.x; i < n4; i += (size_t)kNormParts * blockDim.x) {
            const float4 v = r4[i];
            m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        }
        if (part == 0 && threadIdx.x < (total & 3)) m = fmaxf(m, fabsf(S.rir[(n4 << 2) + threadIdx.x]));
    } else {
        for (size_t i = (size_t)part * blockDim.x + threadIdx.x; i < total; i += (size_t)kNormParts * blockDim.x)
            m = fmaxf(m, fabsf(S.rir[i]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float s_m[8];
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmaxf(m, s_m[w]);
        const_cast<float*>(S.norm_part)[part] = m;
    }
}

// ----------------------------------------------------------------------------- k_prepare
// CTA kinds, flattened per source through `prefix`:
//   [0, nh)            two RIR-partition rows  -> two half spectra (one complex FFT)
//   [nh, nh + nx)      two dry windows         -> two half spectra
//   [nh + nx, ...)     8 blocks each (one warp per block): position range of the block and its
//                      k_render work items (RItem)
// A CTA lives ~9 us, so the chain of dependent loads in front of its first row load matters: for chunks of up
// to kPrepInline sources the source table and its prefix sums travel as kernel parameters (constant bank)
// instead of global memory (binary search + descriptor = 4-5 L2 round trips per CTA).
constexpr int kPrepInline = 24;
struct PrepParams {
    int n_inline;                      // 0: use the global tables
    int prefix[kPrepInline + 1];
    Source srcs[kPrepInline];
};
static_assert(sizeof(PrepParams) <= 4000, "kernel parameter space");

#ifndef SS_PREP_MINB
#define SS_PREP_MINB 2        // resident CTAs per SM the register allocation of k_prepare aims at (3 fits shared memory)
#endif
template <bool INL>
__global__ void __launch_bounds__(kThreads, SS_PREP_MINB)
k_prepare(const Source* __restrict__ srcs_g, const int* __restrict__ prefix_g, int n_src, RItem* __restrict__ items,
          const __grid_constant__ PrepParams pp) {
    extern __shared__ float2 smem[];
    const Tables T{g_tw, g_twB, g_twC};
    const int t = threadIdx.x;
    int si = 0;
    if (INL) { for (int i = 1; i < kPrepInline; ++i) si += (i < n_src && pp.prefix[i] <= (int)blockIdx.x) ? 1 : 0; }
    else si = find_source(prefix_g, n_src, blockIdx.x);
    const Source& S = INL ? pp.srcs[si] : srcs_g[si];
    int local = blockIdx.x - (INL ? pp.prefix[si] : prefix_g[si]);
    Row ra, rb;
    const int nh = spectra_pairs_h(S), nx = spectra_pairs_x(S);
    if (local >= nh + nx) {
        const int blk = (local - nh - nx) * 8 + (t >> 5), lane = t & 31;
        if (blk >= S.counts[0]) return;
        Block bk = S.blocks[blk];
        if (S.mode == MODE_MOVING_BOUNDS && !S.aligned) {
            if (blk == 0 && lane == 0) {                   // one warp per source checks the device-side bounds table
                bool bad = S.bounds[0] != 0 || S.bounds[S.P - 1] != S.N;
                for (int q = 0; q + 1 < S.P && !bad; ++q) bad = S.bounds[q + 1] < S.bounds[q];
                if (bad) atomicOr(&g_dev_err, 2u);
            }
            bk.p_lo = seg_of(S.bounds, S.P - 1, bk.start);
            bk.p_hi = seg_of(S.bounds, S.P - 1, bk.start + bk.len - 1) + 1;
        } else if (S.mode == MODE_MOVING_INDEXED) {
            int pmin = 0x7fffffff, pmax = -1;
            for (int n = bk.start + lane; n < bk.start + bk.len; n += 32) { int v = S.idx[n]; pmin = v < pmin ? v : pmin; pmax = v > pmax ? v : pmax; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                int a = __shfl_xor_sync(0xffffffffu, pmin, o), bm = __shfl_xor_sync(0xffffffffu, pmax, o);
                pmin = a < pmin ? a : pmin; pmax = bm > pmax ? bm : pmax;
            }
            // the reference raises IndexError for idx + 1 >= P (checked on the host path); clamp here, leave a mark
            if (lane == 0 && (pmin < 0 || pmax + 1 > S.P - 1)) atomicOr(&g_dev_err, 1u);
            bk.p_lo = pmin < 0 ? 0 : pmin;
            bk.p_hi = pmax + 1 > S.P - 1 ? S.P - 1 : pmax + 1;
        }
        fill_items(S, items, blk, bk, lane, 32);
        return;
    }
    if (local < nh) { ra = make_row_h(S, 2 * local); rb = make_row_h(S, 2 * local + 1); }
    else { local -= nh; ra = make_row_x(S, 2 * local); rb = make_row_x(S, 2 * local + 1); }
    if (!ra.dst && !rb.dst) return;              // both block slots unused (aligned blocking over-allocates)

    Regs32 R;
    spectra_phase1(t, ra, rb, smem);
    __syncthreads();
    load2(t, smem, R);
    __syncthreads();
    passB2<false>(t, smem, R, T);
    __syncthreads();
    load2(t, smem, R);
    spectra_phase3_compute(t, R, T);
    __syncthreads();
    spectra_phase3_store(t, smem, R);
    __syncthreads();
    spectra_phase4(t, smem, ra, rb);
}

// profiling aid: keeps the stream busy for `ns` nanoseconds
__global__ void k_delay(unsigned ns) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { __nanosleep(2000); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < ns);
}

// ----------------------------------------------------------------------------- k_render
// mbarrier / bulk-copy (TMA engine, 1-D) helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// work item, global -> shared, asynchronously (consumed one transform later)
__device__ __forceinline__ void item_prefetch(RItem* dst, const RItem* src) {
#pragma unroll
    for (int i = 0; i < (int)sizeof(RItem) / 16; ++i)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32((const char*)dst + 16 * i)), "l"((const char*)src + 16 * i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void item_prefetch_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kSpecBytes = kSpec * (int)sizeof(float2);                       // 32 KB
constexpr int kRenderSmem = kPadF * (int)sizeof(float2) + kSpecBytes;         // FFT buffer + Hq staging

#ifndef SS_RENDER_MINB
#define SS_RENDER_MINB 2
#endif
// Persistent CTAs: CTA i renders items i, i + gridDim.x, ...  Thread 0 runs one transform ahead:
// while the CTA computes transform k it has the bulk-copy engine stage transform k+1's spectra —
// Hq into the staging buffer as soon as form_z(k) has consumed it, X and Hp into the FFT buffer itself
// once pass C of transform k has read it out — so form_z never waits on L2.
// LONG: RIR partitions >= 1 are accumulated (L > 4096).  FAST: every item of the chunk is a compact-trajectory
// source under aligned blocking (one transform per block, two positions each, one segment per block).
#ifndef SS_LONG_STAGED
#define SS_LONG_STAGED 1          // 0: partitions >= 1 read with plain loads (form_z_parts), for comparison
#endif
template <bool LONG, bool FAST>
__global__ void __launch_bounds__(kThreads, SS_RENDER_MINB)
k_render(const RItem* __restrict__ items, const int* __restrict__ n_items_ptr, int n_items_host) {
    extern __shared__ __align__(128) float2 smem[];
    __shared__ __align__(16) RItem s_item[2];
    __shared__ __align__(16) XDesc s_desc[2];
    __shared__ __align__(8) uint64_t s_bar[2];          // [0]: X + Hp landed, [1]: Hq landed
    float2* const fftbuf = smem;
    float2* const sX = smem;
    float2* const sHp = smem + kSpec;
    float2* const sHq = smem + kPadF;
    const Tables T{g_tw, g_twB, g_twC};
    const int t = threadIdx.x;
    // table length: known on the host when it built the tables, otherwise written by k_blocks
    const int n_items = n_items_host >= 0 ? n_items_host : *n_items_ptr;

    // thread-0 iterator state
    int it_cur = blockIdx.x;          // item of the transform most recently published
    int slot = 0;                     // s_item slot holding it
    int p_cur = 0;

    if (t == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        XDesc d; d.valid = 0; d.Hq = nullptr;
        if (it_cur < n_items) {
            s_item[0] = items[it_cur];
            if (it_cur + (int)gridDim.x < n_items) item_prefetch(&s_item[1], &items[it_cur + gridDim.x]);
            p_cur = s_item[0].p_lo;
            d = make_xdesc(s_item[0], p_cur);
        }
        s_desc[0] = d;
        fence_proxy_async();
        if (d.valid) {
            mbar_expect_tx(&s_bar[0], 2 * kSpecBytes);
            bulk_g2s(sX, d.X, kSpecBytes, &s_bar[0]);
            bulk_g2s(sHp, d.Hp, kSpecBytes, &s_bar[0]);
            if (d.Hq) { mbar_expect_tx(&s_bar[1], kSpecBytes); bulk_g2s(sHq, d.Hq, kSpecBytes, &s_bar[1]); }
            else mbar_arrive(&s_bar[1]);
        } else { mbar_arrive(&s_bar[0]); mbar_arrive(&s_bar[1]); }
    }
    __syncthreads();

    Regs32 R;
    unsigned ph = 0;                                  // parity of the staging barriers' current phase
    for (int k = 0;; ++k) {
        mbar_wait(&s_bar[0], ph);
        mbar_wait(&s_bar[1], ph);
        ph ^= 1;
        const XDesc& d = s_desc[k & 1];
        if (!d.valid) break;
        if (LONG && SS_LONG_STAGED) {
            // partitions 1 .. kparts-1 follow partition 0 through the same staging buffers
            const float2* const q = d.Hq ? sHq : nullptr;
            DcNy e;
            long_stage<true>(t, sX, sHp, q, R, e);
            const int kparts = d.kparts;
            for (int part = 1; part < kparts; ++part) {
                __syncthreads();                      // partition part-1 consumed by every thread
                if (t == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(&s_bar[0], 2 * kSpecBytes);
                    bulk_g2s(sX, d.X - (size_t)part * kSpec, kSpecBytes, &s_bar[0]);
                    bulk_g2s(sHp, d.Hp + (size_t)part * kSpec, kSpecBytes, &s_bar[0]);
                    if (d.Hq) { mbar_expect_tx(&s_bar[1], kSpecBytes); bulk_g2s(sHq, d.Hq + (size_t)part * kSpec, kSpecBytes, &s_bar[1]); }
                    else mbar_arrive(&s_bar[1]);
                }
                mbar_wait(&s_bar[0], ph);
                mbar_wait(&s_bar[1], ph);
                ph ^= 1;
                long_stage<false>(t, sX, sHp, q, R, e);
            }
            long_finish(t, R, e);
        } else {
            form_z<LONG, FAST>(t, sX, sHp, (FAST || d.Hq) ? sHq : nullptr, d, R);
        }
        __syncthreads();                              // staged spectra consumed, s_desc[k & 1] read by all
        if (t == 0) {
            // publish transform k+1 and start staging its Hq
            XDesc nx; nx.valid = 0; nx.Hq = nullptr;
            const RItem& cur = s_item[slot];
            if (p_cur + 2 <= cur.p_hi) { p_cur += 2; nx = make_xdesc(cur, p_cur); }
            else if (it_cur + (int)gridDim.x < n_items) {
                it_cur += gridDim.x; slot ^= 1;
                item_prefetch_wait();                 // issued one item ago
                p_cur = s_item[slot].p_lo;
                nx = make_xdesc(s_item[slot], p_cur);
                if (it_cur + (int)gridDim.x < n_items) item_prefetch(&s_item[slot ^ 1], &items[it_cur + gridDim.x]);
            }
            s_desc[(k + 1) & 1] = nx;
            fence_proxy_async();
            if (nx.valid && nx.Hq) { mbar_expect_tx(&s_bar[1], kSpecBytes); bulk_g2s(sHq, nx.Hq, kSpecBytes, &s_bar[1]); }
            else mbar_arrive(&s_bar[1]);
        }
        render_phase1(t, fftbuf, R);
        __syncthreads();
        load2(t, fftbuf, R);
        __syncthreads();
        passB2<true>(t, fftbuf, R, T);
        __syncthreads();
        load2(t, fftbuf, R);
        __syncthreads();                              // FFT buffer is free: stage X, Hp of transform k+1 into it
        if (t == 0) {
            const XDesc nx = s_desc[(k + 1) & 1];
            fence_proxy_async();
            if (nx.valid) {
                mbar_expect_tx(&s_bar[0], 2 * kSpecBytes);
                bulk_g2s(sX, nx.X, kSpecBytes, &s_bar[0]);
                bulk_g2s(sHp, nx.Hp, kSpecBytes, &s_bar[0]);
            } else mbar_arrive(&s_bar[0]);
        }
        render_phase3(t, R, T);
        render_epilogue<FAST>(t, d, R);
    }
}

// ----------------------------------------------------------------------------- k_render_fast
// Chunks whose items are one transform each: compact trajectories under aligned blocking (the transform packs the
// positions (s, s + 1) of the block's segment) and static sources (two channels per transform), L <= 4096: every source
// of BASELINE configs[1] / [2] and of a SonicSet scene.  Same phases as k_render; what differs is who
// feeds the bulk-copy engine and how many bytes pass through the SM's L1 / shared-memory data pipe, which together with
// the issue slots is what the kernel runs against (profiles/EXPERIMENTS.md):
//   * a CTA renders a CONTIGUOUS range of items: the channels of a block are neighbours and share the dry spectrum X,
//     which lives in its own 32 KB buffer and is copied only when the next item belongs to another block (one lane
//     of warp 7, right after the staged spectra have been consumed);
//   * the filter spectra Hp, Hq of the next transform land in the FFT buffer itself; their copy is issued by whichever
//     warp is the LAST to finish its pass-C loads (a shared counter tells), so nobody waits for the buffer to drain;
//   * no serial producer section: work items are final (k_prepare wrote pointers, sample range, segment start and
//     1 / length) and prefetched two transforms ahead with cp.async;
//   * inter-pass twiddles: 4 table rows + 11 products per pass (tw_get) instead of 15 loads.
// Measured and rejected (profiles/EXPERIMENTS.md, profiles/r2_exp_fast_knobs.patch): split arrive / wait mbarriers
// instead of two of the four CTA barriers, twiddle loads hoisted above the exchange barriers, twiddles from registers
// only, a fixed producer lane waiting for the buffer - none of them moves the kernel.
constexpr int kItemLane = 7 * 32;     // thread that prefetches work items and issues the X copies (not warp 0, whose
                                      // thread 0 already carries the DC / Nyquist words)
__device__ __forceinline__ unsigned atom_inc_acqrel(unsigned* p) {
    unsigned old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(p)) : "memory");
    return old;
}
// both butterflies of a thread, the 16 loads of butterfly t first so that its arithmetic can start while the
// loads of butterfly t + 256 are still in flight
__device__ __forceinline__ void load2_ab(int t, const float2* s, Regs32& R) {
    const float2* p = s + pad(t);
#pragma unroll
    for (int r = 0; r < 16; ++r) R.a[r] = p[544 * r];
#pragma unroll
    for (int r = 0; r < 16; ++r) R.b[r] = p[544 * r + 272];
}
__global__ void __launch_bounds__(kThreads, SS_RENDER_MINB)
k_render_fast(const RItem* __restrict__ items, const int* __restrict__ n_items_ptr, int n_items_host) {
    extern __shared__ __align__(128) float2 smem[];
    __shared__ __align__(16) RItem s_item[3];
    __shared__ __align__(8) uint64_t s_bar[2];          // [0] Hp + Hq landed, [1] X landed or kept (transaction counts)
    __shared__ unsigned s_drained;                      // warps that have finished their pass-C loads, cumulative
    float2* const fftbuf = smem;
    float2* const sHp = smem;                           // the filter spectra land in the FFT buffer itself
    float2* const sHq = smem + kSpec;
    float2* const sX = smem + kPadF;                    // the dry spectrum has its own buffer: it survives the transform
    const Tables T{g_tw, g_twB, g_twC};
    const int t = threadIdx.x, lane = t & 31;
    // table length: known on the host when it built the tables, otherwise written by k_blocks
    const int n_items = n_items_host >= 0 ? n_items_host : *n_items_ptr;
    const int per = n_items / (int)gridDim.x, rem = n_items % (int)gridDim.x;
    const int n_k = per + ((int)blockIdx.x < rem ? 1 : 0);              // transforms of this CTA: a contiguous range
    if (n_k <= 0) return;
    const RItem* const my_items = items + (size_t)blockIdx.x * per + ((int)blockIdx.x < rem ? (int)blockIdx.x : rem);

    if (t == kItemLane) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        s_drained = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_item[0] = my_items[0];
        if (n_k > 1) s_item[1] = my_items[1];
        fence_proxy_async();
        mbar_expect_tx(&s_bar[0], 2 * kSpecBytes);
        bulk_g2s(sHp, item_hp(s_item[0]), kSpecBytes, &s_bar[0]);
        bulk_g2s(sHq, item_hq(s_item[0]), kSpecBytes, &s_bar[0]);
        mbar_expect_tx(&s_bar[1], kSpecBytes);
        bulk_g2s(sX, s_item[0].X, kSpecBytes, &s_bar[1]);
    }
    __syncthreads();

    Regs32 R;
    float2 w[16];
    XDesc unused; unused.kparts = 1; unused.Hq = nullptr; unused.X = nullptr; unused.Hp = nullptr;
    unsigned ph = 0;
    for (int k = 0; k < n_k; ++k) {
        mbar_wait(&s_bar[0], ph);
        mbar_wait(&s_bar[1], ph);
        form_z<false, true>(t, sX, sHp, sHq, unused, R);
        fft16<true>(R.a);
        fft16<true>(R.b);
        __syncthreads();                                  // B1: staged spectra consumed, pass A may overwrite Hp / Hq
        if (t == kItemLane) {
            if (k + 1 < n_k) {                            // item k + 1 became visible at barrier B4 of transform k - 1
                const RItem& nx = s_item[(k + 1) % 3];
                if (nx.X != s_item[k % 3].X) {            // next item belongs to another block: its dry spectrum
                    fence_proxy_async();
                    mbar_expect_tx(&s_bar[1], kSpecBytes);
                    bulk_g2s(sX, nx.X, kSpecBytes, &s_bar[1]);
                } else mbar_arrive(&s_bar[1]);            // same block, other channel: X stays where it is
            }
            // item k + 2 into the slot of item k - 1 (last read in the output stage of k - 1, which every warp has left)
            if (k + 2 < n_k) item_prefetch(&s_item[(k + 2) % 3], my_items + (k + 2));
        }
        passA_store(fftbuf, passA_jA(t), R.a);
        passA_store(fftbuf, passA_jB(t), R.b);
        __syncthreads();                                  // B2: pass A -> pass B exchange
        load2_ab(t, fftbuf, R);
        tw_get<true, 16>(T.twB + (t & 15), w);
        fft16_w<true>(R.a, w);
        fft16_w<true>(R.b, w);
        __syncthreads();                                  // B3: every warp has loaded its pass-B inputs
        passB_store(fftbuf, t, R.a);
        passB_store(fftbuf, t + 256, R.b);
        if (t == kItemLane) item_prefetch_wait();         // item k + 2 has landed; visible to all behind B4
        __syncthreads();                                  // B4: pass B -> pass C exchange
        load2_ab(t, fftbuf, R);
        tw_get<true, 256>(T.twC + t, w);
        __syncwarp();
        if (lane == 0 && atom_inc_acqrel(&s_drained) == (unsigned)(8 * k + 7) && k + 1 < n_k) {
            // last warp out of the FFT buffer: stage Hp, Hq of transform k + 1 into it
            const RItem& nx = s_item[(k + 1) % 3];
            fence_proxy_async();
            mbar_expect_tx(&s_bar[0], 2 * kSpecBytes);
            bulk_g2s(sHp, item_hp(nx), kSpecBytes, &s_bar[0]);
            bulk_g2s(sHq, item_hq(nx), kSpecBytes, &s_bar[0]);
        }
        ph ^= 1;
        fft16_w<true>(R.a, w);                            // pass C
        fft16_w<true>(R.b, w);
        render_phase3_close(t, R, T);
        render_epilogue_item(t, s_item[k % 3], R);
    }
}

// ============================================================================= host side
thread_local int g_last_cuda = 0;

extern "C" int ss_version(void) { return 100; }
extern "C" int ss_last_cuda_error(void) { return g_last_cuda; }
extern "C" const char* ss_strerror(int st) {
    switch (st) {
        case SS_OK: return "ok";
        case SS_ERR_INVALID: return "invalid argument";
        case SS_ERR_INDEX: return "index out of bounds: trajectory refers to position >= P - 1";
        case SS_ERR_CUDA: return "CUDA runtime error";
        case SS_ERR_NOMEM: return "out of device / pinned memory";
        case SS_ERR_UNSUPPORTED: return "unsupported shape";
        default: return "unknown status";
    }
}


extern "C" void ss_destroy(ss_ctx* c);

static int init_ctx(ss_ctx* c, int device) {
    CK(cudaSetDevice(device));
    c->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    std::vector<float2> tw(kF);
    for (int m = 0; m < kF; ++m) {
        double a = -2.0 * M_PI * (double)m / (double)kF;
        tw[m] = make_float2((float)cos(a), (float)sin(a));
    }
    CK(cudaMemcpyToSymbol(g_tw, tw.data(), sizeof(float2) * kF));
    std::vector<float2> tb(kTabB), tc(kTabC);
    for (int r = 0; r < 16; ++r) {
        for (int k = 0; k < 16; ++k) {
            double a = -2.0 * M_PI * (double)(k * r) / 256.0;
            tb[r * 16 + k] = make_float2((float)cos(a), (float)sin(a));
        }
        for (int k = 0; k < 256; ++k) {
            double a = -2.0 * M_PI * (double)(k * r) / 4096.0;
            tc[r * 256 + k] = make_float2((float)cos(a), (float)sin(a));
        }
    }
    CK(cudaMemcpyToSymbol(g_twB, tb.data(), sizeof(float2) * kTabB));
    CK(cudaMemcpyToSymbol(g_twC, tc.data(), sizeof(float2) * kTabC));
    { const unsigned zero = 0; CK(cudaMemcpyToSymbol(g_dev_err, &zero, sizeof(zero))); }
    CK(cudaFuncSetAttribute(k_prepare<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPadF * (int)sizeof(float2)));
    CK(cudaFuncSetAttribute(k_prepare<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPadF * (int)sizeof(float2)));
    CK(cudaFuncSetAttribute(k_render<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRenderSmem));
    CK(cudaFuncSetAttribute(k_render_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, kRenderSmem));
    CK(cudaFuncSetAttribute(k_render<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRenderSmem));
    for (int i = 0; i < ss_ctx::kRing; ++i) CK(cudaEventCreateWithFlags(&c->desc_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < ss_ctx::kAux; ++i) {
        CK(cudaStreamCreateWithFlags(&c->s_aux[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->s_cmp, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < ss_ctx::kSlots; ++i) {
        CK(cudaEventCreateWithFlags(&c->slot[i].ev_in, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->slot[i].ev_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->slot[i].ev_rend, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->slot[i].ev_free, cudaEventDisableTiming));
    }
    c->single_stream = getenv("SS_SINGLE_STREAM") != nullptr;
    c->no_fast = getenv("SS_NO_FAST") != nullptr;
    c->no_graph = getenv("SS_NO_GRAPH") != nullptr;
    if (getenv("SS_HOST_CHUNK_MB")) c->chunk_bytes_host = (int64_t)atoi(getenv("SS_HOST_CHUNK_MB")) << 20;
    return SS_OK;
}

extern "C" int ss_create(int device, ss_ctx** out) {
    if (!out) return SS_ERR_INVALID;
    *out = nullptr;
    ss_ctx* c = new (std::nothrow) ss_ctx();
    if (!c) return SS_ERR_NOMEM;
    const int st = init_ctx(c, device);
    if (st != SS_OK) { ss_destroy(c); return st; }        // releases whatever was created before the failure
    *out = c;
    return SS_OK;
}

extern "C" void ss_destroy(ss_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < ss_ctx::kAux; ++i) {
        if (c->d_scratch[i]) cudaFree(c->d_scratch[i]);
        if (c->s_aux[i]) cudaStreamDestroy(c->s_aux[i]);
        if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (int i = 0; i < ss_ctx::kRing; ++i) {
        if (c->h_desc[i]) cudaFreeHost(c->h_desc[i]);
        if (c->d_desc[i]) cudaFree(c->d_desc[i]);
        if (c->desc_ev[i]) cudaEventDestroy(c->desc_ev[i]);
    }
    for (int i = 0; i < ss_ctx::kSlots; ++i) {
        if (c->slot[i].d_in) cudaFree(c->slot[i].d_in);
        if (c->slot[i].d_out) cudaFree(c->slot[i].d_out);
        if (c->slot[i].ev_in) cudaEventDestroy(c->slot[i].ev_in);
        if (c->slot[i].ev_done) cudaEventDestroy(c->slot[i].ev_done);
        if (c->slot[i].ev_rend) cudaEventDestroy(c->slot[i].ev_rend);
        if (c->slot[i].ev_free) cudaEventDestroy(c->slot[i].ev_free);
    }
    if (c->h_post) cudaFreeHost(c->h_post);
    if (c->d_post) cudaFree(c->d_post);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_cmp) cudaStreamDestroy(c->s_cmp);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    delete c;
}

// Bits set by the kernels of this context's device since the last call (waits for the device to go idle).
extern "C" int ss_device_errors(ss_ctx* c, uint32_t* bits) {
    if (!c || !bits) return SS_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaDeviceSynchronize());
    unsigned v = 0, zero = 0;
    CK(cudaMemcpyFromSymbol(&v, g_dev_err, sizeof(v)));
    if (v) CK(cudaMemcpyToSymbol(g_dev_err, &zero, sizeof(zero)));
    *bits = v;
    return SS_OK;
}

extern "C" int ss_set_chunk_bytes(ss_ctx* c, int64_t bytes) {
    if (!c || bytes < (1 << 20)) return SS_ERR_INVALID;
    c->chunk_bytes = bytes;
    return SS_OK;
}
extern "C" int64_t ss_launch_count(const ss_ctx* c) { return c ? c->launches : 0; }
extern "C" void ss_reset_stats(ss_ctx* c) { if (c) c->launches = 0; }

extern "C" int ss_set_profiling(ss_ctx* c, int on) {
    if (!c) return SS_ERR_INVALID;
    c->profiling = on != 0;
    return SS_OK;
}
// Sum of per-launch device times since the last call (waits for the recorded work to finish).
extern "C" int ss_get_profile(ss_ctx* c, double* ms_spectra, double* ms_render, int64_t* n_pairs) {
    if (!c) return SS_ERR_INVALID;
    // Launch pairs with the same position inside their render call repeat the same work: each group contributes
    // (median over the calls) x (group size), so a host hiccup that leaves the GPU waiting inside one interval
    // does not leak into the sums.
    std::vector<std::vector<float>> ta, tb;
    for (auto& pf : c->prof) {
        CK(cudaEventSynchronize(pf.e2));
        float t1 = 0, t2 = 0;
        CK(cudaEventElapsedTime(&t1, pf.e0, pf.e1));
        CK(cudaEventElapsedTime(&t2, pf.e1, pf.e2));
        if ((size_t)pf.chunk >= ta.size()) { ta.resize(pf.chunk + 1); tb.resize(pf.chunk + 1); }
        ta[pf.chunk].push_back(t1); tb[pf.chunk].push_back(t2);
        cudaEventDestroy(pf.e0); cudaEventDestroy(pf.e1); cudaEventDestroy(pf.e2);
    }
    auto robust_sum = [](std::vector<std::vector<float>>& groups) {
        double s = 0;
        for (auto& g : groups) {
            if (g.empty()) continue;
            std::sort(g.begin(), g.end());
            const size_t n = g.size();
            const double med = (n & 1) ? g[n / 2] : 0.5 * ((double)g[n / 2 - 1] + g[n / 2]);
            s += med * (double)n;
        }
        return s;
    };
    const double a = robust_sum(ta), b = robust_sum(tb);
    if (ms_spectra) *ms_spectra = a;
    if (ms_render) *ms_render = b;
    if (n_pairs) *n_pairs = (int64_t)c->prof.size();
    c->prof.clear();
    return SS_OK;
}

extern "C" int ss_host_alloc(void** p, int64_t bytes) {
    if (!p || bytes <= 0) return SS_ERR_INVALID;
    CK(cudaHostAlloc(p, (size_t)bytes, cudaHostAllocDefault));
    return SS_OK;
}
extern "C" void ss_host_free(void* p) { if (p) cudaFreeHost(p); }

static int validate_item(const ss_source& it) {
    if (!it.x || !it.rir || !it.out) return SS_ERR_INVALID;
    if (it.N <= 0 || it.C <= 0 || it.L <= 0 || it.P <= 0) return SS_ERR_INVALID;
    // 32-bit sample / table indices inside the kernels: leave one FFT of headroom below 2^31
    if (it.N > 0x7fffffff - 2 * kF) return SS_ERR_UNSUPPORTED;
    if ((int64_t)it.P * it.C * ((it.L + kB - 1) / kB) > (int64_t)1 << 28) return SS_ERR_UNSUPPORTED;
    if (it.flags & ~SS_RIR_NORMALIZE) return SS_ERR_INVALID;
    if (it.mode == SS_STATIC) { if (it.P != 1) return SS_ERR_INVALID; }
    else if (it.mode == SS_MOVING_BOUNDS) { if (it.P < 2 || !it.bounds) return SS_ERR_INVALID; }
    else if (it.mode == SS_MOVING_INDEXED) { if (it.P < 2 || !it.idx || !it.w) return SS_ERR_INVALID; }
    else return SS_ERR_INVALID;
    return SS_OK;
}
// shapes derived from one ss_source
struct Shape { int K, nb, aligned, nblk_max, per, max_items; };
// Aligned blocking costs one transform per block, sum_s ceil(n_s / 4096); grid blocking costs
// ceil(positions touched / 2) per 4096-block.  With the bounds on the host both counts are exact and the
// cheaper plan wins (trajectories with many short segments are better off on the grid); without them the
// plan is aligned when the average segment is at least one block long.
static bool choose_aligned(const ss_source& it, int K, int nb) {
    if (it.mode != SS_MOVING_BOUNDS || K != 1) return false;
    const int32_t* hb = it.bounds_host;
    if (!hb) return (int64_t)(it.P - 1) * kB <= (int64_t)it.N;
    int64_t aligned = 0, grid = 0;
    for (int sg = 0; sg + 1 < it.P; ++sg) aligned += seg_blocks(hb[sg + 1] - hb[sg]);
    int lo = 0;                                          // first segment that reaches into the current block
    for (int b = 0; b < nb; ++b) {
        const int n0 = b * kB, n1 = (n0 + kB < it.N ? n0 + kB : it.N) - 1;
        while (lo + 2 < it.P && hb[lo + 1] <= n0) ++lo;
        int hi = lo;
        while (hi + 2 < it.P && hb[hi + 1] <= n1) ++hi;
        grid += (hi - lo + 2 + 1) / 2;                   // positions lo .. hi + 1, two per transform
    }
    return aligned <= grid;
}
static Shape shape_of(const ss_source& it) {
    Shape s;
    s.K = (it.L + kB - 1) / kB; s.nb = (it.N + kB - 1) / kB;
    s.aligned = choose_aligned(it, s.K, s.nb) ? 1 : 0;
    s.nblk_max = s.aligned ? s.nb + it.P - 1 : s.nb;
    s.per = it.mode == SS_STATIC ? (it.C + 1) / 2 : it.C;
    s.max_items = s.nblk_max * s.per;
    return s;
}
static size_t spectra_bytes(const ss_source& it) {
    const Shape sh = shape_of(it);
    return ((size_t)it.P * it.C * sh.K + sh.nblk_max) * kSpec * sizeof(float2) + (size_t)sh.max_items * sizeof(RItem) +
           align_up(sizeof(Block) * (size_t)sh.nblk_max, 256) + align_up(sizeof(double) * (size_t)it.P, 256) + 256 +
           ((it.flags & SS_RIR_NORMALIZE) ? 512 : 0);
}

// Host-side twin of k_blocks for one source whose trajectory bounds are visible on the host.
// Returns the number of blocks in use.
static int build_blocks_host(const ss_source& it, const Shape& sh, const int32_t* hb, Block* blocks, double* rstep) {
    int nblk = 0;
    if (sh.aligned) {
        for (int sg = 0; sg < it.P - 1; ++sg) {
            const int b0 = hb[sg], n_s = hb[sg + 1] - b0;
            for (int q = 0; q < seg_blocks(n_s); ++q) {
                Block bk; bk.start = b0 + kB * q; bk.len = n_s - kB * q < kB ? n_s - kB * q : kB; bk.p_lo = sg; bk.p_hi = sg + 1;
                blocks[nblk++] = bk;
            }
        }
    } else {
        nblk = sh.nb;
        for (int bi = 0; bi < nblk; ++bi) {
            Block bk; bk.start = bi * kB; bk.len = it.N - bi * kB < kB ? it.N - bi * kB : kB; bk.p_lo = 0; bk.p_hi = 0;
            blocks[bi] = bk;
        }
    }
    for (int bi = nblk; bi < sh.nblk_max; ++bi) { Block z; z.start = 0; z.len = 0; z.p_lo = 0; z.p_hi = 0; blocks[bi] = z; }
    if (it.mode == SS_MOVING_BOUNDS && hb)
        for (int sg = 0; sg < it.P - 1; ++sg) rstep[sg] = 1.0 / (double)(hb[sg + 1] - hb[sg]);
    return nblk;
}

// Test hook (pure host, no CUDA call): the blocking plan the library would use for `item` - the choice between
// waypoint-aligned and 4096-grid blocking and the block table of the host-built path.  blocks_out receives
// (start, len, p_lo, p_hi) per block in use; returns their number or a negative ss_status.
extern "C" int ss_debug_plan(const ss_source* item, int32_t* blocks_out, int32_t max_blocks, int32_t* aligned_out) {
    if (!item || !blocks_out || max_blocks < 0) return SS_ERR_INVALID;
    if (item->N <= 0 || item->C <= 0 || item->L <= 0 || item->P <= 0) return SS_ERR_INVALID;
    if (item->mode == SS_MOVING_BOUNDS && (!item->bounds_host || item->P < 2)) return SS_ERR_INVALID;
    const Shape sh = shape_of(*item);
    if (sh.nblk_max > max_blocks) return SS_ERR_NOMEM;
    std::vector<Block> blocks((size_t)sh.nblk_max);
    std::vector<double> rstep((size_t)item->P);
    const int nblk = build_blocks_host(*item, sh, item->bounds_host, blocks.data(), rstep.data());
    for (int i = 0; i < nblk; ++i) {
        blocks_out[4 * i] = blocks[i].start; blocks_out[4 * i + 1] = blocks[i].len;
        blocks_out[4 * i + 2] = blocks[i].p_lo; blocks_out[4 * i + 3] = blocks[i].p_hi;
    }
    if (aligned_out) *aligned_out = sh.aligned;
    return nblk;
}

// ---- one chunk of sources = one (k_blocks,) k_prepare, k_render launch group
// Everything the launches need besides the descriptor block in device memory.
struct ChunkLaunch {
    int n = 0, ps = 0, grid_r = 0, n_known = -1;
    bool host_tables = true, any_long = false, all_aligned = true, any_norm = false;
    const Source* ds = nullptr; const int* dps = nullptr; RItem* d_items = nullptr; int* d_total = nullptr;
    PrepParams pp;
};
static size_t chunk_scratch_bytes(const ss_source* items, int first, int last) {
    size_t need = 256;
    for (int i = first; i < last; ++i) need += spectra_bytes(items[i]);
    return need;
}
// Block tables on the host when every trajectory is host-visible (always true on the host path)
static bool chunk_host_tables(const ss_source* items, int first, int last) {
    for (int i = first; i < last; ++i)
        if (items[i].mode == SS_MOVING_BOUNDS && !items[i].bounds_host) return false;
    return true;
}
// descriptor block: Source[n] | prefix_prepare[n+1] | total | per source: blocks, rstep, counts
static size_t chunk_desc_bytes(const ss_source* items, int first, int last) {
    const int n = last - first;
    size_t bytes = align_up(sizeof(Source) * n, 16) + align_up(sizeof(int) * (n + 1), 16) + 16;
    if (chunk_host_tables(items, first, last))
        for (int i = first; i < last; ++i) {
            const Shape sh = shape_of(items[i]);
            bytes += align_up(sizeof(Block) * (size_t)sh.nblk_max, 16) + align_up(sizeof(double) * (size_t)items[i].P, 16) + 16;
        }
    return bytes;
}
// Fill the descriptor block of items[first, last) at `hbase` (host memory) for its device address `dbase`, carve the
// chunk's spectra / tables / work items out of `scratch` (device), and note the launch parameters in `L`.
static void chunk_describe(const ss_source* items, int first, int last, char* hbase, char* dbase, char* scratch,
                           int sm_count, ChunkLaunch* L) {
    const int n = last - first;
    const bool host_tables = chunk_host_tables(items, first, last);
    const size_t off_ps = align_up(sizeof(Source) * n, 16);
    const size_t off_tot = off_ps + align_up(sizeof(int) * (n + 1), 16);
    Source* hs = (Source*)hbase;
    int* hps = (int*)(hbase + off_ps);
    int ps = 0, pr = 0, total_items = 0;
    size_t tab_off = off_tot + 16;
    bool any_long = false, all_aligned = true, any_norm = false;
    for (int i = 0; i < n; ++i) {
        const ss_source& it = items[first + i];
        const Shape sh = shape_of(it);
        Source s;
        memset(&s, 0, sizeof(s));
        s.x = it.x; s.rir = it.rir; s.out = it.out;
        s.bounds = it.mode == SS_MOVING_BOUNDS ? it.bounds : nullptr;
        s.idx = it.mode == SS_MOVING_INDEXED ? it.idx : nullptr;
        s.w = it.mode == SS_MOVING_INDEXED ? it.w : nullptr;
        s.N = it.N; s.P = it.P; s.C = it.C; s.L = it.L;
        s.K = sh.K; s.nb = sh.nb; s.mode = it.mode; s.aligned = sh.aligned; s.nblk_max = sh.nblk_max;
        s.hspec = (float2*)scratch; scratch += (size_t)s.P * s.C * s.K * kSpec * sizeof(float2);
        s.xspec = (float2*)scratch; scratch += (size_t)s.nblk_max * kSpec * sizeof(float2);
        if (it.flags & SS_RIR_NORMALIZE) { s.norm_part = (const float*)scratch; scratch += 512; any_norm = true; }
        if (host_tables) {
            // tables live in the descriptor block itself: filled here, copied with it
            Block* hb_blocks = (Block*)(hbase + tab_off);
            s.blocks = (Block*)(dbase + tab_off); tab_off += align_up(sizeof(Block) * (size_t)s.nblk_max, 16);
            double* hb_rstep = (double*)(hbase + tab_off);
            s.rstep = (double*)(dbase + tab_off); tab_off += align_up(sizeof(double) * (size_t)s.P, 16);
            int* hb_counts = (int*)(hbase + tab_off);
            s.counts = (int*)(dbase + tab_off); tab_off += 16;
            const int nblk = build_blocks_host(it, sh, it.bounds_host, hb_blocks, hb_rstep);
            hb_counts[0] = nblk; hb_counts[1] = total_items; hb_counts[2] = 0; hb_counts[3] = 0;
            total_items += nblk * sh.per;
        } else {
            s.blocks = (Block*)scratch; scratch += align_up(sizeof(Block) * (size_t)s.nblk_max, 256);
            s.rstep = (double*)scratch; scratch += align_up(sizeof(double) * (size_t)s.P, 256);
            s.counts = (int*)scratch; scratch += 256;
        }
        hs[i] = s;
        hps[i] = ps;
        ps += spectra_pairs_h(s) + spectra_pairs_x(s) + range_ctas(s);
        pr += sh.max_items;
        any_long = any_long || s.K > 1;
        all_aligned = all_aligned && (s.aligned || (s.mode == MODE_STATIC && s.K == 1));      // one transform per item: k_render_fast
    }
    hps[n] = ps;
    // work-item table of the whole chunk (dense) + its length
    L->d_items = (RItem*)scratch; scratch += (size_t)pr * sizeof(RItem);
    L->d_total = host_tables ? (int*)(dbase + off_tot) : (int*)scratch;
    if (host_tables) { *(int*)(hbase + off_tot) = total_items; pr = total_items; }
    L->n = n; L->ps = ps; L->host_tables = host_tables; L->any_long = any_long; L->all_aligned = all_aligned;
    L->any_norm = any_norm;
    L->n_known = host_tables ? total_items : -1;
    L->grid_r = pr < sm_count * SS_RENDER_MINB ? pr : sm_count * SS_RENDER_MINB;
    L->ds = (const Source*)dbase;
    L->dps = (const int*)(dbase + off_ps);
    L->pp.n_inline = n <= kPrepInline ? n : 0;
    if (L->pp.n_inline) {
        memcpy(L->pp.prefix, hps, sizeof(int) * (size_t)(n + 1));
        memcpy(L->pp.srcs, hs, sizeof(Source) * (size_t)n);
    }
}
// The kernel launches of one described chunk (its descriptor block is, or will be by stream order, in device memory).
static int chunk_enqueue(ss_ctx* c, const ChunkLaunch& L, cudaStream_t stream, ss_ctx::Prof* pf) {
    if (pf) CK(cudaEventRecord(pf->e0, stream));
    if (!L.host_tables) {
        k_blocks<<<1, 256, 0, stream>>>(L.ds, L.n, L.d_total);
        CK(cudaGetLastError());
        c->launches += 1;
    }
    if (L.any_norm) {
        k_rir_absmax<<<L.n * kNormParts, 256, 0, stream>>>(L.ds);
        CK(cudaGetLastError());
        c->launches += 1;
    }
    if (L.ps > 0) {
        if (L.pp.n_inline) k_prepare<true><<<L.ps, kThreads, kPadF * (int)sizeof(float2), stream>>>(L.ds, L.dps, L.n, L.d_items, L.pp);
        else k_prepare<false><<<L.ps, kThreads, kPadF * (int)sizeof(float2), stream>>>(L.ds, L.dps, L.n, L.d_items, L.pp);
        CK(cudaGetLastError());
    }
    if (pf) CK(cudaEventRecord(pf->e1, stream));
    if (L.grid_r > 0) {
        if (L.any_long) k_render<true, false><<<L.grid_r, kThreads, kRenderSmem, stream>>>(L.d_items, L.d_total, L.n_known);
        else if (L.all_aligned && !c->no_fast) k_render_fast<<<L.grid_r, kThreads, kRenderSmem, stream>>>(L.d_items, L.d_total, L.n_known);
        else k_render<false, false><<<L.grid_r, kThreads, kRenderSmem, stream>>>(L.d_items, L.d_total, L.n_known);
        CK(cudaGetLastError());
    }
    if (pf) CK(cudaEventRecord(pf->e2, stream));
    c->launches += 2;
    return SS_OK;
}

// Describe items[first, last) (device pointers) in the next slot of the descriptor ring, copy the block and enqueue
// the launches on `stream`.
static int launch_chunk(ss_ctx* c, const ss_source* items, int first, int last, cudaStream_t stream, int buf = 0) {
    const size_t need = chunk_scratch_bytes(items, first, last);
    if (need > c->scratch_cap[buf]) {
        CK(cudaDeviceSynchronize());
        if (c->d_scratch[buf]) CK(cudaFree(c->d_scratch[buf]));
        c->d_scratch[buf] = nullptr; c->scratch_cap[buf] = 0;
        size_t cap = align_up(need + need / 4, 1 << 20);
        CK(cudaMalloc(&c->d_scratch[buf], cap));
        c->scratch_cap[buf] = cap;
    }
    const size_t bytes = chunk_desc_bytes(items, first, last);
    int slot; char *hblk, *dblk;
    { int st = ring_acquire(c, bytes, &slot, &hblk, &dblk); if (st) return st; }
    static thread_local ChunkLaunch L;                  // holds 3.2 KB of kernel parameters: not on the stack
    chunk_describe(items, first, last, hblk, dblk, c->d_scratch[buf], c->sm_count, &L);
    CK(cudaMemcpyAsync(dblk, hblk, bytes, cudaMemcpyHostToDevice, stream));
    CK(cudaEventRecord(c->desc_ev[slot], stream));
    ss_ctx::Prof pf;
    if (c->profiling) {
        CK(cudaEventCreate(&pf.e0)); CK(cudaEventCreate(&pf.e1)); CK(cudaEventCreate(&pf.e2));
        pf.chunk = c->prof_chunk++;
    }
    const int st = chunk_enqueue(c, L, stream, c->profiling ? &pf : nullptr);
    if (c->profiling) c->prof.push_back(pf);
    return st;
}

// split [0, n) into chunks whose spectra fit the L2-sized budget
static void make_chunks_bytes(const std::vector<size_t>& sb, int64_t budget, std::vector<int>& cuts) {
    const int n = (int)sb.size();
    // greedy pass: how many chunks does the budget need ...
    size_t total = 0, acc = 0;
    int n_chunks = 1;
    for (int i = 0; i < n; ++i) {
        total += sb[i];
        if (acc > 0 && acc + sb[i] > (size_t)budget) { ++n_chunks; acc = 0; }
        acc += sb[i];
    }
    // ... then cut at equal shares of the total, so that the last chunk is not a runt (7,7,7,7,4 -> 6,6,7,6,7):
    // equal chunks keep the streams of the overlap equally busy.  A chunk never exceeds the budget (measured from
    // the bytes accumulated at its own first item) unless a single item does.
    cuts.clear(); cuts.push_back(0);
    acc = 0;
    size_t chunk_start = 0;                            // bytes in front of the current chunk
    for (int i = 0; i < n; ++i) {
        const size_t target = total * cuts.size() / n_chunks;
        if (i > cuts.back() && (acc + sb[i] / 2 > target || acc - chunk_start + sb[i] > (size_t)budget)) {
            cuts.push_back(i);
            chunk_start = acc;
        }
        acc += sb[i];
    }
    cuts.push_back(n);
}
static void make_chunks(const ss_ctx* c, const ss_source* items, int n, std::vector<int>& cuts, int64_t budget = 0) {
    if (budget <= 0) budget = c->chunk_bytes;
    std::vector<size_t> sb(n);
    for (int i = 0; i < n; ++i) sb[i] = spectra_bytes(items[i]);
    make_chunks_bytes(sb, budget, cuts);
}
// Device path: sources are independent, so a batch is regrouped by the kernel variant its chunk would run - one
// transform per item (aligned moving sources and static ones: k_render_fast), long RIRs (k_render<LONG>), everything
// else (k_render<0, 0>) - and each group
// is chunked on its own: one static source no longer sends the moving sources of its chunk through the generic kernel
// (a SonicSet scene is 3 moving + 2 static sources).  `arranged` receives the regrouped items, `cuts` the chunk bounds.
static void arrange_chunks(const ss_ctx* c, const ss_source* items, int n, std::vector<ss_source>& arranged, std::vector<int>& cuts) {
    arranged.clear(); arranged.reserve(n);
    cuts.clear(); cuts.push_back(0);
    for (int cls = 0; cls < 3; ++cls) {
        const int lo = (int)arranged.size();
        for (int i = 0; i < n; ++i) {
            const Shape sh = shape_of(items[i]);
            const int k = sh.K > 1 ? 1 : ((sh.aligned || items[i].mode == SS_STATIC) ? 0 : 2);
            if (k == cls) arranged.push_back(items[i]);
        }
        const int cnt = (int)arranged.size() - lo;
        if (cnt == 0) continue;
        std::vector<int> local;
        make_chunks(c, arranged.data() + lo, cnt, local);
        for (size_t q = 1; q < local.size(); ++q) cuts.push_back(lo + local[q]);
    }
}
// Test hook (pure host): the chunking of items whose scratch needs are `bytes[i]` under `budget`; cuts_out receives
// the first item of every chunk plus n (at most max_cuts values); returns the number of values written or a negative status.
extern "C" int ss_debug_chunks(const int64_t* bytes, int32_t n, int64_t budget, int32_t* cuts_out, int32_t max_cuts) {
    if (!bytes || !cuts_out || n <= 0 || budget <= 0) return SS_ERR_INVALID;
    std::vector<size_t> sb(n);
    for (int i = 0; i < n; ++i) { if (bytes[i] <= 0) return SS_ERR_INVALID; sb[i] = (size_t)bytes[i]; }
    std::vector<int> cuts;
    make_chunks_bytes(sb, budget, cuts);
    if ((int)cuts.size() > max_cuts) return SS_ERR_NOMEM;
    for (size_t i = 0; i < cuts.size(); ++i) cuts_out[i] = cuts[i];
    return (int)cuts.size();
}

static int validate_dev_items(const ss_source* items, int n_items) {
    for (int i = 0; i < n_items; ++i) {
        int st = validate_item(items[i]); if (st) return st;
        if (items[i].mode == SS_MOVING_BOUNDS && items[i].bounds_host) {
            const int32_t* b = items[i].bounds_host;
            if (b[0] != 0 || b[items[i].P - 1] != items[i].N) return SS_ERR_INVALID;
            for (int q = 0; q + 1 < items[i].P; ++q) if (b[q + 1] < b[q]) return SS_ERR_INVALID;
        }
    }
    return SS_OK;
}

extern "C" int ss_render_dev(ss_ctx* c, const ss_source* items, int n_items, void* stream) {
    if (!c || (!items && n_items > 0) || n_items < 0) return SS_ERR_INVALID;
    if (n_items == 0) return SS_OK;
    CK(cudaSetDevice(c->device));
    { int st = validate_dev_items(items, n_items); if (st) return st; }
    std::vector<int> cuts;
    std::vector<ss_source> arranged;
    arrange_chunks(c, items, n_items, arranged, cuts);
    items = arranged.data();
    const size_t n_chunks = cuts.size() - 1;
    if (c->profiling) {
        // let the host enqueue this call's launches while the GPU idles, so no timed interval contains a wait for
        // the host; chunk positions restart for the grouping in ss_get_profile
        c->prof_chunk = 0;
        k_delay<<<1, 1, 0, (cudaStream_t)stream>>>(300000u);
        CK(cudaGetLastError());
    }
    if (n_chunks < 2 || c->single_stream || c->profiling) {      // profiling: kernels serialised -> clean per-kernel times
        for (size_t k = 0; k < n_chunks; ++k) {
            int st = launch_chunk(c, items, cuts[k], cuts[k + 1], (cudaStream_t)stream, 0);
            if (st) return st;
        }
        CK(cudaEventRecord(c->ev_join[0], (cudaStream_t)stream));      // scratch buffer 0 busy until here
        return SS_OK;
    }
    // fork: chunks alternate between two internal streams / scratch buffers; join back into `stream`
    CK(cudaEventRecord(c->ev_fork, (cudaStream_t)stream));
    for (int i = 0; i < ss_ctx::kAux; ++i) CK(cudaStreamWaitEvent(c->s_aux[i], c->ev_fork, 0));
    c->ring_pos = 0;                                    // slot k % kRing <-> stream k % kAux
    int rc = SS_OK;
    for (size_t k = 0; k < n_chunks && rc == SS_OK; ++k)
        rc = launch_chunk(c, items, cuts[k], cuts[k + 1], c->s_aux[k % ss_ctx::kAux], (int)(k % ss_ctx::kAux));
    for (int i = 0; i < ss_ctx::kAux; ++i) {
        CK(cudaEventRecord(c->ev_join[i], c->s_aux[i]));
        CK(cudaStreamWaitEvent((cudaStream_t)stream, c->ev_join[i], 0));
    }
    return rc;
}

// ----------------------------------------------------------------------------- plans (device path)
// A batch of device-resident sources bound once: validation, chunking, block tables and descriptor blocks are done at
// creation and stay resident in device memory together with the plan's own scratch, so a run is kernel launches only -
// captured into a CUDA graph on the first run (fork onto the context's internal streams, join back), one
// cudaGraphLaunch per run afterwards instead of ~25 runtime calls.  The pointers in `items` (and the contents of
// bounds_host) must stay valid and unchanged for the plan's life; buffer *contents* may change between runs.
struct ss_plan {
    ss_ctx* c = nullptr;
    std::vector<ChunkLaunch> chunks;
    char* d_desc = nullptr;
    char* d_scratch[ss_ctx::kAux] = {};
    cudaGraphExec_t exec = nullptr;
    bool graph_failed = false;
    int launches_per_run = 0;
    // fork / join events of the plan's own (events recorded while capturing must not be waited on outside the graph,
    // which the context's events are)
    cudaEvent_t ev_fork = nullptr, ev_join[ss_ctx::kAux] = {};
    cudaStream_t s_cap = nullptr;      // capture origin: the caller's stream may be the legacy default stream, which cannot capture
};

extern "C" void ss_plan_destroy(ss_plan* p) {
    if (!p) return;
    cudaSetDevice(p->c->device);
    cudaDeviceSynchronize();
    if (p->exec) cudaGraphExecDestroy(p->exec);
    if (p->s_cap) cudaStreamDestroy(p->s_cap);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    for (int i = 0; i < ss_ctx::kAux; ++i) if (p->ev_join[i]) cudaEventDestroy(p->ev_join[i]);
    if (p->d_desc) cudaFree(p->d_desc);
    for (int i = 0; i < ss_ctx::kAux; ++i) if (p->d_scratch[i]) cudaFree(p->d_scratch[i]);
    delete p;
}

static int plan_build(ss_plan* p, const ss_source* items, int n_items) {
    ss_ctx* c = p->c;
    std::vector<int> cuts;
    std::vector<ss_source> arranged;
    arrange_chunks(c, items, n_items, arranged, cuts);
    items = arranged.data();
    const size_t n_chunks = cuts.size() - 1;
    std::vector<size_t> off(n_chunks + 1, 0);
    size_t scratch_need[ss_ctx::kAux] = {};
    for (size_t k = 0; k < n_chunks; ++k) {
        off[k + 1] = off[k] + align_up(chunk_desc_bytes(items, cuts[k], cuts[k + 1]), 256);
        const size_t sb = chunk_scratch_bytes(items, cuts[k], cuts[k + 1]);
        size_t& m = scratch_need[n_chunks < 2 ? 0 : k % ss_ctx::kAux];
        m = sb > m ? sb : m;
    }
    CK(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < ss_ctx::kAux; ++i) CK(cudaEventCreateWithFlags(&p->ev_join[i], cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&p->s_cap, cudaStreamNonBlocking));
    CK(cudaMalloc((void**)&p->d_desc, off[n_chunks] + 256));
    for (int i = 0; i < ss_ctx::kAux; ++i)
        if (scratch_need[i]) CK(cudaMalloc((void**)&p->d_scratch[i], align_up(scratch_need[i], 1 << 20)));
    std::vector<char> host(off[n_chunks] + 256);
    p->chunks.resize(n_chunks);
    p->launches_per_run = 0;
    for (size_t k = 0; k < n_chunks; ++k) {
        chunk_describe(items, cuts[k], cuts[k + 1], host.data() + off[k], p->d_desc + off[k],
                       p->d_scratch[n_chunks < 2 ? 0 : k % ss_ctx::kAux], c->sm_count, &p->chunks[k]);
        p->launches_per_run += (p->chunks[k].host_tables ? 2 : 3) + (p->chunks[k].any_norm ? 1 : 0);
    }
    CK(cudaMemcpy(p->d_desc, host.data(), off[n_chunks], cudaMemcpyHostToDevice));
    return SS_OK;
}

extern "C" int ss_plan_create(ss_ctx* c, const ss_source* items, int n_items, ss_plan** out) {
    if (!c || !out || !items || n_items <= 0) return SS_ERR_INVALID;
    *out = nullptr;
    CK(cudaSetDevice(c->device));
    { int st = validate_dev_items(items, n_items); if (st) return st; }
    ss_plan* p = new (std::nothrow) ss_plan();
    if (!p) return SS_ERR_NOMEM;
    p->c = c;
    const int st = plan_build(p, items, n_items);
    if (st != SS_OK) { ss_plan_destroy(p); return st; }
    *out = p;
    return SS_OK;
}

// the launches of every chunk, forked over the context's internal streams and joined back into `stream`
static int plan_enqueue(ss_plan* p, cudaStream_t stream) {
    ss_ctx* c = p->c;
    const size_t n_chunks = p->chunks.size();
    if (n_chunks < 2 || c->single_stream) {
        for (size_t k = 0; k < n_chunks; ++k) { int st = chunk_enqueue(c, p->chunks[k], stream, nullptr); if (st) return st; }
        return SS_OK;
    }
    CK(cudaEventRecord(p->ev_fork, stream));
    for (int i = 0; i < ss_ctx::kAux; ++i) CK(cudaStreamWaitEvent(c->s_aux[i], p->ev_fork, 0));
    int rc = SS_OK;
    for (size_t k = 0; k < n_chunks && rc == SS_OK; ++k) rc = chunk_enqueue(c, p->chunks[k], c->s_aux[k % ss_ctx::kAux], nullptr);
    for (int i = 0; i < ss_ctx::kAux; ++i) {
        CK(cudaEventRecord(p->ev_join[i], c->s_aux[i]));
        CK(cudaStreamWaitEvent(stream, p->ev_join[i], 0));
    }
    return rc;
}

extern "C" int ss_plan_run(ss_plan* p, void* stream_) {
    if (!p) return SS_ERR_INVALID;
    ss_ctx* c = p->c;
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaSetDevice(c->device));
    if (c->profiling) {
        // per-kernel timing: chunks one after the other on the caller's stream, events around every launch
        c->prof_chunk = 0;
        k_delay<<<1, 1, 0, stream>>>(300000u);
        CK(cudaGetLastError());
        for (auto& L : p->chunks) {
            ss_ctx::Prof pf;
            CK(cudaEventCreate(&pf.e0)); CK(cudaEventCreate(&pf.e1)); CK(cudaEventCreate(&pf.e2));
            pf.chunk = c->prof_chunk++;
            int st = chunk_enqueue(c, L, stream, &pf);
            c->prof.push_back(pf);
            if (st) return st;
        }
        return SS_OK;
    }
    if (!p->exec && !p->graph_failed && !c->no_graph) {
        const int64_t launches_before = c->launches;
        cudaGraph_t g = nullptr;
        bool ok = cudaStreamBeginCapture(p->s_cap, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            const int st = plan_enqueue(p, p->s_cap);
            const cudaError_t e = cudaStreamEndCapture(p->s_cap, &g);
            ok = st == SS_OK && e == cudaSuccess && g != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&p->exec, g, 0) == cudaSuccess;
        if (g) cudaGraphDestroy(g);
        c->launches = launches_before;                      // nothing ran while capturing
        if (!ok) { p->exec = nullptr; p->graph_failed = true; (void)cudaGetLastError(); }
    }
    if (p->exec) {
        CK(cudaGraphLaunch(p->exec, stream));
        c->launches += p->launches_per_run;
        return SS_OK;
    }
    return plan_enqueue(p, stream);
}
// 1: the plan runs as an instantiated CUDA graph, 0: not (yet), negative: error
extern "C" int ss_plan_is_graph(const ss_plan* p) { return p ? (p->exec ? 1 : 0) : SS_ERR_INVALID; }

// ----------------------------------------------------------------------------- host path
static int check_traj_host(const ss_source& it) {
    if (it.mode == SS_MOVING_BOUNDS) {
        const int32_t* b = it.bounds;
        if (b[0] != 0 || b[it.P - 1] != it.N) return SS_ERR_INVALID;
        for (int i = 0; i + 1 < it.P; ++i) if (b[i + 1] < b[i]) return SS_ERR_INVALID;
    } else if (it.mode == SS_MOVING_INDEXED) {
        for (int n = 0; n < it.N; ++n) { int v = it.idx[n]; if (v < 0 || v + 1 >= it.P) return SS_ERR_INDEX; }
    }
    return SS_OK;
}

static int ensure(char** p, size_t* cap, size_t need) {
    if (need <= *cap) return SS_OK;
    CK(cudaDeviceSynchronize());
    if (*p) CK(cudaFree(*p));
    *p = nullptr; *cap = 0;
    size_t c2 = align_up(need + need / 8, 1 << 20);
    CK(cudaMalloc((void**)p, c2));
    *cap = c2;
    return SS_OK;
}

extern "C" int ss_loudness_dev(ss_ctx* c, const ss_loud_item* items, int n_items, void* stream);

// Host path: H2D -> k_spectra -> k_render [-> loudness measure + in-place gain] -> D2H, pipelined
// over chunks on three streams with two device slots.
extern "C" int ss_render_host_ex(ss_ctx* c, const ss_source* items, int n_items, const ss_post_lufs* post) {
    if (!c || (!items && n_items > 0) || n_items < 0) return SS_ERR_INVALID;
    if (n_items == 0) return SS_OK;
    CK(cudaSetDevice(c->device));
    for (int i = 0; i < n_items; ++i) {
        int st = validate_item(items[i]); if (st) return st;
        st = check_traj_host(items[i]); if (st) return st;
        if (post && post[i].brk) {
            if (!post[i].blk_lo || !post[i].blk_hi || post[i].n_e <= 0 || post[i].rate <= 0 || items[i].C > 8) return SS_ERR_INVALID;
            if (post[i].rate != post[0].rate && post[0].brk) return SS_ERR_INVALID;
        }
    }
    // an earlier asynchronous ss_render_dev may still be using scratch buffer 0 on its internal streams
    for (int i = 0; i < ss_ctx::kAux; ++i) CK(cudaStreamWaitEvent(c->s_cmp, c->ev_join[i], 0));
    std::vector<int> cuts;
    make_chunks(c, items, n_items, cuts, c->chunk_bytes_host);
    std::vector<ss_source> dev(n_items);
    std::vector<ss_loud_item> loud;
    std::vector<double*> res_dev(n_items, nullptr);
    // loudness: (a) the gating tables of all sources that share them (same host arrays = same clip length, the normal
    // case) are uploaded once per call from a pinned staging copy; (b) the two result doubles of every stem come back
    // into pinned memory - a copy into the caller's pageable array would block the host until the stream gets
    // there and serialise the whole pipeline; (c) the loudness kernels of consecutive chunks run on rotating side
    // streams: their grids are small and latency-bound (~0.5 ms per chunk), three of them overlap each other, the
    // next chunk's render and the copies
    struct Tab { const int32_t *brk, *lo, *hi; int n_e, n_blocks; const int32_t *d_brk, *d_lo, *d_hi; };
    std::vector<Tab> tabs;
    std::vector<int> tab_of(n_items, -1);
    if (post) {
        size_t tab_bytes = 0;
        for (int i = 0; i < n_items; ++i) {
            if (!post[i].brk) continue;
            int t = -1;
            for (size_t q = 0; q < tabs.size(); ++q)
                if (tabs[q].brk == post[i].brk && tabs[q].lo == post[i].blk_lo && tabs[q].hi == post[i].blk_hi &&
                    tabs[q].n_e == post[i].n_e && tabs[q].n_blocks == post[i].n_blocks) { t = (int)q; break; }
            if (t < 0) {
                Tab nt = {post[i].brk, post[i].blk_lo, post[i].blk_hi, post[i].n_e, post[i].n_blocks, nullptr, nullptr, nullptr};
                tabs.push_back(nt); t = (int)tabs.size() - 1;
                tab_bytes += align_up(4 * (size_t)(nt.n_e + 1), 256) + 2 * align_up(4 * (size_t)(nt.n_blocks + 1), 256);
            }
            tab_of[i] = t;
        }
        const size_t res_bytes = align_up(2 * sizeof(double) * (size_t)n_items, 256);
        if (tab_bytes + res_bytes > c->h_post_cap) {
            CK(cudaDeviceSynchronize());
            if (c->h_post) CK(cudaFreeHost(c->h_post));
            if (c->d_post) CK(cudaFree(c->d_post));
            c->h_post = nullptr; c->d_post = nullptr; c->h_post_cap = 0;
            const size_t cap = align_up((tab_bytes + res_bytes) * 2, 4096);
            CK(cudaHostAlloc((void**)&c->h_post, cap, cudaHostAllocDefault));
            CK(cudaMalloc((void**)&c->d_post, cap));
            c->h_post_cap = cap;
        }
        size_t off = res_bytes;                              // [0, res_bytes): results (host side only)
        for (auto& t : tabs) {
            size_t nb = 4 * (size_t)(t.n_e + 1);
            memcpy(c->h_post + off, t.brk, nb); t.d_brk = (const int32_t*)(c->d_post + off); off += align_up(nb, 256);
            nb = 4 * (size_t)t.n_blocks;
            if (nb) memcpy(c->h_post + off, t.lo, nb);
            t.d_lo = (const int32_t*)(c->d_post + off); off += align_up(nb + 4, 256);
            if (nb) memcpy(c->h_post + off, t.hi, nb);
            t.d_hi = (const int32_t*)(c->d_post + off); off += align_up(nb + 4, 256);
        }
        if (off > res_bytes)
            CK(cudaMemcpyAsync(c->d_post + res_bytes, c->h_post + res_bytes, off - res_bytes, cudaMemcpyHostToDevice, c->s_in));
    }
    double* const h_res = post ? (double*)c->h_post : nullptr;
    // the pipeline proper; any failure falls through to the stream synchronisation below, so that no copy is
    // still reading or writing the caller's buffers when this function returns
    auto pipeline = [&]() -> int {
    int rc = SS_OK;
    for (size_t k = 0; k + 1 < cuts.size() && rc == SS_OK; ++k) {
        const int first = cuts[k], last = cuts[k + 1];
        ss_ctx::Slot& sl = c->slot[k % ss_ctx::kSlots];
        size_t in_b = 0, out_b = 0;
        for (int i = first; i < last; ++i) {
            const ss_source& it = items[i];
            in_b += align_up(sizeof(float) * (size_t)it.N, 256) + align_up(sizeof(float) * (size_t)it.P * it.C * it.L, 256);
            if (it.mode == SS_MOVING_BOUNDS) in_b += align_up(sizeof(int32_t) * (size_t)it.P, 256);
            if (it.mode == SS_MOVING_INDEXED) in_b += 2 * align_up(4 * (size_t)it.N, 256);
            out_b += align_up(sizeof(float) * (size_t)it.C * it.N, 256);
            if (post && post[i].brk)
                out_b += align_up(8 * (size_t)SS_LOUD_SCRATCH_DOUBLES * it.C * post[i].n_e, 256) + 256;
        }
        if (k >= (size_t)ss_ctx::kSlots) CK(cudaEventSynchronize(sl.ev_free));      // slot's previous chunk fully drained
        if ((rc = ensure(&sl.d_in, &sl.in_cap, in_b)) != SS_OK) break;
        if ((rc = ensure(&sl.d_out, &sl.out_cap, out_b)) != SS_OK) break;
        char* pi = sl.d_in; char* po = sl.d_out;
        loud.clear();
        for (int i = first; i < last; ++i) {
            const ss_source& it = items[i];
            ss_source d = it;
            size_t nb;
            nb = sizeof(float) * (size_t)it.N;
            CK(cudaMemcpyAsync(pi, it.x, nb, cudaMemcpyHostToDevice, c->s_in)); d.x = (const float*)pi; pi += align_up(nb, 256);
            nb = sizeof(float) * (size_t)it.P * it.C * it.L;
            CK(cudaMemcpyAsync(pi, it.rir, nb, cudaMemcpyHostToDevice, c->s_in)); d.rir = (const float*)pi; pi += align_up(nb, 256);
            if (it.mode == SS_MOVING_BOUNDS) {
                nb = sizeof(int32_t) * (size_t)it.P;
                CK(cudaMemcpyAsync(pi, it.bounds, nb, cudaMemcpyHostToDevice, c->s_in)); d.bounds = (const int32_t*)pi; pi += align_up(nb, 256);
                d.bounds_host = it.bounds;
            } else if (it.mode == SS_MOVING_INDEXED) {
                nb = 4 * (size_t)it.N;
                CK(cudaMemcpyAsync(pi, it.idx, nb, cudaMemcpyHostToDevice, c->s_in)); d.idx = (const int32_t*)pi; pi += align_up(nb, 256);
                CK(cudaMemcpyAsync(pi, it.w, nb, cudaMemcpyHostToDevice, c->s_in)); d.w = (const float*)pi; pi += align_up(nb, 256);
            }
            d.out = (float*)po; po += align_up(sizeof(float) * (size_t)it.C * it.N, 256);
            dev[i] = d;
            if (post && post[i].brk) {
                const ss_post_lufs& pl = post[i];
                ss_loud_item li; memset(&li, 0, sizeof(li));
                const Tab& tb = tabs[tab_of[i]];
                li.brk = tb.d_brk; li.blk_lo = tb.d_lo; li.blk_hi = tb.d_hi;
                li.scratch = (double*)po; po += align_up(8 * (size_t)SS_LOUD_SCRATCH_DOUBLES * it.C * pl.n_e, 256);
                li.result = (double*)po; po += 256;
                res_dev[i] = li.result;
                li.data = d.out; li.out = d.out;                 // (C, N) stem, normalised in place
                li.stride_n = 1; li.stride_c = it.N; li.N = it.N; li.C = it.C; li.n_e = pl.n_e; li.n_blocks = pl.n_blocks;
                li.rate = pl.rate; li.block_size = pl.block_size; li.target_lufs = pl.target_lufs;
                loud.push_back(li);
            }
        }
        CK(cudaEventRecord(sl.ev_in, c->s_in));
        CK(cudaStreamWaitEvent(c->s_cmp, sl.ev_in, 0));
        rc = launch_chunk(c, dev.data(), first, last, c->s_cmp);
        if (rc != SS_OK) break;
        if (!loud.empty()) {
            cudaStream_t s_l = c->s_aux[k % ss_ctx::kAux];
            CK(cudaEventRecord(sl.ev_rend, c->s_cmp));
            CK(cudaStreamWaitEvent(s_l, sl.ev_rend, 0));
            rc = ss_loudness_dev(c, loud.data(), (int)loud.size(), (void*)s_l);
            if (rc != SS_OK) break;
            CK(cudaEventRecord(sl.ev_done, s_l));
        } else {
            CK(cudaEventRecord(sl.ev_done, c->s_cmp));
        }
        CK(cudaStreamWaitEvent(c->s_out, sl.ev_done, 0));
        for (int i = first; i < last; ++i) {
            CK(cudaMemcpyAsync(items[i].out, dev[i].out, sizeof(float) * (size_t)items[i].C * items[i].N,
                               cudaMemcpyDeviceToHost, c->s_out));
            if (res_dev[i])
                CK(cudaMemcpyAsync(h_res + 2 * i, res_dev[i], 2 * sizeof(double), cudaMemcpyDeviceToHost, c->s_out));
        }
        CK(cudaEventRecord(sl.ev_free, c->s_out));
        // the next chunk's H2D into the *other* slot may start now; it must not overtake the
        // render still reading this slot, which the per-slot ev_free wait above guarantees.
    }
    return rc;
    };
    const int rc = pipeline();
    cudaError_t e1 = cudaStreamSynchronize(c->s_in), e2 = cudaStreamSynchronize(c->s_cmp), e3 = cudaSuccess;
    if (post) for (int i = 0; i < ss_ctx::kAux; ++i) { cudaError_t e = cudaStreamSynchronize(c->s_aux[i]); if (e3 == cudaSuccess) e3 = e; }
    cudaError_t e4 = cudaStreamSynchronize(c->s_out);
    if (rc != SS_OK) return rc;
    CK(e1); CK(e2); CK(e3); CK(e4);
    if (post)
        for (int i = 0; i < n_items; ++i)
            if (res_dev[i] && post[i].result) { post[i].result[0] = h_res[2 * i]; post[i].result[1] = h_res[2 * i + 1]; }
    return SS_OK;
}

extern "C" int ss_render_host(ss_ctx* c, const ss_source* items, int n_items) {
    return ss_render_host_ex(c, items, n_items, nullptr);
}

extern "C" int ss_convolve_fixed_receiver(ss_ctx* c, const float* x, const float* rirs, float* out,
                                          int32_t N, int32_t C, int32_t L) {
    ss_source it; memset(&it, 0, sizeof(it));
    it.x = x; it.rir = rirs; it.out = out; it.N = N; it.P = 1; it.C = C; it.L = L; it.mode = SS_STATIC;
    return ss_render_host(c, &it, 1);
}
extern "C" int ss_convolve_moving_receiver(ss_ctx* c, const float* x, const float* rirs, const int32_t* idx,
                                           const float* w, float* out, int32_t N, int32_t P, int32_t C, int32_t L) {
    ss_source it; memset(&it, 0, sizeof(it));
    it.x = x; it.rir = rirs; it.out = out; it.idx = idx; it.w = w;
    it.N = N; it.P = P; it.C = C; it.L = L; it.mode = SS_MOVING_INDEXED;
    return ss_render_host(c, &it, 1);
}
