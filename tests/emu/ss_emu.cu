// tests/emu/ss_emu.cu - CPU emulation of the CUDA kernels, one emulated thread at a time.
//
// TEST INFRASTRUCTURE (not a product path, never shipped as a fallback).  It includes the very
// same per-thread phase functions the kernels call (sonicsim_b200/csrc/ss_phases.cuh) and replaces
// __syncthreads() by "finish the phase for every tid".  This is how index math, butterfly
// constants and the overlap-save / hat-weight logic are validated in the GPU-less authoring
// container.  Built with g++ (`-x c++`) into tests/emu/libss_emu.so by tests/emu/build.py.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../sonicsim_b200/csrc/ss_phases.cuh"
#include "../../sonicsim_b200/csrc/ss_loud.cuh"

using namespace ss;

struct HostTables {
    std::vector<float2> tw, tb, tc;
    Tables T;
    HostTables() : tw(kF), tb(kTabB), tc(kTabC) {
        for (int m = 0; m < kF; ++m) {
            double a = -2.0 * M_PI * (double)m / (double)kF;
            tw[m] = make_float2((float)cos(a), (float)sin(a));
        }
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
        T.tw = tw.data(); T.twB = tb.data(); T.twC = tc.data();
    }
};
static const Tables& tables() { static HostTables h; return h.T; }

static void cta_spectra(const Source& S, int local, const Tables& T) {
    std::vector<float2> smem(kPadF);
    std::vector<Regs32> R(kThreads);
    Row ra, rb;
    const int nh = spectra_pairs_h(S);
    if (local < nh) { ra = make_row_h(S, 2 * local); rb = make_row_h(S, 2 * local + 1); }
    else { local -= nh; ra = make_row_x(S, 2 * local); rb = make_row_x(S, 2 * local + 1); }
    float2* s = smem.data();
    for (int t = 0; t < kThreads; ++t) spectra_phase1(t, ra, rb, s);
    for (int t = 0; t < kThreads; ++t) load2(t, s, R[t]);
    for (int t = 0; t < kThreads; ++t) passB2<false>(t, s, R[t], T);
    for (int t = 0; t < kThreads; ++t) { load2(t, s, R[t]); spectra_phase3_compute(t, R[t], T); }
    for (int t = 0; t < kThreads; ++t) spectra_phase3_store(t, s, R[t]);
    for (int t = 0; t < kThreads; ++t) spectra_phase4(t, s, ra, rb);
}

static void cta_render(const Source& S, int local, const Tables& T) {
    std::vector<float2> smem(kPadF);
    std::vector<Regs32> R(kThreads);
    float2* s = smem.data();
    const int mode = S.mode;
    int b, c, p_lo = 0, p_hi = 0;
    std::vector<int> sg0(kThreads, 0);
    if (mode == MODE_STATIC) {
        const int ncp = (S.C + 1) >> 1;
        b = local / ncp; c = 2 * (local - b * ncp);
    } else {
        b = local / S.C; c = local - b * S.C;
    }
    const int n0 = b * kB;
    if (mode == MODE_MOVING_BOUNDS) {
        const int n_last = (n0 + kB < S.N ? n0 + kB : S.N) - 1;
        p_lo = seg_of(S.bounds, S.P - 1, n0);
        p_hi = seg_of(S.bounds, S.P - 1, n_last) + 1;
        for (int t = 0; t < kThreads; ++t) sg0[t] = seg_of(S.bounds, S.P - 1, n0 + t < S.N ? n0 + t : S.N - 1);
    } else if (mode == MODE_MOVING_INDEXED) {
        int pmin = 0x7fffffff, pmax = -1;
        for (int t = 0; t < kThreads; ++t) {
            int a, bb;
            idx_range(t, n0, S, a, bb);
            pmin = a < pmin ? a : pmin; pmax = bb > pmax ? bb : pmax;
        }
        p_lo = pmin < 0 ? 0 : pmin;
        p_hi = pmax + 1 > S.P - 1 ? S.P - 1 : pmax + 1;
    }
    const float2* X0 = S.xspec + (size_t)b * kSpec;
    float* row = S.out + (size_t)c * S.N;
    for (int p = p_lo; p <= p_hi; p += 2) {
        const float2 *Hp, *Hq;
        if (mode == MODE_STATIC) {
            Hp = S.hspec + (size_t)c * S.K * kSpec;
            Hq = (c + 1 < S.C) ? Hp + (size_t)S.K * kSpec : nullptr;
        } else {
            Hp = S.hspec + ((size_t)p * S.C + c) * S.K * kSpec;
            Hq = (p + 1 <= p_hi) ? Hp + (size_t)S.C * S.K * kSpec : nullptr;
        }
        for (int t = 0; t < kThreads; ++t) form_z(t, X0, b, S.K, Hp, Hq, R[t]);
        for (int t = 0; t < kThreads; ++t) render_phase1(t, s, R[t]);
        for (int t = 0; t < kThreads; ++t) load2(t, s, R[t]);
        for (int t = 0; t < kThreads; ++t) passB2<true>(t, s, R[t], T);
        for (int t = 0; t < kThreads; ++t) {
            load2(t, s, R[t]);
            render_phase3(t, R[t], T);
            if (mode == MODE_MOVING_BOUNDS) { MovingSinkBounds sk(S, row, n0, t, p, sg0[t], p == p_lo); render_epilogue(R[t], sk); }
            else if (mode == MODE_MOVING_INDEXED) { MovingSinkIndexed sk(S, row, n0, t, p, p == p_lo); render_epilogue(R[t], sk); }
            else { StaticSink sk{row, (c + 1 < S.C) ? row + S.N : nullptr, S.N, n0 + t}; render_epilogue(R[t], sk); }
        }
    }
}

extern "C" {

// Emulate k_spectra + k_render for one source.  mode: 0 static, 1 bounds, 2 (idx, w).
int emu_render(const float* x, const float* rir, float* out, const int32_t* bounds, const int32_t* idx,
               const float* w, int N, int P, int C, int L, int mode) {
    const Tables& T = tables();
    Source S;
    memset(&S, 0, sizeof(S));
    S.x = x; S.rir = rir; S.out = out; S.bounds = bounds; S.idx = idx; S.w = w;
    S.N = N; S.P = P; S.C = C; S.L = L; S.K = (L + kB - 1) / kB; S.nb = (N + kB - 1) / kB; S.mode = mode;
    std::vector<float2> hs((size_t)P * C * S.K * kSpec), xs((size_t)S.nb * kSpec);
    S.hspec = hs.data(); S.xspec = xs.data();
    const int ns = spectra_pairs_h(S) + spectra_pairs_x(S);
    for (int i = 0; i < ns; ++i) cta_spectra(S, i, T);
    const int nr = render_ctas(S);
    for (int i = 0; i < nr; ++i) cta_render(S, i, T);
    return 0;
}

// forward 8192-point FFT of z = a + i b via the spectra kernel phases; returns the two half spectra
int emu_spectra_pair(const float* a, const float* b, int len, float* specA, float* specB) {
    const Tables& T = tables();
    Source S; memset(&S, 0, sizeof(S));
    std::vector<float> rir(2 * (size_t)len);
    memcpy(rir.data(), a, sizeof(float) * len); memcpy(rir.data() + len, b, sizeof(float) * len);
    S.rir = rir.data(); S.P = 1; S.C = 2; S.L = len; S.K = (len + kB - 1) / kB; S.N = 1; S.nb = 1;
    std::vector<float2> hs((size_t)2 * S.K * kSpec);
    S.hspec = hs.data();
    for (int i = 0; i < spectra_pairs_h(S); ++i) cta_spectra(S, i, T);
    // return partition 0 of each row
    memcpy(specA, hs.data(), sizeof(float2) * kSpec);
    memcpy(specB, hs.data() + (size_t)S.K * kSpec, sizeof(float2) * kSpec);
    return 0;
}

// loudness: k_kweight_energy + k_loud_gate, one emulated thread per (channel, interval)
int emu_lufs(const float* data, int N, int C, long long stride_n, long long stride_c, double rate, double block_size,
             double target, const int32_t* brk, int n_e, const int32_t* blk_lo, const int32_t* blk_hi, int n_blocks,
             double* result) {
    LoudItem it; memset(&it, 0, sizeof(it));
    std::vector<double> E((size_t)C * n_e);
    it.data = data; it.brk = brk; it.blk_lo = blk_lo; it.blk_hi = blk_hi; it.E = E.data(); it.result = result;
    it.stride_n = stride_n; it.stride_c = stride_c; it.N = N; it.C = C; it.n_e = n_e; it.n_blocks = n_blocks;
    it.warm = (int)ceil(0.128 * rate); it.inv_norm = 1.0 / (block_size * rate); it.target = target;
    KCoef k = make_kcoef(rate);
    for (int c = 0; c < C; ++c) for (int e = 0; e < n_e; ++e) E[(size_t)c * n_e + e] = kweight_interval_energy(it, k, c, e);
    loudness_gate(it);
    return 0;
}

}  // extern "C"
