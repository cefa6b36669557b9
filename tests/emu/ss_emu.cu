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

using namespace ss;

static std::vector<float2> make_tw() {
    std::vector<float2> tw(kF);
    for (int m = 0; m < kF; ++m) {
        double a = -2.0 * M_PI * (double)m / (double)kF;
        tw[m] = make_float2((float)cos(a), (float)sin(a));
    }
    return tw;
}

static void cta_spectra(const Source& S, int local, const float2* tw) {
    std::vector<float2> smem(kPadF);
    std::vector<Regs32> R(kThreads);
    Row ra, rb;
    const int nh = spectra_pairs_h(S);
    if (local < nh) { ra = make_row_h(S, 2 * local); rb = make_row_h(S, 2 * local + 1); }
    else { local -= nh; ra = make_row_x(S, 2 * local); rb = make_row_x(S, 2 * local + 1); }
    float2* s = smem.data();
    for (int t = 0; t < kThreads; ++t) spectra_phase1(t, ra, rb, s);
    for (int t = 0; t < kThreads; ++t) load2(t, s, R[t]);
    for (int t = 0; t < kThreads; ++t) passB2<false>(t, s, R[t], tw);
    for (int t = 0; t < kThreads; ++t) { load2(t, s, R[t]); spectra_phase3_compute(t, R[t], tw); }
    for (int t = 0; t < kThreads; ++t) spectra_phase3_store(t, s, R[t]);
    for (int t = 0; t < kThreads; ++t) spectra_phase4(t, s, ra, rb);
}

static void ifft_passes(float2* s, std::vector<Regs32>& R, const float2* tw) {
    for (int t = 0; t < kThreads; ++t) render_phase1(t, s, R[t]);
    for (int t = 0; t < kThreads; ++t) load2(t, s, R[t]);
    for (int t = 0; t < kThreads; ++t) passB2<true>(t, s, R[t], tw);
    for (int t = 0; t < kThreads; ++t) load2(t, s, R[t]);
}

static void cta_render(const Source& S, int local, const float2* tw) {
    std::vector<float2> smem(kPadF);
    std::vector<Regs32> R(kThreads);
    float2* s = smem.data();
    if (S.mode == MODE_STATIC) {
        const int ncp = (S.C + 1) >> 1;
        const int b = local / ncp, cp = local - b * ncp;
        const int c0 = 2 * cp, c1 = c0 + 1;
        const int n0 = b * kB;
        const float2* X0 = S.xspec + (size_t)b * kSpec;
        const float2* Hp = S.hspec + (size_t)c0 * S.K * kSpec;
        const float2* Hq = (c1 < S.C) ? S.hspec + (size_t)c1 * S.K * kSpec : nullptr;
        for (int t = 0; t < kThreads; ++t) form_z(t, X0, b, S.K, Hp, Hq, R[t]);
        ifft_passes(s, R, tw);
        for (int t = 0; t < kThreads; ++t) {
            passC_compute<true>(t, R[t].a, tw);
            passC_compute<true>(t + 256, R[t].b, tw);
            const float2 wt = dirw<true>(tw[t]);
            for (int r = 0; r < 16; ++r) {
                const int sl = out16(r);
                float2 z = csub(R[t].a[sl], cmul(R[t].b[sl], final_twiddle<true>(t, r, wt)));
                int n = n0 + t + 256 * r;
                if (n < S.N) {
                    S.out[(size_t)c0 * S.N + n] = z.x;
                    if (c1 < S.C) S.out[(size_t)c1 * S.N + n] = z.y;
                }
            }
        }
        return;
    }
    const int b = local / S.C, c = local - b * S.C;
    const int n0 = b * kB;
    std::vector<float> accs(kThreads * 16, 0.f);
    int p_lo, p_hi;
    if (S.mode == MODE_MOVING_BOUNDS) {
        const int n_last = (n0 + kB < S.N ? n0 + kB : S.N) - 1;
        p_lo = seg_of(S.bounds, S.P - 1, n0);
        p_hi = seg_of(S.bounds, S.P - 1, n_last) + 1;
    } else {
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
    for (int p = p_lo; p <= p_hi; p += 2) {
        const float2* Hp = S.hspec + ((size_t)p * S.C + c) * S.K * kSpec;
        const float2* Hq = (p + 1 <= p_hi) ? S.hspec + ((size_t)(p + 1) * S.C + c) * S.K * kSpec : nullptr;
        for (int t = 0; t < kThreads; ++t) form_z(t, X0, b, S.K, Hp, Hq, R[t]);
        ifft_passes(s, R, tw);
        for (int t = 0; t < kThreads; ++t) {
            float (&acc)[16] = *reinterpret_cast<float (*)[16]>(&accs[t * 16]);
            if (S.mode == MODE_MOVING_BOUNDS) { BoundsWeights wf(S, n0, t, p); render_phase3(t, R[t], tw, acc, wf); }
            else { IndexedWeights wf(S, n0, t, p); render_phase3(t, R[t], tw, acc, wf); }
        }
    }
    for (int t = 0; t < kThreads; ++t) {
        float (&acc)[16] = *reinterpret_cast<float (*)[16]>(&accs[t * 16]);
        store_block(t, n0, S, S.out + (size_t)c * S.N, acc, 1.0f);
    }
}

extern "C" {

// Emulate k_spectra + k_render for one source.  mode: 0 static, 1 bounds, 2 (idx, w).
int emu_render(const float* x, const float* rir, float* out, const int32_t* bounds, const int32_t* idx,
               const float* w, int N, int P, int C, int L, int mode) {
    static std::vector<float2> tw = make_tw();
    Source S;
    memset(&S, 0, sizeof(S));
    S.x = x; S.rir = rir; S.out = out; S.bounds = bounds; S.idx = idx; S.w = w;
    S.N = N; S.P = P; S.C = C; S.L = L; S.K = (L + kB - 1) / kB; S.nb = (N + kB - 1) / kB; S.mode = mode;
    std::vector<float2> hs((size_t)P * C * S.K * kSpec), xs((size_t)S.nb * kSpec);
    S.hspec = hs.data(); S.xspec = xs.data();
    const int ns = spectra_pairs_h(S) + spectra_pairs_x(S);
    for (int i = 0; i < ns; ++i) cta_spectra(S, i, tw.data());
    const int nr = render_ctas(S);
    for (int i = 0; i < nr; ++i) cta_render(S, i, tw.data());
    return 0;
}

// forward 8192-point FFT of z = a + i b via the spectra kernel phases; returns the two half spectra
int emu_spectra_pair(const float* a, const float* b, int len, float* specA, float* specB) {
    static std::vector<float2> tw = make_tw();
    Source S; memset(&S, 0, sizeof(S));
    std::vector<float> rir(2 * (size_t)len);
    memcpy(rir.data(), a, sizeof(float) * len); memcpy(rir.data() + len, b, sizeof(float) * len);
    S.rir = rir.data(); S.P = 1; S.C = 2; S.L = len; S.K = (len + kB - 1) / kB; S.N = 1; S.nb = 1;
    std::vector<float2> hs((size_t)2 * S.K * kSpec);
    S.hspec = hs.data();
    for (int i = 0; i < spectra_pairs_h(S); ++i) cta_spectra(S, i, tw.data());
    // return partition 0 of each row
    memcpy(specA, hs.data(), sizeof(float2) * kSpec);
    memcpy(specB, hs.data() + (size_t)S.K * kSpec, sizeof(float2) * kSpec);
    return 0;
}

}  // extern "C"
