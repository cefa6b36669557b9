// sonicsim_b200 :: ss_core.cuh
//
// Per-thread building blocks of the partitioned overlap-save renderer.  Everything here is
// `__host__ __device__` so that tests/emu/ can run the *same* index math and butterflies on
// the CPU, one emulated thread at a time (the authoring container has no GPU).  The CUDA
// kernels in ss_kernels.cu and the emulator in tests/emu/ss_emu.cu share these functions and
// differ only in who loops over `tid` and where the barriers are.
//
// Geometry (fixed): FFT size F = 8192 complex points, partition / block size B = 4096 samples,
// 256 threads per CTA, 32 complex points per thread.  The 8192-point transform is a 3-pass
// Stockham radix-16 x 16 x 16 with the closing radix-2 done in registers (each thread owns
// butterflies j and j+256 of the last pass, whose outputs are exactly a radix-2 pair).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SS_HD __host__ __device__ __forceinline__
#else
#define SS_HD inline
#endif

namespace ss {

constexpr int kF = 8192;          // FFT length (complex points)
constexpr int kB = 4096;          // overlap-save block = RIR partition length (samples / taps)
constexpr int kThreads = 256;     // threads per CTA
constexpr int kSpec = 4096;       // float2 per stored half spectrum; [0] = (DC, Nyquist)
constexpr int kPadF = kF + kF / 16;   // padded smem length in float2

// One float2 of padding per 16: makes the stride-16 scatter of pass A and every other
// half-warp access pattern used below hit 16 distinct 8-byte bank pairs.
SS_HD int pad(int i) { return i + (i >> 4); }

SS_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SS_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SS_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SS_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> SS_HD float2 rot90(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }
// conjugate the twiddle for the inverse transform
template <bool INV> SS_HD float2 dirw(float2 w) { return INV ? make_float2(w.x, -w.y) : w; }

// Streaming 8-byte load: spectra are read once per CTA, keep them out of L1 so the twiddle tables
// stay resident there.
SS_HD float2 ldg_stream(const float2* p) {
#if defined(__CUDA_ARCH__)
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
#else
    return *p;
#endif
}
SS_HD float2 ldg_cached(const float2* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// Fire-and-forget accumulate into global memory (second transform of a block that straddles a
// waypoint).  Exactly two addends ever meet at an address and the first was stored by the same
// thread earlier, so the result is order-independent (a + b == b + a in IEEE arithmetic).
SS_HD void red_add(float* p, float v) {
#if defined(__CUDA_ARCH__)
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
#else
    *p += v;
#endif
}

// (1 - w) * a + w * b with every product and the sum rounded separately, like NumPy evaluates
// SonicSim_moving.py:94 (no FMA contraction), so both trajectory forms give identical bits.
SS_HD float lerp_terms(float fa, float a, float fb, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(__fmul_rn(fa, a), __fmul_rn(fb, b));
#else
    volatile float pa = fa * a, pb = fb * b;
    return pa + pb;
#endif
}
SS_HD float one_minus(float w) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(1.0f, w);
#else
    return 1.0f - w;
#endif
}

// Twiddle tables (forward sign, filled in double precision on the host):
//   tw [m]           = exp(-2 pi i m / 8192)            m < 8192   (closing radix-2)
//   twB[r * 16 + k]  = exp(-2 pi i k r / 256)           r, k < 16  (pass B)
//   twC[r * 256 + k] = exp(-2 pi i k r / 4096)          r < 16, k < 256 (pass C)
// r-major so that a warp (consecutive k) reads consecutive words.
struct Tables { const float2* tw; const float2* twB; const float2* twC; };
constexpr int kTabB = 16 * 16, kTabC = 16 * 256;

template <bool INV>
SS_HD void fft4(float2& v0, float2& v1, float2& v2, float2& v3) {
    float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = rot90<INV>(csub(v1, v3));
    v0 = cadd(a0, a2); v2 = csub(a0, a2); v1 = cadd(a1, a3); v3 = csub(a1, a3);
}

// In-register 16-point DFT (4 x 4).  Input natural order v[n]; output X[a + 4c] is left in
// v[4a + c] (see out16()).
template <bool INV>
SS_HD void fft16(float2 (&v)[16]) {
    const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, H = 0.70710678118654752f;
#pragma unroll
    for (int b = 0; b < 4; ++b) fft4<INV>(v[b], v[b + 4], v[b + 8], v[b + 12]);
    // twiddle v[b + 4a] *= W16^(a b)   (forward W16 = exp(-2 pi i / 16))
    v[5]  = cmul(v[5],  dirw<INV>(make_float2(C1, -S1)));      // ab = 1
    v[6]  = cmul(v[6],  dirw<INV>(make_float2(H, -H)));        // ab = 2
    v[7]  = cmul(v[7],  dirw<INV>(make_float2(S1, -C1)));      // ab = 3
    v[9]  = cmul(v[9],  dirw<INV>(make_float2(H, -H)));        // ab = 2
    v[10] = rot90<INV>(v[10]);                                 // ab = 4
    v[11] = cmul(v[11], dirw<INV>(make_float2(-H, -H)));       // ab = 6
    v[13] = cmul(v[13], dirw<INV>(make_float2(S1, -C1)));      // ab = 3
    v[14] = cmul(v[14], dirw<INV>(make_float2(-H, -H)));       // ab = 6
    v[15] = cmul(v[15], dirw<INV>(make_float2(-C1, S1)));      // ab = 9
#pragma unroll
    for (int a = 0; a < 4; ++a) fft4<INV>(v[4 * a], v[4 * a + 1], v[4 * a + 2], v[4 * a + 3]);
}
// register slot that holds output index r after fft16
SS_HD constexpr int out16(int r) { return 4 * (r & 3) + (r >> 2); }

// v[r] *= tab[r * stride], r = 1..15 (tab points at the entry of this butterfly's k)
template <bool INV, int STRIDE>
SS_HD void twiddle16(float2 (&v)[16], const float2* tab) {
#pragma unroll
    for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], dirw<INV>(ldg_cached(tab + r * STRIDE)));
}

// ---------------------------------------------------------------------------------------
// Stockham passes on a padded shared-memory array `s` (float2[kPadF]).
// `tw` is the table exp(-2 pi i m / 8192), m in [0, 8192).
// Butterfly j of a pass with sub-length Ns reads s[j + 512 r] and writes
// s[(j / Ns) * 16 Ns + (j % Ns) + r Ns].
// ---------------------------------------------------------------------------------------

// Pass A (Ns = 1) store: v holds fft16 outputs of butterfly j (no input twiddles when Ns = 1).
// pad(16 j + r) = 17 j + r.
SS_HD void passA_store(float2* s, int j, const float2 (&v)[16]) {
    float2* d = s + 17 * j;
#pragma unroll
    for (int r = 0; r < 16; ++r) d[r] = v[out16(r)];
}

// pad(j + 512 r) = pad(j) + 544 r
SS_HD void pass_load(const float2* s, int j, float2 (&v)[16]) {
    const float2* p = s + pad(j);
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = p[544 * r];
}

// Pass B (Ns = 16): twiddle exp(-+2 pi i k r / 256), k = j % 16
template <bool INV>
SS_HD void passB_compute(int j, float2 (&v)[16], const Tables& T) {
    twiddle16<INV, 16>(v, T.twB + (j & 15));
    fft16<INV>(v);
}
// destination (j / 16) * 256 + (j % 16) + 16 r, padded = (j / 16) * 272 + (j % 16) + 17 r
SS_HD void passB_store(float2* s, int j, const float2 (&v)[16]) {
    float2* d = s + (j >> 4) * 272 + (j & 15);
#pragma unroll
    for (int r = 0; r < 16; ++r) d[17 * r] = v[out16(r)];
}

// Pass C (Ns = 256): twiddle exp(-+2 pi i k r / 4096), k = j % 256.  Outputs stay in registers:
// butterfly j (< 256) yields natural-order points j + 256 r of the first half-length transform,
// butterfly j + 256 the same points of the second one.
template <bool INV>
SS_HD void passC_compute(int j, float2 (&v)[16], const Tables& T) {
    twiddle16<INV, 256>(v, T.twC + (j & 255));
    fft16<INV>(v);
}

// Closing radix-2 (Ns = 4096) for thread t: lo/hi are the pass-C results of butterflies t and t+256.
// Returns through lo[slot] the points t + 256 r and through hi[slot] the points 4096 + t + 256 r,
// slot = out16(r).
template <bool INV>
SS_HD float2 final_twiddle(int t, int r, float2 wt /* tw[t], already direction-adjusted */) {
    // exp(-+2 pi i (t + 256 r) / 8192) = wt * exp(-+2 pi i r / 32)
    const float c32[16] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                           0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f,
                           0.19509032201612825f, 0.f, -0.19509032201612825f, -0.38268343236508977f,
                           -0.55557023301960218f, -0.70710678118654752f, -0.83146961230254524f,
                           -0.92387953251128674f, -0.98078528040323043f};
    const float s32[16] = {0.f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f,
                           0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f,
                           0.98078528040323043f, 1.f, 0.98078528040323043f, 0.92387953251128674f,
                           0.83146961230254524f, 0.70710678118654752f, 0.55557023301960218f,
                           0.38268343236508977f, 0.19509032201612825f};
    float2 c = make_float2(c32[r], INV ? s32[r] : -s32[r]);
    return cmul(wt, c);
}

// ---------------------------------------------------------------------------------------
// Spectrum-domain helpers.
//   stored half spectrum S[0] = (DC, Nyquist), S[k] = bin k for 1 <= k < 4096.
// For a pair of real sequences (a, b) convolved with real x:  z = x*a + i x*b,
//   Z[k]   = P + iQ,  Z[F-k] = conj(P) + i conj(Q),   P = X[k] A[k], Q = X[k] B[k].
// ---------------------------------------------------------------------------------------
SS_HD float2 z_direct(float2 P, float2 Q) { return make_float2(P.x - Q.y, P.y + Q.x); }
SS_HD float2 z_mirror(float2 P, float2 Q) { return make_float2(P.x + Q.y, Q.x - P.y); }
SS_HD void cmac(float2& acc, float2 a, float2 b) {
    acc.x += a.x * b.x - a.y * b.y;
    acc.y += a.x * b.y + a.y * b.x;
}

// Ownership of pass-A butterflies: thread t (1..255) owns jA = t and jB = 512 - t; thread 0 owns
// jA = 0 and jB = 256.  With that choice every bin k < 4096 a thread loads also supplies the
// mirrored bin F - k of one of its own butterflies, so each spectrum word is read exactly once.
SS_HD int passA_jA(int t) { return t; }
SS_HD int passA_jB(int t) { return t == 0 ? 256 : 512 - t; }

// ---------------------------------------------------------------------------------------
// Trajectory helpers (compact form of SonicSim_moving.setup_dynamic_interp's (idx, w)):
//   bounds[0..S], bounds[0] = 0, bounds[S] = N, segment s covers [bounds[s], bounds[s+1]).
// ---------------------------------------------------------------------------------------
SS_HD int seg_of(const int* bounds, int S, int n) {
    // number of i in [1, S] with bounds[i] <= n  (skips empty segments like np.repeat does)
    int lo = 0, hi = S;            // answer in [lo, hi]
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (bounds[mid] <= n) lo = mid; else hi = mid - 1;
    }
    return lo < S ? lo : S - 1;    // n == N cannot happen for valid samples; clamp anyway
}
// np.linspace(0, 1, num, endpoint=False)[i].astype(float32): float64 i * (1/num), then rounded.
SS_HD float seg_weight(int i, int num) {
    double step = 1.0 / (double)num;
    return (float)((double)i * step);
}

}  // namespace ss
