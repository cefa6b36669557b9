// sonicsim_b200 :: ss_loud.cuh - ITU-R BS.1770 loudness (K-weighting + gated mean square) as
// pyloudnorm 0.1.1 computes it for SonicSim_audio.lufs_norm (SonicSim_audio.py:68-81).
// Host/device per-thread code, shared with the CPU emulation in tests/emu.
#pragma once
#include "ss_core.cuh"
#include <math.h>

namespace ss {

// two RBJ biquads, a0-normalised: high shelf (+4 dB, Q = 1/sqrt 2, 1500 Hz), high pass (Q = 0.5, 38 Hz)
struct KCoef { double b0[2], b1[2], b2[2], a1[2], a2[2]; };

inline KCoef make_kcoef(double rate) {
    KCoef k;
    const double PI = 3.14159265358979323846;
    {   // high shelf
        double G = 4.0, Q = 1.0 / sqrt(2.0), fc = 1500.0;
        double A = pow(10.0, G / 40.0), w0 = 2.0 * PI * (fc / rate), alpha = sin(w0) / (2.0 * Q);
        double b0 = A * ((A + 1) + (A - 1) * cos(w0) + 2 * sqrt(A) * alpha);
        double b1 = -2 * A * ((A - 1) + (A + 1) * cos(w0));
        double b2 = A * ((A + 1) + (A - 1) * cos(w0) - 2 * sqrt(A) * alpha);
        double a0 = (A + 1) - (A - 1) * cos(w0) + 2 * sqrt(A) * alpha;
        double a1 = 2 * ((A - 1) - (A + 1) * cos(w0));
        double a2 = (A + 1) - (A - 1) * cos(w0) - 2 * sqrt(A) * alpha;
        k.b0[0] = b0 / a0; k.b1[0] = b1 / a0; k.b2[0] = b2 / a0; k.a1[0] = a1 / a0; k.a2[0] = a2 / a0;
    }
    {   // high pass
        double Q = 0.5, fc = 38.0;
        double w0 = 2.0 * PI * (fc / rate), alpha = sin(w0) / (2.0 * Q);
        double b0 = (1 + cos(w0)) / 2, b1 = -(1 + cos(w0)), b2 = (1 + cos(w0)) / 2;
        double a0 = 1 + alpha, a1 = -2 * cos(w0), a2 = 1 - alpha;
        k.b0[1] = b0 / a0; k.b1[1] = b1 / a0; k.b2[1] = b2 / a0; k.a1[1] = a1 / a0; k.a2[1] = a2 / a0;
    }
    return k;
}

// One stem to measure (and optionally normalise).  Element (n, c) is data[n * stride_n + c * stride_c].
struct LoudItem {
    const float* data;
    float* out;              // C*N contiguous floats scaled by the gain (may alias data); null = measure only
    const int* brk;          // n_e + 1 ascending sample indices: the gating blocks' distinct bounds
    const int* blk_lo;       // per gating block j: first elementary interval
    const int* blk_hi;       // per gating block j: one past its last elementary interval
    double* E;               // scratch, C * n_e interval energies (sum of squares of the K-weighted signal)
    double* result;          // [0] integrated loudness (LUFS, -inf allowed), [1] linear gain applied
    long long stride_n, stride_c;
    int N, C, n_e, n_blocks, warm, pad_;
    double inv_norm;         // 1 / (T_g * rate)
    double target;           // target LUFS (SonicSim_audio.py:77 `norm`)
};

// Energy of the K-weighted channel over [brk[e], brk[e+1]).  The recursion is started `warm`
// samples early from rest; the high-pass double pole (r ~ 0.985 at 16 kHz) has decayed below
// 1e-10 by then, and when the start clamps to sample 0 the state is exact.
SS_HD double kweight_interval_energy(const LoudItem& it, const KCoef& k, int c, int e) {
    const int start = it.brk[e], end = it.brk[e + 1];
    int n = start - it.warm; if (n < 0) n = 0;
    const float* p = it.data + (long long)c * it.stride_c;
    double z1a = 0, z2a = 0, z1b = 0, z2b = 0, acc = 0;
    for (; n < end; ++n) {
        double x = (double)p[(long long)n * it.stride_n];
        // scipy.signal.lfilter: direct form II transposed, float64; pyloudnorm stores each stage back
        // into the float32 array (meter.py: input_data[:,ch] = filter.apply_filter(...))
        double y = k.b0[0] * x + z1a;
        z1a = k.b1[0] * x - k.a1[0] * y + z2a;
        z2a = k.b2[0] * x - k.a2[0] * y;
        double x2 = (double)(float)y;
        double y2 = k.b0[1] * x2 + z1b;
        z1b = k.b1[1] * x2 - k.a1[1] * y2 + z2b;
        z2b = k.b2[1] * x2 - k.a2[1] * y2;
        if (n >= start) { float yf = (float)y2; acc += (double)(yf * yf); }
    }
    return acc;
}

SS_HD double channel_gain(int c) { return (c == 3 || c == 4) ? 1.41 : 1.0; }   // pyloudnorm G = [1,1,1,1.41,1.41]

// Gating (pyloudnorm meter.py integrated_loudness) for one stem; single thread.
SS_HD void loudness_gate(const LoudItem& it) {
    const int nb = it.n_blocks, C = it.C;
    // pass 1: absolute gate (-70 LUFS) -> per-channel mean of z over the blocks that pass
    double zsum[8]; int cnt = 0;
    for (int c = 0; c < C && c < 8; ++c) zsum[c] = 0;
    for (int pass = 0; pass < 2; ++pass) {
        double gamma_r = -1e300;
        if (pass == 1) {
            double s = 0;
            for (int c = 0; c < C; ++c) s += channel_gain(c) * (cnt ? zsum[c] / cnt : NAN);
            gamma_r = -0.691 + 10.0 * log10(s) - 10.0;        // NaN when no block passes -> nothing passes below
            for (int c = 0; c < C; ++c) zsum[c] = 0;
            cnt = 0;
        }
        for (int j = 0; j < nb; ++j) {
            double zc[8], s = 0;
            for (int c = 0; c < C; ++c) {
                double z = 0;
                for (int e = it.blk_lo[j]; e < it.blk_hi[j]; ++e) z += it.E[(long long)c * it.n_e + e];
                zc[c] = z * it.inv_norm;
                s += channel_gain(c) * zc[c];
            }
            double l = -0.691 + 10.0 * log10(s);
            bool ok = pass == 0 ? (l >= -70.0) : (l > gamma_r && l > -70.0);
            if (ok) { for (int c = 0; c < C; ++c) zsum[c] += zc[c]; ++cnt; }
        }
    }
    double s = 0;
    for (int c = 0; c < C; ++c) s += channel_gain(c) * (cnt ? zsum[c] / cnt : 0.0);   // nan_to_num(mean of empty) = 0
    double lufs = -0.691 + 10.0 * log10(s);                                            // log10(0) = -inf
    double used = isinf(lufs) ? -40.0 : lufs;                                          // SonicSim_audio.py:73-75
    it.result[0] = lufs;
    it.result[1] = pow(10.0, (it.target - used) / 20.0);                              // pyln.normalize.loudness
}

}  // namespace ss
