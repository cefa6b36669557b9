// sonicsim_b200 :: ss_phases.cuh
//
// The kernels of the renderer, cut into barrier-free *phases*.  A phase is a function of one
// thread id, that thread's registers (a struct) and the CTA's shared array.  ss_kernels.cu calls
// the phases with __syncthreads() between them; tests/emu/ss_emu.cu calls the same phases in a
// `for (tid)` loop per phase.  All addressing and arithmetic is therefore tested on the CPU.
#pragma once
#include "ss_core.cuh"

namespace ss {

enum { MODE_STATIC = 0, MODE_MOVING_BOUNDS = 1, MODE_MOVING_INDEXED = 2 };

// One (utterance, source) unit.  Reference shapes: dry (N,), RIRs (P, C, L), output (C, N)
// (SonicSim_moving.py:63-96); static source has P = 1 (SonicSim_moving.py:47-61).
struct Source {
    const float* x;        // dry waveform, N samples
    const float* rir;      // (P, C, L) row-major
    float* out;            // (C, N) row-major, channel-major like torchaudio.save expects
    const int* bounds;     // compact trajectory: S+1 cumulative segment bounds (mode 1), else null
    const int* idx;        // per-sample interp_index  (mode 2), else null
    const float* w;        // per-sample interp_weight (mode 2), else null
    float2* hspec;         // scratch: P*C*K half spectra of the RIR partitions, pre-scaled by 1/F
    float2* xspec;         // scratch: nb half spectra of the dry windows
    struct RItem* items;   // scratch: this source's render work items (render_ctas() of them)
    int N, P, C, L;
    int K;                 // RIR partitions = ceil(L / kB)
    int nb;                // output blocks = ceil(N / kB)
    int mode;              // 0 static, 1 moving (bounds), 2 moving (idx, w)
    int pad_[1];
};

// One unit of k_render work: output block b of channel c (static: channels c, c+1), written by
// k_prepare so that k_render never searches trajectories or prefix tables.
struct RItem {
    const float2* X;       // dry spectrum of block b
    const float2* H0;      // RIR spectrum of (position 0, channel c, partition 0); position p at H0 + p * pstride
    float* row;            // output row of channel c
    int pstride;           // float2 words between positions (moving: C*K*kSpec) / to channel c+1 (static: K*kSpec)
    int si, b, c;
    int p_lo, p_hi;        // transforms p = p_lo, p_lo + 2, ... <= p_hi   (static: 0, 0)
    int flags;             // bit 0: static source; bit 1: (static) channel c+1 exists
    int pad_[2];
};
static_assert(sizeof(RItem) == 64, "RItem must be 64 bytes (copied as 4 x 16 B)");

// One inverse transform, published by the CTA's thread 0 to the other threads through shared memory.
struct XDesc {
    const float2* X;       // global pointers, used for RIR partitions >= 1 (partition 0 is staged by TMA)
    const float2* Hp;
    const float2* Hq;      // null: only one filter packed into this transform
    float* row;
    int si, b, c, p;
    int first;             // first transform of its block: plain store, otherwise read-add-store
    int valid;
    int p_lo;              // first position of the block (start of the per-thread segment walk)
    int pad_;
};
static_assert(sizeof(XDesc) == 56 || sizeof(XDesc) == 64, "XDesc layout");

SS_HD int range_ctas(const Source& s) { return (s.nb + 7) >> 3; }       // k_prepare: one warp per block

// fill the work items of block b once its position range is known
SS_HD void fill_items(const Source& s, int si, int b, int p_lo, int p_hi, int lane, int nlanes) {
    if (s.mode == MODE_STATIC) {
        const int ncp = (s.C + 1) >> 1;
        for (int cp = lane; cp < ncp; cp += nlanes) {
            RItem it;
            it.X = s.xspec + (size_t)b * kSpec;
            it.H0 = s.hspec + (size_t)(2 * cp) * s.K * kSpec;
            it.row = s.out + (size_t)(2 * cp) * s.N;
            it.pstride = s.K * kSpec;
            it.si = si; it.b = b; it.c = 2 * cp; it.p_lo = 0; it.p_hi = 0;
            it.flags = 1 | ((2 * cp + 1 < s.C) ? 2 : 0);
            it.pad_[0] = 0; it.pad_[1] = 0;
            s.items[(size_t)b * ncp + cp] = it;
        }
    } else {
        for (int c = lane; c < s.C; c += nlanes) {
            RItem it;
            it.X = s.xspec + (size_t)b * kSpec;
            it.H0 = s.hspec + (size_t)c * s.K * kSpec;
            it.row = s.out + (size_t)c * s.N;
            it.pstride = s.C * s.K * kSpec;
            it.si = si; it.b = b; it.c = c; it.p_lo = p_lo; it.p_hi = p_hi;
            it.flags = 0;
            it.pad_[0] = 0; it.pad_[1] = 0;
            s.items[(size_t)b * s.C + c] = it;
        }
    }
}

// transform (item, p) -> descriptor
SS_HD XDesc make_xdesc(const RItem& it, int p) {
    XDesc d;
    d.X = it.X;
    d.Hp = it.H0 + (size_t)p * it.pstride;
    const bool has_q = (it.flags & 1) ? ((it.flags & 2) != 0) : (p + 1 <= it.p_hi);
    d.Hq = has_q ? d.Hp + it.pstride : nullptr;
    d.row = it.row;
    d.si = it.si; d.b = it.b; d.c = it.c; d.p = p;
    d.first = (p == it.p_lo);
    d.valid = 1;
    d.p_lo = it.p_lo;
    d.pad_ = 0;
    return d;
}

SS_HD int spectra_pairs_h(const Source& s) { return (s.P * s.C * s.K + 1) >> 1; }
SS_HD int spectra_pairs_x(const Source& s) { return (s.nb + 1) >> 1; }
SS_HD int render_ctas(const Source& s) {
    return s.mode == MODE_STATIC ? s.nb * ((s.C + 1) >> 1) : s.nb * s.C;
}

// largest i with prefix[i] <= v, prefix[0] = 0, prefix has n+1 entries
SS_HD int find_source(const int* prefix, int n, int v) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (prefix[mid] <= v) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// =======================================================================================
// Spectra kernel: two real rows (a, b) -> one complex FFT -> two stored half spectra.
// A "row" is 8192 samples  a[n] = (0 <= g0 + n < len && n < ncap) ? base[g0 + n] : 0.
// =======================================================================================
struct Row {
    const float* base;
    float2* dst;           // null -> row absent (odd count)
    int g0, len, ncap;
    float scale;
};

SS_HD Row make_row_h(const Source& s, int row) {
    Row r;
    if (row >= s.P * s.C * s.K) { r.base = nullptr; r.dst = nullptr; r.g0 = 0; r.len = 0; r.ncap = 0; r.scale = 0.f; return r; }
    int part = row % s.K, pc = row / s.K;
    r.base = s.rir + (size_t)pc * s.L;
    r.dst = s.hspec + (size_t)row * kSpec;
    r.g0 = part * kB; r.len = s.L; r.ncap = kB;
    r.scale = 1.0f / (float)kF;
    return r;
}
SS_HD Row make_row_x(const Source& s, int blk) {
    Row r;
    if (blk >= s.nb) { r.base = nullptr; r.dst = nullptr; r.g0 = 0; r.len = 0; r.ncap = 0; r.scale = 0.f; return r; }
    r.base = s.x;
    r.dst = s.xspec + (size_t)blk * kSpec;
    r.g0 = (blk - 1) * kB; r.len = s.N; r.ncap = kF;
    r.scale = 1.0f;
    return r;
}
SS_HD float row_at(const Row& r, int n) {
    int g = r.g0 + n;
    return (r.dst != nullptr && n < r.ncap && g >= 0 && g < r.len) ? r.base[g] : 0.f;
}

struct Regs32 { float2 a[16]; float2 b[16]; };

// S1: pass A of the forward transform straight from global memory.
SS_HD void spectra_phase1(int t, const Row& ra, const Row& rb, float2* s) {
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        int j = t + 256 * h;
        float2 v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = make_float2(row_at(ra, j + 512 * r), row_at(rb, j + 512 * r));
        fft16<false>(v);
        passA_store(s, j, v);
    }
}
// generic "load both butterflies t and t+256":  pad(t + 256) = pad(t) + 272
SS_HD void load2(int t, const float2* s, Regs32& R) {
    const float2* p = s + pad(t);
#pragma unroll
    for (int r = 0; r < 16; ++r) { R.a[r] = p[544 * r]; R.b[r] = p[544 * r + 272]; }
}

template <bool INV>
SS_HD void passB2(int t, float2* s, Regs32& R, const Tables& T) {
    passB_compute<INV>(t, R.a, T);       passB_store(s, t, R.a);
    passB_compute<INV>(t + 256, R.b, T); passB_store(s, t + 256, R.b);
}

// S3: pass C + closing radix-2, result (natural order) left in R: a[slot] = Zf[t + 256 r],
// b[slot] = Zf[4096 + t + 256 r], slot = out16(r).
SS_HD void spectra_phase3_compute(int t, Regs32& R, const Tables& T) {
    passC_compute<false>(t, R.a, T);
    passC_compute<false>(t + 256, R.b, T);
    float2 wt = ldg_cached(T.tw + t);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int sl = out16(r);
        float2 tmp = cmul(R.b[sl], final_twiddle<false>(t, r, wt));
        float2 lo = R.a[sl];
        R.a[sl] = cadd(lo, tmp);
        R.b[sl] = csub(lo, tmp);
    }
}
// natural order to shared: pad(t + 256 r) = pad(t) + 272 r, pad(4096 + i) = 4352 + pad(i)
SS_HD void spectra_phase3_store(int t, float2* s, const Regs32& R) {
    float2* d = s + pad(t);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        d[272 * r] = R.a[out16(r)];
        d[272 * r + 4352] = R.b[out16(r)];
    }
}
// S4: split Zf into the spectra of a and b (Hermitian parts), scale, store.
SS_HD void spectra_phase4(int t, const float2* s, const Row& ra, const Row& rb) {
#pragma unroll 4
    for (int m = 0; m < 16; ++m) {
        int k = t + 256 * m;
        float2 A, Bv;
        if (k == 0) {
            float2 z0 = s[pad(0)], zn = s[pad(4096)];
            A = make_float2(z0.x, zn.x);
            Bv = make_float2(z0.y, zn.y);
        } else {
            float2 zk = s[pad(k)], zm = s[pad(kF - k)];
            A = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
            Bv = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
        }
        if (ra.dst) ra.dst[k] = make_float2(A.x * ra.scale, A.y * ra.scale);
        if (rb.dst) rb.dst[k] = make_float2(Bv.x * rb.scale, Bv.y * rb.scale);
    }
}

// =======================================================================================
// Render kernel (inverse transform).  One CTA = one output block b of one channel (moving) or
// one channel pair (static).
// =======================================================================================

// Z formation fused with pass A.  sX / sHp / sHq: partition 0 of the dry window and of the two
// real filters packed into this transform, staged in shared memory (linear, 4096 words each) by the
// bulk-copy engine; sHq may be null.  RIR partitions j >= 1 (long RIRs) pair with dry window b - j
// and are streamed from global memory through the pointers in `d`.
SS_HD void form_z(int t, const float2* sX, const float2* sHp, const float2* sHq, const XDesc& d, int K, Regs32& R) {
    const int jB = passA_jB(t);
    {
        const float2 *xa_p = sX + t, *xb_p = sX + jB, *ha_p = sHp + t, *hb_p = sHp + jB;
        if (sHq) {
            const float2 *ga_p = sHq + t, *gb_p = sHq + jB;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                float2 xa = xa_p[512 * m], xb = xb_p[512 * m];
                R.a[m] = cmul(xa, ha_p[512 * m]);            // P_A[m]
                R.b[m] = cmul(xb, hb_p[512 * m]);            // P_B[m]
                R.b[15 - m] = cmul(xa, ga_p[512 * m]);       // Q_A[m]
                R.a[15 - m] = cmul(xb, gb_p[512 * m]);       // Q_B[m]
            }
        } else {
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                float2 xa = xa_p[512 * m], xb = xb_p[512 * m];
                R.a[m] = cmul(xa, ha_p[512 * m]);
                R.b[m] = cmul(xb, hb_p[512 * m]);
                R.b[15 - m] = make_float2(0.f, 0.f);
                R.a[15 - m] = make_float2(0.f, 0.f);
            }
        }
    }
    const int kparts = (d.b + 1 < K) ? d.b + 1 : K;
    for (int part = 1; part < kparts; ++part) {
        const float2* xa_p = d.X - (size_t)part * kSpec + t;
        const float2* xb_p = d.X - (size_t)part * kSpec + jB;
        const float2* ha_p = d.Hp + (size_t)part * kSpec + t;
        const float2* hb_p = d.Hp + (size_t)part * kSpec + jB;
        const float2* ga_p = d.Hq ? d.Hq + (size_t)part * kSpec + t : nullptr;
        const float2* gb_p = d.Hq ? d.Hq + (size_t)part * kSpec + jB : nullptr;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            float2 xa = ldg_stream(xa_p + 512 * m), xb = ldg_stream(xb_p + 512 * m);
            cmac(R.a[m], xa, ldg_stream(ha_p + 512 * m));
            cmac(R.b[m], xb, ldg_stream(hb_p + 512 * m));
            if (ga_p) {
                cmac(R.b[15 - m], xa, ldg_stream(ga_p + 512 * m));
                cmac(R.a[15 - m], xb, ldg_stream(gb_p + 512 * m));
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        float2 PA = R.a[m], QA = R.b[15 - m], PB = R.b[m], QB = R.a[15 - m];
        R.a[m] = z_direct(PA, QA);  R.b[15 - m] = z_mirror(PA, QA);
        R.b[m] = z_direct(PB, QB);  R.a[15 - m] = z_mirror(PB, QB);
    }
    if (t == 0) {
        // thread 0 owns the self-mirrored butterflies 0 and 256 and the packed (DC, Nyquist) word
        float2 x0 = sX[0], h0 = sHp[0];
        float pdc = x0.x * h0.x, pny = x0.y * h0.y, qdc = 0.f, qny = 0.f;
        if (sHq) { float2 g0 = sHq[0]; qdc = x0.x * g0.x; qny = x0.y * g0.y; }
        for (int part = 1; part < kparts; ++part) {
            x0 = (d.X - (size_t)part * kSpec)[0];
            h0 = (d.Hp + (size_t)part * kSpec)[0];
            pdc += x0.x * h0.x; pny += x0.y * h0.y;
            if (d.Hq) { float2 g0 = (d.Hq + (size_t)part * kSpec)[0]; qdc += x0.x * g0.x; qny += x0.y * g0.y; }
        }
        float2 tmp[8];
#pragma unroll
        for (int r = 8; r < 16; ++r) tmp[r - 8] = R.a[r];
#pragma unroll
        for (int r = 9; r < 16; ++r) R.a[r] = R.b[r - 1];
#pragma unroll
        for (int r = 8; r < 16; ++r) R.b[r] = tmp[r - 8];
        R.a[8] = make_float2(pny, qny);
        R.a[0] = make_float2(pdc, qdc);
    }
}
SS_HD void render_phase1(int t, float2* s, Regs32& R) {
    fft16<true>(R.a); passA_store(s, passA_jA(t), R.a);
    fft16<true>(R.b); passA_store(s, passA_jB(t), R.b);
}

// pass C + closing radix-2 (second half only = the alias-free overlap-save samples):
// leaves z[4096 + t + 256 r] in R.a[out16(r)].
SS_HD void render_phase3(int t, Regs32& R, const Tables& T) {
    passC_compute<true>(t, R.a, T);
    passC_compute<true>(t + 256, R.b, T);
    float2 wt = dirw<true>(ldg_cached(T.tw + t));
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int sl = out16(r);
        R.a[sl] = csub(R.a[sl], cmul(R.b[sl], final_twiddle<true>(t, r, wt)));
    }
}
// epilogue: the sink is called with r = 0..15 in order; it may carry state from sample to sample
template <class Sink>
SS_HD void render_epilogue(const Regs32& R, Sink& sink) {
#pragma unroll
    for (int r = 0; r < 16; ++r) sink(r, R.a[out16(r)]);
}

// hat functions of positions (p, p+1) at a sample that lies in segment sg with weight w:
//   reference lerp (SonicSim_moving.py:94):  (1 - w) * conv[sg] + w * conv[sg + 1]
SS_HD void hat_pair(int sg, float w, int p, float& fa, float& fb) {
    const float omw = one_minus(w);
    fa = (sg == p) ? omw : ((sg + 1 == p) ? w : 0.f);
    fb = (sg == p + 1) ? omw : ((sg == p) ? w : 0.f);
}

// Output row of one CTA.  `first` = this is the first transform of the block (plain store),
// otherwise the partial result already in `row` is read back and added to (only blocks that
// straddle a waypoint need a second transform), which keeps 16 accumulators out of the register
// file during the FFT passes.
struct MovingSinkBounds {       // compact trajectory (MODE_MOVING_BOUNDS)
    const int* bounds; float* row; int S, N, nbase, p, sg, b0, b1; double step; bool first;
    SS_HD MovingSinkBounds(const Source& s, float* row_, int n0, int t, int p_, int sg0, bool first_)
        : bounds(s.bounds), row(row_), S(s.P - 1), N(s.N), nbase(n0 + t), p(p_), sg(sg0), first(first_) {
        b0 = bounds[sg]; b1 = bounds[sg + 1]; step = 1.0 / (double)(b1 - b0);
    }
    SS_HD void operator()(int r, float2 z) {
        int n = nbase + 256 * r;
        if (n >= N) return;
        if (n >= b1) {
            do { ++sg; b1 = bounds[sg + 1]; } while (n >= b1);
            b0 = bounds[sg]; step = 1.0 / (double)(b1 - b0);
        }
        float w = (float)((double)(n - b0) * step);      // == np.linspace(0, 1, num, False)[i] as float32
        float fa, fb;
        hat_pair(sg, w, p, fa, fb);
        float v = lerp_terms(fa, z.x, fb, z.y);
        if (first) row[n] = v; else red_add(row + n, v);
    }
};
struct MovingSinkIndexed {      // per-sample arrays (MODE_MOVING_INDEXED)
    const int* idx; const float* w; float* row; int N, nbase, p; bool first;
    SS_HD MovingSinkIndexed(const Source& s, float* row_, int n0, int t, int p_, bool first_)
        : idx(s.idx), w(s.w), row(row_), N(s.N), nbase(n0 + t), p(p_), first(first_) {}
    SS_HD void operator()(int r, float2 z) {
        int n = nbase + 256 * r;
        if (n >= N) return;
        float fa, fb;
        hat_pair(idx[n], w[n], p, fa, fb);
        float v = lerp_terms(fa, z.x, fb, z.y);
        if (first) row[n] = v; else red_add(row + n, v);
    }
};
struct StaticSink {             // Re -> channel c0, Im -> channel c1 (row1 null if C is odd)
    float* row0; float* row1; int N, nbase;
    SS_HD void operator()(int r, float2 z) {
        int n = nbase + 256 * r;
        if (n >= N) return;
        row0[n] = z.x;
        if (row1) row1[n] = z.y;
    }
};

// min / max interp_index over this thread's 16 samples (indexed mode)
SS_HD void idx_range(int t, int n0, const Source& s, int& pmin, int& pmax) {
    pmin = 0x7fffffff; pmax = -1;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        int n = n0 + t + 256 * r;
        if (n < s.N) { int sg = s.idx[n]; pmin = sg < pmin ? sg : pmin; pmax = sg > pmax ? sg : pmax; }
    }
}

}  // namespace ss
