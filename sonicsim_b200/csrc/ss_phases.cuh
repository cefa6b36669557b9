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

// One output block: samples [start, start + len), len <= 4096, rendered by one transform per pair of
// positions in [p_lo, p_hi].  len == 0: unused slot of the table.
//   grid blocking     start = 4096 b                     (long RIRs, indexed trajectories, static sources)
//   aligned blocking  blocks never cross a waypoint: every block lies in one segment s and needs
//                     exactly the pair (s, s + 1) -> one transform, both packed results used
//                     (bounds trajectory and L <= 4096)
struct Block { int start, len, p_lo, p_hi; };

struct RItem;

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
    float2* xspec;         // scratch: nblk_max half spectra of the dry windows
    Block* blocks;         // scratch: nblk_max blocks (k_blocks)
    double* rstep;         // scratch: 1 / (samples in segment s), s < P - 1 (mode 1)
    int* counts;           // scratch: [0] = blocks in use, [1] = index of this source's first render item
    const float* norm_part;// scratch: kNormParts partial abs-maxima of the RIR tensor (k_rir_absmax), or null: taps used as given
    int N, P, C, L;
    int K;                 // RIR partitions = ceil(L / kB)
    int nb;                // ceil(N / kB)
    int mode;              // 0 static, 1 moving (bounds), 2 moving (idx, w)
    int aligned;           // 1: blocks aligned to the trajectory segments (mode 1 and K == 1)
    int nblk_max;          // table size: nb (grid) or nb + P - 1 (aligned)
    int pad_[1];
};
constexpr int kNormParts = 128;    // CTAs (partial maxima) per normalised source in k_rir_absmax

// One unit of k_render work: one block of channel c (static: channels c, c+1), written by k_prepare
// so that k_render never searches trajectories, prefix tables or the Source array.
struct RItem {
    const float2* X;       // dry spectrum of the block's window [start - 4096, start + 4096)
    const float2* H0;      // RIR spectrum of (position 0, channel c, partition 0); position p at H0 + p * pstride
    float* row;            // output row of channel c
    float* row1;           // static source: output row of channel c + 1, or null
    const int* bounds;     // trajectory of the source (mode 1)
    const double* rstep;   //   "
    const int* idx;        // (mode 2)
    const float* w;        // (mode 2)
    int pstride;           // float2 words between positions (moving: C*K*kSpec) / to channel c+1 (static: K*kSpec)
    int n0, n_end;         // output samples [n0, n_end)
    int p_lo, p_hi;        // transforms p = p_lo, p_lo + 2, ... <= p_hi   (static: 0, 0)
    int mode;
    int kparts;            // RIR partitions that reach back into the signal: min(K, n0 / 4096 + 1)
    int b0;                // aligned blocking: first sample of the block's segment, bounds[p_lo] ...
    double step;           // ... and 1 / (samples in that segment), so that k_render_fast's output stage loads no table
    int pad_[2];
};
static_assert(sizeof(RItem) == 112, "RItem is copied as 7 x 16 B");

// One inverse transform, published by the CTA's thread 0 to the other threads through shared memory.
struct XDesc {
    const float2* X;       // global pointers, used for RIR partitions >= 1 (partition 0 is staged by TMA)
    const float2* Hp;
    const float2* Hq;      // null: only one filter packed into this transform
    float* row;
    float* row1;
    const int* bounds;
    const double* rstep;
    const int* idx;
    const float* w;
    int n0, n_end, p, p_lo, p_hi;
    int first;             // first transform of its block: plain store, otherwise accumulate
    int valid;
    int mode;
    int kparts;
    int pad_;
};
static_assert(sizeof(XDesc) == 112, "XDesc layout");

SS_HD int spectra_pairs_h(const Source& s) { return (s.P * s.C * s.K + 1) >> 1; }
SS_HD int spectra_pairs_x(const Source& s) { return (s.nblk_max + 1) >> 1; }
SS_HD int range_ctas(const Source& s) { return (s.nblk_max + 7) >> 3; }       // k_prepare: one warp per block
SS_HD int items_per_block(const Source& s) { return s.mode == MODE_STATIC ? (s.C + 1) >> 1 : s.C; }
SS_HD int max_render_items(const Source& s) { return s.nblk_max * items_per_block(s); }

// blocks of one segment under aligned blocking
SS_HD int seg_blocks(int n_s) { return (n_s + kB - 1) / kB; }

// fill the work items of block `blk` once its position range is known (lanes split the channels)
SS_HD void fill_items(const Source& s, RItem* items, int blk, const Block& bk, int lane, int nlanes) {
    RItem it;
    it.X = s.xspec + (size_t)blk * kSpec;
    it.bounds = s.bounds; it.rstep = s.rstep; it.idx = s.idx; it.w = s.w;
    it.n0 = bk.start; it.n_end = bk.start + bk.len;
    it.mode = s.mode;
    const int reach = bk.start / kB + 1;
    it.kparts = reach < s.K ? reach : s.K;
    it.b0 = 0; it.step = 0.0; it.pad_[0] = 0; it.pad_[1] = 0;
    if (s.aligned) { it.b0 = s.bounds[bk.p_lo]; it.step = s.rstep[bk.p_lo]; }
    const int per = items_per_block(s);
    RItem* dst = items + s.counts[1] + (size_t)blk * per;
    if (s.mode == MODE_STATIC) {
        for (int cp = lane; cp < per; cp += nlanes) {
            it.H0 = s.hspec + (size_t)(2 * cp) * s.K * kSpec;
            it.row = s.out + (size_t)(2 * cp) * s.N;
            it.row1 = (2 * cp + 1 < s.C) ? it.row + s.N : nullptr;
            it.pstride = s.K * kSpec;
            it.p_lo = 0; it.p_hi = 0;
            dst[cp] = it;
        }
    } else {
        for (int c = lane; c < per; c += nlanes) {
            it.H0 = s.hspec + (size_t)c * s.K * kSpec;
            it.row = s.out + (size_t)c * s.N;
            it.row1 = nullptr;
            it.pstride = s.C * s.K * kSpec;
            it.p_lo = bk.p_lo; it.p_hi = bk.p_hi;
            dst[c] = it;
        }
    }
}

// k_render_fast: the two filter spectra of an item's single transform.  Aligned moving source: positions p_lo, p_lo + 1;
// static source: channels c, c + 1 - a static item without a partner channel reuses the first filter (its imaginary
// output is not stored).
SS_HD const float2* item_hp(const RItem& it) { return it.H0 + (size_t)it.p_lo * it.pstride; }
SS_HD const float2* item_hq(const RItem& it) {
    const bool has = it.mode == MODE_STATIC ? it.row1 != nullptr : true;
    return item_hp(it) + (has ? it.pstride : 0);
}

// transform (item, p) -> descriptor
SS_HD XDesc make_xdesc(const RItem& it, int p) {
    XDesc d;
    d.X = it.X;
    d.Hp = it.H0 + (size_t)p * it.pstride;
    const bool has_q = (it.mode == MODE_STATIC) ? (it.row1 != nullptr) : (p + 1 <= it.p_hi);
    d.Hq = has_q ? d.Hp + it.pstride : nullptr;
    d.row = it.row; d.row1 = it.row1;
    d.bounds = it.bounds; d.rstep = it.rstep; d.idx = it.idx; d.w = it.w;
    d.n0 = it.n0; d.n_end = it.n_end; d.p = p; d.p_lo = it.p_lo; d.p_hi = it.p_hi; d.pad_ = 0;
    d.first = (p == it.p_lo);
    d.valid = 1;
    d.mode = it.mode;
    d.kparts = it.kparts;
    return d;
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
    const float* src;      // base + g0: sample n of the row is src[n] for n in [nlo, nhi), zero elsewhere
    float2* dst;           // null -> row absent (odd count / unused block slot)
    int nlo, nspan;        // nspan = nhi - nlo (0: empty)
    float scale;
    float div;             // 0: samples as loaded; else every sample is divided by it first (SS_RIR_NORMALIZE)
};
SS_HD Row make_row(const float* base, float2* dst, int g0, int len, int ncap, float scale) {
    Row r;
    r.dst = dst; r.scale = scale; r.div = 0.f;
    int nlo = g0 < 0 ? -g0 : 0;
    int nhi = len - g0 < ncap ? len - g0 : ncap;
    r.nlo = nlo; r.nspan = nhi > nlo ? nhi - nlo : 0;
    r.src = base + g0;
    return r;
}
SS_HD Row no_row() { Row r; r.src = nullptr; r.dst = nullptr; r.nlo = 0; r.nspan = 0; r.scale = 0.f; r.div = 0.f; return r; }

// global abs-max of a normalised source's RIR tensor from k_rir_absmax's partial maxima
SS_HD float rir_absmax(const Source& s) {
    float m = 0.f;
    for (int i = 0; i < kNormParts; ++i) { const float v = s.norm_part[i]; m = v > m ? v : m; }
    return m;
}
SS_HD Row make_row_h(const Source& s, int row) {
    if (row >= s.P * s.C * s.K) return no_row();
    int part = row % s.K, pc = row / s.K;
    Row r = make_row(s.rir + (size_t)pc * s.L, s.hspec + (size_t)row * kSpec, part * kB, s.L, kB, 1.0f / (float)kF);
    if (s.norm_part) r.div = rir_absmax(s);
    return r;
}
SS_HD Row make_row_x(const Source& s, int blk) {
    if (blk >= s.nblk_max) return no_row();
    const Block bk = s.blocks[blk];
    if (bk.len == 0) return no_row();
    return make_row(s.x, s.xspec + (size_t)blk * kSpec, bk.start - kB, s.N, kF, 1.0f);
}
SS_HD float row_at(const Row& r, int n) {
    return ((unsigned)(n - r.nlo) < (unsigned)r.nspan) ? r.src[n] : 0.f;
}
// IEEE single-precision division, what `ir_output /= ir_output.abs().max()` does per element (SonicSim_audio.py:398)
SS_HD float div_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}

struct Regs32 { float2 a[16]; float2 b[16]; };

// S1: pass A of the forward transform straight from global memory: all 64 loads of the thread's two
// butterflies are issued before the first butterfly is computed (one exposed memory latency, not two).
SS_HD void spectra_phase1(int t, const Row& ra, const Row& rb, float2* s) {
    Regs32 R;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        R.a[r] = make_float2(row_at(ra, t + 512 * r), row_at(rb, t + 512 * r));
        R.b[r] = make_float2(row_at(ra, t + 256 + 512 * r), row_at(rb, t + 256 + 512 * r));
    }
    if (ra.div != 0.f || rb.div != 0.f) {            // CTA-uniform: rows of a source whose RIRs are normalised on the fly
        const float da = ra.div != 0.f ? ra.div : 1.f, db = rb.div != 0.f ? rb.div : 1.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            R.a[r] = make_float2(div_rn(R.a[r].x, da), div_rn(R.a[r].y, db));
            R.b[r] = make_float2(div_rn(R.b[r].x, da), div_rn(R.b[r].y, db));
        }
    }
    fft16<false>(R.a); passA_store(s, t, R.a);
    fft16<false>(R.b); passA_store(s, t + 256, R.b);
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
    float2 u[16];
    final_twiddles<false>(ldg_cached(T.tw + t), u);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int sl = out16(r);
        const float2 lo = R.a[sl], hi = cfma(lo, u[r], R.b[sl]);
        R.a[sl] = hi;
        R.b[sl] = twice_minus(lo, hi);
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

// RIR partitions >= 1 (long RIRs only), streamed from global memory and accumulated into the
// (P_A, P_B, Q_A, Q_B) slots of form_z.  Only instantiated in the LONG variant of k_render, which is
// launched for chunks that contain a source with L > 4096: keeping this loop (or even a call to it) out
// of the common kernel is worth 9 % there (register allocation of the hot loop).
SS_HD void form_z_parts(int t, const XDesc& d, Regs32& R) {
    const int jB = passA_jB(t);
    const int kparts = d.kparts;
    const bool has_q = d.Hq != nullptr;
    for (int part = 1; part < kparts; ++part) {
        const float2* xa_p = d.X - (size_t)part * kSpec + t;
        const float2* xb_p = d.X - (size_t)part * kSpec + jB;
        const float2* ha_p = d.Hp + (size_t)part * kSpec + t;
        const float2* hb_p = d.Hp + (size_t)part * kSpec + jB;
        if (has_q) {
            const float2* ga_p = d.Hq + (size_t)part * kSpec + t;
            const float2* gb_p = d.Hq + (size_t)part * kSpec + jB;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                float2 xa = ldg_stream(xa_p + 512 * m), xb = ldg_stream(xb_p + 512 * m);
                cmac(R.a[m], xa, ldg_stream(ha_p + 512 * m));
                cmac(R.b[m], xb, ldg_stream(hb_p + 512 * m));
                cmac(R.b[15 - m], xa, ldg_stream(ga_p + 512 * m));
                cmac(R.a[15 - m], xb, ldg_stream(gb_p + 512 * m));
            }
        } else {
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                float2 xa = ldg_stream(xa_p + 512 * m), xb = ldg_stream(xb_p + 512 * m);
                cmac(R.a[m], xa, ldg_stream(ha_p + 512 * m));
                cmac(R.b[m], xb, ldg_stream(hb_p + 512 * m));
            }
        }
    }
}

// Z formation fused with pass A.  sX / sHp / sHq: partition 0 of the dry window and of the two
// real filters packed into this transform, staged in shared memory (linear, 4096 words each) by the
// bulk-copy engine; sHq may be null.  RIR partitions j >= 1 (long RIRs) pair with the dry window
// 4096 j samples earlier (grid blocking) and are streamed from global memory through `d`.
template <bool LONG, bool FAST = false>
SS_HD void form_z(int t, const float2* sX, const float2* sHp, const float2* sHq, const XDesc& d, Regs32& R) {
    const int jB = passA_jB(t);
    {
        const float2 *xa_p = sX + t, *xb_p = sX + jB, *ha_p = sHp + t, *hb_p = sHp + jB;
        if (FAST || sHq) {          // FAST: every transform packs two positions (aligned blocking)
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
    const int kparts = LONG ? d.kparts : 1;
    if (LONG) form_z_parts(t, d, R);
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
        if (FAST || sHq) { float2 g0 = sHq[0]; qdc = x0.x * g0.x; qny = x0.y * g0.y; }
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
// ---- long RIRs (LONG): the partitions of a transform pass through the staging buffers one after the other.
// Stage j holds partition j of the two filters and the dry window 4096 j samples earlier (grid blocking);
// stage 0 initialises the product slots (P_A, P_B, Q_A, Q_B) of form_z, later stages accumulate, long_finish
// forms Z.  Thread 0 carries the packed (DC, Nyquist) products of all stages in `e`.
struct DcNy { float pdc, pny, qdc, qny; };
template <bool FIRST>
SS_HD void long_stage(int t, const float2* sX, const float2* sHp, const float2* sHq, Regs32& R, DcNy& e) {
    const int jB = passA_jB(t);
    const float2 *xa_p = sX + t, *xb_p = sX + jB, *ha_p = sHp + t, *hb_p = sHp + jB;
    if (sHq) {
        const float2 *ga_p = sHq + t, *gb_p = sHq + jB;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const float2 xa = xa_p[512 * m], xb = xb_p[512 * m];
            if (FIRST) {
                R.a[m] = cmul(xa, ha_p[512 * m]);       R.b[m] = cmul(xb, hb_p[512 * m]);
                R.b[15 - m] = cmul(xa, ga_p[512 * m]);  R.a[15 - m] = cmul(xb, gb_p[512 * m]);
            } else {
                cmac(R.a[m], xa, ha_p[512 * m]);        cmac(R.b[m], xb, hb_p[512 * m]);
                cmac(R.b[15 - m], xa, ga_p[512 * m]);   cmac(R.a[15 - m], xb, gb_p[512 * m]);
            }
        }
    } else {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const float2 xa = xa_p[512 * m], xb = xb_p[512 * m];
            if (FIRST) {
                R.a[m] = cmul(xa, ha_p[512 * m]);       R.b[m] = cmul(xb, hb_p[512 * m]);
                R.b[15 - m] = make_float2(0.f, 0.f);    R.a[15 - m] = make_float2(0.f, 0.f);
            } else {
                cmac(R.a[m], xa, ha_p[512 * m]);        cmac(R.b[m], xb, hb_p[512 * m]);
            }
        }
    }
    if (t == 0) {
        const float2 x0 = sX[0], h0 = sHp[0];
        if (FIRST) { e.pdc = x0.x * h0.x; e.pny = x0.y * h0.y; e.qdc = 0.f; e.qny = 0.f; }
        else { e.pdc += x0.x * h0.x; e.pny += x0.y * h0.y; }
        if (sHq) {
            const float2 g0 = sHq[0];
            if (FIRST) { e.qdc = x0.x * g0.x; e.qny = x0.y * g0.y; }
            else { e.qdc += x0.x * g0.x; e.qny += x0.y * g0.y; }
        }
    }
}
SS_HD void long_finish(int t, Regs32& R, const DcNy& e) {
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        float2 PA = R.a[m], QA = R.b[15 - m], PB = R.b[m], QB = R.a[15 - m];
        R.a[m] = z_direct(PA, QA);  R.b[15 - m] = z_mirror(PA, QA);
        R.b[m] = z_direct(PB, QB);  R.a[15 - m] = z_mirror(PB, QB);
    }
    if (t == 0) {               // as in form_z: thread 0 owns butterflies 0 and 256 and the packed (DC, Nyquist) word
        float2 tmp[8];
#pragma unroll
        for (int r = 8; r < 16; ++r) tmp[r - 8] = R.a[r];
#pragma unroll
        for (int r = 9; r < 16; ++r) R.a[r] = R.b[r - 1];
#pragma unroll
        for (int r = 8; r < 16; ++r) R.b[r] = tmp[r - 8];
        R.a[8] = make_float2(e.pny, e.qny);
        R.a[0] = make_float2(e.pdc, e.qdc);
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
    float2 u[16];
    final_twiddles<true>(dirw<true>(ldg_cached(T.tw + t)), u);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int sl = out16(r);
        R.a[sl] = cfms(R.a[sl], u[r], R.b[sl]);
    }
}
// the closing radix-2 alone (pass C already done by the caller)
SS_HD void render_phase3_close(int t, Regs32& R, const Tables& T) {
    float2 u[16];
    final_twiddles<true>(dirw<true>(ldg_cached(T.tw + t)), u);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int sl = out16(r);
        R.a[sl] = cfms(R.a[sl], u[r], R.b[sl]);
    }
}
// hat functions of positions (p, p+1) at a sample that lies in segment sg with weight w:
//   reference lerp (SonicSim_moving.py:94):  (1 - w) * conv[sg] + w * conv[sg + 1]
SS_HD void hat_pair(int sg, float w, int p, float& fa, float& fb) {
    const float omw = one_minus(w);
    fa = (sg == p) ? omw : ((sg + 1 == p) ? w : 0.f);
    fb = (sg == p + 1) ? omw : ((sg == p) ? w : 0.f);
}

// Output stage of one thread: samples n = n0 + t + 256 r, r = 0..15, visited in order.  `first` = first
// transform of its block (plain store); the second transform of a block that straddles a waypoint
// (grid blocking only) accumulates with a fire-and-forget RED - two addends per address, the first
// stored by this very thread, so the sum is order-free.
template <bool FAST = false>
SS_HD void render_epilogue(int t, const XDesc& d, const Regs32& R) {
    // `d` lives in shared memory: copy what the loop needs into registers once - after every global
    // store the compiler would otherwise have to reload each field (generic pointers may alias)
    const int nbase = d.n0 + t, n_end = d.n_end, mode = d.mode, p = d.p;
    float* const row = d.row;
    const bool first = d.first != 0;
    if (nbase >= n_end) return;
    if (!FAST && mode == MODE_STATIC) {                      // Re -> channel c, Im -> channel c + 1
        float* const row1 = d.row1;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int n = nbase + 256 * r;
            if (n < n_end) {
                const float2 z = R.a[out16(r)];
                row[n] = z.x;
                if (row1) row1[n] = z.y;
            }
        }
    } else if (FAST || (mode == MODE_MOVING_BOUNDS && d.p_lo + 1 == d.p_hi)) {
        // the block lies inside one segment (always so under aligned blocking): sg = p for every
        // sample, 16 independent weight computations, lerp = (1 - w) Re z + w Im z
        // The 16 weights first, unguarded (no memory access), so that their int -> double -> float chains
        // interleave; (double)(n - b0) = d0 + 256 r exactly.
        const double d0 = (double)(nbase - d.bounds[p]);
        const double step = d.rstep[p];
        float w[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) w[r] = (float)((d0 + 256.0 * r) * step);   // == np.linspace(0, 1, num, False)[i] as float32
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int n = nbase + 256 * r;
            const float2 z = R.a[out16(r)];
            const float v = lerp_terms(one_minus(w[r]), z.x, w[r], z.y);
            if (n < n_end) row[n] = v;
        }
    } else if (mode == MODE_MOVING_BOUNDS) {
        const int* const bounds = d.bounds;
        const double* const rstep = d.rstep;
        int sg = d.p_lo;
        int b1 = bounds[sg + 1];
        while (nbase >= b1) { ++sg; b1 = bounds[sg + 1]; }
        int b0 = bounds[sg];
        double step = rstep[sg];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int n = nbase + 256 * r;
            if (n < n_end) {
                if (n >= b1) {
                    do { ++sg; b1 = bounds[sg + 1]; } while (n >= b1);
                    b0 = bounds[sg]; step = rstep[sg];
                }
                const float w = (float)((double)(n - b0) * step);
                float fa, fb;
                hat_pair(sg, w, p, fa, fb);
                const float2 z = R.a[out16(r)];
                const float v = lerp_terms(fa, z.x, fb, z.y);
                if (first) row[n] = v; else red_add(row + n, v);
            }
        }
    } else {
        const int* const idx = d.idx;
        const float* const wv = d.w;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int n = nbase + 256 * r;
            if (n < n_end) {
                float fa, fb;
                hat_pair(idx[n], wv[n], p, fa, fb);
                const float2 z = R.a[out16(r)];
                const float v = lerp_terms(fa, z.x, fb, z.y);
                if (first) row[n] = v; else red_add(row + n, v);
            }
        }
    }
}

// Output stage of k_render_fast (aligned blocking: the block lies in segment p_lo, one transform per item): as the
// single-segment branch above, with the segment's first sample and 1 / length taken from the work item itself.
SS_HD void render_epilogue_item(int t, const RItem& it, const Regs32& R) {
    const int nbase = it.n0 + t, n_end = it.n_end;
    float* const row = it.row;
    if (nbase >= n_end) return;
    if (it.mode == MODE_STATIC) {                            // Re -> channel c, Im -> channel c + 1 (item-uniform branch)
        float* const row1 = it.row1;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int n = nbase + 256 * r;
            if (n < n_end) {
                const float2 z = R.a[out16(r)];
                row[n] = z.x;
                if (row1) row1[n] = z.y;
            }
        }
        return;
    }
    const double d0 = (double)(nbase - it.b0);
    const double step = it.step;
    float w[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) w[r] = (float)((d0 + 256.0 * r) * step);   // == np.linspace(0, 1, num, False)[i] as float32
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int n = nbase + 256 * r;
        const float v = lerp_pair(w[r], R.a[out16(r)]);
        if (n < n_end) row[n] = v;
    }
}

}  // namespace ss
