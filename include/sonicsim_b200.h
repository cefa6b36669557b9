/* sonicsim_b200.h - C ABI of the B200-native moving-source renderer.
 *
 * The reference (JusperLee/SonicSim) has no FFI layer: its hot path is module-level Python
 * (SonicSim-SonicSet/SonicSim_moving.py, SonicSim_audio.py) called positionally from
 * SonicSet.py:77-101.  This header is the native boundary a maintainer binds instead (ctypes stub
 * in INTEGRATION.md; sonicsim_b200/_lib.py is that stub).  Each entry point names the reference
 * function(s) it replaces.  Plain pointers and sizes only; no torch / Python types.
 *
 * Conventions
 *   - float32 everywhere, row-major.  dry x: (N,), RIRs: (P, C, L), output: (C, N) channel-major.
 *   - `*_dev` entry points take DEVICE pointers and enqueue work on `stream` (a cudaStream_t cast to
 *     void*; NULL = legacy default stream) without synchronising.
 *   - `*_host` entry points take HOST pointers (pinned memory gives full PCIe rate), copy in,
 *     render and copy out; they return after the results are in host memory.
 *   - an ss_ctx is NOT thread-safe: use one context per host thread (one process per GPU is the
 *     intended deployment); calls on one context are serialised by the caller.  The context's scratch and
 *     descriptor buffers are ordered against earlier work only through the stream a call is given: give
 *     all `*_dev` calls of one context the same stream, or synchronise between calls that use different ones.
 *   - every function returns 0 (SS_OK) or a negative ss_status; ss_strerror() describes it.  The
 *     Python shim turns these into the exception types the reference raises (IndexError /
 *     ValueError), see INTEGRATION.md.
 */
#ifndef SONICSIM_B200_H_
#define SONICSIM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ss_ctx ss_ctx;

typedef enum {
    SS_OK = 0,
    SS_ERR_INVALID = -1,      /* bad argument (null pointer, non-positive size, unknown mode)       */
    SS_ERR_INDEX = -2,        /* trajectory refers to a position >= P - 1 (reference: IndexError)   */
    SS_ERR_CUDA = -3,         /* CUDA runtime error; ss_last_cuda_error() has the code              */
    SS_ERR_NOMEM = -4,        /* device or pinned allocation failed                                 */
    SS_ERR_UNSUPPORTED = -5   /* shape outside what this build handles                              */
} ss_status;

/* how one source's trajectory is given */
typedef enum {
    SS_STATIC = 0,            /* P == 1, no interpolation: convolve_fixed_receiver                  */
    SS_MOVING_BOUNDS = 1,     /* `bounds`: P int32 cumulative segment bounds, bounds[0]=0, [P-1]=N   */
    SS_MOVING_INDEXED = 2     /* `idx` (int32, N) and `w` (float32, N): interp_index / interp_weight */
} ss_mode;

/* per-source options (ss_source.flags) */
typedef enum {
    SS_RIR_NORMALIZE = 1      /* `rir` is the raw simulator output, clipped and stacked: divide every tap by the global
                               * abs-max of the source's whole (P, C, L) tensor before convolving, i.e. the
                               * `ir_output /= ir_output.abs().max()` of generate_rir_combination
                               * (SonicSim_audio.py:391-398), fused into the spectra kernel's row loads (one reduction
                               * kernel + an IEEE division per tap; same bits as normalising first) */
} ss_source_flags;

/* One (utterance, source) unit of work = one call of convolve_moving_receiver /
 * convolve_fixed_receiver in the reference.  Pointers are device pointers for ss_render_dev and
 * host pointers for ss_render_host. */
typedef struct {
    const float* x;           /* (N,)       dry source, SonicSim_moving.py:64                        */
    const float* rir;         /* (P, C, L)  per-position RIRs, SonicSim_moving.py:65; (C, L) if static */
    float* out;               /* (C, N)     rendered stem, SonicSim_moving.py:96 / :60               */
    const int32_t* bounds;    /* SS_MOVING_BOUNDS: np.cumsum(samples_per_interval) with leading 0    */
    const int32_t* idx;       /* SS_MOVING_INDEXED: interp_index (SonicSim_moving.py:42)             */
    const float* w;           /* SS_MOVING_INDEXED: interp_weight (SonicSim_moving.py:43)            */
    int32_t N, P, C, L;
    int32_t mode;             /* ss_mode */
    int32_t flags;            /* ss_source_flags, 0 = none */
    const int32_t* bounds_host; /* ss_render_dev only, optional: HOST copy of `bounds` (P ints).  With it the block
                                 * table is built on the host and one small kernel launch per chunk is saved.     */
} ss_source;

/* library / build identification */
int ss_version(void);
const char* ss_strerror(int status);
int ss_last_cuda_error(void);

/* One context per process per GPU: owns the twiddle table, scratch for spectra and the staging
 * buffers / streams of the host path.  `device` is the CUDA ordinal (after CUDA_VISIBLE_DEVICES). */
int ss_create(int device, ss_ctx** out);
void ss_destroy(ss_ctx* ctx);

/* Tunables: scratch budget per launch group (bytes of spectra kept live, sized to stay in L2). */
int ss_set_chunk_bytes(ss_ctx* ctx, int64_t bytes);

/* Replaces SonicSim_moving.convolve_moving_receiver (SonicSim_moving.py:63-96) and
 * convolve_fixed_receiver (:47-61) for a whole batch of sources in two (three without bounds_host) kernel
 * launches per chunk.  Device pointers; asynchronous on `stream`.
 * Preconditions the device path cannot check without a round trip (the host path checks them and returns
 * SS_ERR_INDEX / SS_ERR_INVALID like the reference raises): SS_MOVING_INDEXED: 0 <= idx[n] <= P - 2 for every n
 * (out-of-range samples are rendered as zeros, not reported); SS_MOVING_BOUNDS without bounds_host: bounds is
 * non-decreasing with bounds[0] = 0 and bounds[P - 1] = N (anything else is undefined behaviour). */
int ss_render_dev(ss_ctx* ctx, const ss_source* items, int n_items, void* stream);

/* The device path cannot return SS_ERR_INDEX: kernels that meet a trajectory outside the contract above render the
 * offending samples as silence (clamped positions) and set a bit that this call returns and clears after waiting for
 * the device: bit 0 = an interp_index outside [0, P - 2], bit 1 = a device-side bounds table that is not ascending from
 * 0 to N (checked for grid-blocked sources).  The Python device API raises IndexError / ValueError from it on request
 * (Renderer.check_device_errors). */
int ss_device_errors(ss_ctx* ctx, uint32_t* bits);

/* A batch of device-resident sources bound once (the generation loop of SonicSet.py:180-214 renders the same
 * shapes scene after scene into the same buffers).  ss_plan_create validates, chunks, builds the block tables and
 * keeps the descriptor blocks and the scratch for the spectra resident; ss_plan_run is then kernel launches only,
 * issued as ONE CUDA graph launch on `stream` (captured on the first run; the launches fork over the context's
 * internal streams and join back).  Results are identical to ss_render_dev on the same items.  The pointers in
 * `items` and the values behind bounds_host must stay valid and unchanged while the plan lives; the contents of the
 * x / rir / out buffers may change between runs.  Device-side `bounds` / `idx` are read at run time and must satisfy the
 * same preconditions as for ss_render_dev.  Not thread-safe; destroy plans before their context. */
typedef struct ss_plan ss_plan;
int ss_plan_create(ss_ctx* ctx, const ss_source* items, int n_items, ss_plan** out);
int ss_plan_run(ss_plan* plan, void* stream);
int ss_plan_is_graph(const ss_plan* plan);     /* 1 once the plan runs as an instantiated CUDA graph */
void ss_plan_destroy(ss_plan* plan);

/* Same, host pointers: H2D -> render -> D2H, pipelined over chunks on internal streams.  Returns
 * when every `out` is complete.  This is what the drop-in Python functions call. */
int ss_render_host(ss_ctx* ctx, const ss_source* items, int n_items);

/* Single-source conveniences with the argument order of the reference functions. */
int ss_convolve_fixed_receiver(ss_ctx* ctx, const float* source_audio, const float* rirs, float* out,
                               int32_t N, int32_t C, int32_t L);                        /* :47-61 */
int ss_convolve_moving_receiver(ss_ctx* ctx, const float* source_audio, const float* rirs,
                                const int32_t* interp_index, const float* interp_weight, float* out,
                                int32_t N, int32_t P, int32_t C, int32_t L);            /* :63-96 */

/* ---- loudness: SonicSim_audio.lufs_norm (SonicSim_audio.py:68-81) = pyloudnorm 0.1.1
 * Meter(rate, block_size).integrated_loudness + normalize.loudness.  The gating-block sample
 * bounds are computed by the caller exactly as pyloudnorm does (Python float expressions
 * truncated with int()), and passed as `brk` = the sorted distinct bounds (n_e + 1 values) plus,
 * per gating block j, the range [blk_lo[j], blk_hi[j]) of elementary intervals it covers. */
#define SS_LOUD_SCRATCH_DOUBLES 20
typedef struct {
    const float* data;        /* element (n, c) at data[n * stride_n + c * stride_c]                   */
    float* out;               /* N*C contiguous floats = gain * data (may alias data); NULL = measure  */
    const int32_t* brk;       /* n_e + 1                                                               */
    const int32_t* blk_lo;    /* n_blocks                                                              */
    const int32_t* blk_hi;    /* n_blocks                                                              */
    double* scratch;          /* SS_LOUD_SCRATCH_DOUBLES * C * n_e doubles                             */
    double* result;           /* 2 doubles: integrated loudness (LUFS, may be -inf), linear gain       */
    int64_t stride_n, stride_c;
    int32_t N, C, n_e, n_blocks;
    double rate;              /* sample rate                                                           */
    double block_size;        /* T_g in seconds (0.4, or N / rate for short clips, SonicSim_audio.py:69) */
    double target_lufs;       /* `norm` of lufs_norm                                                   */
} ss_loud_item;

/* Device pointers, asynchronous on `stream`; all items share one sample rate. */
int ss_loudness_dev(ss_ctx* ctx, const ss_loud_item* items, int n_items, void* stream);

/* One stem in host memory (any layout through the strides); returns when `out` (if given), the
 * measured loudness and the linear gain are available. */
int ss_lufs_norm_host(ss_ctx* ctx, const float* data, float* out, int32_t N, int32_t C,
                      int64_t stride_n, int64_t stride_c, double rate, double block_size,
                      double target_lufs, const int32_t* brk, int32_t n_e, const int32_t* blk_lo,
                      const int32_t* blk_hi, int32_t n_blocks, double* loudness, double* gain);

/* Optional per-source post-processing of ss_render_host_ex: measure the rendered (C, N) stem and
 * normalise it to `target_lufs` on the device before it is copied out (SonicSet.py:97-101 applies
 * get_lufs_norm_audio to every stem right after rendering it).  brk == NULL skips the source. */
typedef struct {
    const int32_t* brk; const int32_t* blk_lo; const int32_t* blk_hi;   /* host arrays, see ss_loud_item */
    int32_t n_e, n_blocks;
    double rate, block_size, target_lufs;
    double* result;           /* host, 2 doubles: measured LUFS, linear gain (may be NULL)             */
} ss_post_lufs;

/* ss_render_host + optional loudness normalisation of each stem while it is still in HBM. */
int ss_render_host_ex(ss_ctx* ctx, const ss_source* items, int n_items, const ss_post_lufs* post);

/* ---- mixture assembly of the training dataloader: separation/look2hear/datas/movingdatamodule.py
 * :29-32 (compute_mch_rms_dB) and :105-124 (SIR gains of the interferers, SNR gain of the summed noise,
 * both clamped at +40 dB, sums).  E = elements per stem (channels * samples). */
typedef struct {
    const float* speakers;    /* (S, E)  speaker_wav, speaker 0 is the reference                         */
    const float* noises;      /* (M, E)  noise_wav                                                        */
    const float* sirs;        /* (S - 1) SIR of each interferer in dB (reference: U(-6, 6))              */
    float* mix;               /* (E)     mix_wav                                                          */
    float* speakers_out;      /* (S, E)  speakers after their gains (may alias speakers; NULL = skip)    */
    double* scratch;          /* ss_mix_scratch_doubles() doubles                                        */
    int64_t E;
    int32_t S, M;
    float snr;                /* dB (reference: U(10, 20); enhancement variant U(-10, 15))               */
    int32_t noise_delay;      /* 0, or overlap_audio's shift D in elements applied to the summed noise:
                                 n[e] = (s[e-D] + s[e+D]) + s[e]  (enhancement/.../movingdatamodule.py:34-48,240) */
} ss_mix_item;

int64_t ss_mix_scratch_doubles(void);
int ss_mix_dev(ss_ctx* ctx, const ss_mix_item* items, int n_items, void* stream);      /* device pointers, async */
int ss_mix_host(ss_ctx* ctx, const float* speakers, const float* noises, const float* sirs, float snr,
                float* mix, float* speakers_out, int32_t S, int32_t M, int64_t E);      /* host pointers */
int ss_mix_host_ex(ss_ctx* ctx, const float* speakers, const float* noises, const float* sirs, float snr,
                   float* mix, float* speakers_out, int32_t S, int32_t M, int64_t E, int32_t noise_delay);

/* overlap_audio (enhancement/look2hear/datas/movingdatamodule.py:34-48): y[r][n] = (x[r][n-D] + x[r][n+D]) + x[r][n],
 * zeros outside [0, T); x != y. */
int ss_overlap_dev(ss_ctx* ctx, const float* x, float* y, int32_t rows, int64_t T, int64_t delay, void* stream);
int ss_overlap_host(ss_ctx* ctx, const float* x, float* y, int32_t rows, int64_t T, int64_t delay);

/* ---- dry-stream assembly: the arithmetic of SonicSim_audio.create_long_audio (SonicSim_audio.py:231-279) and
 * create_background_audio (:281-340) once the host has drawn the clips and their places (`random` stream as in the
 * reference): resampling with torchaudio.transforms.Resample's polyphase filter bank (:253-256), stereo -> mono mean
 * (:311-312) and `long_audio[:, a:b] += clip[...]` (:268, :326, :332), one kernel, the stream stays in HBM. */
typedef struct {
    const float* src;         /* (channels, src_len) device: the decoded clip at its own sample rate            */
    const float* kernel_t;    /* (taps, new_rate) device: Resample.kernel[:, 0, :] transposed; NULL = same rate  */
    int64_t dst_start;        /* first sample of the stream written                                             */
    int64_t src_start;        /* first sample of the (resampled, mono) clip used                                 */
    int64_t count;            /* samples added                                                                   */
    int32_t channels;         /* 1, or 2 (averaged after resampling)                                             */
    int32_t src_len;
    int32_t orig, new_rate;   /* sample rates divided by their gcd (Resample.orig_freq / gcd, new_freq / gcd)    */
    int32_t width, taps;      /* Resample.width; taps = 2 * width + orig                                         */
} ss_dry_clip;
/* zero-fills out[0, total) and adds every clip; device pointers, asynchronous on `stream` */
int ss_dry_assemble_dev(ss_ctx* ctx, const ss_dry_clip* clips, int n_clips, float* out, int64_t total, void* stream);

/* Counters since ss_create / ss_reset_stats: kernels launched and device time is NOT measured
 * here (bench.py uses CUDA events); this is the launch count bench.py reports as gpu_launches. */
int64_t ss_launch_count(const ss_ctx* ctx);
void ss_reset_stats(ss_ctx* ctx);

/* Per-kernel device timing with CUDA events recorded on the launching stream around k_spectra and
 * k_render (bench.py's roofline numbers).  ss_get_profile waits for the recorded launches, returns
 * the summed milliseconds of each kernel and the number of (k_spectra, k_render) launch pairs since
 * the previous call, and clears the record.  While profiling is on, chunks run one after the other on the
 * caller's stream behind a 0.3 ms idle kernel (so no interval contains a wait for the host), and the sums
 * are built from the median over calls of each launch position (repeat the same call to use this). */
int ss_set_profiling(ss_ctx* ctx, int on);
int ss_get_profile(ss_ctx* ctx, double* ms_spectra, double* ms_render, int64_t* n_pairs);

/* Test hook, pure host code (no CUDA call): the blocking plan of one source whose trajectory bounds are on the host
 * (`bounds_host`): blocks_out receives (start, len, p_lo, p_hi) per block, *aligned_out whether the blocks follow
 * the trajectory's waypoints (one transform each) or the 4096-sample grid.  Returns the number of blocks, or a
 * negative ss_status (SS_ERR_NOMEM: max_blocks too small). */
int ss_debug_plan(const ss_source* item, int32_t* blocks_out, int32_t max_blocks, int32_t* aligned_out);

/* Test hook, pure host code: how a batch whose items need bytes[i] of scratch is cut into launch groups under `budget`
 * (ss_set_chunk_bytes): cuts_out receives the first item of every chunk followed by n.  Returns the number of values
 * written, or a negative ss_status (SS_ERR_NOMEM: max_cuts too small). */
int ss_debug_chunks(const int64_t* bytes, int32_t n, int64_t budget, int32_t* cuts_out, int32_t max_cuts);

/* pinned host memory helpers (cudaHostAlloc) for callers without torch */
int ss_host_alloc(void** ptr, int64_t bytes);
void ss_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* SONICSIM_B200_H_ */
