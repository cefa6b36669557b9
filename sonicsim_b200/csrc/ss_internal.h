// sonicsim_b200 :: ss_internal.h - host-side state shared by the translation units of the library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#include "../../include/sonicsim_b200.h"

struct ss_ctx {
    int device = 0;
    int sm_count = 148;
    int64_t chunk_bytes = 96ll << 20;        // device path: launch granularity (bigger = better amortised)
    int64_t chunk_bytes_host = 28ll << 20;   // host path: copy/compute/copy pipeline granularity (PCIe-bound)
    // scratch for spectra
    // kAux scratch buffers: consecutive chunks of ss_render_dev rotate through them (and through as many
    // internal streams) so that chunk i+1's descriptor copy + k_prepare overlap chunk i's k_render tail
#ifndef SS_AUX_STREAMS
#define SS_AUX_STREAMS 3
#endif
    static const int kAux = SS_AUX_STREAMS;
    char* d_scratch[kAux] = {}; size_t scratch_cap[kAux] = {};
    cudaStream_t s_aux[kAux] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[kAux] = {};
    // descriptor ring (pinned host + device)
    static const int kRing = 2 * SS_AUX_STREAMS;      // multiple of kAux: a slot is always reused on the same stream
    char* h_desc[kRing] = {};
    char* d_desc[kRing] = {};
    size_t desc_cap[kRing] = {};
    cudaEvent_t desc_ev[kRing] = {};
    int ring_pos = 0;
    // host path
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
    struct Slot { char* d_in = nullptr; size_t in_cap = 0; char* d_out = nullptr; size_t out_cap = 0;
                  cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_free = nullptr, ev_rend = nullptr; } slot[4];
    // loudness post-processing of the host path: pinned block (results | gating tables) and its device twin
    char* h_post = nullptr; char* d_post = nullptr; size_t h_post_cap = 0;
    static const int kSlots = 4;   // H2D may run up to 3 chunks ahead of the D2H that frees a slot
    int64_t launches = 0;
    bool single_stream = false;   // experiment knob (SS_SINGLE_STREAM=1): no chunk overlap
    bool no_fast = false;         // experiment knob (SS_NO_FAST=1): never pick the all-aligned kernel variant
    bool no_graph = false;        // experiment knob (SS_NO_GRAPH=1): plans enqueue their launches instead of a CUDA graph
    // optional per-kernel timing (CUDA events on the launching stream)
    bool profiling = false;
    struct Prof { cudaEvent_t e0, e1, e2; int chunk; };   // chunk = position of the launch pair inside its render call
    int prof_chunk = 0;
    std::vector<Prof> prof;
};

extern thread_local int g_last_cuda;
#define CK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { g_last_cuda = (int)e_; \
    return e_ == cudaErrorMemoryAllocation ? SS_ERR_NOMEM : SS_ERR_CUDA; } } while (0)


inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Take the next slot of the descriptor ring (pinned host block + its device twin), grown to `bytes`.
// Blocks the host until the slot's previous host->device copy has completed.  The caller fills `*h`,
// copies it to `*d` on its stream and records c->desc_ev[*slot] right after the copy.
inline int ring_acquire(ss_ctx* c, size_t bytes, int* slot_out, char** h, char** d) {
    const int slot = c->ring_pos; c->ring_pos = (c->ring_pos + 1) % ss_ctx::kRing;
    CK(cudaEventSynchronize(c->desc_ev[slot]));
    if (bytes > c->desc_cap[slot]) {
        if (c->h_desc[slot]) CK(cudaFreeHost(c->h_desc[slot]));
        if (c->d_desc[slot]) { CK(cudaDeviceSynchronize()); CK(cudaFree(c->d_desc[slot])); }
        c->h_desc[slot] = nullptr; c->d_desc[slot] = nullptr; c->desc_cap[slot] = 0;
        const size_t cap = align_up(bytes * 2, 4096);
        CK(cudaHostAlloc((void**)&c->h_desc[slot], cap, cudaHostAllocDefault));
        CK(cudaMalloc((void**)&c->d_desc[slot], cap));
        c->desc_cap[slot] = cap;
    }
    *slot_out = slot; *h = c->h_desc[slot]; *d = c->d_desc[slot];
    return SS_OK;
}
