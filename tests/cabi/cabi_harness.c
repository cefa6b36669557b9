/* tests/cabi/cabi_harness.c - a plain C program against include/sonicsim_b200.h, the way a maintainer of the
 * reference would bind the library without Python.  TEST INFRASTRUCTURE.
 *
 *   cabi_harness <in.bin> <out.bin>
 * in.bin : int32 N, P, C, L | x[N] f32 | rirs[P*C*L] f32 | idx[N] i32 | w[N] f32 | bounds[P] i32
 * out.bin: moving (indexed) C*N f32 | fixed (rirs[0]) C*N f32 | moving (bounds, via ss_render_host) C*N f32
 * Exercises ss_create, ss_host_alloc, ss_convolve_moving_receiver, ss_convolve_fixed_receiver, ss_render_host,
 * ss_strerror and the error status of an out-of-range trajectory.  Exit code 0 = every call behaved. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sonicsim_b200.h"

#define CHECK(call) do { int st_ = (call); if (st_ != SS_OK) { fprintf(stderr, "%s -> %d (%s)\n", #call, st_, ss_strerror(st_)); return 2; } } while (0)

static int read_all(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n ? 0 : -1; }

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 1; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("in"); return 1; }
    int32_t hdr[4];
    if (read_all(f, hdr, sizeof hdr)) return 1;
    const int32_t N = hdr[0], P = hdr[1], C = hdr[2], L = hdr[3];
    ss_ctx* ctx = NULL;
    CHECK(ss_create(0, &ctx));
    float *x, *h, *w, *out;
    int32_t *idx, *bounds;
    CHECK(ss_host_alloc((void**)&x, sizeof(float) * (size_t)N));              /* pinned: full PCIe rate */
    CHECK(ss_host_alloc((void**)&h, sizeof(float) * (size_t)P * C * L));
    CHECK(ss_host_alloc((void**)&w, sizeof(float) * (size_t)N));
    CHECK(ss_host_alloc((void**)&idx, sizeof(int32_t) * (size_t)N));
    CHECK(ss_host_alloc((void**)&bounds, sizeof(int32_t) * (size_t)P));
    CHECK(ss_host_alloc((void**)&out, sizeof(float) * (size_t)C * N * 3));
    if (read_all(f, x, sizeof(float) * (size_t)N) || read_all(f, h, sizeof(float) * (size_t)P * C * L) ||
        read_all(f, idx, sizeof(int32_t) * (size_t)N) || read_all(f, w, sizeof(float) * (size_t)N) ||
        read_all(f, bounds, sizeof(int32_t) * (size_t)P)) { fprintf(stderr, "short input\n"); return 1; }
    fclose(f);
    const size_t cn = (size_t)C * N;
    /* SonicSim_moving.convolve_moving_receiver(source_audio, rirs, interp_index, interp_weight), SonicSim_moving.py:63-96 */
    CHECK(ss_convolve_moving_receiver(ctx, x, h, idx, w, out, N, P, C, L));
    /* SonicSim_moving.convolve_fixed_receiver(source_audio, rirs[0]), :47-61 */
    CHECK(ss_convolve_fixed_receiver(ctx, x, h, out + cn, N, C, L));
    /* the batch entry point with the compact trajectory */
    ss_source item;
    memset(&item, 0, sizeof item);
    item.x = x; item.rir = h; item.out = out + 2 * cn; item.bounds = bounds;
    item.N = N; item.P = P; item.C = C; item.L = L; item.mode = SS_MOVING_BOUNDS;
    CHECK(ss_render_host(ctx, &item, 1));
    /* error path: an index that needs position P (the reference raises IndexError) */
    {
        const int32_t keep = idx[N / 2];
        idx[N / 2] = P - 1;
        const int st = ss_convolve_moving_receiver(ctx, x, h, idx, w, out, N, P, C, L);
        idx[N / 2] = keep;
        if (st != SS_ERR_INDEX) { fprintf(stderr, "expected SS_ERR_INDEX, got %d\n", st); return 3; }
        CHECK(ss_convolve_moving_receiver(ctx, x, h, idx, w, out, N, P, C, L));   /* and the context still works */
    }
    if (ss_convolve_fixed_receiver(ctx, NULL, h, out, N, C, L) != SS_ERR_INVALID) { fprintf(stderr, "expected SS_ERR_INVALID\n"); return 3; }
    f = fopen(argv[2], "wb");
    if (!f || fwrite(out, sizeof(float), 3 * cn, f) != 3 * cn) { perror("out"); return 1; }
    fclose(f);
    printf("cabi_harness ok: version %d, %lld launches\n", ss_version(), (long long)ss_launch_count(ctx));
    ss_host_free(x); ss_host_free(h); ss_host_free(w); ss_host_free(idx); ss_host_free(bounds); ss_host_free(out);
    ss_destroy(ctx);
    return 0;
}
