/* examples/slab_loop.c -- a slab-decomposed run driven entirely through the C ABI, in plain C, one process
 * per GPU: what a C/C++ host code needs to run BASELINE config c5 on the GPUs of one box (INTEGRATION.md
 * section 5).  No Python, no MPI: the processes are started by any launcher that sets RANK / WORLD_SIZE /
 * LOCAL_RANK (torchrun --no-python does) and trade their CUDA IPC handles through files in a directory.
 *
 *   gcc -std=c99 -I include examples/slab_loop.c -L drstencil_b200 -ldrstencil -Wl,-rpath,$PWD/drstencil_b200 -o slab_loop
 *   python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
 *          --master-port 29612 ./slab_loop stc/3d7pt_star.stc /tmp/rendezvous_dir [step]
 *
 * Every rank also sweeps the whole (small) grid alone on its own GPU and compares its planes of the slab run
 * with that, bit for bit.  Prints "SLAB_LOOP_OK rank r" and exits 0 on success.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "drstencil.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != DRS_OK) {                                                          \
            fprintf(stderr, "rank %d: %s -> %d: %s\n", rank, #call, rc_, drs_last_error()); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

static int rank = 0, world = 1;

static void nap(void) {
    struct timespec ts = {0, 2000000};
    nanosleep(&ts, NULL);
}

/* every rank writes <dir>/<tag>.<rank> (atomically) and waits until all of them exist */
static int publish(const char *dir, const char *tag, const void *data, size_t bytes) {
    char tmp[512], path[512];
    snprintf(tmp, sizeof tmp, "%s/.%s.%d.tmp", dir, tag, rank);
    snprintf(path, sizeof path, "%s/%s.%d", dir, tag, rank);
    FILE *f = fopen(tmp, "wb");
    if (!f) return 1;
    if (bytes && fwrite(data, 1, bytes, f) != bytes) { fclose(f); return 1; }
    fclose(f);
    return rename(tmp, path);
}

static int fetch(const char *dir, const char *tag, int r, void *data, size_t bytes) {
    char path[512];
    snprintf(path, sizeof path, "%s/%s.%d", dir, tag, r);
    for (int tries = 0; tries < 30000; ++tries) {          /* 60 s */
        FILE *f = fopen(path, "rb");
        if (f) {
            size_t got = bytes ? fread(data, 1, bytes, f) : 0;
            fclose(f);
            if (got == bytes) return 0;
        }
        nap();
    }
    return 1;
}

static int barrier(const char *dir, const char *tag) {
    if (publish(dir, tag, NULL, 0)) return 1;
    for (int r = 0; r < world; ++r)
        if (fetch(dir, tag, r, NULL, 0)) return 1;
    return 0;
}

/* input value of global point (k, j, i): any deterministic function will do */
static double value(long long k, long long j, long long i) {
    unsigned long long s = (unsigned long long)((k * 1315423911LL) ^ (j * 2654435761LL) ^ (i * 97531LL)) + 12345ULL;
    s ^= s >> 33; s *= 0xff51afd7ed558ccdULL; s ^= s >> 33;
    return (double)(s >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv) {
    if (argc < 3) { puts("usage: slab_loop <3d .stc> <rendezvous dir> [step]"); return 2; }
    const char *dir = argv[2];
    if (getenv("RANK")) rank = atoi(getenv("RANK"));
    if (getenv("WORLD_SIZE")) world = atoi(getenv("WORLD_SIZE"));
    const int local = getenv("LOCAL_RANK") ? atoi(getenv("LOCAL_RANK")) : rank;
    drs_knobs k;
    drs_knobs_default(&k);
    if (argc > 3) { k.step = atoi(argv[3]); k.explicit_mask |= 1; }
    k.sn = 8; k.explicit_mask |= 1 << 5;                     /* several chunks per slab */
    const long long L = 24LL * world + 5, M = 96, N = 200;
    const int timesteps = 16 * k.step;

    CHECK(drs_set_device(local));
    drs_stencil *st = NULL;
    CHECK(drs_stencil_from_file(argv[1], 1, &st));

    /* ---- the undecomposed run on this rank's GPU ---- */
    CHECK(drs_stencil_set_size(st, L, M, N, timesteps));
    drs_plan *whole = NULL;
    CHECK(drs_plan_create(st, &k, &whole));
    drs_plan_info info;
    CHECK(drs_plan_get_info(whole, &info));
    const long long ghost = info.halo;
    const size_t plane = (size_t)M * N, pbytes = plane * sizeof(double);
    double *h_full = (double *)malloc(L * pbytes), *h_zero = (double *)calloc((size_t)L * plane, sizeof(double));
    for (long long z = 0; z < L; ++z)
        for (long long j = 0; j < M; ++j)
            for (long long i = 0; i < N; ++i) h_full[(z * M + j) * N + i] = value(z, j, i);
    void *fa, *fb;
    CHECK(drs_device_malloc(L * pbytes, &fa));
    CHECK(drs_device_malloc(L * pbytes, &fb));
    CHECK(drs_device_upload(fa, h_full, L * pbytes));
    CHECK(drs_device_upload(fb, h_zero, L * pbytes));
    int sweeps_whole = 0;
    CHECK(drs_run(whole, fa, fb, timesteps, NULL, &sweeps_whole));
    CHECK(drs_plan_sync_check(whole, NULL));
    double *h_ref = (double *)malloc(L * pbytes);
    CHECK(drs_device_download(h_ref, fa, L * pbytes));
    drs_device_free(fa); drs_device_free(fb);
    drs_plan_destroy(whole);

    /* ---- this rank's slab: planes [lo, hi) + `ghost` planes on each side ---- */
    const long long lo = rank * L / world, hi = (rank + 1) * L / world, local_planes = hi - lo + 2 * ghost;
    const long long org = lo - ghost;                        /* global index of local plane 0 */
    CHECK(drs_stencil_set_size(st, local_planes, M, N, timesteps));
    drs_plan *plan = NULL;
    CHECK(drs_plan_create(st, &k, &plan));
    CHECK(drs_plan_set_slab(plan, L, lo, hi));
    void *bases[2], *flags;
    CHECK(drs_device_malloc(local_planes * pbytes, &bases[0]));
    CHECK(drs_device_malloc(local_planes * pbytes, &bases[1]));
    CHECK(drs_device_malloc(16, &flags));
    double *h_loc = (double *)calloc((size_t)local_planes * plane, sizeof(double));
    for (long long zl = 0; zl < local_planes; ++zl)
        if (org + zl >= 0 && org + zl < L) memcpy(h_loc + zl * plane, h_full + (org + zl) * plane, pbytes);
    CHECK(drs_device_upload(bases[0], h_loc, local_planes * pbytes));      /* A: own planes and ghost planes */
    CHECK(drs_device_upload(bases[1], h_zero, local_planes * pbytes));     /* B = 0 (common.hpp:34-45) */
    CHECK(drs_device_upload(flags, h_zero, 16));

    /* ---- trade IPC handles with the neighbours ---- */
    unsigned char mine[3][64], theirs[3][64];
    CHECK(drs_ipc_export(bases[0], mine[0]));
    CHECK(drs_ipc_export(bases[1], mine[1]));
    CHECK(drs_ipc_export(flags, mine[2]));
    if (publish(dir, "handles", mine, sizeof mine)) { fprintf(stderr, "rank %d: cannot write to %s\n", rank, dir); return 1; }
    void *lower[3] = {NULL, NULL, NULL}, *upper[3] = {NULL, NULL, NULL};
    if (rank > 0) {
        if (fetch(dir, "handles", rank - 1, theirs, sizeof theirs)) { fprintf(stderr, "rank %d: no handles from below\n", rank); return 1; }
        for (int x = 0; x < 3; ++x) CHECK(drs_ipc_import(theirs[x], &lower[x]));
    }
    if (rank + 1 < world) {
        if (fetch(dir, "handles", rank + 1, theirs, sizeof theirs)) { fprintf(stderr, "rank %d: no handles from above\n", rank); return 1; }
        for (int x = 0; x < 3; ++x) CHECK(drs_ipc_import(theirs[x], &upper[x]));
    }
    CHECK(drs_plan_set_peers(plan, bases, lower, upper, rank > 0 ? (rank - 1) * L / world : 0,
                             rank + 1 < world ? (rank + 1) * L / world : 0));
    /* a rank writes slot 1 of its lower neighbour's flag array and slot 0 of its upper neighbour's */
    CHECK(drs_plan_set_flags(plan, flags, lower[2] ? (char *)lower[2] + 8 : NULL, upper[2]));
    if (barrier(dir, "ready")) { fprintf(stderr, "rank %d: barrier failed\n", rank); return 1; }

    /* ---- the emitted host loop, in two calls (the flag values carry over) ---- */
    int s1 = 0, s2 = 0;
    CHECK(drs_run_slab(plan, timesteps / 2, NULL, &s1));
    CHECK(drs_run_slab(plan, timesteps / 2, NULL, &s2));
    CHECK(drs_plan_sync_check(plan, NULL));
    const long long launches = drs_plan_launch_count(plan);
    CHECK(drs_device_download(h_loc, bases[0], local_planes * pbytes));
    const int same = memcmp(h_loc + ghost * plane, h_ref + lo * plane, (size_t)(hi - lo) * pbytes) == 0;
    printf("rank %d/%d: planes [%lld, %lld) of %lld, Halo %lld, %d + %d sweeps in %lld launches (whole grid: %d sweeps) -> %s\n",
           rank, world, lo, hi, L, ghost, s1, s2, launches, sweeps_whole, same ? "bit-exact" : "MISMATCH");
    if (barrier(dir, "done")) return 1;                       /* nobody unmaps while a neighbour still pushes */
    for (int x = 0; x < 3; ++x) {
        if (lower[x]) drs_ipc_close(lower[x]);
        if (upper[x]) drs_ipc_close(upper[x]);
    }
    drs_plan_destroy(plan);
    drs_device_free(bases[0]); drs_device_free(bases[1]); drs_device_free(flags);
    drs_stencil_destroy(st);
    free(h_full); free(h_zero); free(h_ref); free(h_loc);
    if (same && s1 + s2 == sweeps_whole && launches == s1 + s2) {
        printf("SLAB_LOOP_OK rank %d\n", rank);
        return 0;
    }
    return 1;
}
