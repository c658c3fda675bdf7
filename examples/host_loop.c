/* examples/host_loop.c -- the reference's emitted host loop (codegen_2d.hpp:568-663) written
 * against the C ABI, in plain C: what INTEGRATION.md section 2 describes.
 *
 *   gcc -std=c99 -I include examples/host_loop.c -L drstencil_b200 -ldrstencil -Wl,-rpath,$PWD/drstencil_b200 -o host_loop
 *   ./host_loop stc/2d9pt_box.stc [step] [M N]
 *
 * Without a GPU it stops after the analysis (plan creation needs none) and reports DRS_E_NOGPU.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "drstencil.h"

int main(int argc, char **argv) {
    if (argc < 2) { puts("Please specify the .stc file."); return 0; }
    const int is3d = strstr(argv[1], "3d") != NULL;
    drs_knobs k;
    drs_knobs_default(&k);
    if (argc > 2) { k.step = atoi(argv[2]); k.explicit_mask |= 1; }
    drs_stencil *st = NULL;
    if (drs_stencil_from_file(argv[1], is3d, &st) != DRS_OK) { puts(drs_last_error()); return 255; }
    long long dims[3];
    int iterations = 0;
    if (argc > 4) {
        drs_stencil_size(st, dims, &iterations);
        drs_stencil_set_size(st, dims[0] > 1 ? atoll(argv[3]) : 1, atoll(argv[3]), atoll(argv[4]), iterations);
    }
    drs_stencil_size(st, dims, &iterations);
    int halo, dist, range, sizes[4];
    int rc = drs_stencil_analyze(st, k.dist, k.merge_forward, &halo, &dist, &range, sizes);
    printf("%s: %lld x %lld x %lld, %d iterations, %d points; reference macros Halo %d Dist %d%s\n", argv[1], dims[0],
           dims[1], dims[2], iterations, drs_stencil_terms(st, NULL, NULL, 0), halo, dist,
           rc == DRS_E_NOREUSE ? " (reference: no data to reuse)" : "");
    drs_plan *plan = NULL;
    if (drs_plan_create(st, &k, &plan) != DRS_OK) { puts(drs_last_error()); return 255; }
    drs_plan_info info;
    drs_plan_get_info(plan, &info);
    printf("plan: kernel %s, %d timestep(s) per sweep, tile %d x %d, grid %d x %d threads, %d B shared\n", info.kernel_name,
           info.timesteps_per_sweep, info.tile_x, info.tile_y, info.grid_x, info.block, info.smem_bytes);
    if (drs_device_count() == 0) {
        const size_t n = (size_t)dims[0] * dims[1] * dims[2];
        double *tmp = (double *)calloc(n, sizeof(double));
        rc = drs_run_host(plan, tmp, NULL, iterations, NULL);
        printf("no GPU: drs_run_host -> %d (%s)\n", rc, drs_last_error());
        free(tmp);
        return rc == DRS_E_NOGPU ? 0 : 1;
    }
    puts("Initiating ...");
    const size_t n = (size_t)dims[0] * dims[1] * dims[2], bytes = n * sizeof(double);
    double *h_in = (double *)malloc(bytes), *h_out = (double *)calloc(n, sizeof(double));
    for (size_t x = 0; x < n; ++x) h_in[x] = (double)rand() / (double)(RAND_MAX - 1);   /* common.hpp:9-11 */
    void *in, *out, *g_in, *g_out;
    if (drs_device_malloc(bytes, &in) || drs_device_malloc(bytes, &out) || drs_device_malloc(bytes, &g_in) ||
        drs_device_malloc(bytes, &g_out)) { puts(drs_last_error()); return 255; }
    drs_device_upload(in, h_in, bytes);   drs_device_upload(out, h_out, bytes);
    drs_device_upload(g_in, h_in, bytes); drs_device_upload(g_out, h_out, bytes);
    puts("GPU computing ...");
    for (int i = 0; i < 10; i++) drs_sweep(plan, in, out, NULL);                 /* warm up */
    int sweeps = 0;
    if (drs_run(plan, in, out, iterations, NULL, &sweeps) || drs_plan_sync_check(plan, NULL)) { puts(drs_last_error()); return 255; }
    puts("GPU finished computing.");
    puts("Checking error ...");
    drs_gold_run(plan, g_in, g_out, iterations, NULL, NULL);
    double res[2];
    if (drs_check_error(plan, in, g_in, res)) { puts(drs_last_error()); return 255; }
    printf("[Test] Max Error : %e\n[Test] RMS Error: %e\n", res[0], res[1]);
    printf("%d sweeps, %lld kernel launches\n", sweeps, drs_plan_launch_count(plan));
    drs_device_free(in); drs_device_free(out); drs_device_free(g_in); drs_device_free(g_out);
    free(h_in); free(h_out);
    drs_plan_destroy(plan);
    drs_stencil_destroy(st);
    return res[0] < 1e-9 ? 0 : 1;
}
