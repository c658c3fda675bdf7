// TEST INFRASTRUCTURE ONLY -- wraps ONE program emitted by the reference generator
// (oracle/_ref/drstencil_ref, built from /root/reference/main.cpp) so that tests and bench.py can
// drive the reference's own gold_<name> and dr_<name> kernels on caller-provided data.
//
// Built by oracle/build_ref.py with
//   -DDRS_REF_CU="<emitted file>" -DDRS_REF_NAME=<stencil name> -DDRS_REF_3D=0|1
//   -DDRS_REF_STREAMING=0|1 -DDRS_REF_MX=<merge x> -DDRS_REF_MY=<merge y> -DDRS_REF_STEP=<step>
// and the reference's own nvcc flags (benchmarks/*/compile_run.sh:4) retargeted to sm_100a.
// The emitted text is #included from oracle/_ref/ (git-ignored); no reference source is copied.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define main drs_ref_emitted_main
#include DRS_REF_CU
#undef main

#define DRS_CAT2(a, b) a##b
#define DRS_CAT(a, b) DRS_CAT2(a, b)
#define GOLD_KERNEL DRS_CAT(gold_, DRS_REF_NAME)
#define DR_KERNEL DRS_CAT(dr_, DRS_REF_NAME)

#if DRS_REF_3D
#define DRS_L L
#else
#define DRS_L 1
#endif

// grid/block of the emitted host code: codegen_2d.hpp:585-598, codegen.hpp:566-571
static void dr_launch_shape(dim3 &grid, dim3 &block) {
#if DRS_REF_3D
    block = dim3(Bx, By, 1);
    grid = dim3(ceil(N, DRS_REF_MX * Bx - Halo * 2), ceil(M, DRS_REF_MY * By - Halo * 2), ceil(L, Sn));
#elif DRS_REF_STREAMING
    block = dim3(Bx, 1);
    grid = dim3(ceil(N, DRS_REF_MX * Bx - Halo * 2), ceil(M, Sn));
#else
    block = dim3(Bx, By);
    grid = dim3(ceil(N, DRS_REF_MX * Bx - Halo * 2), ceil(M, DRS_REF_MY * By - Halo * 2));
#endif
}

static void gold_launch_shape(dim3 &grid, dim3 &block) {
#if DRS_REF_3D
    block = dim3(8, 8, 8);
    grid = dim3(ceil(N, 8), ceil(M, 8), ceil(L, 8));
#else
    block = dim3(8, 8);
    grid = dim3(ceil(N, 8), ceil(M, 8));
#endif
}

extern "C" {

// info[0..2] = L, M, N ; info[3] = Halo ; info[4] = Iterations ; info[5] = step ; info[6] = Dist
void drs_ref_info(long long *info) {
    info[0] = DRS_L; info[1] = M; info[2] = N; info[3] = Halo; info[4] = Iterations;
    info[5] = DRS_REF_STEP; info[6] = Dist;
}

// Runs `sweeps` alternating launches (A->B, B->A, ...) of gold_<name> (which = 0) or dr_<name>
// (which = 1) from host arrays a (input) and b (initial contents of the other buffer), and copies
// both buffers back.  Returns 0 or a CUDA error code.
int drs_ref_run(int which, int sweeps, double *a, double *b) {
    size_t nbytes = sizeof(double) * (size_t)DRS_L * M * N;
    double *d[2];
    if (cudaMalloc(&d[0], nbytes) != cudaSuccess) return (int)cudaGetLastError();
    if (cudaMalloc(&d[1], nbytes) != cudaSuccess) return (int)cudaGetLastError();
    cudaMemcpy(d[0], a, nbytes, cudaMemcpyHostToDevice);
    cudaMemcpy(d[1], b, nbytes, cudaMemcpyHostToDevice);
    dim3 grid, block;
    if (which) dr_launch_shape(grid, block); else gold_launch_shape(grid, block);
    for (int s = 0; s < sweeps; ++s) {
        if (which) DR_KERNEL<<<grid, block>>>(d[s & 1], d[(s & 1) ^ 1]);
        else GOLD_KERNEL<<<grid, block>>>(d[s & 1], d[(s & 1) ^ 1]);
    }
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(a, d[0], nbytes, cudaMemcpyDeviceToHost);
    cudaMemcpy(b, d[1], nbytes, cudaMemcpyDeviceToHost);
    cudaFree(d[0]); cudaFree(d[1]);
    return (int)e;
}

// Times `sweeps` ping-pong launches of dr_<name> (which = 1) or gold_<name> (0) on device-resident
// zero-initialised data with CUDA events, after `warm` untimed launches; returns ms or < 0.
float drs_ref_time(int which, int sweeps, int warm) {
    size_t nbytes = sizeof(double) * (size_t)DRS_L * M * N;
    double *d[2];
    if (cudaMalloc(&d[0], nbytes) != cudaSuccess) return -1.f;
    if (cudaMalloc(&d[1], nbytes) != cudaSuccess) { cudaFree(d[0]); return -1.f; }
    cudaMemset(d[0], 0, nbytes); cudaMemset(d[1], 0, nbytes);
    dim3 grid, block;
    if (which) dr_launch_shape(grid, block); else gold_launch_shape(grid, block);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int s = 0; s < warm; ++s) {
        if (which) DR_KERNEL<<<grid, block>>>(d[s & 1], d[(s & 1) ^ 1]);
        else GOLD_KERNEL<<<grid, block>>>(d[s & 1], d[(s & 1) ^ 1]);
    }
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int s = 0; s < sweeps; ++s) {
        if (which) DR_KERNEL<<<grid, block>>>(d[s & 1], d[(s & 1) ^ 1]);
        else GOLD_KERNEL<<<grid, block>>>(d[s & 1], d[(s & 1) ^ 1]);
    }
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaGetLastError();      // e.g. an invalid launch shape: nothing ran
    float ms = -1.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d[0]); cudaFree(d[1]);
    return ms;
}

// The reference's own acceptance test (its --check: dr_ against gold_ through the same schedule, common.hpp:47-102)
// on device-generated input, compared on the device: returns the largest |dr - gold| over the whole array after
// `sweeps` ping-pong launches each, or < 0 when a launch fails.  Used by oracle/tune_ref.py to reject candidates of
// the reference's search space that do not run (or do not compute the stencil) on this GPU.
__global__ void drs_ref_fill(double *a, size_t n) {
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (size_t)gridDim.x * blockDim.x) {
        unsigned long long s = (unsigned long long)x * 6364136223846793005ULL + 1442695040888963407ULL;
        s ^= s >> 33; s *= 0xff51afd7ed558ccdULL; s ^= s >> 33;
        a[x] = (double)(s >> 11) * (1.0 / 9007199254740992.0);
    }
}
__global__ void drs_ref_maxdiff(const double *a, const double *b, size_t n, unsigned long long *res) {
    double m = 0.0;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (size_t)gridDim.x * blockDim.x) {
        double d = a[x] - b[x];
        d = d < 0.0 ? -d : d;
        if (!(d <= m)) m = d;          // NaN counts as a failure
    }
    if (m != m) m = 1e300;
    atomicMax(res, (unsigned long long)__double_as_longlong(m));
}
double drs_ref_check(int sweeps) {
    size_t n = (size_t)DRS_L * M * N, nbytes = sizeof(double) * n;
    double *d[4] = {0, 0, 0, 0};
    unsigned long long *res = 0;
    for (int x = 0; x < 4; ++x)
        if (cudaMalloc(&d[x], nbytes) != cudaSuccess) { for (int y = 0; y < x; ++y) cudaFree(d[y]); cudaGetLastError(); return -1.0; }
    cudaMalloc(&res, sizeof *res);
    cudaMemset(res, 0, sizeof *res);
    drs_ref_fill<<<1184, 256>>>(d[0], n);
    cudaMemcpy(d[2], d[0], nbytes, cudaMemcpyDeviceToDevice);
    cudaMemset(d[1], 0, nbytes); cudaMemset(d[3], 0, nbytes);
    dim3 grid, block, ggrid, gblock;
    dr_launch_shape(grid, block);
    gold_launch_shape(ggrid, gblock);
    for (int s = 0; s < sweeps; ++s) {
        DR_KERNEL<<<grid, block>>>(d[s & 1], d[(s & 1) ^ 1]);
        GOLD_KERNEL<<<ggrid, gblock>>>(d[2 + (s & 1)], d[2 + ((s & 1) ^ 1)]);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    double out = -1.0;
    if (e == cudaSuccess) {
        drs_ref_maxdiff<<<1184, 256>>>(d[sweeps & 1], d[2 + (sweeps & 1)], n, res);
        unsigned long long bits = 0;
        if (cudaMemcpy(&bits, res, sizeof bits, cudaMemcpyDeviceToHost) == cudaSuccess) memcpy(&out, &bits, sizeof out);
    }
    cudaGetLastError();
    for (int x = 0; x < 4; ++x) cudaFree(d[x]);
    cudaFree(res);
    return out;
}

// The emitted program, verbatim (prints the reference's own stdout lines).
int drs_ref_main(void) { return drs_ref_emitted_main(0, nullptr); }

}  // extern "C"
