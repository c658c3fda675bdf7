// Micro-probe (development aid): which box start coordinates does a tiled TMA load accept?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/tma_align_probe tools/tma_align_probe.cu -lcuda
//   ./tma_align_probe <element bytes: 8|4> <x> [rank: 1|2]
// One thread loads a box of 32 elements starting at element x of a 4096-element array described as a rank-2
// {4096, 1} (or rank-1) tensor, and the host checks what arrived.  Run once per x: a faulting load kills the context.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

struct __align__(64) TensorMap { unsigned long long opaque[16]; };

__global__ void probe(const __grid_constant__ TensorMap tmap, int x, int rank, unsigned char* out, int bytes) {
    __shared__ __align__(128) unsigned char buf[1024];
    __shared__ unsigned long long bar;
    if (threadIdx.x == 0) {
        unsigned b = (unsigned)__cvta_generic_to_shared(&bar), d = (unsigned)__cvta_generic_to_shared(buf);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        if (rank == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(d), "l"(&tmap), "r"(x), "r"(0), "r"(b) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];"
                         ::"r"(d), "l"(&tmap), "r"(x), "r"(b) : "memory");
        unsigned ok = 0;
        for (int spin = 0; spin < 100000000 && !ok; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
        out[1023] = (unsigned char)ok;
        for (int i = 0; i < bytes; ++i) out[i] = buf[i];
    }
}

int main(int argc, char** argv) {
    const int es = argc > 1 ? atoi(argv[1]) : 8, x = argc > 2 ? atoi(argv[2]) : 0, rank = argc > 3 ? atoi(argv[3]) : 2;
    const int n = 4096, box = 32, bytes = box * es;
    cudaFree(0);
    typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    encode_t encode = 0;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q);
    unsigned char* h = (unsigned char*)malloc((size_t)n * es);
    for (int i = 0; i < n; ++i) {
        if (es == 8) ((double*)h)[i] = (double)i; else ((float*)h)[i] = (float)i;
    }
    void* d;
    unsigned char* out;
    cudaMalloc(&d, (size_t)n * es);
    cudaMalloc(&out, 1024);
    cudaMemset(out, 0xff, 1024);
    cudaMemcpy(d, h, (size_t)n * es, cudaMemcpyHostToDevice);
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)n, 1};
    cuuint64_t strides[1] = {(cuuint64_t)n * es};
    cuuint32_t bx[2] = {(cuuint32_t)box, 1}, estr[2] = {1, 1};
    CUresult r = encode(&m, es == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims,
                        rank == 2 ? strides : (argc > 4 ? strides : nullptr), bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("es %d x %d rank %d: encode failed (%d)\n", es, x, rank, (int)r); return 2; }
    TensorMap tm;
    memcpy(&tm, &m, sizeof tm);
    probe<<<1, 32>>>(tm, x, rank, out, bytes);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("es %d x %d rank %d: kernel failed: %s\n", es, x, rank, cudaGetErrorString(e)); return 1; }
    unsigned char res[1024];
    cudaMemcpy(res, out, 1024, cudaMemcpyDeviceToHost);
    int good = res[1023] == 1;
    for (int i = 0; i < box && good; ++i) {
        const double want = (x + i >= 0 && x + i < n) ? (double)(x + i) : 0.0;
        const double got = es == 8 ? ((double*)res)[i] : (double)((float*)res)[i];
        if (got != want) { printf("es %d x %d rank %d: element %d = %g, expected %g\n", es, x, rank, i, got, want); good = 0; }
    }
    printf("es %d x %d rank %d: %s\n", es, x, rank, good ? "OK" : (res[1023] == 1 ? "WRONG DATA" : "TIMEOUT"));
    return good ? 0 : 3;
}
