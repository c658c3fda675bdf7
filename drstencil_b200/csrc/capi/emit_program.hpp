// `drstencil -o out.cu`: a standalone CUDA program = the specialised translation unit + a main()
// that mirrors the reference's emitted host code (/root/reference/codegen_2d.hpp:564-664,
// codegen.hpp:547-635): rand()/(RAND_MAX-1) input, zero output, 10 warm-up launches, the
// `for (t = 0; t < Iterations; t += 2*step)` ping-pong loop, "GPU computation time: %f ms", and
// with --check the gold kernel run through the same schedule plus the max/RMS error lines of
// /root/reference/common.hpp:47-102.  Scripts that grep the reference's stdout keep working.
// Differences: sizes are 64-bit (the reference overflows `unsigned int nbytes` above 4 GiB,
// codegen_2d.hpp:575), the timed region is bracketed by device synchronisation, and the build
// line needs  -I <repo>/drstencil_b200/csrc/kernels  for the kernel templates.
#pragma once
#include <sstream>
#include <string>

#include "../core/generate.hpp"

namespace drs {

inline std::string emit_program_text(const Stencil& st, const drs_knobs& k, const KernelSpec& s) {
    std::ostringstream o;
    const char* T = s.dtype == DRS_F64 ? "double" : "float";
    o << "// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false \\\n"
         "//        -I <repo>/drstencil_b200/csrc/kernels out.cu -o out\n";
    o << "#include <cstdio>\n#include <cstdlib>\n#include <cstring>\n#include <cmath>\n#include <sys/time.h>\n"
         "#include <cuda.h>\n#include <cuda_runtime.h>\n";
    o << generate_tu(s);
    o << "\n#define GridL " << st.L << "LL\n#define GridM " << st.M << "LL\n#define GridN " << st.N << "LL\n";
    o << "#define Iterations " << st.iterations << "\n#define Halo " << s.halo << "\n#define Step " << s.step << "\n";
    o << "typedef " << T << " real_t;\n";
    o << R"(
static void check_error(const char* message) {
    cudaError_t error = cudaGetLastError();
    if (error != cudaSuccess) { printf("CUDA error : %s, %s\n", message, cudaGetErrorString(error)); exit(-1); }
}
static double get_time() {
    struct timeval tv; gettimeofday(&tv, 0);
    return tv.tv_sec + (double)tv.tv_usec * 1e-6;
}
static real_t* random_array(size_t n) {   // common.hpp:9-32
    real_t* a = (real_t*)malloc(n * sizeof(real_t));
    for (size_t x = 0; x < n; ++x) a[x] = (real_t)((double)rand() / (double)(RAND_MAX - 1));
    return a;
}
static double check_result(const real_t* out, const real_t* ref) {   // common.hpp:47-102
    double error = 0.0, max_error = 1e-13; long long mk = 0, mj = 0, mi = 0;
    const long long k0 = DRS_DIM == 3 ? Halo : 0, k1 = DRS_DIM == 3 ? GridL - Halo : 1;
    for (long long kk = k0; kk < k1; kk++)
        for (long long j = Halo; j < GridM - Halo; j++)
            for (long long i = Halo; i < GridN - Halo; i++) {
                double d = (double)out[(kk * GridM + j) * GridN + i] - (double)ref[(kk * GridM + j) * GridN + i];
                d = d < 0.0 ? -d : d;
                error += d * d;
                if (d > max_error) {
                    if (DRS_DIM == 3) printf("Values at index (%lld,%lld,%lld) differ : %.6f and %.6f\n", kk, j, i,
                                             (double)ref[(kk * GridM + j) * GridN + i], (double)out[(kk * GridM + j) * GridN + i]);
                    else printf("Values at index (%lld,%lld) differ : %.6f and %.6f\n", j, i,
                                (double)ref[(kk * GridM + j) * GridN + i], (double)out[(kk * GridM + j) * GridN + i]);
                    max_error = d; mk = kk; mj = j; mi = i;
                }
            }
    if (DRS_DIM == 3) printf("[Test] Max Error : %e @ (%lld,%lld,%lld)\n", max_error, mk, mj, mi);
    else printf("[Test] Max Error : %e @ (,%lld,%lld)\n", max_error, mj, mi);
    return sqrt(error / ((double)(k1 - k0) * (double)(GridM - 2 * Halo) * (double)(GridN - 2 * Halo)));
}
)";
    // `ring` = frozen ring of this launch: Halo, or a sub-step's share of it when a temporal depth runs as one
    // single-step launch per sub-step (KernelSpec::sub_launches)
    o << "static drs::Params make_params(const real_t* in, real_t* out, int ring = Halo) {\n"
         "    drs::Params p; memset(&p, 0, sizeof p);\n"
         "    p.in = in; p.out = out; p.L = GridL; p.M = GridM; p.N = GridN; p.halo = ring;\n"
         "    p.slow_lo = ring; p.slow_hi = (DRS_DIM == 3 ? GridL : GridM) - ring;\n";
    o << "    const long long a0 = (ring / " << s.vec() << ") * " << s.vec() << ";\n";
    o << "    p.nxs = (int)(((GridN - ring) - a0 + " << s.wu() - 1 << ") / " << s.wu() << ");\n";
    o << "    p.chunk = " << s.chunk << ";\n";
    o << "    const long long nslow = (p.slow_hi - p.slow_lo + p.chunk - 1) / p.chunk;\n";
    if (s.dim == 2) o << "    p.nys = (int)nslow; p.nzs = 1;\n";
    else o << "    p.nys = (int)((GridM - 2 * ring + " << s.tile_rows_useful() - 1 << ") / " << s.tile_rows_useful() << "); p.nzs = (int)nslow;\n";
    o << "    return p;\n}\n";
    o << R"(
static int* g_fault = 0;
static void gold_launch(const real_t* in, real_t* out) {
    drs::Params p = make_params(in, out);
    dim3 block(32, 8, 1), grid((unsigned)((GridN + 31) / 32), (unsigned)((GridM + 7) / 8), (unsigned)GridL);
    DRS_GOLD_NAME<<<grid, block>>>(p);
}
)";
    if (s.reuse) {
        // the reference's launch shape: overlapped bx x by tiles, one block per tile and chunk (codegen_2d.hpp:585-598)
        o << "static void dr_launch(const real_t* in, real_t* out) {\n"
             "    drs::Params p = make_params(in, out);\n";
        o << "    dim3 block(" << s.rbx << ", " << s.rby << ", 1);\n";
        o << "    const long long nslow = (p.slow_hi - p.slow_lo + p.chunk - 1) / p.chunk;\n";
        o << "    const unsigned gx = (unsigned)((GridN - 2 * Halo + " << (s.rbx - 2 * s.halo - 1) << ") / " << (s.rbx - 2 * s.halo) << ");\n";
        if (s.dim == 3)
            o << "    dim3 grid(gx, (unsigned)((GridM - 2 * Halo + " << (s.rby - 2 * s.halo - 1) << ") / " << (s.rby - 2 * s.halo) << "), (unsigned)nslow);\n";
        else
            o << "    dim3 grid(gx, (unsigned)nslow, 1);\n";
        o << "    DRS_NAME<<<grid, block>>>(p);\n}\n";
    } else if (s.tma_ok) {
        if (s.flat == 2) {
            // row pitch not a multiple of 16 bytes, cp.async ring (DRS_FLAT 2): the kernel ignores the map
            o << "static CUtensorMap make_map(const real_t* base) { (void)base; CUtensorMap m; memset(&m, 0, sizeof m); return m; }\n";
        } else {
            o << "static CUtensorMap make_map(const real_t* base) {\n"
                 "    typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,\n"
                 "        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,\n"
                 "        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);\n"
                 "    static encode_t encode = 0;\n"
                 "    if (!encode) { cudaDriverEntryPointQueryResult q; cudaFree(0);\n"
                 "        cudaGetDriverEntryPoint(\"cuTensorMapEncodeTiled\", (void**)&encode, cudaEnableDefault, &q);\n"
                 "        if (!encode) { printf(\"CUDA error : cuTensorMapEncodeTiled unavailable\\n\"); exit(-1); } }\n"
                 "    CUtensorMap m;\n";
            o << "    const CUtensorMapDataType dt = " << (s.dtype == DRS_F64 ? "CU_TENSOR_MAP_DATA_TYPE_FLOAT64" : "CU_TENSOR_MAP_DATA_TYPE_FLOAT32") << ";\n";
            if (s.flat == 1) {
                // row pitch not a multiple of 16 bytes, per-row TMA (DRS_FLAT 1): the array as one row of a {total, 1} tensor
                o << "    cuuint64_t dims[2] = {(cuuint64_t)GridL * GridM * GridN, 1};\n"
                     "    cuuint64_t strides[1] = {(dims[0] * sizeof(real_t) + 15) / 16 * 16};\n";
                o << "    cuuint32_t box[2] = {" << s.wb() + s.vec() << ", 1}; cuuint32_t es[2] = {1, 1};\n";
                o << "    CUresult r = encode(&m, dt, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,\n";
            } else if (s.dim == 2) {
                o << "    cuuint64_t dims[2] = {(cuuint64_t)GridN, (cuuint64_t)GridM}; cuuint64_t strides[1] = {(cuuint64_t)GridN * sizeof(real_t)};\n";
                o << "    cuuint32_t box[2] = {" << s.wb() << ", " << s.rb << "}; cuuint32_t es[2] = {1, 1};\n";
                o << "    CUresult r = encode(&m, dt, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,\n";
            } else {
                o << "    cuuint64_t dims[3] = {(cuuint64_t)GridN, (cuuint64_t)GridM, (cuuint64_t)GridL};\n"
                     "    cuuint64_t strides[2] = {(cuuint64_t)GridN * sizeof(real_t), (cuuint64_t)GridN * GridM * sizeof(real_t)};\n";
                o << "    cuuint32_t box[3] = {" << s.wb() << ", " << s.box_rows() << ", 1}; cuuint32_t es[3] = {1, 1, 1};\n";
                o << "    CUresult r = encode(&m, dt, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,\n";
            }
            o << "        CU_TENSOR_MAP_SWIZZLE_NONE, " << (s.dim == 3 ? "CU_TENSOR_MAP_L2_PROMOTION_L2_256B" : "CU_TENSOR_MAP_L2_PROMOTION_L2_128B")
              << ", CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);\n"
                 "    if (r != CUDA_SUCCESS) { printf(\"CUDA error : cuTensorMapEncodeTiled failed (%d)\\n\", (int)r); exit(-1); }\n"
                 "    return m;\n}\n";
        }
        o << "static void dr_launch_one(const real_t* in, real_t* out, int ring) {\n"
             "    static CUtensorMap maps[4]; static const real_t* bases[4] = {0, 0, 0, 0}; static int used = 0;\n"
             "    int b = 0;\n"
             "    while (b < used && bases[b] != in) ++b;\n"
             "    if (b == used) { b = used < 4 ? used++ : 0; maps[b] = make_map(in); bases[b] = in; }\n"
             "    drs::Params p = make_params(in, out, ring); p.fault = g_fault;\n"
             "    const long long tiles = (long long)p.nxs * p.nys * p.nzs;\n"
             "    if (tiles <= 0) return;\n";
        if (s.share3d)
            o << "    const unsigned ctas = (unsigned)(((p.nxs + " << s.sx - 1 << ") / " << s.sx << ") * ((p.nys + " << s.sy - 1 << ") / " << s.sy << ") * p.nzs);\n";
        else
            o << "    const unsigned ctas = (unsigned)((tiles + " << s.tiles_per_cta() - 1 << ") / " << s.tiles_per_cta() << ");\n";
        o << "    drs::TensorMap tm; memcpy(&tm, &maps[b], sizeof tm);\n";
        o << "    DRS_NAME<<<ctas, " << s.nw * 32 << ", " << s.smem_bytes() << ">>>(tm, p);\n}\n";
        if (s.sub_launches > 1) {
            // one sweep = sub_launches single-step launches with frozen rings of r, 2r, ... through scratch buffers
            // whose rings are never read (the library does the same: capi.cpp launch_sweep)
            o << "static void dr_launch(const real_t* in, real_t* out) {\n"
                 "    static real_t* scratch[2] = {0, 0};\n"
                 "    const size_t nbytes = (size_t)GridL * GridM * GridN * sizeof(real_t);\n";
            o << "    for (int i = 0; i < " << (s.sub_launches > 2 ? 2 : 1) << "; ++i)\n"
                 "        if (!scratch[i]) { cudaMalloc(&scratch[i], nbytes); check_error(\"Failed to allocate the temporal scratch buffer.\\n\"); }\n";
            o << "    const real_t* src = in;\n"
                 "    for (int s = 1; s <= " << s.sub_launches << "; ++s) {\n"
                 "        real_t* dst = s == " << s.sub_launches << " ? out : scratch[(s - 1) & 1];\n"
                 "        dr_launch_one(src, dst, s * " << s.base_order << ");\n"
                 "        src = dst;\n"
                 "    }\n}\n";
        } else {
            o << "static void dr_launch(const real_t* in, real_t* out) { dr_launch_one(in, out, Halo); }\n";
        }
    } else {
        o << "static void dr_launch(const real_t* in, real_t* out) { gold_launch(in, out); }\n";
    }
    o << R"(
int main(int argc, char** argv)
{
    puts("Initiating ...");
    const size_t count = (size_t)GridL * GridM * GridN, nbytes = count * sizeof(real_t);
    real_t* h_in = random_array(count);
    real_t* h_out = (real_t*)calloc(count, sizeof(real_t));
    real_t *in, *out;
    cudaMalloc(&in, nbytes);
    check_error("Failed to allocate device memory for in.\n");
    cudaMemcpy(in, h_in, nbytes, cudaMemcpyHostToDevice);
    cudaMalloc(&out, nbytes);
    check_error("Failed to allocate device memory for out.\n");
    cudaMemcpy(out, h_out, nbytes, cudaMemcpyHostToDevice);
    cudaMalloc(&g_fault, sizeof(int)); cudaMemset(g_fault, 0, sizeof(int));
)";
    if (s.tma_ok && !s.reuse)
        o << "    cudaFuncSetAttribute(DRS_NAME, cudaFuncAttributeMaxDynamicSharedMemorySize, " << s.smem_bytes() << ");\n";
    o << R"(
    puts("GPU computing ...");
    for (int i = 0; i < 10; i++) dr_launch(in, out);   // warm up
    cudaDeviceSynchronize();
    double startTime = get_time();
    for (int t = 0; t < Iterations; t += 2 * Step) {
        dr_launch(in, out);
        dr_launch(out, in);
    }
    cudaDeviceSynchronize();
    double endTime = get_time();
    check_error("Kernel error");
    puts("GPU finished computing.");
    printf("GPU computation time: %f ms\n", 1000 * (endTime - startTime));
)";
    if (k.check) {
        o << R"(
    puts("Checking error ...");
    real_t *g_in, *g_out;
    cudaMalloc(&g_in, nbytes);
    check_error("Failed to allocate device memory for g_in.\n");
    cudaMemcpy(g_in, h_in, nbytes, cudaMemcpyHostToDevice);
    cudaMalloc(&g_out, nbytes);
    check_error("Failed to allocate device memory for g_out.\n");
    cudaMemcpy(g_out, h_out, nbytes, cudaMemcpyHostToDevice);
    for (int t = 0; t < Iterations; t += 2 * Step) {
        gold_launch(g_in, g_out);
        gold_launch(g_out, g_in);
    }
    cudaDeviceSynchronize();
    check_error("Kernel(gold) error");
    cudaMemcpy(h_out, in, nbytes, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_in, g_in, nbytes, cudaMemcpyDeviceToHost);
    double error = check_result(h_out, h_in);
    printf("[Test] RMS Error: %e\n", error);
    cudaFree(g_in);
    cudaFree(g_out);
)";
    }
    o << R"(
    free(h_in);
    free(h_out);
    cudaFree(in);
    cudaFree(out);
    return 0;
}
)";
    return o.str();
}

}  // namespace drs
