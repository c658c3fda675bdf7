// drs_gold.cuh -- naive one-thread-per-point evaluation of the (composed) operator and the
// device-side error metric.  Stands in for the reference's gold_<name> kernels
// (/root/reference/codegen_2d.hpp:666-688, codegen.hpp:637-660) and checkError2D/3D
// (/root/reference/common.hpp:47-102).  Also the engine's fallback sweep for grids the TMA path
// cannot describe (row pitch not a multiple of 16 bytes).
//
// Generated translation unit must define: DRS_T DRS_GOLD_NAME DRS_CHECK_NAME DRS_GOLD_CHAIN(MUL,FMA).
#pragma once
#include "drs_common.cuh"

extern "C" __global__ void __launch_bounds__(256)
DRS_GOLD_NAME(const __grid_constant__ drs::Params p) {
    using namespace drs;
    const drs_i64 i = (drs_i64)blockIdx.x * blockDim.x + threadIdx.x;
    const drs_i64 j = (drs_i64)blockIdx.y * blockDim.y + threadIdx.y;
    const drs_i64 k = (drs_i64)blockIdx.z * blockDim.z + threadIdx.z;
    const int H = p.halo;
#if DRS_DIM == 3
    if (k < p.slow_lo || k >= p.slow_hi || j < H || j >= p.M - H) return;
#else
    if (k != 0 || j < p.slow_lo || j >= p.slow_hi) return;
#endif
    if (i < H || i >= p.N - H) return;
    const real* c = p.in + (k * p.M + j) * p.N + i;
    const drs_i64 sj = p.N, sk = p.M * p.N;
    real acc;
#define DRS_MUL_(dk, dj, di, cf) acc = rmul(c[(dk) * sk + (dj) * sj + (di)], (real)(cf));
#define DRS_FMA_(dk, dj, di, cf) acc = rfma(c[(dk) * sk + (dj) * sj + (di)], (real)(cf), acc);
    DRS_GOLD_CHAIN(DRS_MUL_, DRS_FMA_)
#undef DRS_MUL_
#undef DRS_FMA_
    p.out[(k * p.M + j) * p.N + i] = acc;
}

// res[0] = max |a - b| as raw bits of a non-negative double (atomicMax on the bit pattern),
// res[1] = sum of squared differences, over the same interior the gold kernel writes.
extern "C" __global__ void __launch_bounds__(256)
DRS_CHECK_NAME(const __grid_constant__ drs::Params p, const drs::real* __restrict__ ref, double* res) {
    using namespace drs;
    const drs_i64 i = (drs_i64)blockIdx.x * blockDim.x + threadIdx.x;
    const drs_i64 j = (drs_i64)blockIdx.y * blockDim.y + threadIdx.y;
    const drs_i64 k = (drs_i64)blockIdx.z * blockDim.z + threadIdx.z;
    const int H = p.halo;
    bool inside = i >= H && i < p.N - H;
#if DRS_DIM == 3
    inside = inside && k >= p.slow_lo && k < p.slow_hi && j >= H && j < p.M - H;
#else
    inside = inside && k == 0 && j >= p.slow_lo && j < p.slow_hi;
#endif
    double d = 0.0;
    if (inside) {
        const drs_i64 x = (k * p.M + j) * p.N + i;
        d = (double)p.in[x] - (double)ref[x];
        d = d < 0.0 ? -d : d;
    }
    double mx = d, sq = d * d;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double m2 = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = m2 > mx ? m2 : mx;
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    const int lane = (threadIdx.x + threadIdx.y * blockDim.x + threadIdx.z * blockDim.x * blockDim.y) & 31;
    if (lane == 0 && mx > 0.0) {
        atomicMax(reinterpret_cast<unsigned long long*>(res), (unsigned long long)__double_as_longlong(mx));
        atomicAdd(res + 1, sq);
    }
}

// ---- cross-GPU step flags for slab runs (one process per GPU, buffers mapped with CUDA IPC) ----
// After sweep s a rank stores s+1 into a slot of each neighbour's flag array (system-scope
// release, after a system fence that orders the sweep's peer stores before it); before sweep s+1
// it waits until both of its own slots hold >= s+1.  Ranks run on different GPUs, so the waiting
// kernel never shares a device with the kernel it waits for.
extern "C" __global__ void DRS_SIGNAL_NAME(long long* peer_a, long long* peer_b, long long value) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        __threadfence_system();
        if (peer_a) asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(peer_a), "l"(value) : "memory");
        if (peer_b) asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(peer_b), "l"(value) : "memory");
    }
}

extern "C" __global__ void DRS_WAIT_NAME(const long long* mine, int use_a, int use_b, long long value, int* fault) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const drs_u64 t0 = drs::global_ns();
        for (int slot = 0; slot < 2; ++slot) {
            if (!(slot == 0 ? use_a : use_b)) continue;
            for (;;) {
                long long v;
                asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(mine + slot) : "memory");
                if (v >= value) break;
                if (drs::global_ns() - t0 > 5000000000ull) {   // 5 s: a neighbour died
                    if (fault) atomicExch(fault, 2);
                    return;
                }
                __nanosleep(200);
            }
        }
    }
}
