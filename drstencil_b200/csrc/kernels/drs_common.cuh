// drs_common.cuh -- device-side building blocks shared by the sm_100a sweep kernels.
// Compiles under NVRTC (no host headers) and under nvcc (emitted standalone programs).
//
// The sweep kernels replace the reference's emitted dr_<name> kernels
// (/root/reference/codegen_2d.hpp:149-561, codegen.hpp:143-544).  Where the reference stages
// rows/planes with LDG -> STS -> __syncthreads -> LDS and writes every output twice
// (STG + atomicAdd), these kernels let the TMA unit fill a per-warp ring of shared-memory
// stages (mbarrier-tracked, no block-wide barrier anywhere), keep the slow-axis window in
// registers, exchange fast-axis neighbours of intermediate time levels with warp shuffles, and
// store each output exactly once with 128-bit stores.
#pragma once

#ifndef DRS_T
#error "DRS_T (element type) must be defined by the generated translation unit"
#endif

// DRS_FLAT != 0: the grid's row pitch is not a multiple of 16 bytes (odd N in fp64, N % 4 != 0 in fp32).  No
// tensor map can describe such an array by rows (strides must be multiples of 16 bytes), and a TMA box must START on
// a 16-byte boundary (measured: tools/tma_align_probe.cu), which every second row of an odd-N fp64 grid misses.  The
// ring of shared-memory stages is kept -- same mbarriers, same consumers, same chain -- and filled in one of two ways:
//   DRS_FLAT == 1  the array is described as ONE row of a {total, 1} tensor and a tile arrives as one TMA request per
//                  ROW, each starting at the 16-byte boundary at or below the row's first element; the row's data then
//                  sits `shift` (0 .. kVec-1) elements into its slot, and the consumers add that shift (their 128-bit
//                  shared loads become scalar ones).  Rows are 128 bytes apart in shared memory (a TMA destination
//                  must be 128-byte aligned).  A flat coordinate is a signed 32-bit element index: arrays below 2^31
//                  elements only.
//   DRS_FLAT == 2  (larger arrays, and the fused 3D temporal kernel) the warp fills the stages itself with
//                  element-sized cp.async (LDGSTS, zero fill outside the grid) whose completion arrives on the stage's
//                  mbarrier (cp.async.mbarrier.arrive.noinc, one arrival per lane); layout as in the aligned case.
// Stores are scalar in both, since rows start at alternating 16-byte offsets.  With DRS_FLAT == 1 a box that hangs
// over a row end reads the neighbouring row instead of zeros -- harmless: every output whose cone leaves the grid
// lies in the frozen ring and is never stored.  The reference's emitted kernels take any N through their i_ok guards
// (/root/reference/codegen_2d.hpp:192-207); this is the engine's equivalent.
#ifndef DRS_FLAT
#define DRS_FLAT 0
#endif

typedef unsigned int drs_u32;
typedef unsigned long long drs_u64;
typedef long long drs_i64;

namespace drs {

typedef DRS_T real;
constexpr int kVec = 16 / (int)sizeof(real);  // elements per 128-bit access: 2 (f64) or 4 (f32)

struct __align__(64) TensorMap { drs_u64 opaque[16]; };  // CUtensorMap, encoded on the host

// DRS_FLAT == 1: width of the per-row TMA box (the tile's box plus room for the shift) and the row pitch of a stage
__host__ __device__ constexpr int flat_box(int wb) { return wb + kVec; }
__host__ __device__ constexpr int smem_row_pitch(int wb) {
    return DRS_FLAT == 1 ? (int)(((flat_box(wb) * sizeof(real) + 127) / 128 * 128) / sizeof(real)) : wb;
}
// elements between the 16-byte boundary at or below flat element index f and f itself
__host__ __device__ constexpr int flat_shift(drs_i64 f) { return (int)(f & (drs_i64)(kVec - 1)); }

// Kernel parameters common to the 2D and 3D sweeps (sizes are runtime values: unlike the
// reference, which bakes L/M/N in as macros, one compiled plan serves any grid size).
struct Params {
    const real* in;
    real* out;
    drs_i64 L, M, N;        // local array extents (3D: L planes of M rows of N), N contiguous
    drs_i64 slow_lo, slow_hi;  // output range along the slow axis (rows in 2D, planes in 3D)
    int halo;                  // frozen ring width in the other axes (reference macro Halo)
    int nxs, nys, nzs;         // tiles per axis (x strips, y tiles / row chunks, plane chunks)
    int chunk;                 // slow-axis outputs per tile (the reference's Sn)
    int* fault;                // device flag set when a pipeline wait times out
    // slab mode (multi-GPU, slow-axis decomposition): output planes z in [push_lo0, push_lo1) are
    // also stored into the lower neighbour's ghost planes, at plane z + peer_lo_shift of its
    // array (peer_lo = that array's base, mapped over NVLink); likewise push_hi* / peer_hi.
    real* peer_lo;
    real* peer_hi;
    drs_i64 push_lo0, push_lo1, peer_lo_shift;
    drs_i64 push_hi0, push_hi1, peer_hi_shift;
    // in-kernel slab protocol (drs_run_slab; all null / zero otherwise): the step flags of a slab run are
    // handled by the sweep itself.  A work unit (a warp's tile, or a CTA's) whose chunk reads ghost planes
    // or pushes into a neighbour's waits until that neighbour's slot of `my_flags` holds >= seq (the
    // neighbour's face units of the previous sweep are done: its pushes have landed here and it no longer
    // reads the ghost planes this sweep overwrites); when the last unit of a face has finished, seq + 1 is
    // released into the neighbour's slot.  Face chunks are scheduled first, so a sweep signals early and
    // its successor practically never waits.  seq = *seq_base + seq_off (seq_base is device memory, so a
    // captured launch sequence can be replayed with a new base).
    const drs_i64* my_flags;       // [0] written by the lower neighbour, [1] by the upper one
    drs_i64* lower_flag;           // lower neighbour's slot 1 / upper neighbour's slot 0 (peer mapped), or null
    drs_i64* upper_flag;
    const drs_i64* seq_base;
    unsigned int* face_cnt;        // [0] lower face, [1] upper face: units finished so far (self-resetting)
    drs_i64 face_lo_end;           // a chunk [za, zb) touches the lower face iff za < face_lo_end
    drs_i64 face_hi_begin;         //                     ... the upper face iff zb > face_hi_begin
    int seq_off;
    int face_lo_chunks, face_hi_chunks;   // how many chunks those are (a prefix / a suffix of the chunk list)
    unsigned int face_lo_units, face_hi_units;   // work units per face = arrivals that complete it
};

__device__ __forceinline__ drs_u32 smem_u32(const void* p) {
    return (drs_u32)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(drs_u64* bar, drs_u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// orders this thread's earlier generic-proxy shared-memory accesses before later async-proxy
// (TMA) accesses to the same locations
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(drs_u64* bar, drs_u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(drs_u64* bar, drs_u32 parity) {
    drs_u32 ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ drs_u64 global_ns() {
    drs_u64 t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait: a TMA that never lands (bad descriptor, driver fault) raises the fault flag after
// ~1 s and lets the kernel run to completion instead of hanging the GPU.  Callers return at once
// when this yields false.
__device__ __forceinline__ bool mbar_wait(drs_u64* bar, drs_u32 parity, int* fault) {
    if (mbar_try_wait(bar, parity)) return true;
    const drs_u64 t0 = global_ns();
    for (;;) {
        #pragma unroll 1
        for (int spin = 0; spin < 64; ++spin)
            if (mbar_try_wait(bar, parity)) return true;
        if (global_ns() - t0 > 1000000000ull || (fault && *(volatile int*)fault)) break;
    }
    if (fault) atomicCAS(fault, 0, 1);
    return false;
}

// ---- in-kernel slab protocol (see Params) ----------------------------------------------------
// Chunk order of a slab launch: the chunks of the lower face, then those of the upper face, then the
// interior -- a sweep's boundary planes are produced (and pushed, and signalled) first.
__device__ __forceinline__ int slab_chunk_order(const Params& p, int i) {
    if (p.my_flags == nullptr) return i;
    if (i < p.face_lo_chunks) return i;
    if (i < p.face_lo_chunks + p.face_hi_chunks) return p.nzs - p.face_hi_chunks + (i - p.face_lo_chunks);
    return i - p.face_hi_chunks;
}
__device__ __forceinline__ bool slab_touches_lo(const Params& p, drs_i64 za) {
    return p.my_flags != nullptr && p.face_lo_units != 0 && za < p.face_lo_end;
}
__device__ __forceinline__ bool slab_touches_hi(const Params& p, drs_i64 zb) {
    return p.my_flags != nullptr && p.face_hi_units != 0 && zb > p.face_hi_begin;
}
// bit 0 / bit 1: the chunk with index zc (after slab_chunk_order) touches the lower / upper face.  Cheap and
// uniform: recomputed where it is needed instead of being kept live across the plane loop.
__device__ __forceinline__ int slab_face(const Params& p, int zc) {
    if (p.my_flags == nullptr) return 0;
    const drs_i64 za = p.slow_lo + (drs_i64)zc * p.chunk;
    const drs_i64 zb = (za + p.chunk < p.slow_hi) ? za + p.chunk : p.slow_hi;
    return (slab_touches_lo(p, za) ? 1 : 0) | (slab_touches_hi(p, zb) ? 2 : 0);
}
// One thread of the unit, before the unit's first TMA request.  5 s watchdog (fault code 2).
__device__ __forceinline__ bool slab_wait(const Params& p, bool lo, bool hi) {
    const drs_i64 want = *p.seq_base + p.seq_off;
    const drs_u64 t0 = global_ns();
    for (int slot = 0; slot < 2; ++slot) {
        if (!(slot == 0 ? lo : hi)) continue;
        for (;;) {
            drs_i64 v;
            asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p.my_flags + slot) : "memory");
            if (v >= want) break;
            if (global_ns() - t0 > 5000000000ull || (p.fault && *(volatile int*)p.fault)) {
                if (p.fault) atomicCAS(p.fault, 0, 2);
                return false;
            }
            __nanosleep(100);
        }
    }
    // the ghost planes were written through the generic proxy (peer stores); the TMA unit reads them
    // through the async proxy
    asm volatile("fence.proxy.async.global;" ::: "memory");
    return true;
}
// One thread of the unit, after the unit has synchronised (__syncwarp / __syncthreads: the unit's stores, own planes
// and pushed ghost planes, are then ordered before this thread's next operation).  The arrival is a RELEASE at GPU
// scope on the face counter and the last arriver ACQUIRES it, so every unit's stores are in causality order before
// the last arriver's system-scope release of the flag, which the neighbour acquires -- one system-scope operation
// per face and sweep instead of a system fence in every unit (measured at 8 GPUs: that fence held each boundary
// warp's slot for the NVLink round trip of its pushes and cost 3.5 % of the sweep).
__device__ __forceinline__ unsigned int slab_count(unsigned int* cnt) {
    unsigned int old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(cnt) : "memory");
    return old;
}
__device__ __forceinline__ void slab_arrive(const Params& p, bool lo, bool hi) {
    const drs_i64 next = *p.seq_base + p.seq_off + 1;
    if (lo && slab_count(&p.face_cnt[0]) + 1u == p.face_lo_units) {
        p.face_cnt[0] = 0u;                      // the next launch starts after this one has drained
        __threadfence_system();
        if (p.lower_flag) asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p.lower_flag), "l"(next) : "memory");
    }
    if (hi && slab_count(&p.face_cnt[1]) + 1u == p.face_hi_units) {
        p.face_cnt[1] = 0u;
        __threadfence_system();
        if (p.upper_flag) asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p.upper_flag), "l"(next) : "memory");
    }
}

// TMA tile loads: global -> shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void tma_load_2d(void* dst, const TensorMap* map, int x, int y, drs_u64* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
// L2 policy experiments (DRS_EXTRA_DEFINES; 0 = none, the shipped setting -- DESIGN.md 3.5 has the measurements):
//   DRS_LD_HINT  1 / 2: the 3D plane loads carry an L2 evict_last / evict_first policy
//   DRS_ST_HINT  1: output vectors are stored with st.global.cs (streaming)
#ifndef DRS_LD_HINT
#define DRS_LD_HINT 0
#endif
#ifndef DRS_ST_HINT
#define DRS_ST_HINT 0
#endif
__device__ __forceinline__ void tma_load_3d(void* dst, const TensorMap* map, int x, int y, int z, drs_u64* bar) {
#if DRS_LD_HINT
    // the fixed encodings of createpolicy.fractional.L2::evict_last / evict_first with fraction 1.0
    const drs_u64 policy = DRS_LD_HINT == 1 ? 0x14F0000000000000ull : 0x12F0000000000000ull;
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
#else
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
#endif
}
// DRS_FLAT == 1: one row of a tile; x = flat element index of the box start (a multiple of kVec)
__device__ __forceinline__ void tma_load_row(void* dst, const TensorMap* map, int x, drs_u64* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(0), "r"(smem_u32(bar))
        : "memory");
}
// DRS_FLAT == 2: one element global -> shared, asynchronously; !valid writes a zero (what the tensor map's out-of-bounds
// fill does).  cp_async_arrive makes this lane's earlier copies arrive on `bar` when they have landed.
__device__ __forceinline__ void cp_async_elem(real* dst, const real* src, bool valid) {
    const drs_u32 n = valid ? (drs_u32)sizeof(real) : 0u;
    if constexpr (sizeof(real) == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(drs_u64* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Rows [0, ROWS) x columns [0, WBX) of a box whose corner is (y0, x0) of the plane at `plane` (row pitch N, M rows),
// into shared memory at `dst` (row pitch WBX): the share of thread `t` of NT cooperating threads.  A warp (NT <= WBX)
// goes row by row, thread t taking columns t, t + NT, ... (column validity is the same for every row and is
// computed once); a whole CTA (NT > WBX) walks the box as one run of ROWS*WBX elements.
template <int ROWS, int WBX, int NT>
__device__ __forceinline__ void flat_fill(real* dst, const real* plane, bool plane_ok, drs_i64 M, drs_i64 N, int y0, int x0, int t) {
    if constexpr (NT <= WBX) {
        constexpr int KC = (WBX + NT - 1) / NT;
        bool cok[KC];
#pragma unroll
        for (int j = 0; j < KC; ++j) cok[j] = x0 + t + j * NT >= 0 && x0 + t + j * NT < N;
        const real* src = plane + (drs_i64)y0 * N + x0 + t;
        real* d = dst + t;
#pragma unroll 2
        for (int r = 0; r < ROWS; ++r) {
            const bool row_ok = plane_ok && y0 + r >= 0 && y0 + r < M;
#pragma unroll
            for (int j = 0; j < KC; ++j) {
                if (KC * NT <= WBX || t + j * NT < WBX) {
                    const bool ok = row_ok && cok[j];
                    cp_async_elem(d + j * NT, ok ? src + j * NT : plane, ok);
                }
            }
            src += N;
            d += WBX;
        }
    } else {
        constexpr int TOTAL = ROWS * WBX, DR = NT / WBX, DC = NT % WBX;
        int r = t / WBX, c = t % WBX;
        const real* src = plane + ((drs_i64)y0 + r) * N + x0 + c;
#pragma unroll 2
        for (int q = t; q < TOTAL; q += NT) {
            const bool ok = plane_ok && y0 + r >= 0 && y0 + r < M && x0 + c >= 0 && x0 + c < N;
            cp_async_elem(dst + q, ok ? src : plane, ok);
            c += DC; r += DR; src += (drs_i64)DR * N + DC;
            if (c >= WBX) { c -= WBX; r += 1; src += N - WBX; }
        }
    }
}
__device__ __forceinline__ void tma_prefetch_desc(const TensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// Explicitly rounded arithmetic: the evaluation order of a stencil expression is part of the
// result, so nothing is left to the compiler's contraction rules.
__device__ __forceinline__ double rmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double rfma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double radd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float radd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float rmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float rfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// 128-bit shared loads / global stores of kVec elements
__device__ __forceinline__ void lds_vec(real (&v)[kVec], const real* p) {
    if constexpr (DRS_FLAT == 1) {
        // the row's shift (warp-uniform) may break the 16-byte alignment of the vector: scalar loads then
        if ((smem_u32(p) & 15u) != 0u) {
#pragma unroll
            for (int x = 0; x < kVec; ++x) v[x] = p[x];
            return;
        }
    }
    if constexpr (sizeof(real) == 8) {
        double2 t = *reinterpret_cast<const double2*>(p);
        v[0] = t.x; v[1] = t.y;
    } else {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
}
__device__ __forceinline__ void stg_vec(real* p, const real (&v)[kVec]) {
    if constexpr (DRS_FLAT != 0) {
        // rows start at any element: whether this row's vectors are 16-byte aligned is a (warp-uniform) run-time fact
        if ((reinterpret_cast<drs_u64>(p) & 15ull) != 0ull) {
#pragma unroll
            for (int x = 0; x < kVec; ++x) p[x] = v[x];
            return;
        }
    }
#if DRS_ST_HINT
    if constexpr (sizeof(real) == 8) __stcs(reinterpret_cast<double2*>(p), make_double2(v[0], v[1]));
    else __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
#else
    if constexpr (sizeof(real) == 8) {
        *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    } else {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
#endif
}

}  // namespace drs
