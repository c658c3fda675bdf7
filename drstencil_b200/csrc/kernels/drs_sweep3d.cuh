// drs_sweep3d.cuh -- 3D sweep for sm_100a: planes streamed along k, one warp per
// (x strip, y band, plane chunk).
//
// Stands in for the reference's emitted 3D kernel (/root/reference/codegen.hpp:143-544), which
// keeps `in_shm[Range][my*By][mx*Bx]` planes in shared memory behind two __syncthreads per plane
// and assembles each output from a store plus atomicAdds (codegen.hpp:391-427).  Here:
//   * a warp owns a tile of 32*kVec columns x DRS_RY rows and marches along k; each thread owns
//     one 128-bit vector of columns in every row of the tile;
//   * planes (tile + halo rows/columns) arrive through a per-warp ring of DRS_ST stages, one
//     TMA box (cp.async.bulk.tensor.3d) and one mbarrier per plane; the TMA unit zero-fills
//     whatever lies outside the grid;
//   * the thread's own column of every plane in the k window sits in a register queue
//     (`q[2*RK+1][RY][kVec]`, rotated statically); operands with di != 0, or in the halo rows of
//     the tile, are read from the staged planes, which stay resident until the window has
//     passed them;
//   * one explicitly ordered mul/fma chain per output (gold order, drstencil.hpp:182-196), one
//     store per output; in slab mode boundary planes are stored a second time, straight into
//     the neighbour GPU's ghost planes over NVLink (fused halo push).
//
// Generated translation unit must define: DRS_T DRS_NAME DRS_RK DRS_RJ DRS_E
// DRS_CHAIN(MUL,FMA) DRS_NW DRS_ST DRS_RY DRS_MINB.
#pragma once
#include "drs_common.cuh"

// Ablation switch (DRS_EXTRA_DEFINES="DRS_S3_OWNREG=1"): operands at di != 0 that fall inside the
// thread's own 128-bit vector come from the register queue instead of the staged plane.  B200:
// 3d9pt_cross 768^3 338 -> 367 GStencil/s, 3d7pt_star 768^3 unchanged, 1536^3 376 -> 367; off by default.
#ifndef DRS_S3_OWNREG
#define DRS_S3_OWNREG 0
#endif

namespace drs {
namespace s3d {

constexpr int RK = DRS_RK, RJ = DRS_RJ, E = DRS_E;
constexpr int K2 = 2 * RK + 1;        // register queue depth (planes)
constexpr int RY = DRS_RY;            // rows per warp tile
constexpr int E0 = ((E + kVec - 1) / kVec) * kVec;
constexpr int WT = 32 * kVec;         // columns per warp tile
constexpr int WB = WT + 2 * E0;       // box width
constexpr int YB = RY + 2 * RJ;       // box height
constexpr int ST = DRS_ST, NW = DRS_NW;
constexpr int LA = ST - 2 * RK;       // planes requested ahead of the one being consumed
constexpr int RP = smem_row_pitch(WB);   // row pitch inside a staged plane (== WB unless DRS_FLAT == 1)
constexpr int STAGE_BYTES = (DRS_FLAT == 1 ? flat_box(WB) : WB) * YB * (int)sizeof(real);   // bytes the TMA unit delivers per plane
constexpr int STAGE_STRIDE = (RP * YB * (int)sizeof(real) + 127) / 128 * 128;
constexpr int WARP_SMEM = ST * STAGE_STRIDE;
static_assert((ST & (ST - 1)) == 0, "stage count is a power of two");
static_assert(LA >= 1, "ring must hold the whole k window plus at least one plane in flight");

template <int PH>
__device__ __forceinline__ constexpr int slot(int dk) { return (PH + dk - RK + 2 * K2) % K2; }
__device__ __forceinline__ constexpr int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct Tile {
    int lane;
    int x_first, v_lo, v_hi;   // as in the 2D kernel
    int y_first;               // global row of tile row 0
    int ny;                    // storable tile rows: rows [0, ny) (rows below the tile start are always interior)
    drs_i64 z_out0;            // output plane produced at iteration 0
    int n_first, n_end;        // iterations whose output plane lies inside the chunk
    drs_i64 M, N;
    real* out;
    // slab mode
    real* peer_lo; real* peer_hi;
    drs_i64 lo0, lo1, lo_shift, hi0, hi1, hi_shift;
};

struct Stream {
    unsigned char* wbase;
    drs_u64* bars;
    const TensorMap* tmap;
    int* fault;
    int x_box, y_box, z0;      // TMA coordinates of iteration 0
    int NIT;
    int lane;
    const real* in;            // DRS_FLAT: the warp fills its stages itself (cp.async), straight from the array
    drs_i64 L, M, N;
    // one stage = one plane of the tile (+ halo); called by lane 0 (TMA) or by every lane of the warp (DRS_FLAT)
    __device__ __forceinline__ void issue(int n) const {
        const int s = n & (ST - 1);
#if DRS_FLAT == 2
        const drs_i64 z = (drs_i64)z0 + n;
        flat_fill<YB, WB, 32>(reinterpret_cast<real*>(wbase + s * STAGE_STRIDE), in + z * M * N, z >= 0 && z < L, M, N, y_box, x_box, lane);
        cp_async_arrive(&bars[s]);
#elif DRS_FLAT == 1
        mbar_expect_tx(&bars[s], STAGE_BYTES);
        const drs_i64 f0 = flat0(n);
#pragma unroll
        for (int r = 0; r < YB; ++r) {
            const drs_i64 f = f0 + r * N;
            tma_load_row(wbase + s * STAGE_STRIDE + r * RP * (int)sizeof(real), tmap, (int)(f - flat_shift(f)), &bars[s]);
        }
#else
        mbar_expect_tx(&bars[s], STAGE_BYTES);
        tma_load_3d(wbase + s * STAGE_STRIDE, tmap, x_box, y_box, z0 + n, &bars[s]);
#endif
    }
    __device__ __forceinline__ const real* plane(int n) const {
        return reinterpret_cast<const real*>(wbase + (n & (ST - 1)) * STAGE_STRIDE);
    }
    // DRS_FLAT == 1: flat element index of the first element of box row 0 of plane n; box row r of that plane sits
    // flat_shift(flat0(n) + r * N) elements into its slot
    __device__ __forceinline__ drs_i64 flat0(int n) const { return (((drs_i64)z0 + n) * M + y_box) * N + x_box; }
};

template <int PH>
__device__ __forceinline__ bool iteration(real (&q)[K2][RY][kVec], const Stream& st, const Tile& t, int n) {
    if (!mbar_wait(&st.bars[n & (ST - 1)], (drs_u32)((n / ST) & 1), st.fault)) return false;
#if DRS_FLAT == 1
    // row (y + RJ) of window plane d holds its data DRS_SH_(d, y) elements into its slot
    int fl[K2];                                           // shift of box row 0 of each window plane
#pragma unroll
    for (int d = 0; d < K2; ++d) fl[d] = flat_shift(st.flat0(n - 2 * RK + d));
    const int nm = (int)(st.N & (drs_i64)(kVec - 1));     // the shift advances by N mod kVec per row
#define DRS_SH_(d, yy) ((fl[d] + ((yy) + RJ) * nm) & (kVec - 1))
#else
#define DRS_SH_(d, yy) 0
#endif
    // newest plane: own vectors of every tile row into the queue
    {
        const real* pl = st.plane(n) + RJ * RP + E0 + t.lane * kVec;
#pragma unroll
        for (int y = 0; y < RY; ++y) lds_vec(q[PH][y], pl + y * RP + DRS_SH_(K2 - 1, y));
    }
    if (n >= t.n_first && n < t.n_end) {
        // staged planes of the window: sp[dk + RK] -> this thread's element 0 of tile row 0
        const real* sp[K2];
#pragma unroll
        for (int d = 0; d < K2; ++d) sp[d] = st.plane(n - 2 * RK + d) + RJ * RP + E0 + t.lane * kVec;
        const drs_i64 z = t.z_out0 + n;
        real* orow = t.out + (z * t.M + t.y_first) * t.N + t.x_first;
        const bool push_lo = t.peer_lo != nullptr && z >= t.lo0 && z < t.lo1;
        const bool push_hi = t.peer_hi != nullptr && z >= t.hi0 && z < t.hi1;
        real* plo = push_lo ? t.peer_lo + ((z + t.lo_shift) * t.M + t.y_first) * t.N + t.x_first : nullptr;
        real* phi = push_hi ? t.peer_hi + ((z + t.hi_shift) * t.M + t.y_first) * t.N + t.x_first : nullptr;
#pragma unroll
        for (int y = 0; y < RY; ++y) {
            real o[kVec];
#pragma unroll
            for (int v = 0; v < kVec; ++v) {
                real acc;
                // operand: own register queue when it is this thread's column inside the tile rows,
                // else the staged plane (halo rows / neighbouring columns)
#if DRS_S3_OWNREG
#define DRS_OPERAND_(dk, dj, di)                                                               \
    (((v + (di)) >= 0 && (v + (di)) < kVec && (y + (dj)) >= 0 && (y + (dj)) < RY)              \
         ? q[slot<PH>(dk)][clampi(y + (dj), 0, RY - 1)][clampi(v + (di), 0, kVec - 1)]         \
         : sp[(dk) + RK][(y + (dj)) * RP + v + (di) + DRS_SH_((dk) + RK, y + (dj))])
#else
#define DRS_OPERAND_(dk, dj, di)                                                               \
    (((di) == 0 && (y + (dj)) >= 0 && (y + (dj)) < RY)                                         \
         ? q[slot<PH>(dk)][clampi(y + (dj), 0, RY - 1)][v]                                     \
         : sp[(dk) + RK][(y + (dj)) * RP + v + (di) + DRS_SH_((dk) + RK, y + (dj))])
#endif
#define DRS_MUL_(dk, dj, di, c) acc = rmul(DRS_OPERAND_(dk, dj, di), (real)(c));
#define DRS_FMA_(dk, dj, di, c) acc = rfma(DRS_OPERAND_(dk, dj, di), (real)(c), acc);
                DRS_CHAIN(DRS_MUL_, DRS_FMA_)
#undef DRS_MUL_
#undef DRS_FMA_
#undef DRS_OPERAND_
                o[v] = acc;
            }
            if (y < t.ny) {
                real* dst = orow + (drs_i64)y * t.N;
                if (t.v_lo <= 0 && t.v_hi >= kVec) {
                    stg_vec(dst, o);
                    if (push_lo) stg_vec(plo + (drs_i64)y * t.N, o);
                    if (push_hi) stg_vec(phi + (drs_i64)y * t.N, o);
                } else {
#pragma unroll
                    for (int v = 0; v < kVec; ++v)
                        if (v >= t.v_lo && v < t.v_hi) {
                            dst[v] = o[v];
                            if (push_lo) plo[(drs_i64)y * t.N + v] = o[v];
                            if (push_hi) phi[(drs_i64)y * t.N + v] = o[v];
                        }
                }
            }
        }
    }
#undef DRS_SH_
    // the oldest plane of the window is no longer needed: refill its stage
    __syncwarp();
#if DRS_FLAT == 2
    if (n + LA < st.NIT) st.issue(n + LA);
#else
    if (st.lane == 0 && n + LA < st.NIT) {
        fence_proxy_async();
        st.issue(n + LA);
    }
#endif
    return true;
}

template <int PH>
__device__ __forceinline__ bool phases(real (&q)[K2][RY][kVec], const Stream& st, const Tile& t, int n0) {
    if constexpr (PH < K2) {
        if (!iteration<PH>(q, st, t, n0 + PH)) return false;
        return phases<PH + 1>(q, st, t, n0);
    } else {
        return true;
    }
}

// SLAB: the entry point of drs_run_slab (in-kernel step flags, face chunks first)
template <bool SLAB>
__device__ __forceinline__ void sweep(const TensorMap& tmap, const Params& p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const drs_i64 tile = (drs_i64)blockIdx.x * NW + warp;
    const drs_i64 per_chunk = (drs_i64)p.nxs * p.nys;
    if (tile >= per_chunk * p.nzs) return;
    const int zc = SLAB ? slab_chunk_order(p, (int)(tile / per_chunk)) : (int)(tile / per_chunk);
    const int rem = (int)(tile % per_chunk);
    const int ys = rem / p.nxs;
    const int xs = rem % p.nxs;

    Stream st;
    st.wbase = smem_raw + warp * WARP_SMEM;
    st.bars = reinterpret_cast<drs_u64*>(smem_raw + NW * WARP_SMEM) + warp * ST;
    st.tmap = &tmap;
    st.fault = p.fault;
    st.lane = lane;
    st.in = p.in;
    st.L = p.L;
    st.M = p.M;
    st.N = p.N;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < ST; ++s) mbar_init(&st.bars[s], DRS_FLAT == 2 ? 32 : 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncwarp();

    const int H = p.halo;
    const int X0 = (H / kVec) * kVec + xs * WT;
    const int Y0 = H + ys * RY;
    const drs_i64 za = p.slow_lo + (drs_i64)zc * p.chunk;
    const drs_i64 zb = (za + p.chunk < p.slow_hi) ? za + p.chunk : p.slow_hi;
    const int n_end = (int)(zb - za) + 2 * RK;
    st.NIT = (n_end + K2 - 1) / K2 * K2;   // whole queue rotations: the unrolled phases need no guard
    st.z0 = (int)(za - RK);
    st.x_box = X0 - E0;
    st.y_box = Y0 - RJ;

    Tile t;
    t.lane = lane;
    t.x_first = X0 + lane * kVec;
    {
        const drs_i64 lo = H;
        const drs_i64 hi = (p.N - H < X0 + WT) ? p.N - H : X0 + WT;
        t.v_lo = (int)(lo - t.x_first);
        t.v_hi = (int)(hi - t.x_first);
    }
    t.y_first = Y0;
    {
        const drs_i64 rows = p.M - H - Y0;
        t.ny = rows < RY ? (int)rows : RY;
    }
    t.n_first = 2 * RK;
    t.n_end = n_end;
    t.z_out0 = za - 2 * RK;
    t.M = p.M;
    t.N = p.N;
    t.out = p.out;
    t.peer_lo = p.peer_lo;
    t.peer_hi = p.peer_hi;
    t.lo0 = p.push_lo0; t.lo1 = p.push_lo1; t.lo_shift = p.peer_lo_shift;
    t.hi0 = p.push_hi0; t.hi1 = p.push_hi1; t.hi_shift = p.peer_hi_shift;

    // slab runs (drs_run_slab): a tile next to a neighbour's slab waits for that neighbour's previous sweep
    if constexpr (SLAB) {
        const int face = slab_face(p, zc);       // warp-uniform
        if (face) {
            int ok = 1;
            if (lane == 0) ok = slab_wait(p, face & 1, face & 2) ? 1 : 0;
            if (!__shfl_sync(0xffffffffu, ok, 0)) return;
        }
    }

    if (DRS_FLAT == 2 || lane == 0) {
        for (int n = 0; n < LA && n < st.NIT; ++n) st.issue(n);
    }

    real q[K2][RY][kVec];
#pragma unroll
    for (int d = 0; d < K2; ++d)
#pragma unroll
        for (int y = 0; y < RY; ++y)
#pragma unroll
            for (int v = 0; v < kVec; ++v) q[d][y][v] = (real)0;

#pragma unroll 1
    for (int n0 = 0; n0 < st.NIT; n0 += K2) {
        if (!phases<0>(q, st, t, n0)) return;
    }
    if constexpr (SLAB) {
        const int face = slab_face(p, slab_chunk_order(p, (int)(((drs_i64)blockIdx.x * NW + warp) / ((drs_i64)p.nxs * p.nys))));
        if (face) {
            __syncwarp();
            if (lane == 0) slab_arrive(p, face & 1, face & 2);
        }
    }
}

}  // namespace s3d
}  // namespace drs

extern "C" __global__ void __launch_bounds__(DRS_NW * 32, DRS_MINB)
DRS_NAME(const __grid_constant__ drs::TensorMap tmap, const __grid_constant__ drs::Params p) {
    drs::s3d::sweep<false>(tmap, p);
}
extern "C" __global__ void __launch_bounds__(DRS_NW * 32, DRS_MINB)
DRS_SLAB_NAME(const __grid_constant__ drs::TensorMap tmap, const __grid_constant__ drs::Params p) {
    drs::s3d::sweep<true>(tmap, p);
}
