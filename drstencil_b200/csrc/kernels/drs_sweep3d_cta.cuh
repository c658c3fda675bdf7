// drs_sweep3d_cta.cuh -- single-step 3D sweep with ONE input ring per CTA (engine override share_x / share_y;
// the c4 / c5 presets since round 2: 2 x 2 warps per ring, 9.32 vs 9.50 ms per 1536^3 sweep for private rings under
// sustained load; bit-exact on B200, tests/test_shared_ring.py).
//
// Same arithmetic, register queue and stores as drs_sweep3d.cuh (which stays the default and is not
// touched by this file); what changes is who fetches the input.  There every warp owns a private
// ring and fetches its own 64 x RY tile plus halo, so the halo rows/columns between two tiles are
// fetched twice; ncu shows that on 1536^3 about a third of those second fetches miss the L2
// (33.5 GB read for 29.0 GB of grid, profiles/r01_ncu_summary.md).  Here the SX x SY warps of a CTA
// share one TMA box of (SX*64 + halo) x (SY*RY + halo) per plane: the halo between warps of the same
// CTA is fetched once by construction.
//   * one ring of DRS_ST stages per CTA; per stage a FULL mbarrier (TMA completion, expect_tx) and an
//     EMPTY mbarrier (one arrival per warp once the stage has left that warp's k window);
//   * warp 0 / lane 0 is the producer: before re-filling a stage it waits on its EMPTY barrier; there
//     is no __syncthreads -- warps drift by up to the ring depth;
//   * warps whose tile lies outside the grid still take part in the barrier protocol and store nothing.
//
// Generated translation unit must define: DRS_T DRS_NAME DRS_RK DRS_RJ DRS_E DRS_CHAIN(MUL,FMA)
// DRS_SX DRS_SY DRS_NW (= DRS_SX * DRS_SY) DRS_ST DRS_RY DRS_MINB.
#pragma once
#include "drs_common.cuh"

namespace drs {
namespace s3c {

constexpr int RK = DRS_RK, RJ = DRS_RJ, E = DRS_E;
constexpr int K2 = 2 * RK + 1;        // register queue depth (planes)
constexpr int RY = DRS_RY;            // rows per warp tile
constexpr int SX = DRS_SX, SY = DRS_SY, NW = SX * SY;
constexpr int E0 = ((E + kVec - 1) / kVec) * kVec;
constexpr int WT = 32 * kVec;         // columns per warp tile
constexpr int WB = SX * WT + 2 * E0;  // box width (whole CTA)
constexpr int YB = SY * RY + 2 * RJ;  // box height (whole CTA)
constexpr int ST = DRS_ST;
constexpr int LA = ST - 2 * RK;       // planes requested ahead of the one being consumed
constexpr int STAGE_BYTES = WB * YB * (int)sizeof(real);
constexpr int STAGE_STRIDE = (STAGE_BYTES + 127) / 128 * 128;
static_assert(DRS_NW == NW, "warps per CTA = share_x * share_y");
static_assert((ST & (ST - 1)) == 0, "stage count is a power of two");
static_assert(LA >= 1, "ring must hold the whole k window plus at least one plane in flight");
static_assert(WB <= 256 && YB <= 256, "TMA box extents are at most 256 elements");

template <int PH>
__device__ __forceinline__ constexpr int slot(int dk) { return (PH + dk - RK + 2 * K2) % K2; }
__device__ __forceinline__ constexpr int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ void mbar_arrive(drs_u64* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Tile {
    int lane;
    int x_first, v_lo, v_hi;
    int y_first;
    int ny;
    drs_i64 z_out0;
    int n_first, n_end;
    drs_i64 M, N;
    real* out;
    real* peer_lo; real* peer_hi;
    drs_i64 lo0, lo1, lo_shift, hi0, hi1, hi_shift;
};

struct Stream {
    unsigned char* ring;       // the CTA's ST stages
    drs_u64* full;             // ST barriers: TMA bytes landed
    drs_u64* empty;            // ST barriers: NW arrivals = every warp is done with the stage
    const TensorMap* tmap;
    int* fault;
    int x_box, y_box, z0;      // TMA coordinates of iteration 0 (CTA box)
    int NIT;
    int lane;
    int own;                   // element offset of this thread's vector in tile row 0 inside a staged plane
    bool producer;             // warp 0 of the CTA
    __device__ __forceinline__ void issue(int n) const {
        const int s = n & (ST - 1);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        tma_load_3d(ring + s * STAGE_STRIDE, tmap, x_box, y_box, z0 + n, &full[s]);
    }
    __device__ __forceinline__ const real* plane(int n) const {
        return reinterpret_cast<const real*>(ring + (n & (ST - 1)) * STAGE_STRIDE) + own;
    }
};

template <int PH>
__device__ __forceinline__ bool iteration(real (&q)[K2][RY][kVec], const Stream& st, const Tile& t, int n) {
    if (!mbar_wait(&st.full[n & (ST - 1)], (drs_u32)((n / ST) & 1), st.fault)) return false;
    {
        const real* pl = st.plane(n);
#pragma unroll
        for (int y = 0; y < RY; ++y) lds_vec(q[PH][y], pl + y * WB);
    }
    if (n >= t.n_first && n < t.n_end) {
        const real* sp[K2];
#pragma unroll
        for (int d = 0; d < K2; ++d) sp[d] = st.plane(n - 2 * RK + d);
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
#define DRS_OPERAND_(dk, dj, di)                                                               \
    (((di) == 0 && (y + (dj)) >= 0 && (y + (dj)) < RY)                                         \
         ? q[slot<PH>(dk)][clampi(y + (dj), 0, RY - 1)][v]                                     \
         : sp[(dk) + RK][(y + (dj)) * WB + v + (di)])
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
    // the oldest plane of the window (plane n - 2*RK) has left this warp's window
    __syncwarp();
    const int done = n - 2 * RK;
    if (st.lane == 0 && done >= 0) mbar_arrive(&st.empty[done & (ST - 1)]);
    if (st.producer) {          // warp-uniform
        int ok = 1;
        if (st.lane == 0 && n + LA < st.NIT) {
            // stage of plane n + LA = stage of plane `done`: wait until every warp has released it
            if (done >= 0) ok = mbar_wait(&st.empty[done & (ST - 1)], (drs_u32)((done / ST) & 1), st.fault) ? 1 : 0;
            if (ok) {
                fence_proxy_async();
                st.issue(n + LA);
            }
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) return false;
    }
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
    const int wx = warp % SX, wy = warp / SX;
    const int nxc = (p.nxs + SX - 1) / SX, nyc = (p.nys + SY - 1) / SY;   // CTA tiles per plane
    const drs_i64 per_chunk = (drs_i64)nxc * nyc;
    const drs_i64 cta = blockIdx.x;
    const int zc = SLAB ? slab_chunk_order(p, (int)(cta / per_chunk)) : (int)(cta / per_chunk);
    const int rem = (int)(cta % per_chunk);
#ifdef DRS_S3C_YFAST      // ablation (DRS_EXTRA_DEFINES): y-adjacent CTA tiles consecutive in launch order instead of x-adjacent
    const int cys = rem % nyc, cxs = rem / nyc;
#elif defined(DRS_S3C_BLOCKED)   // ablation: 2 x 2 groups of CTA tiles consecutive in launch order
    const int g = rem >> 2, gx = g % ((nxc + 1) >> 1), gy = g / ((nxc + 1) >> 1);
    const int cxs = 2 * gx + (rem & 1), cys = 2 * gy + ((rem >> 1) & 1);
#else
    const int cys = rem / nxc, cxs = rem % nxc;
#endif

    Stream st;
    st.ring = smem_raw;
    st.full = reinterpret_cast<drs_u64*>(smem_raw + ST * STAGE_STRIDE);
    st.empty = st.full + ST;
    st.tmap = &tmap;
    st.fault = p.fault;
    st.lane = lane;
    st.producer = warp == 0;
    st.own = (wy * RY + RJ) * WB + E0 + wx * WT + lane * kVec;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < ST; ++s) {
            mbar_init(&st.full[s], 1);
            mbar_init(&st.empty[s], NW);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();             // barriers initialised before any warp waits on them (once per CTA)

    const int H = p.halo;
    const int X0c = (H / kVec) * kVec + cxs * SX * WT;
    const int Y0c = H + cys * SY * RY;
    const int X0 = X0c + wx * WT;
    const int Y0 = Y0c + wy * RY;
    const drs_i64 za = p.slow_lo + (drs_i64)zc * p.chunk;
    const drs_i64 zb = (za + p.chunk < p.slow_hi) ? za + p.chunk : p.slow_hi;
    const int n_end = (int)(zb - za) + 2 * RK;
    st.NIT = (n_end + K2 - 1) / K2 * K2;
    st.z0 = (int)(za - RK);
    st.x_box = X0c - E0;
    st.y_box = Y0c - RJ;

    Tile t;
    t.lane = lane;
    t.x_first = X0 + lane * kVec;
    {
        const drs_i64 lo = H;
        const drs_i64 hi = (p.N - H < X0 + WT) ? p.N - H : X0 + WT;
        t.v_lo = (int)(lo - t.x_first);
        t.v_hi = (int)(hi - t.x_first);      // <= 0 for a warp beyond the right edge: stores nothing
    }
    t.y_first = Y0;
    {
        const drs_i64 rows = p.M - H - Y0;   // <= 0 for a warp below the bottom edge: stores nothing
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

    // slab runs (drs_run_slab): only the producer waits for the neighbour's previous sweep -- everything the
    // other warps do (including their pushes) depends on planes it requests afterwards.  If the wait fails
    // nothing is requested and every warp leaves through the fault check of its first mbarrier wait.
    if (threadIdx.x == 0) {
        bool go = true;
        if constexpr (SLAB) {
            const int face = slab_face(p, zc);
            go = !face || slab_wait(p, face & 1, face & 2);
        }
        if (go)
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
    if constexpr (SLAB) {            // the unit of the slab protocol is the warp: NW arrivals per CTA
        const int nxc2 = (p.nxs + SX - 1) / SX, nyc2 = (p.nys + SY - 1) / SY;
        const int face = slab_face(p, slab_chunk_order(p, (int)(blockIdx.x / ((drs_i64)nxc2 * nyc2))));
        if (face) {
            __syncwarp();
            if (lane == 0) slab_arrive(p, face & 1, face & 2);
        }
    }
}

}  // namespace s3c
}  // namespace drs

extern "C" __global__ void __launch_bounds__(DRS_NW * 32, DRS_MINB)
DRS_NAME(const __grid_constant__ drs::TensorMap tmap, const __grid_constant__ drs::Params p) {
    drs::s3c::sweep<false>(tmap, p);
}
extern "C" __global__ void __launch_bounds__(DRS_NW * 32, DRS_MINB)
DRS_SLAB_NAME(const __grid_constant__ drs::TensorMap tmap, const __grid_constant__ drs::Params p) {
    drs::s3c::sweep<true>(tmap, p);
}
