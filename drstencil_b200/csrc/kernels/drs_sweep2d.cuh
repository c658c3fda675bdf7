// drs_sweep2d.cuh -- 2D sweep for sm_100a: rows streamed along j, one warp per (x strip, row chunk).
//
// Stands in for the reference's emitted 2D kernels, streaming and non-streaming alike
// (/root/reference/codegen_2d.hpp:149-454 and :456-561).  Differences by design:
//   * the unit of work is a WARP, not a block: no __syncthreads, no shared tile between warps;
//   * input rows arrive through a per-warp ring of DRS_ST shared-memory stages of DRS_RB rows,
//     filled by TMA (cp.async.bulk.tensor.2d) and tracked by one mbarrier per stage -- the
//     reference's `--prefetch` register double-buffer (codegen_2d.hpp:325-337) taken to its
//     hardware form; out-of-grid rows/columns are zero-filled by the TMA unit, so there is
//     no bounds logic on the load side;
//   * the row window the reference keeps in shared memory (`in_shm[Range][...]`,
//     codegen_2d.hpp:166-172) lives in registers, rotated statically (the row loop is
//     unrolled by the window height), each thread owning DRS_VT 128-bit vectors of adjacent
//     columns (the reference's --block-merge-x);
//   * `--step n` in temporal mode runs n sub-steps per sweep: level s of the register
//     window holds rows of time level s; x neighbours of levels >= 1 come from the adjacent
//     lanes by warp shuffle, so a warp loses E columns per level at each edge and strips
//     overlap by 2*HW columns; the levels of one iteration are evaluated top-down so that
//     their chains are independent of each other;
//   * every output is produced by one explicitly ordered mul/fma chain (DRS_CHAIN, gold
//     order: drstencil_2d.hpp:164-178) and stored once (no STG + atomicAdd double touch,
//     codegen_2d.hpp:345-366), with a 128-bit store where the vector is fully interior.
//
// Generated translation unit must define: DRS_T DRS_NAME DRS_RJ DRS_E DRS_TS
// DRS_CHAIN(MUL,FMA) DRS_NW DRS_ST DRS_RB DRS_VT DRS_MINB.
#pragma once
#include "drs_common.cuh"

namespace drs {
namespace s2d {

constexpr int RJ = DRS_RJ;          // max |dj| of the sub-step operator
constexpr int E = DRS_E;            // max |di|
constexpr int TS = DRS_TS;          // sub-steps (time levels) per sweep
constexpr int R2 = 2 * RJ + 1;      // register window height
constexpr int VT = DRS_VT;          // 128-bit vectors per thread (the reference's --block-merge-x)
constexpr int C = VT * kVec;        // consecutive columns a thread owns
constexpr int VW = C + 2 * E;       // register window width per thread
constexpr int E0 = ((E + kVec - 1) / kVec) * kVec;               // smem halo columns per side
constexpr int HW = (((TS - 1) * E + kVec - 1) / kVec) * kVec;    // strip overlap per side
constexpr int WT = 32 * C;          // columns a warp computes per row
constexpr int WU = WT - 2 * HW;     // columns it stores
constexpr int WB = WT + 2 * E0;     // TMA box width
constexpr int ST = DRS_ST, RB = DRS_RB, NW = DRS_NW;
constexpr int STAGE_BYTES = RB * WB * (int)sizeof(real);
constexpr int STAGE_STRIDE = (STAGE_BYTES + 127) / 128 * 128;
constexpr int WARP_SMEM = ST * STAGE_STRIDE;
static_assert((ST & (ST - 1)) == 0 && (RB & (RB - 1)) == 0, "stages and rows/stage are powers of two");
static_assert(TS == 1 || E <= C, "shuffle exchange reaches one lane");
static_assert(WB <= 256, "TMA box is at most 256 elements wide");
static_assert(WU > 0, "strip overlap leaves no useful columns");

// physical slot of the window row at offset dj when the newest row sits in slot PH
template <int PH>
__device__ __forceinline__ constexpr int slot(int dj) { return (PH + dj - RJ + 2 * R2) % R2; }

struct Tile {
    int lane;
    int x_first;       // global column of this thread's element 0
    int v_lo, v_hi;    // storable elements of the vector: v_lo <= v < v_hi
    drs_i64 y_out0;    // output row produced at iteration 0 (may lie before the chunk)
    int n_first;       // first iteration whose output row is inside the chunk
    int n_end;         // one past the last such iteration
    drs_i64 N;
    real* out;
};

// One iteration of the row pipeline.  Levels are evaluated TOP-DOWN: level s reads the window of
// level s-1 as the previous iteration left it, so the TS chains of one iteration are mutually
// independent (instruction-level parallelism across time levels) and the shared-memory loads of
// the new input row are off the critical path.  The price is one iteration of delay per level:
// the row produced for level s at iteration n is  yrow0 + n - s*(RJ + 1).
template <int PH>
__device__ __forceinline__ void row_step(real (&w)[TS][R2][VW], const real* __restrict__ srow, const Tile& t,
                                         int n) {
    constexpr int PP = (PH + R2 - 1) % R2;   // phase the windows were left in by the previous iteration
#pragma unroll
    for (int s = TS; s >= 1; --s) {
        real o[C];
#pragma unroll
        for (int v = 0; v < C; ++v) {
            real acc;
#define DRS_MUL_(dk, dj, di, c) acc = rmul(w[s - 1][slot<PP>(dj)][E + v + (di)], (real)(c));
#define DRS_FMA_(dk, dj, di, c) acc = rfma(w[s - 1][slot<PP>(dj)][E + v + (di)], (real)(c), acc);
            DRS_CHAIN(DRS_MUL_, DRS_FMA_)
#undef DRS_MUL_
#undef DRS_FMA_
            o[v] = acc;
        }
        if (s < TS) {
            // becomes the newest row of level s; x neighbours from the adjacent lanes
#pragma unroll
            for (int v = 0; v < C; ++v) w[s < TS ? s : 0][PH][E + v] = o[v];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                w[s < TS ? s : 0][PH][e] = __shfl_up_sync(0xffffffffu, o[C - E + e], 1);
                w[s < TS ? s : 0][PH][E + C + e] = __shfl_down_sync(0xffffffffu, o[e], 1);
            }
        } else if (n >= t.n_first && n < t.n_end) {
            real* dst = t.out + (t.y_out0 + n) * t.N + t.x_first;
            if (t.v_lo <= 0 && t.v_hi >= C) {
#pragma unroll
                for (int j = 0; j < VT; ++j) {
                    real ov[kVec];
#pragma unroll
                    for (int v = 0; v < kVec; ++v) ov[v] = o[j * kVec + v];
                    stg_vec(dst + j * kVec, ov);
                }
            } else {
#pragma unroll
                for (int v = 0; v < C; ++v)
                    if (v >= t.v_lo && v < t.v_hi) dst[v] = o[v];
            }
        }
    }
    // ---- level 0: this thread's vector plus E halo columns each side, from the staged row ----
    {
        const real* own = srow + E0 + t.lane * C;
#pragma unroll
        for (int j = 0; j < VT; ++j) {
            real tmp[kVec];
            lds_vec(tmp, own + j * kVec);
#pragma unroll
            for (int v = 0; v < kVec; ++v) w[0][PH][E + j * kVec + v] = tmp[v];
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            w[0][PH][e] = own[e - E];
            w[0][PH][E + C + e] = own[C + e];
        }
    }
}

// Per-warp streaming state that does not change across iterations.
struct Stream {
    unsigned char* wbase;   // this warp's ring of ST stages
    drs_u64* bars;          // one mbarrier per stage
    const TensorMap* tmap;
    int* fault;
    int x_box, yrow0;       // TMA coordinates: box column, level-0 row of iteration 0
    int NIT, NCH;           // input rows / stages this tile streams
    int lane;
    __device__ __forceinline__ void issue(int c) const {
        const int s = c & (ST - 1);
        mbar_expect_tx(&bars[s], STAGE_BYTES);
        tma_load_2d(wbase + s * STAGE_STRIDE, tmap, x_box, yrow0 + c * RB, &bars[s]);
    }
};

// One input row: wait for its stage if it is the stage's first row, run the row through every
// time level, hand the stage back to the TMA unit if it was the stage's last row.
template <int PH>
__device__ __forceinline__ bool iteration(real (&w)[TS][R2][VW], const Stream& st, const Tile& t, int n) {
    const int rr = n & (RB - 1);
    const int c = n / RB;
    const int s = c & (ST - 1);
    if (rr == 0) {
        if (!mbar_wait(&st.bars[s], (drs_u32)((c / ST) & 1), st.fault)) return false;
    }
    const real* srow = reinterpret_cast<const real*>(st.wbase + s * STAGE_STRIDE) + rr * WB;
    row_step<PH>(w, srow, t, n);
    if (rr == RB - 1) {
        __syncwarp();
        if (st.lane == 0 && c + ST < st.NCH) {
            fence_proxy_async();
            st.issue(c + ST);
        }
    }
    return true;
}

// R2 consecutive rows with the window phase known at compile time (static register rotation)
template <int PH>
__device__ __forceinline__ bool phases(real (&w)[TS][R2][VW], const Stream& st, const Tile& t, int n0) {
    if constexpr (PH < R2) {
        if (!iteration<PH>(w, st, t, n0 + PH)) return false;
        return phases<PH + 1>(w, st, t, n0);
    } else {
        return true;
    }
}

__device__ __forceinline__ void sweep(const TensorMap& tmap, const Params& p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const drs_i64 tile = (drs_i64)blockIdx.x * NW + warp;
    if (tile >= (drs_i64)p.nxs * p.nys) return;
    const int xs = (int)(tile % p.nxs);
    const int yc = (int)(tile / p.nxs);

    Stream st;
    st.wbase = smem_raw + warp * WARP_SMEM;
    st.bars = reinterpret_cast<drs_u64*>(smem_raw + NW * WARP_SMEM) + warp * ST;
    st.tmap = &tmap;
    st.fault = p.fault;
    st.lane = lane;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < ST; ++s) mbar_init(&st.bars[s], 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncwarp();

    const int H = p.halo;
    const int X0 = (H / kVec) * kVec + xs * WU - HW;   // column of lane 0, element 0
    const drs_i64 ya = p.slow_lo + (drs_i64)yc * p.chunk;
    const drs_i64 yb = (ya + p.chunk < p.slow_hi) ? ya + p.chunk : p.slow_hi;
    // iterations: chunk rows + pipeline depth (one window height per level), rounded up to whole
    // window rotations so that the unrolled phases need no guard; the surplus iterations stream
    // rows past the chunk (or zero fill) and store nothing
    const int n_end = (int)(yb - ya) + TS * R2;
    st.NIT = (n_end + R2 - 1) / R2 * R2;
    st.NCH = (st.NIT + RB - 1) / RB;
    st.yrow0 = (int)(ya - TS * RJ);
    st.x_box = X0 - E0;

    Tile t;
    t.lane = lane;
    t.x_first = X0 + lane * C;
    {
        const drs_i64 lo = (H > X0 + HW) ? H : X0 + HW;
        const drs_i64 hi = (p.N - H < X0 + HW + WU) ? p.N - H : X0 + HW + WU;
        t.v_lo = (int)(lo - t.x_first);
        t.v_hi = (int)(hi - t.x_first);
    }
    t.n_first = TS * R2;
    t.n_end = n_end;
    t.y_out0 = ya - TS * R2;
    t.N = p.N;
    t.out = p.out;

    if (lane == 0) {
        for (int c = 0; c < ST && c < st.NCH; ++c) st.issue(c);
    }

    real w[TS][R2][VW];
#pragma unroll
    for (int s = 0; s < TS; ++s)
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
            for (int x = 0; x < VW; ++x) w[s][r][x] = (real)0;

#pragma unroll 1
    for (int n0 = 0; n0 < st.NIT; n0 += R2) {
        if (!phases<0>(w, st, t, n0)) return;
    }
}

}  // namespace s2d
}  // namespace drs

extern "C" __global__ void __launch_bounds__(DRS_NW * 32, DRS_MINB)
DRS_NAME(const __grid_constant__ drs::TensorMap tmap, const __grid_constant__ drs::Params p) {
    drs::s2d::sweep(tmap, p);
}
