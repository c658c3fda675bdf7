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
//   * each thread owns DRS_VT 128-bit vectors of adjacent columns (the reference's
//     --block-merge-x); every output is stored once, 128 bits at a time where the vector is
//     fully interior (no STG + atomicAdd double touch, codegen_2d.hpp:345-366).
//
// Two evaluation schemes, chosen by the temporal depth DRS_TS:
//
//   DRS_TS == 1  GATHER, bit-exact.  The row window the reference keeps in shared memory
//     (`in_shm[Range][...]`, codegen_2d.hpp:166-172) lives in registers, rotated statically (the
//     row loop is unrolled by the window height); each output is one explicitly ordered mul/fma
//     chain over that window (DRS_CHAIN, gold order: drstencil_2d.hpp:164-178), so results equal
//     the reference's gold kernel bit for bit.
//
//   DRS_TS > 1   SCATTER, temporal blocking (`--step n` as n in-kernel sub-steps, results within
//     1e-12 of the composed operator).  A row of time level s-1 is pushed into the partial sums
//     of the 2*RJ+1 level-s rows it touches -- the reference's forward/backward accumulation idea
//     (drstencil_2d.hpp:180-228) done in registers instead of through global atomics.  State per
//     level is just those partial sums (`pw[s][2*RJ+1][C]`); the slot completed in one iteration is
//     consumed by level s+1 in the next one and then re-initialised, levels are evaluated top-down
//     so the TS chains of an iteration are independent, and x neighbours of levels >= 1 come from
//     the adjacent lanes by warp shuffle (a warp loses E columns per level at each edge: strips
//     overlap by 2*HW columns).  With DRS_FACTORED the generator has found rows of the operator
//     that are multiples of one row vector h (generate.hpp: factorise_rows): H = h . u is computed
//     once per row (DRS_HROW) and scattered with the row weights (2d9pt_box: 6 operations per
//     update instead of 9).
//
// Generated translation unit must define: DRS_T DRS_NAME DRS_RJ DRS_E DRS_TS DRS_NW DRS_ST DRS_RB
// DRS_VT DRS_MINB, DRS_CHAIN(MUL,FMA) and, for DRS_TS > 1, DRS_SCATTER(P,U) [+ DRS_FACTORED,
// DRS_HROW(U)].
#pragma once
#include "drs_common.cuh"

#ifndef DRS_FACTORED
#define DRS_FACTORED 0
#endif

namespace drs {
namespace s2d {

constexpr int RJ = DRS_RJ;          // max |dj| of the sub-step operator
constexpr int E = DRS_E;            // max |di|
constexpr int TS = DRS_TS;          // sub-steps (time levels) per sweep
constexpr int R2 = 2 * RJ + 1;      // rows an output depends on / rows an input contributes to
constexpr int VT = DRS_VT;          // 128-bit vectors per thread
constexpr int C = VT * kVec;        // consecutive columns a thread owns
constexpr int VW = C + 2 * E;       // a row as one thread sees it: own columns + E neighbours per side
constexpr int E0 = ((E + kVec - 1) / kVec) * kVec;               // smem halo columns per side
constexpr int HW = (((TS - 1) * E + kVec - 1) / kVec) * kVec;    // strip overlap per side
constexpr int WT = 32 * C;          // columns a warp computes per row
constexpr int WU = WT - 2 * HW;     // columns it stores
constexpr int WB = WT + 2 * E0;     // TMA box width
constexpr int ST = DRS_ST, RB = DRS_RB, NW = DRS_NW;
constexpr int RP = smem_row_pitch(WB);   // row pitch inside a stage (== WB unless DRS_FLAT == 1)
constexpr int STAGE_BYTES = RB * (DRS_FLAT == 1 ? flat_box(WB) : WB) * (int)sizeof(real);   // bytes the TMA unit delivers per stage
constexpr int STAGE_STRIDE = (RB * RP * (int)sizeof(real) + 127) / 128 * 128;
constexpr int WARP_SMEM = ST * STAGE_STRIDE;
// iterations between a row entering and the output that completes with it leaving
constexpr int DEPTH = TS == 1 ? R2 : 2 * TS * RJ + TS - 1;
// register state: gather keeps R2 input rows, scatter keeps R2 partial-sum rows per level
constexpr int NLV = TS;
constexpr int SW = TS == 1 ? VW : C;
static_assert((ST & (ST - 1)) == 0 && (RB & (RB - 1)) == 0, "stages and rows/stage are powers of two");
static_assert(TS == 1 || E <= C, "shuffle exchange reaches one lane");
static_assert(WB <= 256, "TMA box is at most 256 elements wide");
static_assert(WU > 0, "strip overlap leaves no useful columns");

__device__ __forceinline__ constexpr int mod_r2(int v) { return ((v % R2) + R2) % R2; }

struct Tile {
    int lane;
    int x_first;       // global column of this thread's element 0
    int v_lo, v_hi;    // storable elements of the thread's columns: v_lo <= v < v_hi
    drs_i64 y_out0;    // output row produced at iteration 0 (lies before the chunk)
    int n_first;       // first iteration whose output row is inside the chunk
    int n_end;         // one past the last such iteration
    drs_i64 N;
    real* out;
};

__device__ __forceinline__ void store_row(const Tile& t, int n, const real (&o)[C]) {
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

// this thread's view of a staged input row: own columns + E neighbours per side
__device__ __forceinline__ void load_row(real (&u)[VW], const real* __restrict__ srow, int lane) {
    const real* own = srow + E0 + lane * C;
#pragma unroll
    for (int j = 0; j < VT; ++j) {
        real tmp[kVec];
        lds_vec(tmp, own + j * kVec);
#pragma unroll
        for (int v = 0; v < kVec; ++v) u[E + j * kVec + v] = tmp[v];
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
        u[e] = own[e - E];
        u[E + C + e] = own[C + e];
    }
}

// One iteration of the row pipeline; PH = iteration mod R2 (compile-time: static rotation).
template <int PH>
__device__ __forceinline__ void row_step(real (&w)[NLV][R2][SW], const real* __restrict__ srow, const Tile& t, int n) {
#if DRS_TS == 1
    {
        // ---- gather: output row from the window as the previous iteration left it, then the new
        //      row replaces the oldest one (its loads are off the chain's critical path) ----
        constexpr int PP = mod_r2(PH - 1);     // slot of the newest row before this iteration's insert
        real o[C];
#pragma unroll
        for (int v = 0; v < C; ++v) {
            real acc;
#define DRS_W_(dj) w[0][mod_r2(PP + (dj) - RJ)]
#define DRS_MUL_(dk, dj, di, c) acc = rmul(DRS_W_(dj)[E + v + (di)], (real)(c));
#define DRS_FMA_(dk, dj, di, c) acc = rfma(DRS_W_(dj)[E + v + (di)], (real)(c), acc);
            DRS_CHAIN(DRS_MUL_, DRS_FMA_)
#undef DRS_MUL_
#undef DRS_FMA_
#undef DRS_W_
            o[v] = acc;
        }
        if (n >= t.n_first && n < t.n_end) store_row(t, n, o);
        load_row(w[0][PH], srow, t.lane);
    }
#else
    {
        // ---- scatter, levels top-down ----
#pragma unroll
        for (int s = TS; s >= 1; --s) {
            real u[VW];
            if (s == 1) {
                load_row(u, srow, t.lane);
            } else {
                // the level s-1 row completed by the previous iteration; neighbours by shuffle
                const real (&src)[SW] = w[s >= 2 ? s - 2 : 0][mod_r2(PH - 1 - RJ)];
#pragma unroll
                for (int v = 0; v < C; ++v) u[E + v] = src[v];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    u[e] = __shfl_up_sync(0xffffffffu, src[C - E + e], 1);
                    u[E + C + e] = __shfl_down_sync(0xffffffffu, src[e], 1);
                }
            }
#pragma unroll
            for (int v = 0; v < C; ++v) {
#define DRS_U_(di) u[E + v + (di)]
#define DRS_P_(dj) w[s - 1][mod_r2(PH - (dj))][v]
#if DRS_FACTORED
                real hacc;
                DRS_HROW(DRS_U_)
#endif
                DRS_SCATTER(DRS_P_, DRS_U_)
#undef DRS_U_
#undef DRS_P_
            }
            if (s == TS && n >= t.n_first && n < t.n_end) {
                real o[C];
#pragma unroll
                for (int v = 0; v < C; ++v) {
#ifdef DRS_OUT_SCALE
                    // the factorised sub-steps evaluated K / eta: level TS carries eta^-TS (generate.hpp: fscale)
                    o[v] = rmul(w[TS - 1][mod_r2(PH - RJ)][v], (real)DRS_OUT_SCALE);
#else
                    o[v] = w[TS - 1][mod_r2(PH - RJ)][v];
#endif
                }
                store_row(t, n, o);
            }
        }
    }
#endif
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
    const real* in;         // DRS_FLAT: the warp fills its stages itself (cp.async), straight from the array
    drs_i64 M, N;
    // one stage = RB input rows; called by lane 0 (TMA) or by every lane of the warp (DRS_FLAT)
    __device__ __forceinline__ void issue(int c) const {
        const int s = c & (ST - 1);
#if DRS_FLAT == 2
        flat_fill<RB, WB, 32>(reinterpret_cast<real*>(wbase + s * STAGE_STRIDE), in, true, M, N, yrow0 + c * RB, x_box, lane);
        cp_async_arrive(&bars[s]);
#elif DRS_FLAT == 1
        mbar_expect_tx(&bars[s], STAGE_BYTES);
        const drs_i64 f0 = (drs_i64)(yrow0 + c * RB) * N + x_box;
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const drs_i64 f = f0 + r * N;
            tma_load_row(wbase + s * STAGE_STRIDE + r * RP * (int)sizeof(real), tmap, (int)(f - flat_shift(f)), &bars[s]);
        }
#else
        mbar_expect_tx(&bars[s], STAGE_BYTES);
        tma_load_2d(wbase + s * STAGE_STRIDE, tmap, x_box, yrow0 + c * RB, &bars[s]);
#endif
    }
};

// One input row: wait for its stage if it is the stage's first row, run the row through every
// time level, hand the stage back to the TMA unit if it was the stage's last row.
template <int PH>
__device__ __forceinline__ bool iteration(real (&w)[NLV][R2][SW], const Stream& st, const Tile& t, int n) {
    const int rr = n & (RB - 1);
    const int c = n / RB;
    const int s = c & (ST - 1);
    if (rr == 0) {
        if (!mbar_wait(&st.bars[s], (drs_u32)((c / ST) & 1), st.fault)) return false;
    }
#if DRS_FLAT == 1
    const real* srow = reinterpret_cast<const real*>(st.wbase + s * STAGE_STRIDE) + rr * RP +
                       flat_shift((drs_i64)(st.yrow0 + n) * st.N + st.x_box);
#else
    const real* srow = reinterpret_cast<const real*>(st.wbase + s * STAGE_STRIDE) + rr * RP;
#endif
    row_step<PH>(w, srow, t, n);
    if (rr == RB - 1) {
        __syncwarp();
#if DRS_FLAT == 2
        if (c + ST < st.NCH) st.issue(c + ST);
#else
        if (st.lane == 0 && c + ST < st.NCH) {
            fence_proxy_async();
            st.issue(c + ST);
        }
#endif
    }
    return true;
}

// R2 consecutive rows with the phase known at compile time (static register rotation)
template <int PH>
__device__ __forceinline__ bool phases(real (&w)[NLV][R2][SW], const Stream& st, const Tile& t, int n0) {
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
    st.in = p.in;
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
    const int X0 = (H / kVec) * kVec + xs * WU - HW;   // column of lane 0, element 0
    const drs_i64 ya = p.slow_lo + (drs_i64)yc * p.chunk;
    const drs_i64 yb = (ya + p.chunk < p.slow_hi) ? ya + p.chunk : p.slow_hi;
    // iterations: chunk rows + pipeline depth, rounded up to whole rotations so that the unrolled
    // phases need no guard; the surplus iterations stream rows past the chunk (or zero fill) and
    // store nothing
    const int n_end = (int)(yb - ya) + DEPTH;
    st.NIT = (n_end + R2 - 1) / R2 * R2;
    st.NCH = (st.NIT + RB - 1) / RB;
    st.yrow0 = (int)(ya - TS * RJ);                     // first input row the chunk depends on
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
    t.n_first = DEPTH;
    t.n_end = n_end;
    t.y_out0 = ya - DEPTH;
    t.N = p.N;
    t.out = p.out;

    if (DRS_FLAT == 2 || lane == 0) {
        for (int c = 0; c < ST && c < st.NCH; ++c) st.issue(c);
    }

    real w[NLV][R2][SW];
#pragma unroll
    for (int s = 0; s < NLV; ++s)
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
            for (int x = 0; x < SW; ++x) w[s][r][x] = (real)0;

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
