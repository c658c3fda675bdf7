// drs_sweep3d_t.cuh -- 3D sweep with IN-KERNEL TEMPORAL BLOCKING for sm_100a (`--step n`, n >= 2).
//
// The reference realises `--step n` by multiplying the operator out on the host (3d7pt_star ->
// 25 points at n = 2, /root/reference/drstencil.hpp:262-282) and sweeping that once.  Here n
// sub-steps of the BASE operator run inside one kernel, so the grid makes one HBM round trip per
// n timesteps and the work per update stays that of the small operator.
//
// Unlike the single-step kernel (drs_sweep3d.cuh, warp-private tiles) the unit of work is a CTA:
// DRS_NW warps stacked along y share one x-y tile and march along k together.
//   * level 0 (input) planes arrive through a CTA-wide ring of DRS_ST stages, one TMA box
//     (cp.async.bulk.tensor.3d, tile + halo) and one mbarrier per plane;
//   * the sub-steps are evaluated in SCATTER form, as in the 2D temporal kernel: a plane of time
//     level s-1 is pushed into the partial sums of the 2*RK+1 level-s planes it touches; per-thread
//     state is `pw[TS][2*RK+1][RY][V]`, rotated statically;
//   * a completed plane of an intermediate level stays in the registers of the thread that
//     produced it (the partial-sum slot is not re-initialised before the next level has consumed
//     it); its column halo travels by warp shuffle, and only the rows at the top and bottom of
//     each warp's band are published to a shared-memory plane buffer (double-buffered by
//     iteration parity) for the warp above/below; one __syncthreads per plane orders publication,
//     consumption and the hand-back of the TMA stage;
//   * levels are evaluated top-down inside an iteration, so their chains are independent and the
//     wait for the newest input plane comes last;
//   * the tile loses RJ rows / E columns per level at its edges (overlapped tiling): a CTA of
//     TY = NW*RY rows stores TY - 2*(TS-1)*RJ of them.
//   * slab mode: like the single-step kernel, boundary output planes are stored a second time into
//     the neighbour GPUs' ghost planes (n*r of them per side).
// Results equal the composed operator up to rounding (<= 1e-12 relative, tests), the frozen ring of
// width n*r is never written.
//
// Generated translation unit must define: DRS_T DRS_NAME DRS_RK DRS_RJ DRS_E DRS_TS DRS_NW DRS_RY
// DRS_ST DRS_MINB DRS_SCATTER3(P,U).
#pragma once
#include "drs_common.cuh"

#if DRS_FLAT == 1
#error "the fused 3D temporal kernel has no per-row TMA form: the generator selects DRS_FLAT 2 for it"
#endif

// How much of a source plane a thread takes from registers / its neighbour lanes instead of shared
// memory (the kernel is shared-memory-bandwidth bound; B200, c4 depth 2: 589 / 646 / 657 / 667 / 687
// GStencil/s for 0..4).  Kept as a switch for ablation (DRS_EXTRA_DEFINES="DRS_T3_OWNREG=n").
//   0  every operand is read back from shared memory
//   1  levels >= 2: own rows/columns from the partial-sum registers that still hold them
//   2  + their column halo by warp shuffle
//   3  + only the edge rows of a completed plane are published (nobody else reads the rest)
//   4  + level 1: own vectors loaded once (128-bit), column halo shuffled as well
// Depth >= 4: the partial sums alone fill the 128-register budget of a 16-warp CTA and the register-resident
// variant spills more than the read-back one (ptxas 12.9: 44 vs 0 bytes, 12.8: 16 vs 8): everything is read
// back from shared memory there, as before.
#ifndef DRS_T3_OWNREG
#if DRS_TS >= 4
#define DRS_T3_OWNREG 0
#else
#define DRS_T3_OWNREG 4
#endif
#endif

// Store path of the top level (ablation switch, DRS_EXTRA_DEFINES="DRS_T3_LEANSTORE=0|1").
//   0  per row: range tests on y and on the vector's columns, 64-bit address from (z, row, column)
//   1  one per-thread output pointer and one packed mask (storable rows, first / last storable column of the vector):
//      the tile loses whole vectors per level, so away from the grid edge a thread stores whole vectors or nothing --
//      predicated 128-bit stores at uniform row offsets, no branches
#ifndef DRS_T3_LEANSTORE
#define DRS_T3_LEANSTORE 0
#endif

namespace drs {
namespace s3t {

constexpr int RK = DRS_RK, RJ = DRS_RJ, E = DRS_E, TS = DRS_TS;
constexpr int K2 = 2 * RK + 1;
constexpr int RY = DRS_RY, NW = DRS_NW, ST = DRS_ST;
constexpr int E0 = ((E + kVec - 1) / kVec) * kVec;
constexpr int HW = (((TS - 1) * E + kVec - 1) / kVec) * kVec;   // columns lost per side
constexpr int HY = (TS - 1) * RJ;                               // rows lost per side
constexpr int WT = 32 * kVec, WU = WT - 2 * HW, WB = WT + 2 * E0;
constexpr int TY = NW * RY, TYU = TY - 2 * HY, YB = TY + 2 * RJ;
constexpr int RP = WB;               // row pitch of every plane in shared memory
constexpr int PLANE_BYTES = WB * YB * (int)sizeof(real);           // bytes the TMA unit delivers per plane
constexpr int PLANE_STRIDE = (RP * YB * (int)sizeof(real) + 127) / 128 * 128;
constexpr int NLV = TS - 1;                                     // intermediate levels kept in shared memory
constexpr int DEPTH = 2 * TS * RK + TS - 1;
static_assert(ST >= 2, "at least two stages");
// ring slot and mbarrier phase of input plane n (any stage count; n is CTA-uniform, so this is uniform-datapath work)
__device__ __forceinline__ int ring_slot(int n) { return (ST & (ST - 1)) == 0 ? (n & (ST - 1)) : n % ST; }
__device__ __forceinline__ drs_u32 ring_phase(int n) { return (drs_u32)((n / ST) & 1); }
static_assert(TS >= 2, "single-step sweeps use drs_sweep3d.cuh");
static_assert(WU > 0 && TYU > 0, "tile too small for this depth");

__device__ __forceinline__ constexpr int mod_k2(int v) { return ((v % K2) + K2) % K2; }
__device__ __forceinline__ constexpr int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct Ctx {
    unsigned char* ring;     // ST input planes
    unsigned char* lv;       // NLV x 2 published planes
    drs_u64* bars;
    const TensorMap* tmap;
    int* fault;
    int x_box, y_box, z0;    // TMA coordinates of iteration 0
    int NIT;
    int warp, lane;
    // output
    int x_first, v_lo, v_hi;
    int y_first;             // global row of this warp's tile row 0
    int y_lo, y_hi;          // storable tile rows of this warp: y_lo <= y < y_hi
    drs_i64 z_out0;
    int n_first, n_end;
    real* optr;              // DRS_T3_LEANSTORE: where tile row 0 / element 0 of this thread lands for iteration 0
    unsigned int smask;      // bits 0-7 storable rows, 8-15 first storable element of the vector, 16-23 one past the last
    drs_i64 M, N;
    real* out;
    // slab mode: boundary output planes are also stored into the neighbours' ghost planes (NVLink)
    real* peer_lo; real* peer_hi;
    drs_i64 lo0, lo1, lo_shift, hi0, hi1, hi_shift;
    const real* in;          // DRS_FLAT: the CTA fills its stages itself (cp.async), straight from the array
    drs_i64 L;
    // one stage = one plane of the tile (+ halo); called by thread 0 (TMA) or by every thread of the CTA (DRS_FLAT)
    __device__ __forceinline__ void issue(int n) const {
        const int s = ring_slot(n);
#if DRS_FLAT
        const drs_i64 z = (drs_i64)z0 + n;
        flat_fill<YB, WB, NW * 32>(reinterpret_cast<real*>(ring + s * PLANE_STRIDE), in + z * M * N, z >= 0 && z < L, M, N, y_box, x_box,
                                   (int)threadIdx.x);
        cp_async_arrive(&bars[s]);
#else
        mbar_expect_tx(&bars[s], PLANE_BYTES);
        tma_load_3d(ring + s * PLANE_STRIDE, tmap, x_box, y_box, z0 + n, &bars[s]);
#endif
    }
};

template <int PH>
__device__ __forceinline__ bool iteration(real (&pw)[TS][K2][RY][kVec], const Ctx& c, int n) {
#pragma unroll
    for (int s = TS; s >= 1; --s) {
        // ---- source plane of level s-1 as this thread sees it: own rows/columns + neighbours ----
        const real* src;
        if (s == 1) {
            if (!mbar_wait(&c.bars[ring_slot(n)], ring_phase(n), c.fault)) return false;
            src = reinterpret_cast<const real*>(c.ring + ring_slot(n) * PLANE_STRIDE);
        } else {
            src = reinterpret_cast<const real*>(c.lv + ((s >= 2 ? s - 2 : 0) * 2 + ((n + 1) & 1)) * PLANE_STRIDE);
        }
        const real* mine = src + (c.warp * RY + RJ) * RP + E0 + c.lane * kVec;   // tile row 0, element 0
        // ---- scatter into the partial sums of level s ----
#if DRS_T3_OWNREG
        // Levels >= 2: the thread's own part of the source plane is still in its registers -- the slot
        // of level s-1 completed by the previous iteration is only re-initialised by level s-1's scatter,
        // which runs after this one (levels go top-down).  Only the halo (rows above/below the thread's
        // rows, columns left/right of its vector) is read back from the published plane.
#if DRS_T3_OWNREG >= 2
        // ... and the column halo comes from the adjacent lanes by shuffle (lanes 0 / 31 get their own
        // value: those columns are lost at this level anyway)
        real xl[RY][E > 0 ? E : 1], xr[RY][E > 0 ? E : 1];
        if (s >= 2) {
#pragma unroll
            for (int y = 0; y < RY; ++y)
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    xl[y][e] = __shfl_up_sync(0xffffffffu, pw[s >= 2 ? s - 2 : 0][mod_k2(PH - 1 - RK)][y][kVec - E + e], 1);
                    xr[y][e] = __shfl_down_sync(0xffffffffu, pw[s >= 2 ? s - 2 : 0][mod_k2(PH - 1 - RK)][y][e], 1);
                }
        }
#if DRS_T3_OWNREG >= 4
        // level 1: the input plane's own vectors are loaded once (128-bit) and its column halo is
        // shuffled as well; the edge lanes take theirs from the staged box when the lost columns do
        // not already cover them
        real u0[RY][kVec];
        if (s == 1) {
#pragma unroll
            for (int y = 0; y < RY; ++y) lds_vec(u0[y], mine + y * RP);
#pragma unroll
            for (int y = 0; y < RY; ++y)
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    xl[y][e] = __shfl_up_sync(0xffffffffu, u0[y][kVec - E + e], 1);
                    xr[y][e] = __shfl_down_sync(0xffffffffu, u0[y][e], 1);
                    if constexpr (HW < TS * E) {
                        if (c.lane == 0) xl[y][e] = mine[y * RP - E + e];
                        if (c.lane == 31) xr[y][e] = mine[y * RP + kVec + e];
                    }
                }
        }
#endif
#define DRS_XNBR_(yy, xx) ((xx) < 0 ? xl[yy][E + (xx)] : xr[yy][(xx) - kVec])
#else
#define DRS_XNBR_(yy, xx) mine[(yy) * RP + (xx)]
#endif
#define DRS_OWN_(yy, xx) (((xx) >= 0 && (xx) < kVec) ? pw[s >= 2 ? s - 2 : 0][mod_k2(PH - 1 - RK)][clampi(yy, 0, RY - 1)][clampi(xx, 0, kVec - 1)] \
                                                     : DRS_XNBR_(clampi(yy, 0, RY - 1), xx))
#if DRS_T3_OWNREG >= 4
#define DRS_OWN0_(yy, xx) (((xx) >= 0 && (xx) < kVec) ? u0[clampi(yy, 0, RY - 1)][clampi(xx, 0, kVec - 1)] \
                                                      : DRS_XNBR_(clampi(yy, 0, RY - 1), xx))
#define DRS_U_(dj, di) (((y + (dj)) >= 0 && (y + (dj)) < RY) ? (s >= 2 ? DRS_OWN_(y + (dj), v + (di)) : DRS_OWN0_(y + (dj), v + (di))) \
                                                             : mine[(y + (dj)) * RP + v + (di)])
#else
#define DRS_U_(dj, di) ((s >= 2 && (y + (dj)) >= 0 && (y + (dj)) < RY) ? DRS_OWN_(y + (dj), v + (di)) : mine[(y + (dj)) * RP + v + (di)])
#endif
#else
#define DRS_U_(dj, di) mine[(y + (dj)) * RP + v + (di)]
#endif
#pragma unroll
        for (int y = 0; y < RY; ++y) {
#pragma unroll
            for (int v = 0; v < kVec; ++v) {
#define DRS_P_(dk) pw[s - 1][mod_k2(PH - (dk))][y][v]
                DRS_SCATTER3(DRS_P_, DRS_U_)
#undef DRS_P_
            }
        }
#undef DRS_U_
#if DRS_T3_OWNREG
#undef DRS_OWN_
#undef DRS_XNBR_
#endif
#if DRS_T3_OWNREG >= 4
#undef DRS_OWN0_
#endif
        // ---- the plane of level s completed by this iteration ----
        if (s < TS) {
            real* dst = reinterpret_cast<real*>(c.lv + ((s - 1) * 2 + (n & 1)) * PLANE_STRIDE) +
                        (c.warp * RY + RJ) * RP + E0 + c.lane * kVec;
#pragma unroll
            for (int y = 0; y < RY; ++y) {
#if DRS_T3_OWNREG >= 3
                if (y >= RJ && y < RY - RJ) continue;   // only other warps read the plane back: edge rows suffice
#endif
                real o[kVec];
#pragma unroll
                for (int v = 0; v < kVec; ++v) o[v] = pw[s < TS ? s - 1 : 0][mod_k2(PH - RK)][y][v];
                if constexpr (sizeof(real) == 8) *reinterpret_cast<double2*>(dst + y * RP) = make_double2(o[0], o[1]);
                else *reinterpret_cast<float4*>(dst + y * RP) = make_float4(o[0], o[1], o[2], o[3]);
            }
        } else if (n >= c.n_first && n < c.n_end) {
            const drs_i64 z = c.z_out0 + n;
#if DRS_T3_LEANSTORE
            unsigned int sm = c.smask;
            asm volatile("" : "+r"(sm));      // tested bit by bit where it is used: unpacked into one register per row it costs spills
            auto scaled = [&](int y, int v) {
#ifdef DRS_OUT_SCALE
                return rmul(pw[TS - 1][mod_k2(PH - RK)][y][v], (real)DRS_OUT_SCALE);
#else
                return pw[TS - 1][mod_k2(PH - RK)][y][v];
#endif
            };
            auto store_plane = [&](real* base) {
                if ((sm >> 8) == (unsigned int)(kVec << 8)) {           // whole vectors: every tile away from the grid edge
#pragma unroll
                    for (int y = 0; y < RY; ++y) {
                        real o[kVec];
#pragma unroll
                        for (int v = 0; v < kVec; ++v) o[v] = scaled(y, v);
                        if ((sm >> y) & 1u) stg_vec(base + (drs_i64)y * c.N, o);
                    }
                } else {
                    const int vl = (int)((sm >> 8) & 0xffu), vh = (int)(sm >> 16);
#pragma unroll
                    for (int y = 0; y < RY; ++y)
#pragma unroll
                        for (int v = 0; v < kVec; ++v)
                            if (((sm >> y) & 1u) && v >= vl && v < vh) base[(drs_i64)y * c.N + v] = scaled(y, v);
                }
            };
            real* const own = c.optr + (drs_i64)n * (c.M * c.N);
            store_plane(own);
            // the same cell of a neighbour's array: byte distance between the two allocations + the plane shift
            auto in_peer = [&](real* peer, drs_i64 shift) {
                return reinterpret_cast<real*>(reinterpret_cast<char*>(own) +
                                               (reinterpret_cast<const char*>(peer) - reinterpret_cast<const char*>(c.out))) + shift * c.M * c.N;
            };
            if (c.peer_lo != nullptr && z >= c.lo0 && z < c.lo1) store_plane(in_peer(c.peer_lo, c.lo_shift));
            if (c.peer_hi != nullptr && z >= c.hi0 && z < c.hi1) store_plane(in_peer(c.peer_hi, c.hi_shift));
#else
            const drs_i64 row0 = c.y_first * c.N + c.x_first;
            // one pass per destination (own array, then the neighbours' ghost planes on the boundary planes of a slab):
            // a single base pointer is live at a time -- with all three live the slab entry point spilled in this loop
            auto store_plane = [&](real* base) {
#pragma unroll
                for (int y = 0; y < RY; ++y) {
                    if (y >= c.y_lo && y < c.y_hi) {
                        real o[kVec];
#pragma unroll
                        for (int v = 0; v < kVec; ++v) {
#ifdef DRS_OUT_SCALE
                            // the sub-steps evaluated K / eta: level TS carries eta^-TS (generate.hpp: fscale)
                            o[v] = rmul(pw[TS - 1][mod_k2(PH - RK)][y][v], (real)DRS_OUT_SCALE);
#else
                            o[v] = pw[TS - 1][mod_k2(PH - RK)][y][v];
#endif
                        }
                        real* dst = base + (drs_i64)y * c.N;
                        if (c.v_lo <= 0 && c.v_hi >= kVec) {
                            stg_vec(dst, o);
                        } else {
#pragma unroll
                            for (int v = 0; v < kVec; ++v)
                                if (v >= c.v_lo && v < c.v_hi) dst[v] = o[v];
                        }
                    }
                }
            };
            store_plane(c.out + z * c.M * c.N + row0);
            if (c.peer_lo != nullptr && z >= c.lo0 && z < c.lo1) store_plane(c.peer_lo + (z + c.lo_shift) * c.M * c.N + row0);
            if (c.peer_hi != nullptr && z >= c.hi0 && z < c.hi1) store_plane(c.peer_hi + (z + c.hi_shift) * c.M * c.N + row0);
#endif
        }
    }
    // published planes visible to every warp; the input stage is consumed by all of them
    __syncthreads();
#if DRS_FLAT
    if (n + ST < c.NIT) c.issue(n + ST);
#else
    if (threadIdx.x == 0 && n + ST < c.NIT) {
        fence_proxy_async();
        c.issue(n + ST);
    }
#endif
    return true;
}

template <int PH>
__device__ __forceinline__ bool phases(real (&pw)[TS][K2][RY][kVec], const Ctx& c, int n0) {
    if constexpr (PH < K2) {
        if (!iteration<PH>(pw, c, n0 + PH)) return false;
        return phases<PH + 1>(pw, c, n0);
    } else {
        return true;
    }
}

// SLAB: the entry point of drs_run_slab (in-kernel step flags, face chunks first); the plain entry point carries
// none of it (its register allocation sits at the 128-register cap)
template <bool SLAB>
__device__ __forceinline__ void sweep(const TensorMap& tmap, const Params& p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Ctx c;
    c.warp = threadIdx.x >> 5;
    c.lane = threadIdx.x & 31;
    c.ring = smem_raw;
    c.lv = smem_raw + ST * PLANE_STRIDE;
    c.bars = reinterpret_cast<drs_u64*>(smem_raw + (ST + 2 * NLV) * PLANE_STRIDE);
    c.tmap = &tmap;
    c.fault = p.fault;

    // one tile per CTA and at most 2^31 - 1 CTAs per launch (capi.cpp): 32-bit index arithmetic
    const unsigned int tile = blockIdx.x;
    const unsigned int per_chunk = (unsigned int)p.nxs * (unsigned int)p.nys;
    const int zc = SLAB ? slab_chunk_order(p, (int)(tile / per_chunk)) : (int)(tile / per_chunk);
    const int rem = (int)(tile % per_chunk);
    const int cy = rem / p.nxs, cx = rem % p.nxs;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < ST; ++s) mbar_init(&c.bars[s], DRS_FLAT ? NW * 32 : 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    // the published-plane buffers start out zeroed so that never-written halo cells are finite
    for (int x = threadIdx.x; x < 2 * NLV * PLANE_STRIDE / 16; x += blockDim.x)
        reinterpret_cast<uint4*>(c.lv)[x] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();

    const int H = p.halo;
    const int X0 = (H / kVec) * kVec + cx * WU - HW;
    const int Y0 = H + cy * TYU - HY;
    const drs_i64 za = p.slow_lo + (drs_i64)zc * p.chunk;
    const drs_i64 zb = (za + p.chunk < p.slow_hi) ? za + p.chunk : p.slow_hi;
    const int n_end = (int)(zb - za) + DEPTH;
    c.NIT = (n_end + K2 - 1) / K2 * K2;
    c.z0 = (int)(za - TS * RK);
    c.x_box = X0 - E0;
    c.y_box = Y0 - RJ;
    c.x_first = X0 + c.lane * kVec;
    {
        const drs_i64 lo = (H > X0 + HW) ? H : X0 + HW;
        const drs_i64 hi = (p.N - H < X0 + HW + WU) ? p.N - H : X0 + HW + WU;
        c.v_lo = (int)(lo - c.x_first);
        c.v_hi = (int)(hi - c.x_first);
    }
    c.y_first = Y0 + c.warp * RY;
    {
        const drs_i64 lo = (H > Y0 + HY) ? H : Y0 + HY;
        const drs_i64 hi = (p.M - H < Y0 + HY + TYU) ? p.M - H : Y0 + HY + TYU;
        c.y_lo = (int)(lo - c.y_first);
        c.y_hi = (int)(hi - c.y_first);
    }
    c.n_first = DEPTH;
    c.n_end = n_end;
    c.z_out0 = za - DEPTH;
    c.M = p.M;
    c.N = p.N;
    c.in = p.in;
    c.L = p.L;
    c.out = p.out;
    {
        unsigned int rows = 0u;
#pragma unroll
        for (int y = 0; y < RY; ++y)
            if (y >= c.y_lo && y < c.y_hi) rows |= 1u << y;
        const int vl = clampi(c.v_lo, 0, kVec), vh = clampi(c.v_hi, 0, kVec);
        c.smask = vl < vh ? (rows | ((unsigned int)vl << 8) | ((unsigned int)vh << 16)) : 0u;
        c.optr = p.out + (c.z_out0 * p.M + c.y_first) * p.N + c.x_first;
    }
    c.peer_lo = p.peer_lo;
    c.peer_hi = p.peer_hi;
    c.lo0 = p.push_lo0; c.lo1 = p.push_lo1; c.lo_shift = p.peer_lo_shift;
    c.hi0 = p.push_hi0; c.hi1 = p.push_hi1; c.hi_shift = p.peer_hi_shift;

    // slab runs (drs_run_slab): the thread that requests the planes waits for the neighbour's previous sweep
    // (on failure nothing is requested and the CTA leaves through the fault check of its first wait)
#if DRS_FLAT
    {
        if constexpr (SLAB) {       // every thread fills the ring: all of them wait for thread 0's verdict
            int go = 1;
            if (threadIdx.x == 0) {
                const int face = slab_face(p, zc);
                go = (!face || slab_wait(p, face & 1, face & 2)) ? 1 : 0;
            }
            if (!__syncthreads_and(go)) return;
        }
        for (int n = 0; n < ST && n < c.NIT; ++n) c.issue(n);
    }
#else
    if (threadIdx.x == 0) {
        bool go = true;
        if constexpr (SLAB) {
            const int face = slab_face(p, zc);
            go = !face || slab_wait(p, face & 1, face & 2);
        }
        if (go)
            for (int n = 0; n < ST && n < c.NIT; ++n) c.issue(n);
    }
#endif

    real pw[TS][K2][RY][kVec];
#pragma unroll
    for (int s = 0; s < TS; ++s)
#pragma unroll
        for (int d = 0; d < K2; ++d)
#pragma unroll
            for (int y = 0; y < RY; ++y)
#pragma unroll
                for (int v = 0; v < kVec; ++v) pw[s][d][y][v] = (real)0;

#pragma unroll 1
    for (int n0 = 0; n0 < c.NIT; n0 += K2) {
        if (!phases<0>(pw, c, n0)) return;       // watchdog: every warp of the CTA fails the same wait
    }
    if constexpr (SLAB) {                        // the unit of the slab protocol is the CTA
        const int face = slab_face(p, slab_chunk_order(p, (int)(blockIdx.x / ((unsigned int)p.nxs * (unsigned int)p.nys))));
        if (face) {                              // CTA-uniform
            __syncthreads();
            if (threadIdx.x == 0) slab_arrive(p, face & 1, face & 2);
        }
    }
}

}  // namespace s3t
}  // namespace drs

extern "C" __global__ void __launch_bounds__(DRS_NW * 32, DRS_MINB)
DRS_NAME(const __grid_constant__ drs::TensorMap tmap, const __grid_constant__ drs::Params p) {
    drs::s3t::sweep<false>(tmap, p);
}
extern "C" __global__ void __launch_bounds__(DRS_NW * 32, DRS_MINB)
DRS_SLAB_NAME(const __grid_constant__ drs::TensorMap tmap, const __grid_constant__ drs::Params p) {
    drs::s3t::sweep<true>(tmap, p);
}
