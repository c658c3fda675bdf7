// Time-skewed block schedule of the host-buffer run (drs_run_host): pure host logic, no CUDA.
//
// The emitted main() of the reference copies the grid in, runs the ping-pong loop, copies it out
// (/root/reference/codegen_2d.hpp:572-583,604-619,647).  This planner orders the same work so that
// the three phases can overlap.  The grid is cut into blocks along the slow axis; block b runs ALL
// n sweeps before block b+1 starts, its output range sliding down one halo per sweep:
//
//     sweep s (1-based) of block b produces  [edge[b] - s*H, edge[b+1] - s*H)  clamped to [H, slow-H)
//
// (the first block is pinned to H at the bottom, the last one to slow-H at the top).  With every
// block at least 2*H thick this order honours each dependency of the plain schedule in place, on the
// same two buffers:
//   * what sweep s reads -- [edge[b] - s*H - H, edge[b+1] - s*H + H) of level s-1 -- was produced by
//     blocks b-1 and b at sweep s-1, and block b-1's later sweeps (s+1, s+3, ...) write strictly below;
//   * what it overwrites (level s-2) is no longer needed: block b+1 reads level s-2 only from
//     edge[b+1] - s*H upwards.
// A sweep is a pure function of its input array, so the result equals the plain schedule bit for bit.
// Block b needs nothing above edge[b+1] on the device, and its final planes never change afterwards:
// the upload of later blocks and the download of earlier ones can run under the sweeps.
// tests/test_host_schedule.py replays the step list on the CPU (not-yet-uploaded planes poisoned)
// for both extreme interleavings of the copies and compares with the plain schedule.
#pragma once
#include <algorithm>
#include <vector>

namespace drs {

struct HostStep {
    enum Kind { UPLOAD = 0, SWEEP = 1, DOWNLOAD = 2 };
    int kind, block, sweep;     // sweep: 1..n for SWEEP (odd: A -> B, even: B -> A), else 0
    long long lo, hi;           // slow-axis range [lo, hi): planes copied / output planes of the launch
};

struct HostSchedule {
    std::vector<long long> edge;     // blocks + 1 increasing values, edge[0] = 0, edge[blocks] = slow
    std::vector<HostStep> steps;     // per block: UPLOAD, its non-empty SWEEPs in order, DOWNLOAD
    int blocks() const { return (int)edge.size() - 1; }
};

// slow = slow-axis extent, H = halo of one sweep, S = nominal block thickness (>= 2*H), n = sweeps (even).
// The first and the last block are thin (a quarter of S, at least 2*H): the first one's upload and the
// last one's download (its thickness + n*H planes) are the only copies that nothing overlaps.
inline HostSchedule plan_host_schedule(long long slow, long long H, long long S, int n) {
    HostSchedule hs;
    S = std::max<long long>(S, std::max<long long>(2 * H, 1));
    const long long thin = std::max<long long>(2 * H, std::max<long long>(S / 4, 1));
    hs.edge.push_back(0);
    long long at = std::min(thin, slow);
    while (at < slow) {
        hs.edge.push_back(at);
        at += (slow - at - thin > S) ? S : ((slow - at > 2 * thin) ? slow - at - thin : slow - at);
    }
    hs.edge.push_back(slow);
    const int B = hs.blocks();
    auto cut = [&](int b, int s) -> long long {   // first output plane of block b at sweep s
        if (b <= 0) return H;
        if (b >= B) return slow - H;
        return std::min(std::max(hs.edge[b] - (long long)s * H, H), slow - H);
    };
    for (int b = 0; b < B; ++b) {
        hs.steps.push_back({HostStep::UPLOAD, b, 0, hs.edge[b], hs.edge[b + 1]});
        for (int s = 1; s <= n; ++s) {
            const long long lo = cut(b, s), hi = cut(b + 1, s);
            if (hi > lo) hs.steps.push_back({HostStep::SWEEP, b, s, lo, hi});
        }
        const long long lo = b == 0 ? 0 : cut(b, n), hi = b == B - 1 ? slow : cut(b + 1, n);
        hs.steps.push_back({HostStep::DOWNLOAD, b, 0, lo, std::max(lo, hi)});
    }
    return hs;
}

}  // namespace drs
