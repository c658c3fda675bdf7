// Host-side stencil model: .stc description -> point table -> (optional) operator
// composition -> literal coefficients -> halo / reuse-distance / partition analysis.
//
// This is the B200 engine's replacement for the reference's analysis layer
//   /root/reference/drstencil_2d.hpp  (class DRStencil_2d)
//   /root/reference/drstencil.hpp     (class DRStencil)
// 2D and 3D share one implementation here: a point is always (k, j, i) with k == 0 in 2D,
// so std::map ordering over the triple equals the reference's ordering over (j, i).
//
// Parity-critical behaviours kept on purpose (each cited where implemented):
//   * token-stream .stc grammar and its quirks          drstencil_2d.hpp:48-73, drstencil.hpp:52-78
//   * depth-first composition order of `--step n`       drstencil_2d.hpp:231-251, drstencil.hpp:262-282
//   * coefficients reach the kernel as 6-significant-digit decimal literals
//                                                       drstencil_2d.hpp:174, drstencil.hpp:192
//   * Halo = largest positive offset along the slowest axis only
//                                                       drstencil_2d.hpp:82-97, drstencil.hpp:88-103
//   * forward / backward partition and Range            drstencil_2d.hpp:180-228,254-269
#pragma once
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <set>
#include <string>
#include <tuple>
#include <vector>

namespace drs {

using Point = std::tuple<int, int, int>;  // (k, j, i); slowest axis first
using PointMap = std::map<Point, double>;
using PointSet = std::set<Point>;

struct Term {
    int dk, dj, di;
    double coef;  // value of the decimal literal the reference would print
};

struct Analysis {
    int halo = 0;        // reference macro `Halo`  (order)
    int dist = 0;        // reference macro `Dist`
    int low = 1;         // lowest slow-axis offset touched by any partition set
    int high = -1;       // highest; `Range` = high - low + 1
    bool has_forward_slow = false;  // forward_j (2D) / forward_k (3D) non-empty
    PointSet forward_slow, forward_mid, forward_fast, backward;
    int range() const { return high - low + 1; }
};

// printf("%g") of a double -- the formatting std::ostream applies by default, which is how
// the reference turns a coefficient into kernel source text.
inline std::string coef_literal_text(double c) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%g", c);
    return buf;
}

// The double that nvcc reads back from that text: what dr_/gold_ kernels really multiply by.
inline double coef_literal_value(double c) { return std::strtod(coef_literal_text(c).c_str(), nullptr); }

class Stencil {
   public:
    int dim = 2;                       // 2 or 3
    long long L = 1, M = 0, N = 0;     // extents, N contiguous
    int iterations = 0;                // `iterations` key; stays 0 when the key is missing/misspelt
    int step = 1;                      // how many times the base operator has been composed
    PointMap base;                     // as parsed
    PointMap points;                   // after compose(); equals base when step == 1

    // --- .stc reader --------------------------------------------------------------------
    // Keys: L (3D only) M N iterations stencil.  Unknown tokens are skipped, so a 2D read of
    // a file holding `L 512` ignores both tokens, and `iteratioins 4` leaves iterations
    // unset.  After `stencil`, tuples `[k] j i coef` run to the first token that is not a
    // number; a repeated point keeps the last coefficient.
    // Deviation: the reference spins forever when anything non-numeric follows the tuple
    // list (stream stuck in fail state, drstencil_2d.hpp:59-69); this reader stops there.
    // Returns false when the file cannot be opened (reference prints
    // "Error opening stencil file." and main exits 255).
    bool read_stc(const std::string& path, bool is3d) {
        std::ifstream f(path);
        if (!f) return false;
        dim = is3d ? 3 : 2;
        L = 1;
        std::string tok;
        while (f >> tok) {
            if (is3d && tok == "L") f >> L;
            else if (tok == "M") f >> M;
            else if (tok == "N") f >> N;
            else if (tok == "iterations") f >> iterations;
            else if (tok == "stencil") {
                int k = 0, j = 0, i = 0;
                double c = 0;
                for (;;) {
                    if (is3d) { if (!(f >> k >> j >> i >> c)) break; }
                    else      { if (!(f >> j >> i >> c)) break; }
                    base[Point(k, j, i)] = c;
                }
                break;
            }
            if (!f) break;
        }
        points = base;
        step = 1;
        return true;
    }

    void set_points(int dim_, const int* offs, const double* coefs, int n) {
        dim = dim_;
        base.clear();
        for (int p = 0; p < n; ++p) {
            int k = dim == 3 ? offs[3 * p] : 0;
            int j = dim == 3 ? offs[3 * p + 1] : offs[2 * p];
            int i = dim == 3 ? offs[3 * p + 2] : offs[2 * p + 1];
            base[Point(k, j, i)] = coefs[p];
        }
        points = base;
        step = 1;
    }

    // --- operator composition (`--step n`) ---------------------------------------------
    // The composed operator's coefficient at offset p is the sum, over every length-n path
    // of base offsets ending at p, of the product of the base coefficients along the path.
    // Floating-point parity needs the same evaluation order as drstencil_2d.hpp:231-251:
    // paths enumerated depth-first with the base map iterated in ascending key order at every
    // level, products formed left to right starting from 1.0, sums accumulated in visit order.
    void compose(int nstep) {
        step = nstep;
        PointMap acc;
        walk(acc, Point(0, 0, 0), 1.0, nstep);
        points = acc;
    }

    // Terms in evaluation order (ascending map order == the gold expression's order,
    // drstencil_2d.hpp:164-178) with coefficients passed through the decimal literal.
    std::vector<Term> terms() const {
        std::vector<Term> t;
        for (const auto& [p, c] : points)
            t.push_back({std::get<0>(p), std::get<1>(p), std::get<2>(p), coef_literal_value(c)});
        return t;
    }
    std::vector<Term> base_terms() const {
        std::vector<Term> t;
        for (const auto& [p, c] : base)
            t.push_back({std::get<0>(p), std::get<1>(p), std::get<2>(p), coef_literal_value(c)});
        return t;
    }

    // Largest positive offset along the slowest axis of `pts` (drstencil_2d.hpp:82-92).
    int order_of(const PointMap& pts) const {
        int hi = 0;
        for (const auto& [p, c] : pts) hi = std::max(hi, slow(p));
        return hi;
    }
    // Largest |offset| over every axis.
    static int radius_of(const PointMap& pts) {
        int r = 0;
        for (const auto& [p, c] : pts) {
            r = std::max(r, std::abs(std::get<0>(p)));
            r = std::max(r, std::abs(std::get<1>(p)));
            r = std::max(r, std::abs(std::get<2>(p)));
        }
        return r;
    }

    // --- Halo / Dist / forward-backward partition / Range -------------------------------
    // `dist_opt` 0 means "derive": (high - low) >> 1 over slow-axis offsets.
    // Returns false when no point can be forwarded along the slow axis -- the case where the
    // reference prints "No data to reuse. You can try another dist." and exits 1
    // (drstencil_2d.hpp:217-220).  The B200 kernels never need the partition (they write
    // each output once); it is computed so the CLI reports the same macros and exit codes.
    bool analyze(int dist_opt, int merge_forward, Analysis& a) const {
        int hi = 0, lo = 0;
        for (const auto& [p, c] : points) { hi = std::max(hi, slow(p)); lo = std::min(lo, slow(p)); }
        a = Analysis();
        a.halo = hi;
        a.dist = dist_opt != 0 ? dist_opt : ((hi - lo) >> 1);
        const int d = a.dist;
        PointSet done;
        auto has = [&](const Point& p) { return points.find(p) != points.end(); };
        // A point whose slow-axis predecessor (offset - d) is also a stencil point can be
        // accumulated `d` rows/planes early: it joins the forward set and marks the predecessor done.
        for (const auto& [p, c] : points) {
            Point q = shifted(p, 0, -d);
            if (has(q)) { a.forward_slow.insert(p); done.insert(q); }
        }
        if (dim == 3) {
            for (const auto& [p, c] : points) {
                Point q = shifted(p, 1, -d);
                if (has(q) && !done.count(q)) { a.forward_mid.insert(p); done.insert(q); }
            }
        }
        for (const auto& [p, c] : points) {
            Point q = shifted(p, 2, -d);
            if (has(q) && !done.count(q)) { a.forward_fast.insert(p); done.insert(q); }
        }
        for (const auto& [p, c] : points)
            if (!done.count(p)) { a.backward.insert(p); done.insert(p); }
        a.has_forward_slow = !a.forward_slow.empty();
        if (!a.has_forward_slow) return false;
        // Small secondary forward sets are folded back (their predecessors become backward terms).
        if (dim == 3 && (int)a.forward_mid.size() < merge_forward) {
            for (const auto& p : a.forward_mid) a.backward.insert(shifted(p, 1, -d));
            a.forward_mid.clear();
        }
        if ((int)a.forward_fast.size() < merge_forward) {
            for (const auto& p : a.forward_fast) a.backward.insert(shifted(p, 2, -d));
            a.forward_fast.clear();
        }
        a.low = 1; a.high = -1;
        for (const PointSet* s : {&a.forward_slow, &a.forward_mid, &a.forward_fast, &a.backward})
            for (const auto& p : *s) { a.low = std::min(a.low, slow(p)); a.high = std::max(a.high, slow(p)); }
        return true;
    }

    // The partial sums of the partition as the reference prints them (gen_forward_j / gen_forward_i /
    // gen_backward, drstencil_2d.hpp:120-162): a forward term reads the input at the point's offset but carries
    // the coefficient of the point `dist` earlier along its axis; set order == evaluation order.
    // which: 0 forward_slow, 1 forward_mid, 2 forward_fast, 3 backward.
    std::vector<Term> partial_sum(const Analysis& a, int which) const {
        const PointSet& set = which == 0 ? a.forward_slow : which == 1 ? a.forward_mid : which == 2 ? a.forward_fast : a.backward;
        std::vector<Term> t;
        for (const Point& p : set) {
            const Point q = which == 3 ? p : shifted(p, which, -a.dist);
            const auto it = points.find(q);
            const double c = it == points.end() ? 0.0 : it->second;
            t.push_back({std::get<0>(p), std::get<1>(p), std::get<2>(p), coef_literal_value(c)});
        }
        return t;
    }

   private:
    int slow(const Point& p) const { return dim == 3 ? std::get<0>(p) : std::get<1>(p); }
    // axis: 0 slow, 1 middle (3D only), 2 fast
    Point shifted(const Point& p, int axis, int by) const {
        auto [k, j, i] = p;
        if (axis == 2) i += by;
        else if (axis == 1) j += by;
        else if (dim == 3) k += by;
        else j += by;
        return Point(k, j, i);
    }
    void walk(PointMap& acc, Point at, double prod, int left) const {
        if (left == 0) {
            auto it = acc.find(at);
            if (it == acc.end()) acc[at] = prod; else it->second += prod;
            return;
        }
        for (const auto& [p, c] : base)
            walk(acc,
                 Point(std::get<0>(at) + std::get<0>(p), std::get<1>(at) + std::get<1>(p),
                       std::get<2>(at) + std::get<2>(p)),
                 prod * c, left - 1);
    }
};

}  // namespace drs
