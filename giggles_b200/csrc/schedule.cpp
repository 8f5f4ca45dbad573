// Checkpoint schedule — see schedule.hpp.
#include "schedule.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>

namespace gg {
namespace {

constexpr uint64_t kAlign = 256;
inline uint64_t up(uint64_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

// A planned range: a leaf, a split into parts (ends[] = last column of every part, ascending, the last one = hi), or
// the quadratic last resort.  `cost` = bytes of backward columns produced while processing the range.
struct Node {
    uint32_t lo = 0, hi = 0, depth = 1;
    enum Kind : uint8_t { LEAF, SPLIT, QUAD, EXACT } kind = LEAF;
    double cost = 0;
    uint64_t avail = 0;      // EXACT: bytes the range may use beside its (resident) last backward column
    bool virt = false;       // EXACT: the last backward column is the all-ones column and is re-created where needed
    std::vector<uint32_t> ends;
    std::vector<Node> parts;
};

struct Builder {
    const std::vector<uint64_t> &colb, &fgb;
    std::vector<uint64_t> pcol, pfg;   // prefix sums of the aligned sizes
    uint64_t budget, top = 0, high = 0;
    uint64_t alpha[2] = {0, 0};
    Schedule &s;
    uint32_t max_depth = 0;

    Builder(const std::vector<uint64_t> &cb, const std::vector<uint64_t> &fb, uint64_t bud, Schedule &out)
        : colb(cb), fgb(fb), budget(bud), s(out) {
        const size_t C = cb.size();
        pcol.assign(C + 1, 0);
        pfg.assign(C + 1, 0);
        for (size_t c = 0; c < C; ++c) {
            pcol[c + 1] = pcol[c] + up(cb[c]);
            pfg[c + 1] = pfg[c] + fb[c];     // tables are packed inside one region
        }
    }
    uint64_t alloc(uint64_t n) {
        uint64_t off = top;
        top += up(n);
        if (top > budget)
            throw std::runtime_error("state budget too small for this problem: need at least " + std::to_string(top) +
                                     " bytes, budget is " + std::to_string(budget));
        high = std::max(high, top);
        return off;
    }
    // bytes a leaf over [lo, hi] allocates: betas lo..hi-1 and the F/G tables lo..hi
    uint64_t leaf_need(uint32_t lo, uint32_t hi) const { return (pcol[hi] - pcol[lo]) + up(pfg[hi + 1] - pfg[lo]); }
    // bytes of the two rolling backward columns and their tables while [lo, hi] is swept
    uint64_t roll_need(uint32_t lo, uint32_t hi) const {
        uint64_t maxcol = 0, maxfg = 0;
        for (uint32_t c = lo; c <= hi; ++c) {
            maxcol = std::max(maxcol, up(colb[c]));
            maxfg = std::max(maxfg, up(fgb[c]));
        }
        return 2 * maxcol + 2 * maxfg;
    }
    double span_bytes(uint32_t lo, uint32_t hi) const { return lo > hi ? 0.0 : (double)(pcol[hi + 1] - pcol[lo]); }

    // ------------------------------------------------------------------ planning (no ops emitted)
    // smallest s >= lo such that [s, e] is a leaf within cap; false if not even [e, e] is
    bool leaf_start(uint32_t lo, uint32_t e, uint64_t cap, uint32_t &s0) const {
        if (leaf_need(e, e) > cap) return false;
        uint32_t a = lo, b = e;
        while (a < b) {
            const uint32_t m = a + (b - a) / 2;
            if (leaf_need(m, e) <= cap) b = m; else a = m + 1;
        }
        s0 = a;
        return true;
    }

    // Two-level plan of [lo, hi]: one backward sweep that keeps checkpoints, then leaves left to right.  Part j is
    // processed while the checkpoints of parts j..m-2 are still resident (they are released one by one), so parts grow
    // towards the right: built right to left, every part the longest leaf that fits beside what is resident then.  A
    // checkpoint costs every part to its left its size, so small columns are preferred as part ends: the end p of the
    // next part maximises (columns gained) - mu * (columns still to cover) * bytes(p).
    bool plan_two_level(uint32_t lo, uint32_t hi, uint64_t avail, double mu, uint64_t roll, std::vector<uint32_t> &ends,
                        uint64_t *ck_bytes = nullptr) const {
        ends.assign(1, hi);
        uint64_t used = 0;
        uint32_t e = hi;
        for (;;) {
            if (used > avail) return false;
            uint32_t s0;
            if (!leaf_start(lo, e, avail - used, s0)) return false;
            if (s0 <= lo) break;
            uint32_t best = e - 1;
            double best_score = -1e300;
            for (uint32_t p = e - 1;; --p) {
                const double score = (double)(e - p) - mu * (double)(p - lo + 1) * (double)up(colb[p]);
                if (score > best_score) { best_score = score; best = p; }
                if (p == s0 - 1) break;
            }
            used += up(colb[best]);
            if (used + roll > avail) return false;
            ends.push_back(best);
            e = best;
        }
        std::reverse(ends.begin(), ends.end());
        if (ck_bytes) *ck_bytes = used;
        return true;
    }

    // The sweep of a two-level plan stops at the end of the FIRST part (its columns are only ever produced by its own
    // leaf), so the first part should be as long as the memory left beside all checkpoints allows.  Given a feasible
    // plan, move the end p0 of the first part to the right as far as [lo, p0] stays a leaf beside the checkpoint at p0
    // and those of a two-level plan of (p0, hi], preferring a small column for p0.
    void anchor_first_part(uint32_t lo, uint32_t hi, uint64_t avail, double mu, uint64_t roll, std::vector<uint32_t> &ends) const {
        if (ends.size() < 2) return;
        auto feasible = [&](uint32_t p0, std::vector<uint32_t> &right) {
            uint64_t kr = 0;
            if (!plan_two_level(p0 + 1, hi, avail, mu, roll, right, &kr)) return false;
            const uint64_t k = kr + up(colb[p0]);
            return k + roll <= avail && leaf_need(lo, p0) + k <= avail;
        };
        uint32_t a = ends[0], b = hi - 1;        // a is feasible by construction
        std::vector<uint32_t> right, best_right;
        while (a < b) {
            const uint32_t m = a + (b - a + 1) / 2;
            if (feasible(m, right)) { a = m; best_right = right; } else b = m - 1;
        }
        if (a == ends[0]) return;
        // a smaller column a little to the left is worth a few columns of sweep
        uint32_t p0 = a;
        const uint32_t back = std::min<uint32_t>((a - ends[0]) / 4, 64);
        for (uint32_t p = a; p + back >= a && p > ends[0]; --p)
            if (up(colb[p]) < up(colb[p0]) && (double)(a - p) * (double)up(colb[a]) < 4.0 * (double)(up(colb[p0]) - up(colb[p]))) p0 = p;
        if (p0 != a && !feasible(p0, best_right)) { p0 = a; if (!feasible(p0, best_right)) return; }
        if (best_right.empty() && !feasible(p0, best_right)) return;
        ends.assign(1, p0);
        ends.insert(ends.end(), best_right.begin(), best_right.end());
    }

    // Parts of about `len` columns, each ending at the smallest column within a quarter of `len` of the even split.
    void even_ends(uint32_t lo, uint32_t hi, uint32_t len, std::vector<uint32_t> &ends) const {
        ends.assign(1, hi);
        uint32_t e = hi;
        const uint32_t slack = std::max<uint32_t>(len / 4, 1);
        while (e - lo + 1 > len + len / 2) {
            const uint32_t mid = e - len;                               // nominal end of the next part to the left
            const uint32_t a = mid > lo + slack ? mid - slack : lo, b = std::min(e - 1, mid + slack);
            uint32_t best = b;
            for (uint32_t p = a; p <= b; ++p)
                if (colb[p] < colb[best] || (colb[p] == colb[best] && (p > mid ? p - mid : mid - p) < (best > mid ? best - mid : mid - best))) best = p;
            ends.push_back(best);
            e = best;
        }
        std::reverse(ends.begin(), ends.end());
    }

    bool plan(uint32_t lo, uint32_t hi, uint64_t avail, uint32_t depth, Node &out, bool shallow_only = false) const {
        out = Node();
        out.lo = lo; out.hi = hi; out.depth = depth;
        if (leaf_need(lo, hi) <= avail) {
            out.kind = Node::LEAF;
            out.cost = span_bytes(lo, hi > lo ? hi - 1 : lo) * (hi > lo ? 1.0 : 0.0);
            return true;
        }
        if (lo == hi) return false;
        const uint64_t roll = roll_need(lo, hi);
        const uint32_t n = hi - lo + 1;
        const double mean_col = span_bytes(lo, hi) / n;
        // (1) two levels, if the range is short enough for them: every column at most twice
        {
            const double leaf_cols = std::max(1.0, (double)avail / std::max(mean_col, 1.0));
            for (double m : {1.0, 4.0, 0.25, 0.0}) {
                std::vector<uint32_t> ends;
                if (!plan_two_level(lo, hi, avail, m / (leaf_cols * mean_col), roll, ends)) continue;
                anchor_first_part(lo, hi, avail, m / (leaf_cols * mean_col), roll, ends);
                out.kind = Node::SPLIT;
                // the sweep produces columns ends[0] .. hi-1, the leaves every column but their own last one
                out.cost = (ends.size() > 1 ? span_bytes(ends[0], hi - 1) : 0.0) + span_bytes(lo, hi);
                for (uint32_t e : ends) out.cost -= (double)up(colb[e]);
                out.ends.swap(ends);
                return true;                 // parts stay empty: all leaves, built by the emitter
            }
        }
        // (2) more levels: parts of about n/m columns ending at small columns, each planned recursively.  The number of
        //     parts is searched on a geometric grid from the largest the checkpoint memory allows downwards: more parts
        //     = fewer leaves per part = a larger share of columns that the part's own sweep skips, until the part-end
        //     checkpoints crowd out the leaves.  Plans whose parts all get by with two levels are preferred; deeper
        //     ones are costed only when there is no such plan.
        if (shallow_only) return false;
        std::vector<std::vector<uint32_t>> cands;
        for (uint32_t m = 2; m <= n / 2 && m <= (1u << 16); m = m + std::max<uint32_t>(1, m / 4)) {
            std::vector<uint32_t> ends;
            even_ends(lo, hi, std::max<uint32_t>(2, n / m), ends);
            if (ends.size() < 2) continue;
            uint64_t ck = 0;
            for (size_t j = 0; j + 1 < ends.size(); ++j) ck += up(colb[ends[j]]);
            if (ck + roll > avail) break;               // more parts only need more checkpoint memory
            cands.push_back(std::move(ends));
        }
        bool found = false, found_shallow = false;
        uint32_t worse_in_a_row = 0, deep_costed = 0;
        for (size_t ci = cands.size(); ci-- > 0;) {
            const std::vector<uint32_t> &ends = cands[ci];
            Node cand;
            cand.lo = lo; cand.hi = hi; cand.depth = depth; cand.kind = Node::SPLIT;
            cand.ends = ends;
            cand.parts.resize(ends.size());
            cand.cost = span_bytes(ends[0], hi - 1);    // the sweep produces columns ends[0] .. hi-1
            uint64_t resident = 0;
            for (size_t j = 0; j + 1 < ends.size(); ++j) resident += up(colb[ends[j]]);
            bool ok = true, deep = false;
            uint32_t s0 = lo;
            for (size_t j = 0; j < ends.size() && ok; ++j) {
                Node &pn = cand.parts[j];
                const uint64_t pav = avail - resident;
                // once a plan whose parts get by with two levels exists, deeper parts are not worth costing
                ok = plan(s0, ends[j], pav, depth + 1, pn, found_shallow);
                if (ok) {
                    cand.cost += pn.cost;
                    if (pn.kind == Node::SPLIT && !pn.parts.empty()) deep = true;
                    if (pn.kind == Node::QUAD) deep = true;
                }
                if (j + 1 < ends.size()) resident -= up(colb[ends[j]]);
                s0 = ends[j] + 1;
                if (found && cand.cost >= out.cost) ok = false;       // already worse than the best plan
            }
            if (std::getenv("GG_SCHED_TRACE") && depth <= 2)
                std::fprintf(stderr, "[sched] depth %u [%u,%u] cand m=%zu ok=%d deep=%d cost/span=%.3f\n", depth, lo, hi, ends.size(), (int)ok, (int)deep, cand.cost / span_bytes(lo, hi));
            if (ok && (!found || cand.cost < out.cost)) {
                out = std::move(cand);
                found = true;
                found_shallow = found_shallow || !deep;
                worse_in_a_row = 0;
            } else if (found) {
                if (++worse_in_a_row >= 4) break;
            }
            if (ok && deep && ++deep_costed >= 3) break;
        }
        if (found) return true;
        // (3) last resort (columns of a fifth of HBM): two rolling columns only, the backward chain is recomputed from
        //     the range's last column for every forward column.  Quadratic in the range length.
        if (avail >= roll) {
            out.kind = Node::QUAD;
            out.ends.clear();
            out.parts.clear();
            double cost = 0;
            for (uint32_t c = lo; c < hi; ++c) cost += span_bytes(c, hi - 1);
            out.cost = cost;
            return true;
        }
        return false;
    }

    // ------------------------------------------------------------------ exact planner (few, huge columns)
    // Ranges of a few dozen columns of which only a handful fit (2^20 bipartitions at R = 88: 33 GB per column, five
    // in HBM) are planned by exhaustive recursion instead of the heuristics above:
    //   T(lo, hi, avail) = bytes(lo..hi-1)                                            if the range is a leaf
    //                    = min_p  bytes(p..hi-1) + T(lo, p, avail - col_p) + T(p+1, hi, avail)
    // i.e. sweep from beta_hi down to column p, keep beta_p, do the left part (whose own sweep carries on from
    // beta_p: several checkpoints per sweep come out as left-nested splits), drop beta_p, do the right part.  The
    // sweep alternates between a temporary and the destination itself (an output two steps before the last is dead
    // by the time the last is written); the temporary is the alpha buffer the next forward step will write, idle
    // while backward columns are being recomputed, so a sweep costs no room beyond its destination.  `virt`: the
    // range ends at the chain's last column, whose beta is all ones: instead of keeping it resident from the first
    // step of the sweep to the leaf of the last part, every use re-creates it (one write-only launch, costed at half
    // a step) — with five columns of room that is the difference between no spare checkpoint and one.
    struct XKey {
        uint32_t lo, hi; uint64_t avail; bool virt;
        bool operator<(const XKey &o) const {
            if (lo != o.lo) return lo < o.lo;
            if (hi != o.hi) return hi < o.hi;
            if (avail != o.avail) return avail < o.avail;
            return virt < o.virt;
        }
    };
    struct XVal { double cost; int64_t p; };       // p = -1: leaf, -2: infeasible
    mutable std::map<XKey, XVal> xmemo;
    static constexpr uint32_t kExactMaxCols = 64;
    static constexpr size_t kExactMaxStates = 400000;   // memo entries; beyond that the heuristics take over
    static constexpr uint64_t kExactMaxWork = 4000000;  // candidate splits looked at (a few ms); deterministic cut-off
    mutable uint64_t xwork = 0;
    struct ExactOverflow {};

    bool exact_eligible(uint32_t lo, uint32_t hi, uint64_t avail) const {
        if (hi - lo + 1 > kExactMaxCols) return false;
        uint64_t maxcol = 0;
        for (uint32_t c = lo; c <= hi; ++c) maxcol = std::max(maxcol, up(colb[c]));
        // columns of 4 GiB and more of which fewer than six fit.  (A SMALL column under a budget of a few columns — the
        // benchmark's way of imposing a full-length recompute ratio on a short window — is left to the heuristics: the
        // recursion costs tens of milliseconds when column sizes vary and would sit on the end-to-end path of every
        // chain.  GG_EXACT_MIN_COL overrides the size threshold; the tests set it to 0.)
        const char *e = std::getenv("GG_EXACT_MIN_COL");
        const uint64_t min_col = e ? std::strtoull(e, nullptr, 10) : (1ull << 32);
        return maxcol >= min_col && maxcol > 0 && avail / maxcol < 6;
    }
    // temporaries of the sweep that produces [hi if virt,] hi-1 .. p with beta_p as the destination
    void sweep_need(uint32_t p, uint32_t hi, bool virt, uint64_t &t1, uint64_t &t2, uint64_t &fg) const {
        t1 = t2 = 0;
        const uint64_t dst = up(colb[p]);
        const uint32_t first = virt ? hi : hi - 1;          // produced columns first, first-1, .., p
        for (uint32_t c = first;; --c) {
            const uint32_t k = c - p;                       // steps before the last one
            const uint64_t b = up(colb[c]);
            if (k & 1u) t1 = std::max(t1, b);
            else if (k >= 2 && b > dst) t2 = std::max(t2, b);
            if (c == p) break;
        }
        uint64_t maxfg = 0;
        for (uint32_t c = p; c <= hi; ++c) maxfg = std::max(maxfg, up(fgb[c]));
        fg = 2 * maxfg;
    }
    XVal exact(uint32_t lo, uint32_t hi, uint64_t avail, bool virt) const {
        const XKey key{lo, hi, avail, virt};
        auto it = xmemo.find(key);
        if (it != xmemo.end()) return it->second;
        if (xmemo.size() > kExactMaxStates) throw ExactOverflow{};
        XVal best{1e300, -2};
        const double init_cost = virt ? 0.5 * (double)up(colb[hi]) : 0.0;
        const uint64_t own = virt ? up(colb[hi]) : 0;
        if (leaf_need(lo, hi) + own <= avail) {
            best = XVal{(hi > lo ? span_bytes(lo, hi - 1) : 0.0) + init_cost, -1};
        } else if (hi > lo) {
            // p from hi-1 downwards; running maxima over the produced columns above p, by parity of the column index:
            // those an odd number of steps before the last land in the temporary, the others in the destination
            // itself unless they are larger than it (sweep_need, in O(1) per p)
            uint64_t mpar[2] = {0, 0}, maxfg = up(fgb[hi]);
            if (virt) mpar[hi & 1u] = up(colb[hi]);
            for (uint32_t p = hi - 1;; --p) {
                if (++xwork > kExactMaxWork) throw ExactOverflow{};
                maxfg = std::max(maxfg, up(fgb[p]));
                const uint64_t ck = up(colb[p]);
                const uint64_t same = mpar[p & 1u];
                const uint64_t t2 = same > ck ? same : 0;
                if (ck + t2 + 2 * maxfg <= avail) {       // t1 lives in the idle alpha buffer (see run_exact)
                    const double sweep = span_bytes(p, hi - 1) + init_cost;
                    if (sweep < best.cost) {
                        const XVal l = exact(lo, p, avail - ck, false);
                        if (l.p != -2 && sweep + l.cost < best.cost) {
                            const XVal r = exact(p + 1, hi, avail, virt);
                            if (r.p != -2 && sweep + l.cost + r.cost < best.cost) best = XVal{sweep + l.cost + r.cost, (int64_t)p};
                        }
                    }
                }
                mpar[p & 1u] = std::max(mpar[p & 1u], ck);
                if (p == lo) break;
            }
        }
        xmemo[key] = best;
        return best;
    }

    // ------------------------------------------------------------------ emission
    void emit(uint32_t lo, uint32_t hi, uint64_t region) {
        Op o{};
        o.kind = OP_EMIT; o.c = lo; o.hi = hi; o.out_off = region;
        s.ops.push_back(o);
        ++s.n_emit;
    }
    void bwd(uint32_t c_out, uint64_t in_off, uint64_t out_off, uint64_t fg_in, uint64_t fg_out) {
        Op o{};
        o.kind = OP_BWD; o.c = c_out; o.in_off = in_off; o.out_off = out_off; o.fg_in_off = fg_in; o.fg_out_off = fg_out;
        s.ops.push_back(o);
        ++s.n_bwd;
        s.bwd_col_bytes += colb[c_out];
    }
    void fwd(uint32_t c, uint64_t beta_off, uint64_t fg_out) {
        Op o{};
        o.kind = OP_FWD; o.c = c; o.in_off = alpha[(c + 1) & 1]; o.out_off = alpha[c & 1]; o.beta_off = beta_off; o.fg_out_off = fg_out;
        s.ops.push_back(o);
        ++s.n_fwd;
    }

    void bwd_init(uint32_t c, uint64_t out_off, uint64_t fg_out) {
        Op o{};
        o.kind = OP_BWD_INIT; o.c = c; o.out_off = out_off; o.fg_out_off = fg_out;
        s.ops.push_back(o);
        ++s.n_bwd;
    }

    void leaf(uint32_t lo, uint32_t hi, uint64_t beta_hi, bool virt = false) {
        const uint64_t mark = top;
        const uint64_t fg = alloc(pfg[hi + 1] - pfg[lo]);
        emit(lo, hi, fg);
        if (virt) {
            beta_hi = alloc(colb[hi]);
            bwd_init(hi, beta_hi, fg + (pfg[hi] - pfg[lo]));
        }
        std::vector<uint64_t> offs(hi - lo + 1);
        offs[hi - lo] = beta_hi;
        for (uint32_t c = hi; c > lo; --c) {
            offs[c - 1 - lo] = alloc(colb[c - 1]);
            bwd(c - 1, offs[c - lo], offs[c - 1 - lo], fg + (pfg[c] - pfg[lo]), fg + (pfg[c - 1] - pfg[lo]));
        }
        for (uint32_t c = lo; c <= hi; ++c) fwd(c, offs[c - lo], fg + (pfg[c] - pfg[lo]));
        top = mark;
    }

    // emits the plan exact() found (its argmin is re-read from the memo)
    void run_exact(uint32_t lo, uint32_t hi, uint64_t avail, bool virt, uint64_t beta_hi, uint32_t depth) {
        max_depth = std::max(max_depth, depth);
        const XVal x = exact(lo, hi, avail, virt);
        if (x.p == -2) throw std::runtime_error("state budget too small for this problem (exact plan lost)");
        if (x.p == -1) {
            leaf(lo, hi, beta_hi, virt);
            return;
        }
        max_depth = std::max(max_depth, depth + 1);
        const uint32_t p = (uint32_t)x.p;
        const uint64_t mark0 = top;
        const uint64_t ck = alloc(colb[p]);
        const uint64_t mark1 = top;
        uint64_t t1, t2, fgn;
        sweep_need(p, hi, virt, t1, t2, fgn);
        // every forward column before lo is done and alpha_{lo-1} sits in alpha[(lo+1)&1]: the other alpha buffer (the
        // one forward step lo will write) is idle until then and serves as the sweep's temporary
        const uint64_t T1 = alpha[lo & 1], T2 = t2 ? alloc(t2) : 0;
        uint64_t fbuf[2];
        fbuf[0] = alloc(fgn / 2); fbuf[1] = alloc(fgn / 2);
        auto buf_of = [&](uint32_t c) {
            const uint32_t k = c - p;
            if (k == 0) return ck;
            if (k & 1u) return T1;
            return up(colb[c]) <= up(colb[p]) ? ck : T2;
        };
        emit(hi, hi, fbuf[hi & 1]);
        uint64_t cur = beta_hi;
        if (virt) {
            cur = buf_of(hi);
            bwd_init(hi, cur, fbuf[hi & 1]);
        }
        for (uint32_t c = hi; c > p; --c) {
            emit(c - 1, c - 1, fbuf[(c - 1) & 1]);
            const uint64_t dst = buf_of(c - 1);
            bwd(c - 1, cur, dst, fbuf[c & 1], fbuf[(c - 1) & 1]);
            cur = dst;
        }
        top = mark1;
        run_exact(lo, p, avail - up(colb[p]), false, ck, depth + 1);
        top = mark0;
        run_exact(p + 1, hi, avail, virt, beta_hi, depth + 1);
    }

    void run(const Node &nd, uint64_t beta_hi) {
        max_depth = std::max(max_depth, nd.depth);
        const uint32_t lo = nd.lo, hi = nd.hi;
        if (nd.kind == Node::EXACT) {
            run_exact(lo, hi, nd.avail, nd.virt, beta_hi, nd.depth);
            return;
        }
        if (nd.kind == Node::LEAF) {
            leaf(lo, hi, beta_hi);
            return;
        }
        uint64_t maxcol = 0, maxfg = 0;
        for (uint32_t c = lo; c <= hi; ++c) {
            maxcol = std::max(maxcol, up(colb[c]));
            maxfg = std::max(maxfg, up(fgb[c]));
        }
        const uint64_t mark = top;
        if (nd.kind == Node::QUAD) {
            max_depth = std::max(max_depth, nd.depth + 1);   // it recomputes: never report "everything fitted"
            uint64_t rbuf[2], fbuf[2];
            rbuf[0] = alloc(maxcol); rbuf[1] = alloc(maxcol);
            fbuf[0] = alloc(maxfg); fbuf[1] = alloc(maxfg);
            for (uint32_t c = lo; c <= hi; ++c) {
                uint64_t cur = beta_hi;
                emit(hi, hi, fbuf[hi & 1]);
                for (uint32_t x = hi; x > c; --x) {
                    emit(x - 1, x - 1, fbuf[(x - 1) & 1]);
                    bwd(x - 1, cur, rbuf[(x - 1) & 1], fbuf[x & 1], fbuf[(x - 1) & 1]);
                    cur = rbuf[(x - 1) & 1];
                }
                fwd(c, cur, fbuf[c & 1]);
            }
            top = mark;
            return;
        }
        const std::vector<uint32_t> &ends = nd.ends;
        const uint32_t m = (uint32_t)ends.size();
        // checkpoints are stacked so that the one consumed first (the leftmost part's) lies on top and every part
        // gets back the room of the parts before it
        std::vector<uint64_t> ck(m, 0);
        for (int32_t j = (int32_t)m - 2; j >= 0; --j) ck[j] = alloc(colb[ends[j]]);
        ck[m - 1] = beta_hi;
        if (m > 1) {
            const uint64_t mark2 = top;
            uint64_t rbuf[2], fbuf[2];
            rbuf[0] = alloc(maxcol); rbuf[1] = alloc(maxcol);
            fbuf[0] = alloc(maxfg); fbuf[1] = alloc(maxfg);
            uint64_t cur = beta_hi;
            emit(hi, hi, fbuf[hi & 1]);
            int32_t j = (int32_t)m - 2;
            for (uint32_t c = hi; c > ends[0]; --c) {
                emit(c - 1, c - 1, fbuf[(c - 1) & 1]);
                uint64_t dst = rbuf[(c - 1) & 1];
                if (j >= 0 && ends[j] == c - 1) dst = ck[j--];
                bwd(c - 1, cur, dst, fbuf[c & 1], fbuf[(c - 1) & 1]);
                cur = dst;
            }
            top = mark2;
        }
        uint32_t s0 = lo;
        for (uint32_t j = 0; j < m; ++j) {
            if (!nd.parts.empty()) {
                run(nd.parts[j], ck[j]);
            } else {
                max_depth = std::max(max_depth, nd.depth + 1);
                leaf(s0, ends[j], ck[j]);
            }
            if (j + 1 < m) top = ck[j];       // release this part's checkpoint
            s0 = ends[j] + 1;
        }
        top = mark;
    }
};

}  // namespace

void build_schedule(const std::vector<uint64_t> &colb, const std::vector<uint64_t> &fgb, uint64_t budget, Schedule &out) {
    out = Schedule();
    const uint32_t C = (uint32_t)colb.size();
    if (C == 0) return;
    Builder b(colb, fgb, budget, out);
    uint64_t maxcol = 0;
    for (uint64_t x : colb) maxcol = std::max(maxcol, x);
    b.alpha[0] = b.alloc(maxcol);
    b.alpha[1] = b.alloc(maxcol);
    // few, huge columns: is the chain cheaper when its last (all-ones) backward column is re-created on demand
    // instead of staying resident?  (see the exact planner)
    if (C > 1 && b.exact_eligible(0, C - 1, budget > b.top ? budget - b.top : 0) && budget > b.top) {
        const uint64_t avail_v = budget - b.top;
        const uint64_t res = up(colb[C - 1]) + up(fgb[C - 1]);
        bool use_virt = false, use_exact = true;
        try {
            const auto xv = b.exact(0, C - 1, avail_v, true);
            double cost_r = 1e300;
            if (avail_v > res) {
                const auto xr = b.exact(0, C - 1, avail_v - res, false);
                if (xr.p != -2) cost_r = xr.cost + 0.5 * (double)up(colb[C - 1]);
            }
            use_virt = xv.p != -2 && xv.cost < cost_r;
            if (!use_virt && cost_r >= 1e300) use_exact = false;
        } catch (const Builder::ExactOverflow &) {
            b.xmemo.clear();
            use_exact = false;
        }
        if (use_exact && !use_virt) {
            // resident last column, exact plan
            const uint64_t beta_last = b.alloc(colb[C - 1]);
            const uint64_t fg_last = b.alloc(fgb[C - 1]);
            b.emit(C - 1, C - 1, fg_last);
            b.bwd_init(C - 1, beta_last, fg_last);
            b.run_exact(0, C - 1, budget - b.top, false, beta_last, 1);
            out.arena_bytes = b.high;
            out.levels = b.max_depth;
            return;
        }
        if (use_virt) {
            b.run_exact(0, C - 1, avail_v, true, 0, 1);
            out.arena_bytes = b.high;
            out.levels = b.max_depth;
            return;
        }
    }
    const uint64_t beta_last = b.alloc(colb[C - 1]);
    const uint64_t fg_last = b.alloc(fgb[C - 1]);
    b.emit(C - 1, C - 1, fg_last);
    Op o{};
    o.kind = OP_BWD_INIT; o.c = C - 1; o.out_off = beta_last; o.fg_out_off = fg_last;
    out.ops.push_back(o);
    ++out.n_bwd;
    Node root;
    if (!b.plan(0, C - 1, budget - b.top, 1, root))
        throw std::runtime_error("state budget too small for this problem: " + std::to_string(budget - b.top) +
                                 " bytes left for the backward columns, the largest column takes " + std::to_string(maxcol));
    b.run(root, beta_last);
    out.arena_bytes = b.high;
    out.levels = b.max_depth;
}

void validate_schedule(const std::vector<uint64_t> &colb, const std::vector<uint64_t> &fgb, const Schedule &s) {
    // live buffers: offset -> (end, tag) ; tag = kind*2^32 + column, kinds: 1 alpha, 2 beta, 3 fg
    struct Buf { uint64_t end; uint64_t tag; };
    std::map<uint64_t, Buf> live;
    auto tag = [](uint64_t kind, uint64_t c) { return (kind << 32) | c; };
    auto write = [&](uint64_t off, uint64_t bytes, uint64_t t) {
        const uint64_t end = off + bytes;
        auto it = live.lower_bound(off);
        if (it != live.begin()) {
            auto p = std::prev(it);
            if (p->second.end > off) it = p;
        }
        while (it != live.end() && it->first < end) it = live.erase(it);   // overwritten buffers die
        live[off] = Buf{end, t};
    };
    auto expect = [&](uint64_t off, uint64_t t, const char *what, uint32_t c) {
        auto it = live.find(off);
        if (it == live.end() || it->second.tag != t)
            throw std::runtime_error(std::string("schedule self-check failed: ") + what + " of column " + std::to_string(c) +
                                     " is not resident where the op expects it");
    };
    std::vector<uint64_t> pfg(fgb.size() + 1, 0);
    for (size_t c = 0; c < fgb.size(); ++c) pfg[c + 1] = pfg[c] + fgb[c];
    uint32_t next_fwd = 0;
    for (const Op &o : s.ops) {
        switch (o.kind) {
        case OP_RUN:
            break;   // engine-side fusion happens after validation
        case OP_EMIT:
            for (uint32_t c = o.c; c <= o.hi; ++c) write(o.out_off + (pfg[c] - pfg[o.c]), fgb[c], tag(3, c));
            break;
        case OP_BWD_INIT:
            expect(o.fg_out_off, tag(3, o.c), "F/G table", o.c);
            write(o.out_off, colb[o.c], tag(2, o.c));
            break;
        case OP_BWD:
            expect(o.in_off, tag(2, o.c + 1), "beta", o.c + 1);
            expect(o.fg_in_off, tag(3, o.c + 1), "F/G table", o.c + 1);
            expect(o.fg_out_off, tag(3, o.c), "F/G table", o.c);
            write(o.out_off, colb[o.c], tag(2, o.c));
            // the write must not have destroyed its own inputs
            expect(o.in_off, tag(2, o.c + 1), "beta (clobbered)", o.c + 1);
            expect(o.fg_in_off, tag(3, o.c + 1), "F/G table (clobbered)", o.c + 1);
            expect(o.fg_out_off, tag(3, o.c), "F/G table (clobbered)", o.c);
            break;
        case OP_FWD:
            if (o.c != next_fwd) throw std::runtime_error("schedule self-check failed: forward columns out of order");
            ++next_fwd;
            if (o.c > 0) expect(o.in_off, tag(1, o.c - 1), "alpha", o.c - 1);
            expect(o.beta_off, tag(2, o.c), "beta", o.c);
            expect(o.fg_out_off, tag(3, o.c), "F/G table", o.c);
            write(o.out_off, colb[o.c], tag(1, o.c));
            if (o.c > 0) expect(o.in_off, tag(1, o.c - 1), "alpha (clobbered)", o.c - 1);
            expect(o.beta_off, tag(2, o.c), "beta (clobbered)", o.c);
            expect(o.fg_out_off, tag(3, o.c), "F/G table (clobbered)", o.c);
            break;
        }
    }
    if (next_fwd != colb.size()) throw std::runtime_error("schedule self-check failed: not every column got its forward step");
}

}  // namespace gg
