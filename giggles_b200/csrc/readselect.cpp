// Read selection (coverage capping) — host C++ behind gg_readselection().
//
// Replaces the reference's readselection() (giggles/readselect.pyx:231-271, called from
// giggles/utils.py:71-76 before every GenotypeHMM): greedy slices that each try to cover
// every variant once with the best-scoring reads (readselect.pyx:106-170), optional
// bridging reads that join the slice's connected components (:203-229), repeated until
// no read is undecided, all under a per-variant coverage cap (giggles/coverage.py).
//
// The selected index set must equal the reference's exactly: it decides which reads are
// active in every HMM column.  Equal scores are common, and the reference ranks them by
// the order in which it walks Python sets and one libstdc++ unordered_set, so this port
// keeps those containers' iteration orders (PySetInt in pyset.hpp; std::unordered_set
// used with the same call sequence) and the reference's own binary heap discipline
// (giggles/priorityqueue.pyx:64-190).  Everything whose order cannot matter is a flat
// array.  Host only, no CUDA.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/giggles_b200.h"
#include "pyset.hpp"

namespace {

using gg::PySetInt;

struct Score {
    int s[3];
};
inline bool lower(const Score &a, const Score &b) {   // priorityqueue.pyx:12-22 on equal-length vectors
    for (int i = 0; i < 3; ++i) {
        if (a.s[i] < b.s[i]) return true;
        if (a.s[i] > b.s[i]) return false;
    }
    return false;
}

// Max-heap with item -> slot index; sift rules of priorityqueue.pyx:98-128,142-190.
class Heap {
  public:
    explicit Heap(uint32_t n_items) : where_(n_items, -1) {}
    bool empty() const { return h_.empty(); }
    void push(const Score &sc, int item) {
        int i = (int)h_.size();
        h_.push_back({sc, item});
        where_[item] = i;
        sift_up(i);
    }
    int pop(Score *sc = nullptr) {
        Ent first = h_[0];
        if (sc) *sc = first.sc;
        if (h_.size() == 1) {
            h_.pop_back();
            where_[first.item] = -1;
        } else {
            h_[0] = h_.back();
            h_.pop_back();
            where_[h_[0].item] = 0;
            where_[first.item] = -1;
            sift_down(0);
        }
        return first.item;
    }
    const Score *score_of(int item) const { return where_[item] < 0 ? nullptr : &h_[where_[item]].sc; }
    void change(int item, const Score &ns) {
        int i = where_[item];
        Score old = h_[i].sc;
        h_[i].sc = ns;
        if (lower(old, ns)) sift_up(i);
        else sift_down(i);
    }

  private:
    struct Ent {
        Score sc;
        int item;
    };
    std::vector<Ent> h_;
    std::vector<int> where_;
    void swap_(int a, int b) {
        std::swap(h_[a], h_[b]);
        where_[h_[a].item] = a;
        where_[h_[b].item] = b;
    }
    void sift_up(int i) {
        while (i > 0) {
            int p = (i - 1) / 2;
            if (!lower(h_[p].sc, h_[i].sc)) break;
            swap_(p, i);
            i = p;
        }
    }
    void sift_down(int i) {
        const int n = (int)h_.size();
        for (;;) {
            int l = 2 * i + 1, r = 2 * i + 2, c;
            if (r < n) c = lower(h_[l].sc, h_[r].sc) ? r : l;
            else if (l < n) c = l;
            else break;
            if (!lower(h_[i].sc, h_[c].sc)) break;
            swap_(c, i);
            i = c;
        }
    }
};

// a set whose order never matters: membership + members
struct FlatSet {
    std::vector<uint8_t> in;
    std::vector<int64_t> items;
    explicit FlatSet(size_t n) : in(n, 0) {}
    void add(int64_t k) {
        if (!in[k]) {
            in[k] = 1;
            items.push_back(k);
        }
    }
    size_t size() const { return items.size(); }
    bool contains(int64_t k) const { return in[k] != 0; }
    template <class F> void for_each(F &&f) const {
        for (int64_t k : items) f(k);
    }
};

struct Selector {
    const gg_readset_view &rs;
    uint32_t max_cov;
    bool bridging;
    uint32_t N;
    std::vector<int32_t> positions;            // sorted unique (readset.cpp:56-64)
    std::vector<uint32_t> vidx;                // per read variant: index into positions
    std::vector<uint64_t> v2r_off;             // variant -> reads covering it, in read order (readselect.pyx:19-34)
    std::vector<uint32_t> v2r;
    std::vector<int32_t> coverage;             // coverage.py
    std::vector<uint8_t> selected;

    uint64_t lo(uint32_t r) const { return rs.read_variant_offset[r]; }
    uint64_t hi(uint32_t r) const { return rs.read_variant_offset[r + 1]; }

    Selector(const gg_readset_view &v, uint32_t mc, bool br) : rs(v), max_cov(mc), bridging(br), N(v.n_reads) {
        const uint64_t nv = N ? rs.read_variant_offset[N] : 0;
        positions.assign(rs.variant_position, rs.variant_position + nv);
        std::sort(positions.begin(), positions.end());
        positions.erase(std::unique(positions.begin(), positions.end()), positions.end());
        vidx.resize(nv);
        std::vector<uint64_t> cnt(positions.size() + 1, 0);
        for (uint64_t i = 0; i < nv; ++i) {
            vidx[i] = (uint32_t)(std::lower_bound(positions.begin(), positions.end(), rs.variant_position[i]) - positions.begin());
            ++cnt[vidx[i] + 1];
        }
        for (size_t i = 1; i < cnt.size(); ++i) cnt[i] += cnt[i - 1];
        v2r_off = cnt;
        v2r.resize(nv);
        std::vector<uint64_t> cur(cnt.begin(), cnt.end() - 1);
        for (uint32_t r = 0; r < N; ++r)
            for (uint64_t i = lo(r); i < hi(r); ++i) v2r[cur[vidx[i]]++] = r;
        coverage.assign(positions.size(), 0);
        selected.assign(N, 0);
    }

    // readselect.pyx:54-94: (covered - gaps, covered - gaps, min quality)
    Score initial_score(uint32_t r) const {
        int minq = -1, good = 0;
        for (uint64_t i = lo(r); i < hi(r); ++i) {
            int q = rs.variant_quality[i];
            minq = (i == lo(r)) ? q : std::min(minq, q);
            ++good;
        }
        int span = (int)vidx[hi(r) - 1] - (int)vidx[lo(r)] + 1;
        int bad = (good != span) ? span - good : 0;
        return Score{{good - bad, good - bad, minq}};
    }

    Heap make_queue(const PySetInt &reads) const {    // readselect.pyx:97-103: pushed in set order
        Heap pq(N);
        reads.for_each([&](int64_t r) { pq.push(initial_score((uint32_t)r), (int)r); });
        return pq;
    }

    int max_coverage(uint32_t b, uint32_t e) const { return *std::max_element(coverage.begin() + b, coverage.begin() + e); }
    void add_read(uint32_t b, uint32_t e) {
        for (uint32_t i = b; i < e; ++i) ++coverage[i];
    }

    // readselect.pyx:106-170
    void slice(Heap &pq, FlatSet &in_slice, FlatSet &violating) {
        std::vector<uint8_t> covered(positions.size(), 0);
        std::unordered_set<int> fresh;    // positions newly covered by the read at hand; walked in this container's order (:152)
        while (!pq.empty()) {
            fresh.clear();
            const int r = pq.pop();
            bool covers_new = false;
            for (uint64_t i = lo(r); i < hi(r); ++i) {
                if (covered[vidx[i]]) continue;
                covers_new = true;
                fresh.insert(rs.variant_position[i]);
            }
            const uint32_t b = vidx[lo(r)], e = vidx[hi(r) - 1] + 1;
            if (max_coverage(b, e) >= (int)max_cov) {
                violating.add(r);
            } else if (covers_new) {
                add_read(b, e);
                in_slice.add(r);
                PySetInt touched;
                for (int pos : fresh) {
                    uint32_t v = (uint32_t)(std::lower_bound(positions.begin(), positions.end(), pos) - positions.begin());
                    covered[v] = 1;
                    touched.update_seq(v2r.begin() + v2r_off[v], v2r.begin() + v2r_off[v + 1]);
                }
                PySetInt todo = touched.difference(in_slice);
                todo.for_each([&](int64_t el) {
                    const Score *old = pq.score_of((int)el);
                    if (!old) return;
                    Score ns = *old;                       // :37-51: first component drops by the number of the
                    for (uint64_t i = lo(el); i < hi(el); ++i)   // read's variants that are NOT among the fresh ones
                        if (fresh.find(rs.variant_position[i]) == fresh.end()) --ns.s[0];
                    pq.change((int)el, ns);
                });
            }
        }
    }

    // union-find over variant indices (giggles/graph.py:15-62; only the partition matters here)
    std::vector<uint32_t> parent;
    uint32_t find(uint32_t x) {
        uint32_t r = x;
        while (parent[r] != r) r = parent[r];
        while (parent[x] != r) {
            uint32_t n = parent[x];
            parent[x] = r;
            x = n;
        }
        return r;
    }
    void merge(uint32_t a, uint32_t b) {
        a = find(a);
        b = find(b);
        if (a == b) return;
        if (a < b) parent[b] = a;
        else parent[a] = b;
    }

    // readselect.pyx:185-229
    void helper(PySetInt &undecided) {
        while (undecided.size() > 0) {
            Heap pq = make_queue(undecided);
            FlatSet in_slice(N), violating(N);
            slice(pq, in_slice, violating);
            for (int64_t r : in_slice.items) selected[r] = 1;
            undecided.difference_update(in_slice);
            undecided.difference_update(violating);

            parent.resize(positions.size());
            for (uint32_t i = 0; i < parent.size(); ++i) parent[i] = i;
            for (int64_t r : in_slice.items)
                for (uint64_t i = lo(r) + 1; i < hi(r); ++i) merge(vidx[lo(r)], vidx[i]);

            if (bridging) {
                Heap bq = make_queue(undecided);
                while (!bq.empty()) {
                    const int r = bq.pop();
                    uint32_t first_block = find(vidx[lo(r)]);
                    bool two_blocks = false;
                    for (uint64_t i = lo(r) + 1; i < hi(r); ++i) two_blocks |= find(vidx[i]) != first_block;
                    const uint32_t b = vidx[lo(r)], e = vidx[hi(r) - 1] + 1;
                    if (max_coverage(b, e) >= (int)max_cov) {
                        undecided.discard(r);
                        continue;
                    }
                    if (!two_blocks) continue;
                    selected[r] = 1;
                    add_read(b, e);
                    undecided.discard(r);
                    for (uint64_t i = lo(r) + 1; i < hi(r); ++i) merge(vidx[lo(r)], vidx[i]);
                }
            }
        }
    }
};

}  // namespace

extern "C" int gg_readselection(const gg_readset_view *rs, uint32_t max_cov, const int32_t *preferred_source_ids,
                                uint32_t n_preferred, int bridging, uint32_t *selected, uint32_t *n_selected, char *err,
                                size_t errlen) {
    auto fail = [&](int code, const std::string &m) {
        if (err && errlen) snprintf(err, errlen, "%s", m.c_str());
        return code;
    };
    try {
        if (!rs || !selected || !n_selected) return fail(1, "gg_readselection: null argument");
        *n_selected = 0;
        for (uint32_t r = 0; r < rs->n_reads; ++r)
            if (rs->read_variant_offset[r + 1] - rs->read_variant_offset[r] < 2)
                return fail(2, "readselection expects reads that cover at least two variants");   // readselect.pyx:249-252
        Selector S(*rs, max_cov, bridging != 0);
        PySetInt preferred;                                                                      // readselect.pyx:24-30
        if (preferred_source_ids)
            for (uint32_t r = 0; r < rs->n_reads; ++r)
                if (std::find(preferred_source_ids, preferred_source_ids + n_preferred, rs->read_source_id[r]) !=
                    preferred_source_ids + n_preferred)
                    preferred.add(r);
        PySetInt undecided;                                                                      // set(range(N)), :255
        for (uint32_t r = 0; r < rs->n_reads; ++r) undecided.add(r);
        if (preferred.size() > 0) {
            S.helper(preferred);                 // consumes `preferred`, as the reference's helper consumes its argument …
            undecided.difference_update(preferred);   // … so this subtracts the EMPTY set (:257-260): preferred reads get a second pass
        }
        S.helper(undecided);
        uint32_t n = 0;
        for (uint32_t r = 0; r < rs->n_reads; ++r)
            if (S.selected[r]) selected[n++] = r;
        *n_selected = n;
        return 0;
    } catch (const std::exception &e) {
        return fail(1, std::string("gg_readselection: ") + e.what());
    }
}

// Test hook: replays a script of set operations on eight PySetInt registers and dumps iteration orders, so that
// tests/test_readselect.py can compare them with the running interpreter's own `set` step by step.
// ops: 0 d | 1 d v (add) | 2 d v (discard) | 3 d n v.. (update list) | 4 d s (update set) | 5 d s (d = set(s)) |
//      6 d s (d -= s) | 7 d a b (d = a.difference(b)) | 8 d n (d = set(range(n))) | 9 s (dump: count, items..)
extern "C" int64_t gg_pyset_replay(const int64_t *ops, int64_t n, int64_t *out, int64_t cap) {
    PySetInt S[8];
    int64_t o = 0, i = 0;
    while (i < n) {
        const int64_t op = ops[i++];
        switch (op) {
        case 0: S[ops[i]] = PySetInt(); i += 1; break;
        case 1: S[ops[i]].add(ops[i + 1]); i += 2; break;
        case 2: S[ops[i]].discard(ops[i + 1]); i += 2; break;
        case 3: S[ops[i]].update_seq(ops + i + 2, ops + i + 2 + ops[i + 1]); i += 2 + ops[i + 1]; break;
        case 4: S[ops[i]].merge(S[ops[i + 1]]); i += 2; break;
        case 5: { PySetInt t = S[ops[i + 1]].copy(); S[ops[i]] = t; i += 2; break; }
        case 6: S[ops[i]].difference_update(S[ops[i + 1]]); i += 2; break;
        case 7: { PySetInt t = S[ops[i + 1]].difference(S[ops[i + 2]]); S[ops[i]] = t; i += 3; break; }
        case 8: { PySetInt t; for (int64_t k = 0; k < ops[i + 1]; ++k) t.add(k); S[ops[i]] = t; i += 2; break; }
        case 9: {
            std::vector<int64_t> v = S[ops[i]].items();
            i += 1;
            if (o + 1 + (int64_t)v.size() > cap) return -1;
            out[o++] = (int64_t)v.size();
            for (int64_t k : v) out[o++] = k;
            break;
        }
        default: return -2;
        }
    }
    return o;
}
