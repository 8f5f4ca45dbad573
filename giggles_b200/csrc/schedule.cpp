// Checkpoint schedule — see schedule.hpp.
#include "schedule.hpp"

#include <algorithm>
#include <map>
#include <stdexcept>
#include <string>

namespace gg {
namespace {

constexpr uint64_t kAlign = 256;
inline uint64_t up(uint64_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

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
    }
    void fwd(uint32_t c, uint64_t beta_off, uint64_t fg_out) {
        Op o{};
        o.kind = OP_FWD; o.c = c; o.in_off = alpha[(c + 1) & 1]; o.out_off = alpha[c & 1]; o.beta_off = beta_off; o.fg_out_off = fg_out;
        s.ops.push_back(o);
        ++s.n_fwd;
    }

    void leaf(uint32_t lo, uint32_t hi, uint64_t beta_hi) {
        const uint64_t mark = top;
        const uint64_t fg = alloc(pfg[hi + 1] - pfg[lo]);
        emit(lo, hi, fg);
        std::vector<uint64_t> offs(hi - lo + 1);
        offs[hi - lo] = beta_hi;
        for (uint32_t c = hi; c > lo; --c) {
            offs[c - 1 - lo] = alloc(colb[c - 1]);
            bwd(c - 1, offs[c - lo], offs[c - 1 - lo], fg + (pfg[c] - pfg[lo]), fg + (pfg[c - 1] - pfg[lo]));
        }
        for (uint32_t c = lo; c <= hi; ++c) fwd(c, offs[c - lo], fg + (pfg[c] - pfg[lo]));
        top = mark;
    }

    // largest e >= s0 such that leaf_need(s0, e) <= cap (assumes leaf_need(s0, s0) <= cap)
    uint32_t extend(uint32_t s0, uint32_t hi, uint64_t cap) const {
        uint32_t a = s0, b = hi;
        while (a < b) {
            uint32_t m = a + (b - a + 1) / 2;
            if (leaf_need(s0, m) <= cap) a = m; else b = m - 1;
        }
        return a;
    }

    void process(uint32_t lo, uint32_t hi, uint64_t beta_hi, uint32_t depth) {
        max_depth = std::max(max_depth, depth);
        const uint64_t avail = budget - top;
        if (leaf_need(lo, hi) <= avail) {
            leaf(lo, hi, beta_hi);
            return;
        }
        if (lo == hi) throw std::runtime_error("state budget too small for this problem (one column's tables do not fit)");
        uint64_t maxcol = 0, maxfg = 0;
        for (uint32_t c = lo; c <= hi; ++c) {
            maxcol = std::max(maxcol, up(colb[c]));
            maxfg = std::max(maxfg, up(fgb[c]));
        }
        const uint64_t roll = 2 * maxcol + 2 * maxfg;
        const uint32_t ncols = hi - lo + 1;
        std::vector<uint32_t> ends;   // last column of every part
        // smallest number of parts m such that the parts are leaves beside their m-1 checkpoints
        uint32_t m0 = (uint32_t)std::min<uint64_t>(ncols, std::max<uint64_t>(2, leaf_need(lo, hi) / std::max<uint64_t>(avail, 1)));
        for (uint32_t m = m0; m <= ncols; ++m) {
            const uint64_t ck = (uint64_t)(m - 1) * maxcol;
            if (avail < ck + roll) break;
            const uint64_t cap = avail - ck;
            std::vector<uint32_t> e;
            uint32_t s0 = lo;
            bool ok = true;
            while (s0 <= hi) {
                if (leaf_need(s0, s0) > cap) { ok = false; break; }
                uint32_t x = extend(s0, hi, cap);
                e.push_back(x);
                if (e.size() > m) { ok = false; break; }
                s0 = x + 1;
            }
            if (ok) { ends.swap(e); break; }
        }
        if (ends.empty()) {
            // too tight for one extra level: spend about half of what is left on checkpoints and recurse into the
            // parts, always leaving them room for two rolling columns plus two more (a part can then split again)
            const uint64_t M = avail > 2 * maxfg ? (avail - 2 * maxfg) / std::max<uint64_t>(maxcol, 1) : 0;
            if (M < 5 && avail >= roll) {
                // Room for two rolling columns only (columns of tens of GB): recompute the backward chain from the
                // range's last column for every forward column.  Quadratic in the range length, which is a handful
                // of columns when a single column takes a fifth of HBM.
                max_depth = std::max(max_depth, depth + 1);   // it recomputes: never report "everything fitted"
                const uint64_t mark = top;
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
            if (M < 5)
                throw std::runtime_error("state budget too small for this problem: " + std::to_string(avail) +
                                         " bytes left for a range whose largest column takes " + std::to_string(maxcol));
            uint64_t nck = M / 2;
            if (M - nck < 4) nck = M - 4;
            uint32_t m = (uint32_t)std::min<uint64_t>(nck + 1, ncols);
            for (uint32_t j = 1; j <= m; ++j) ends.push_back(lo + (uint32_t)((uint64_t)ncols * j / m) - 1);
        }
        const uint32_t m = (uint32_t)ends.size();
        const uint64_t mark = top;
        std::vector<uint64_t> ck(m, 0);
        for (uint32_t j = 0; j + 1 < m; ++j) ck[j] = alloc(colb[ends[j]]);
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
            process(s0, ends[j], ck[j], depth + 1);
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
    const uint64_t beta_last = b.alloc(colb[C - 1]);
    const uint64_t fg_last = b.alloc(fgb[C - 1]);
    b.emit(C - 1, C - 1, fg_last);
    Op o{};
    o.kind = OP_BWD_INIT; o.c = C - 1; o.out_off = beta_last; o.fg_out_off = fg_last;
    out.ops.push_back(o);
    ++out.n_bwd;
    b.process(0, C - 1, beta_last, 1);
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
