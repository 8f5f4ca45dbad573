// Iteration-order-exact model of a CPython `set` holding small non-negative ints.
//
// Why this exists: the reference's read selection (giggles/readselect.pyx:106-271)
// keeps its working sets as Python `set`s and walks them to fill and update a
// binary heap (readselect.pyx:97-101, 166-170).  Reads with equal scores are
// therefore ranked by the ORDER in which CPython happens to iterate those sets,
// and that order decides which reads survive — i.e. it fixes the untagged read
// set U_c of every HMM column.  To stay bit-identical the host port has to
// reproduce that order, so this class restates the open-addressing scheme CPython
// documents for its set type (table of 2^k slots, 9 linear probes then the
// i*5+1+perturb recurrence, fill/used accounting with dummy slots, the growth and
// dummy-purge rules of add / merge / difference / difference_update).  It is
// pinned empirically, not by reading: tests/test_readselect.py replays random
// operation sequences against the interpreter's own `set` and compares the
// iteration order after every step.
//
// Keys are ints k >= 0 (hash(k) == k for k < 2^61-1).  Plain host C++.
#pragma once
#include <cstdint>
#include <vector>

namespace gg {

class PySetInt {
  public:
    PySetInt() : tab_(kMinSize) {}

    size_t size() const { return used_; }
    bool empty() const { return used_ == 0; }

    bool contains(int64_t k) const { return tab_[lookup(k)].st == ACTIVE; }

    // set.add(k)
    void add(int64_t k) {
        const size_t mask = tab_.size() - 1;
        size_t i = (size_t)k & mask, perturb = (size_t)k;
        long freeslot = -1;
        for (;;) {
            size_t e = i;
            int probes = (i + kLinearProbes <= mask) ? kLinearProbes : 0;
            do {
                Slot &s = tab_[e];
                if (s.st == UNUSED) {
                    if (freeslot >= 0) {               // reuse a dummy seen on the way
                        tab_[freeslot] = {k, ACTIVE};
                        ++used_;
                        return;
                    }
                    s = {k, ACTIVE};
                    ++fill_;
                    ++used_;
                    if (fill_ * 5 >= mask * 3) resize(used_ > 50000 ? used_ * 2 : used_ * 4);
                    return;
                }
                if (s.st == ACTIVE) {
                    if (s.key == k) return;
                } else {
                    freeslot = (long)e;
                }
                ++e;
            } while (probes--);
            perturb >>= kPerturbShift;
            i = (i * 5 + 1 + perturb) & mask;
        }
    }

    // set.discard(k) / set.remove(k): the slot becomes a dummy, nothing moves
    bool discard(int64_t k) {
        Slot &s = tab_[lookup(k)];
        if (s.st != ACTIVE) return false;
        s.st = DUMMY;
        --used_;
        return true;
    }

    // set.update(<list>) : plain insertions in list order
    template <class It> void update_seq(It b, It e) {
        for (; b != e; ++b) add((int64_t)*b);
    }

    // set.update(<set>) and the body of set(<set>) / set.copy()
    void merge(const PySetInt &o) {
        if (&o == this || o.used_ == 0) return;
        if ((fill_ + o.used_) * 5 >= (tab_.size() - 1) * 3) resize((used_ + o.used_) * 2);
        if (fill_ == 0 && tab_.size() == o.tab_.size() && o.fill_ == o.used_) {
            tab_ = o.tab_;
            fill_ = o.fill_;
            used_ = o.used_;
            return;
        }
        if (fill_ == 0) {
            fill_ = used_ = o.used_;
            for (const Slot &s : o.tab_)
                if (s.st == ACTIVE) insert_clean(tab_, s.key);
            return;
        }
        for (size_t i = 0; i < o.tab_.size(); ++i)
            if (o.tab_[i].st == ACTIVE) add(o.tab_[i].key);
    }

    PySetInt copy() const {
        PySetInt r;
        r.merge(*this);
        return r;
    }

    // a -= b.  `o` is anything with size() / contains() / for_each(): the subtrahend's own order never matters.
    template <class O> void difference_update(const O &o) {
        if ((const void *)&o == (const void *)this) {
            *this = PySetInt();
            return;
        }
        o.for_each([&](int64_t k) { discard(k); });
        purge_dummies();
    }

    // a.difference(b)
    template <class O> PySetInt difference(const O &o) const {
        if ((used_ >> 2) > o.size()) {          // "copy and discard" branch
            PySetInt r = copy();
            r.difference_update(o);
            return r;
        }
        PySetInt r;
        for (const Slot &s : tab_)
            if (s.st == ACTIVE && !o.contains(s.key)) r.add(s.key);
        return r;
    }

    // iteration order = slot order
    template <class F> void for_each(F &&f) const {
        for (const Slot &s : tab_)
            if (s.st == ACTIVE) f(s.key);
    }
    std::vector<int64_t> items() const {
        std::vector<int64_t> v;
        v.reserve(used_);
        for_each([&](int64_t k) { v.push_back(k); });
        return v;
    }

  private:
    enum : uint8_t { UNUSED = 0, ACTIVE = 1, DUMMY = 2 };
    struct Slot {
        int64_t key = 0;
        uint8_t st = UNUSED;
    };
    static constexpr size_t kMinSize = 8;
    static constexpr int kLinearProbes = 9;
    static constexpr int kPerturbShift = 5;

    std::vector<Slot> tab_;
    size_t fill_ = 0, used_ = 0;   // fill = active + dummy

    size_t lookup(int64_t k) const {
        const size_t mask = tab_.size() - 1;
        size_t i = (size_t)k & mask, perturb = (size_t)k;
        for (;;) {
            size_t e = i;
            int probes = (i + kLinearProbes <= mask) ? kLinearProbes : 0;
            do {
                const Slot &s = tab_[e];
                if (s.st == UNUSED) return e;
                if (s.st == ACTIVE && s.key == k) return e;
                ++e;
            } while (probes--);
            perturb >>= kPerturbShift;
            i = (i * 5 + 1 + perturb) & mask;
        }
    }

    static void insert_clean(std::vector<Slot> &t, int64_t k) {
        const size_t mask = t.size() - 1;
        size_t i = (size_t)k & mask, perturb = (size_t)k;
        for (;;) {
            if (t[i].st == UNUSED) {
                t[i] = {k, ACTIVE};
                return;
            }
            if (i + kLinearProbes <= mask) {
                for (int j = 1; j <= kLinearProbes; ++j)
                    if (t[i + j].st == UNUSED) {
                        t[i + j] = {k, ACTIVE};
                        return;
                    }
            }
            perturb >>= kPerturbShift;
            i = (i * 5 + 1 + perturb) & mask;
        }
    }

    void resize(size_t minused) {
        size_t n = kMinSize;
        while (n <= minused) n <<= 1;
        if (n == kMinSize && tab_.size() == kMinSize && fill_ == used_) return;   // small table, nothing to purge
        std::vector<Slot> t(n);
        for (const Slot &s : tab_)
            if (s.st == ACTIVE) insert_clean(t, s.key);
        tab_.swap(t);
        fill_ = used_;
    }

    void purge_dummies() {
        if (fill_ - used_ <= (tab_.size() - 1) / 4) return;
        resize(used_ > 50000 ? used_ * 2 : used_ * 4);
    }
};

}  // namespace gg
