// Host-side column planner — see plan.hpp.
#include "plan.hpp"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <stdexcept>

namespace gg {

void entry_emission(const double *scores, uint32_t n, int32_t reg_const, double base_const, double *out) {
    // Same arithmetic, in the same precision, as reference src/entry.cpp:50-64:
    // the per-allele value is a long double, the normaliser a double.
    std::vector<long double> em(n);
    double normalization = 0.0;
    for (uint32_t i = 0; i < n; ++i) {
        em[i] = std::pow(10, -reg_const);
        long double score = 0.0L;
        for (uint32_t j = 0; j < n; ++j) score += std::pow(base_const, scores[j] - scores[i]);
        em[i] += 1 / score;
        normalization += em[i];
    }
    for (uint32_t i = 0; i < n; ++i) out[i] = (double)(em[i] / normalization);
}

void transition_scalars(float recombcost, uint32_t n_haps, double *pr, double *qr) {
    double s = (double)n_haps;
    float r = recombcost;
    *pr = (1.0 - std::exp(-(r / s))) / s;
    *qr = std::exp(-(r / s)) + *pr;
}

static void require(bool ok, const char *msg) {
    if (!ok) throw std::runtime_error(msg);
}

void build_plan(const gg_problem &p, Plan &pl) {
    const uint32_t C = p.n_columns, R = p.n_haps, N = p.n_reads;
    require(R >= 1, "gg_problem: n_haps must be >= 1");
    require(R <= 1020, "gg_problem: n_haps > 1020 is not supported");
    require(C == 0 || (p.n_alleles && p.allele_ref), "gg_problem: n_alleles / allele_ref missing");
    require(C <= 1 || p.recomb, "gg_problem: recomb missing");
    require(N == 0 || (p.read_tag && p.read_entry_offset && p.entry_column && p.entry_allele &&
                       p.entry_em_offset && p.entry_em),
            "gg_problem: read arrays missing");
    pl = Plan();
    pl.n_columns = C;
    // An odd panel is laid out with phantom haplotypes (one; up to three above 256) whose allele is missing at every column: their states have
    // emission 0, so they carry no mass in either direction and change no sum, but rows stay 8-byte aligned and the
    // vectorised kernels apply.  The transition keeps the true haplotype count (TransitionProbabilityComputer's s).
    const uint32_t RL = R > 256 ? ((R + 3u) & ~3u) : R + (R & 1u);   // > 256: pad to a multiple of 4 (float4 tiles)
    pl.n_haps = RL;
    pl.n_haps_true = R;
    pl.n_reads = N;
    if (p.positions)
        for (uint32_t c = 1; c < C; ++c)
            require(p.positions[c] > p.positions[c - 1], "gg_problem: positions must be strictly increasing");

    // ---- reads: first / last column, sortedness (reference src/columniterator.cpp:28-35, src/read.cpp:74-88)
    std::vector<uint32_t> first(N), last(N);
    for (uint32_t k = 0; k < N; ++k) {
        uint32_t b = p.read_entry_offset[k], e = p.read_entry_offset[k + 1];
        require(e > b, "Read: firstPosition/lastPosition of a read without variants");
        for (uint32_t i = b; i < e; ++i) {
            require(p.entry_column[i] < C, "gg_problem: entry_column out of range");
            if (i > b) require(p.entry_column[i] > p.entry_column[i - 1],
                               "ColumnIterator: encountered read with unsorted variants.");
        }
        first[k] = p.entry_column[b];
        last[k] = p.entry_column[e - 1];
        if (k > 0) require(first[k] >= first[k - 1], "ColumnIterator: reads in ReadSet are not sorted.");
        require(p.read_tag[k] >= -1 && p.read_tag[k] <= 1, "gg_problem: read_tag must be -1, 0 or 1");
    }

    pl.cols.resize(C);
    pl.active_off.assign(C + 1, 0);
    pl.untagged_off.assign(C + 1, 0);
    pl.gl_offsets.assign(C + 1, 0);
    pl.allele1.assign((size_t)C * RL, 0);
    pl.row_order.resize((size_t)C * RL);

    std::vector<uint32_t> active, next_active, cursor(N, 0), prev_untagged, untagged;
    for (uint32_t k = 0; k < N; ++k) cursor[k] = p.read_entry_offset[k];
    uint32_t next_read = 0;
    for (uint32_t c = 0; c < C; ++c) {
        ColumnPlan &cp = pl.cols[c];
        const uint32_t nA = p.n_alleles[c];
        require(nA >= 1 && nA <= GG_MAX_ALLELES, "gg_problem: n_alleles out of range (1..16)");
        cp.n_alleles = nA;
        pl.max_alleles = std::max(pl.max_alleles, nA);
        cp.gl_off = pl.gl_offsets[c];
        pl.gl_offsets[c + 1] = pl.gl_offsets[c] + (uint64_t)nA * (nA + 1) / 2;

        // panel alleles (+1) and allele-sorted haplotype order
        for (uint32_t h = 0; h < R; ++h) {
            int32_t a = p.allele_ref[(size_t)c * R + h];
            require(a >= -1 && a < (int32_t)nA, "gg_problem: allele_ref out of range");
            pl.allele1[(size_t)c * RL + h] = (uint8_t)(a + 1);
        }
        // stable counting sort of the haplotypes by allele class (<= 17 classes)
        const uint8_t *al = &pl.allele1[(size_t)c * RL];
        uint32_t cnt[GG_MAX_ALLELES + 2] = {0};
        for (uint32_t h = 0; h < RL; ++h) ++cnt[al[h] + 1];
        for (uint32_t k = 1; k <= GG_MAX_ALLELES + 1; ++k) cnt[k] += cnt[k - 1];
        uint16_t *ro = &pl.row_order[(size_t)c * RL];
        for (uint32_t h = 0; h < RL; ++h) ro[cnt[al[h]]++] = (uint16_t)h;

        // active set: keep spanning reads (in id order), append reads starting here
        next_active.clear();
        for (uint32_t k : active)
            if (last[k] >= c) next_active.push_back(k);
        while (next_read < N && first[next_read] == c) next_active.push_back(next_read++);
        active.swap(next_active);

        prev_untagged.swap(untagged);
        untagged.clear();
        for (uint32_t k : active)
            if (p.read_tag[k] < 0) untagged.push_back(k);
        cp.u = (uint32_t)untagged.size();
        if (cp.u > GG_MAX_UNTAGGED)
            throw std::runtime_error("gg_problem: more than " + std::to_string(GG_MAX_UNTAGGED) +
                                     " untagged reads cover one column (lower --max-coverage)");
        pl.max_u = std::max(pl.max_u, cp.u);
        pl.state_columns += ((uint64_t)1 << cp.u) * R * R;

        // shared / terminating masks over the previous column's index
        for (uint32_t i = 0; i < prev_untagged.size(); ++i) {
            if (last[prev_untagged[i]] >= c) {
                cp.shared_mask_prev |= 1u << i;
                ++cp.n_sh;
            } else {
                cp.term_mask_prev |= 1u << i;
                ++cp.n_term_prev;
            }
        }
        // the shared reads must be the first n_sh untagged reads of this column, in the same order
        for (uint32_t i = 0, j = 0; i < prev_untagged.size(); ++i)
            if (last[prev_untagged[i]] >= c) require(untagged[j++] == prev_untagged[i], "planner: shared-read order broken");

        // transition c-1 -> c
        if (c > 0) {
            double pr, qr;
            transition_scalars(p.recomb[c - 1], R, &pr, &qr);
            double d = qr - pr;  // = exp(-r/s)
            cp.tA = d * d;
            cp.tB = pr * d;
            cp.tC = pr * pr;
        }

        // emission entries of the active reads
        cp.em_off = pl.em_side.size();
        cp.em_prob_off = pl.em_prob.size();
        pl.active_off[c] = pl.active_ids.size();
        pl.untagged_off[c] = pl.untagged_ids.size();
        uint32_t bit = 0;
        for (uint32_t k : active) {
            pl.active_ids.push_back(k);
            const bool untag = p.read_tag[k] < 0;
            if (untag) pl.untagged_ids.push_back(k);
            uint32_t &cur = cursor[k];
            const uint32_t end = p.read_entry_offset[k + 1];
            while (cur < end && p.entry_column[cur] < c) ++cur;
            if (cur < end && p.entry_column[cur] == c && p.entry_allele[cur] != -1) {
                const uint64_t eo = p.entry_em_offset[cur];
                require(p.entry_em_offset[cur + 1] - eo >= nA, "gg_problem: entry has fewer emission values than alleles");
                const double *em = p.entry_em + eo;
                double mx = 0.0;
                for (uint32_t i = 0; i < nA; ++i) mx = std::max(mx, em[i]);
                require(mx > 0.0 && std::isfinite(mx), "gg_problem: emission vector must have a positive maximum");
                for (uint32_t i = 0; i < nA; ++i) pl.em_prob.push_back(em[i] / mx);
                pl.em_side.push_back(untag ? (uint8_t)bit : (p.read_tag[k] == 0 ? SIDE_FIXED0 : SIDE_FIXED1));
                ++cp.em_count;
            }
            if (untag) ++bit;
        }
    }
    pl.active_off[C] = pl.active_ids.size();
    pl.untagged_off[C] = pl.untagged_ids.size();
}

}  // namespace gg
