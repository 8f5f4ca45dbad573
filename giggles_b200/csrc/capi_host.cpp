// Host-only half of the C ABI (include/giggles_b200.h): column structure export, result access,
// the scalar helpers that define the kernel inputs, and the derived GT/GQ/GL calls.  No CUDA here.
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/giggles_b200.h"
#include "plan.hpp"
#include "result.hpp"
#include "schedule.hpp"

struct gg_plan {
    gg::Plan plan;
};

static void set_err(char *err, size_t errlen, const char *m) {
    if (err && errlen) {
        std::strncpy(err, m, errlen - 1);
        err[errlen - 1] = 0;
    }
}

extern "C" {

const char *gg_version(void) { return "giggles_b200 0.1 (sm_100a)"; }

// ---- column structure ----------------------------------------------------------------------
int gg_plan_build(const gg_problem *p, gg_plan **out, char *err, size_t errlen) {
    if (out) *out = nullptr;
    if (!p || !out) {
        set_err(err, errlen, "gg_plan_build: null argument");
        return GG_ERR_INVALID;
    }
    gg_plan *pl = new gg_plan();
    try {
        gg::build_plan(*p, pl->plan);
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        delete pl;
        return std::string(e.what()).find("untagged reads cover") != std::string::npos ? GG_ERR_UNSUPPORTED : GG_ERR_INVALID;
    }
    *out = pl;
    return GG_OK;
}

uint32_t gg_plan_n_columns(const gg_plan *pl) { return pl ? pl->plan.n_columns : 0; }

const uint32_t *gg_plan_active(const gg_plan *pl, uint32_t column, uint32_t *n) {
    if (!pl || column >= pl->plan.n_columns) {
        if (n) *n = 0;
        return nullptr;
    }
    const auto &P = pl->plan;
    if (n) *n = (uint32_t)(P.active_off[column + 1] - P.active_off[column]);
    return P.active_ids.data() + P.active_off[column];
}

const uint32_t *gg_plan_untagged(const gg_plan *pl, uint32_t column, uint32_t *n) {
    if (!pl || column >= pl->plan.n_columns) {
        if (n) *n = 0;
        return nullptr;
    }
    const auto &P = pl->plan;
    if (n) *n = (uint32_t)(P.untagged_off[column + 1] - P.untagged_off[column]);
    return P.untagged_ids.data() + P.untagged_off[column];
}

uint32_t gg_plan_shared_mask_prev(const gg_plan *pl, uint32_t column, uint32_t *n_shared) {
    if (!pl || column >= pl->plan.n_columns) {
        if (n_shared) *n_shared = 0;
        return 0;
    }
    if (n_shared) *n_shared = pl->plan.cols[column].n_sh;
    return pl->plan.cols[column].shared_mask_prev;
}

uint32_t gg_plan_compatible_prev(const gg_plan *pl, uint32_t column, uint32_t b, uint32_t *out, uint32_t out_cap) {
    if (!pl || column == 0 || column >= pl->plan.n_columns) return 0;
    const gg::ColumnPlan &cp = pl->plan.cols[column];
    const uint32_t base = gg::pdep32(b & ((1u << cp.n_sh) - 1u), cp.shared_mask_prev);
    const uint32_t n = 1u << cp.n_term_prev;
    uint32_t sub = 0;
    for (uint32_t k = 0; k < n; ++k) {      // submasks of term_mask in ascending order
        if (k < out_cap) out[k] = base | sub;
        sub = (sub - cp.term_mask_prev) & cp.term_mask_prev;
    }
    return n;
}

void gg_plan_free(gg_plan *pl) { delete pl; }

// ---- results ---------------------------------------------------------------------------------
const double *gg_result_likelihoods(const gg_result *r, uint32_t column, uint32_t *n_genotypes) {
    if (!r || (size_t)column + 1 >= r->offsets.size()) {
        if (n_genotypes) *n_genotypes = 0;
        return nullptr;
    }
    if (n_genotypes) *n_genotypes = (uint32_t)(r->offsets[column + 1] - r->offsets[column]);
    return r->values.data() + r->offsets[column];
}

uint32_t gg_result_n_columns(const gg_result *r) { return (r && !r->offsets.empty()) ? (uint32_t)(r->offsets.size() - 1) : 0; }

const double *gg_result_flat(const gg_result *r, const uint64_t **offsets) {
    if (!r) return nullptr;
    if (offsets) *offsets = r->offsets.data();
    return r->values.data();
}

void gg_result_free(gg_result *r) { delete r; }
int gg_result_reran_fp64(const gg_result *r) { return r && r->reran_fp64 ? 1 : 0; }

gg_result *gg_result_from_host(uint32_t n_columns, const uint64_t *offsets, const double *values) {
    gg_result *r = new gg_result();
    r->offsets.assign(offsets, offsets + n_columns + 1);
    r->values.assign(values, values + offsets[n_columns]);
    return r;
}

// determine_genotype (reference giggles/cli/genotype.py:82-95) and GL/GQ (giggles/vcf.py:850-872)
int gg_result_call(const gg_result *r, double gt_qual_threshold, int32_t *gt_index, int32_t *gq, double *gl_log10) {
    if (!r) return GG_ERR_INVALID;
    const uint32_t C = gg_result_n_columns(r);
    for (uint32_t c = 0; c < C; ++c) {
        const double *l = r->values.data() + r->offsets[c];
        const uint32_t n = (uint32_t)(r->offsets[c + 1] - r->offsets[c]);
        // largest and second largest value; on ties Python's stable sort keeps the higher index last
        int32_t best = -1;
        double v1 = -INFINITY, v2 = -INFINITY;
        for (uint32_t g = 0; g < n; ++g) {
            if (l[g] >= v1) { v2 = v1; v1 = l[g]; best = (int32_t)g; }
            else if (l[g] > v2) v2 = l[g];
        }
        int32_t call = -1;
        if (n >= 2 && v1 > v2 && v1 - v2 > gt_qual_threshold) call = best;   // NaN compares false -> no call
        if (gt_index) gt_index[c] = call;
        if (gl_log10)
            for (uint32_t g = 0; g < n; ++g) {
                double x = l[g] > 0 ? std::log10(l[g]) : -1000.0;
                gl_log10[r->offsets[c] + g] = x > -1000.0 ? x : -1000.0;
            }
        if (gq) {
            if (call < 0) {
                gq[c] = -1;
            } else {
                double q = 0.0;
                for (uint32_t g = 0; g < n; ++g)
                    if ((int32_t)g != call) q += l[g];
                if (q > 0) {
                    double ph = std::nearbyint(-10.0 * std::log10(q));   // Python round(): half to even
                    gq[c] = ph < 10000.0 ? (int32_t)ph : 10000;
                } else {
                    gq[c] = 10000;
                }
            }
        }
    }
    return GG_OK;
}

// ---- scalar helpers --------------------------------------------------------------------------
void gg_entry_emission(const double *scores, uint32_t n, int32_t reg_const, double base_const, double *out) {
    gg::entry_emission(scores, n, reg_const, base_const, out);
}

void gg_entry_emission_batch(const double *scores, const uint64_t *offsets, uint64_t n_entries, int32_t reg_const,
                             double base_const, double *out) {
    for (uint64_t e = 0; e < n_entries; ++e)
        gg::entry_emission(scores + offsets[e], (uint32_t)(offsets[e + 1] - offsets[e]), reg_const, base_const, out + offsets[e]);
}

void gg_transition_scalars(float recombcost, uint32_t n_haps, double *pr, double *qr) {
    gg::transition_scalars(recombcost, n_haps, pr, qr);
}

// reference src/binomial.cpp:5-15 (integer arithmetic, same evaluation order)
int32_t gg_binomial_coefficient(int32_t n, int32_t k) {
    if (k < 0 || n < 0 || n < k) return 0;
    if (k > n - k) k = n - k;
    int32_t result = 1;
    for (int32_t i = 0; i < k; ++i) {
        result *= (n - i);
        result /= (i + 1);
    }
    return result;
}

// ---- schedule dry run (host only; used by the CPU tests) -----------------------------------------
int gg_schedule_dry_run(uint32_t n_columns, const uint64_t *col_bytes, const uint64_t *fg_bytes, uint64_t budget,
                        uint64_t *arena_bytes, uint32_t *levels, uint64_t *n_bwd, uint64_t *n_fwd, uint64_t *bwd_col_bytes,
                        char *err, size_t errlen) {
    try {
        std::vector<uint64_t> cb(col_bytes, col_bytes + n_columns), fb(fg_bytes, fg_bytes + n_columns);
        gg::Schedule s;
        gg::build_schedule(cb, fb, budget, s);
        gg::validate_schedule(cb, fb, s);
        if (arena_bytes) *arena_bytes = s.arena_bytes;
        if (levels) *levels = s.levels;
        if (n_bwd) *n_bwd = s.n_bwd;
        if (n_fwd) *n_fwd = s.n_fwd;
        if (bwd_col_bytes) *bwd_col_bytes = s.bwd_col_bytes;
        return GG_OK;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return std::string(e.what()).find("budget") != std::string::npos ? GG_ERR_NOMEM : GG_ERR_INVALID;
    }
}

}  // extern "C"
