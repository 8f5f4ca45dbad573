// Host-side column planner of the B200 GenotypeHMM.
//
// Turns one flat gg_problem (include/giggles_b200.h) into the per-column
// structure the reference derives on the fly with ColumnIterator
// (reference src/columniterator.cpp:93-141), Column (src/column.cpp:10-39) and
// Column::get_backward_compatible_bipartitions (src/column.cpp:72-122), but as
// closed-form bit masks instead of per-bipartition vectors:
//
//   * active reads of column c  = reads k with first_k <= c <= last_k, by id
//   * untagged reads U_c        = active reads without haplotag; bit i of the
//                                 bipartition index b is the side of U_c[i]
//   * reads of U_{c-1} still in U_c keep their relative order and occupy the
//     LOW n_sh bits of b_c (new reads have larger ids, they sort last), so
//       compatible(b_c) = pdep(b_c & (2^n_sh-1), shared_mask_prev) | any subset of term_mask_prev
//
// Everything here is plain host C++ (no CUDA), so the structure tests run on a
// CPU-only box.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/giggles_b200.h"

namespace gg {

// side code of an emission entry: 0..GG_MAX_UNTAGGED-1 = bit index in b, or a fixed side
enum : uint8_t { SIDE_FIXED0 = 64, SIDE_FIXED1 = 65 };

struct ColumnPlan {
    uint32_t u = 0;                 // |U_c|
    uint32_t n_alleles = 0;
    uint32_t n_sh = 0;              // untagged reads shared with column c-1 (low bits of b_c)
    uint32_t n_term_prev = 0;       // untagged reads of c-1 that end there
    uint32_t shared_mask_prev = 0;  // their bit positions in b_{c-1}
    uint32_t term_mask_prev = 0;
    uint32_t em_count = 0;          // active reads with a usable entry (allele != -1) at c
    uint64_t em_off = 0;            // first emission entry (index into em_side)
    uint64_t em_prob_off = 0;       // offset into em_prob; entry j allele i at em_prob_off + j*n_alleles + i
    uint64_t gl_off = 0;            // offset of this column's genotype table in the result
    double tA = 1.0, tB = 0.0, tC = 0.0;  // transition c-1 -> c: T(M) = tA*M + tB*(row+col) + tC*tot
};

struct Plan {
    uint32_t n_columns = 0, n_haps = 0, n_reads = 0;   // n_haps: LAYOUT width (even: an odd panel gets one phantom haplotype)
    uint32_t n_haps_true = 0;                          // the panel's haplotype count
    std::vector<ColumnPlan> cols;
    // CSR of active / untagged read ids per column (for the structure parity test)
    std::vector<uint64_t> active_off, untagged_off;
    std::vector<uint32_t> active_ids, untagged_ids;
    // emission entries, column-major
    std::vector<uint8_t> em_side;   // side code per entry
    std::vector<double> em_prob;    // per entry: em[i] / max_i em[i], i < n_alleles(c)
    // panel alleles shifted by one (0 = missing), [C*R]; and per column the
    // haplotypes sorted by allele (stable), [C*R]
    std::vector<uint8_t> allele1;
    std::vector<uint16_t> row_order;
    std::vector<uint64_t> gl_offsets;  // [C+1]
    uint32_t max_u = 0, max_alleles = 0;
    uint64_t state_columns = 0;        // sum_c 2^u_c R^2
};

// Throws std::runtime_error (message mirrors the reference's where it has one).
void build_plan(const gg_problem &p, Plan &out);

// software bit deposit / extract (no BMI2 on the device; tiny on the host)
inline uint32_t pdep32(uint32_t x, uint32_t mask) {
    uint32_t r = 0;
    for (uint32_t m = mask; m; m &= m - 1) {
        if (x & 1u) r |= m & (0u - m);
        x >>= 1;
    }
    return r;
}
inline uint32_t pext32(uint32_t x, uint32_t mask) {
    uint32_t r = 0, k = 0;
    for (uint32_t m = mask; m; m &= m - 1, ++k)
        if (x & (m & (0u - m))) r |= 1u << k;
    return r;
}

// Entry::set_emission_score (reference src/entry.cpp:50-64)
void entry_emission(const double *scores, uint32_t n, int32_t reg_const, double base_const, double *out);
// TransitionProbabilityComputer (reference src/transitionprobabilitycomputer.cpp:10-15)
void transition_scalars(float recombcost, uint32_t n_haps, double *pr, double *qr);

}  // namespace gg
