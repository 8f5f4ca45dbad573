// TEST INFRASTRUCTURE — not part of the product path.
//
// Thin C driver around the UNMODIFIED reference classes, compiled by
// oracle/build_ref.sh together with the reference's own sources where they lie
// under /root/reference/src (nothing is copied into this repo).  It feeds a
// flat problem (the layout of include/giggles_b200.h, but with RAW alignment
// scores so that the reference's own Entry::set_emission_score runs) through
// Read/ReadSet/Pedigree/GenotypeHMM exactly the way giggles/cli/genotype.py:288-330
// does, and returns what GenotypeHMM::get_genotype_likelihoods returns.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// arm may load the resulting oracle/_ref/libgiggles_ref.so.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "column.h"
#include "columniterator.h"
#include "genotype.h"
#include "genotypehmm.h"
#include "pedigree.h"
#include "phredgenotypelikelihoods.h"
#include "read.h"
#include "readset.h"

namespace {

ReadSet *build_readset(uint32_t n_reads, const uint32_t *positions, const int8_t *read_tag,
                       const uint32_t *read_entry_offset, const uint32_t *entry_column,
                       const int32_t *entry_allele, const uint64_t *entry_score_offset,
                       const double *entry_scores, int reg_const, double base_const) {
    ReadSet *rs = new ReadSet();
    for (uint32_t k = 0; k < n_reads; ++k) {
        Read *r = new Read("read" + std::to_string(k), 60, 0, 0, -1, "", reg_const, base_const);
        for (uint32_t e = read_entry_offset[k]; e < read_entry_offset[k + 1]; ++e) {
            std::vector<double> sc(entry_scores + entry_score_offset[e],
                                   entry_scores + entry_score_offset[e + 1]);
            r->addVariant((int)positions[entry_column[e]], entry_allele[e], sc, 30);
        }
        if (read_tag[k] == 0) r->addHaplotag("H1", -1);
        else if (read_tag[k] == 1) r->addHaplotag("H2", -1);
        else r->addHaplotag("none", -1);
        rs->add(r);  // ReadSet takes ownership; NOT sorted here: the flat order is the ReadSet order
    }
    return rs;
}

}  // namespace

extern "C" {

// Returns 0 on success.  out_likelihoods is flat with gl_offset[c]..gl_offset[c+1]
// (n_alleles(n_alleles+1)/2 each).  *ctor_seconds = wall time of the GenotypeHMM
// constructor alone (the span giggles/cli/genotype.py:305-322 times as "genotyping").
int ref_hmm_run(uint32_t n_columns, uint32_t n_haps, const uint32_t *positions,
                const uint32_t *n_alleles, const int32_t *allele_ref, const float *recomb,
                uint32_t n_reads, const int8_t *read_tag, const uint32_t *read_entry_offset,
                const uint32_t *entry_column, const int32_t *entry_allele,
                const uint64_t *entry_score_offset, const double *entry_scores, int reg_const,
                double base_const, const uint64_t *gl_offset, double *out_likelihoods,
                double *ctor_seconds, char *err, size_t errlen) {
    try {
        ReadSet *rs = build_readset(n_reads, positions, read_tag, read_entry_offset, entry_column,
                                    entry_allele, entry_score_offset, entry_scores, reg_const,
                                    base_const);
        std::vector<unsigned int> pos(positions, positions + n_columns);
        std::vector<unsigned int> nall(n_alleles, n_alleles + n_columns);
        std::vector<std::vector<int>> arefs(n_columns);
        for (uint32_t c = 0; c < n_columns; ++c)
            arefs[c].assign(allele_ref + (size_t)c * n_haps, allele_ref + (size_t)(c + 1) * n_haps);
        std::vector<float> rc(recomb, recomb + (n_columns > 0 ? n_columns - 1 : 0));

        Pedigree ped;
        std::vector<Genotype *> gts;
        std::vector<PhredGenotypeLikelihoods *> gls;
        for (uint32_t c = 0; c < n_columns; ++c) {
            gts.push_back(new Genotype());
            gls.push_back(nullptr);
        }
        ped.addIndividual(0, gts, gls);

        auto t0 = std::chrono::steady_clock::now();
        GenotypeHMM hmm(rs, rc, &ped, n_haps, &pos, &nall, &arefs);
        auto t1 = std::chrono::steady_clock::now();
        if (ctor_seconds) *ctor_seconds = std::chrono::duration<double>(t1 - t0).count();

        for (uint32_t c = 0; c < n_columns; ++c) {
            std::vector<long double> l = hmm.get_genotype_likelihoods(0, c);
            uint64_t n = gl_offset[c + 1] - gl_offset[c];
            for (uint64_t g = 0; g < n && g < l.size(); ++g)
                out_likelihoods[gl_offset[c] + g] = (double)l[g];  // long double -> double, core.pyx:506
        }
        delete rs;
        return 0;
    } catch (const std::exception &e) {
        if (err && errlen) {
            std::strncpy(err, e.what(), errlen - 1);
            err[errlen - 1] = 0;
        }
        return 1;
    }
}

// Column structure as the reference sees it: for every column the active read
// ids (ColumnIterator::get_next, src/columniterator.cpp:93-141) and the untagged
// subset (Column::Column, src/column.cpp:13-17), flattened CSR; and, for every
// column c >= 1 and every bipartition index b of c (only when the column has at
// most `max_enum_bits` untagged reads), the list returned by
// hmm_columns[c-1]->get_backward_compatible_bipartitions(b) (src/column.cpp:72-122),
// concatenated in b order.  Vectors are returned through caller-provided
// buffers of the given capacities; the function returns the needed sizes.
int ref_structure(uint32_t n_columns, const uint32_t *positions, uint32_t n_reads,
                  const int8_t *read_tag, const uint32_t *read_entry_offset,
                  const uint32_t *entry_column, uint32_t n_haps, uint32_t max_enum_bits,
                  uint32_t *active_offset /*[C+1]*/, uint32_t *active_ids, uint64_t active_cap,
                  uint32_t *untagged_offset /*[C+1]*/, uint32_t *untagged_ids, uint64_t untagged_cap,
                  uint64_t *compat_offset /*[C+1]*/, uint32_t *compat, uint64_t compat_cap,
                  uint64_t *needed /*[3]*/, char *err, size_t errlen) {
    try {
        // raw scores are irrelevant for the structure: give every entry a 2-allele dummy vector
        std::vector<uint64_t> soff(read_entry_offset[n_reads] + 1);
        for (size_t i = 0; i < soff.size(); ++i) soff[i] = 2 * i;
        std::vector<double> sc(2 * (size_t)read_entry_offset[n_reads], 0.0);
        std::vector<int32_t> al(read_entry_offset[n_reads], 0);
        ReadSet *rs = build_readset(n_reads, positions, read_tag, read_entry_offset, entry_column,
                                    al.data(), soff.data(), sc.data(), 10, 2.718);
        rs->reassignReadIds();
        std::vector<unsigned int> pos(positions, positions + n_columns);
        ColumnIterator it(*rs, &pos);
        std::vector<std::vector<unsigned int>> ids(n_columns);
        for (uint32_t c = 0; c < n_columns; ++c) {
            auto col = it.get_next();
            for (const Entry *e : *col) ids[c].push_back(e->get_read_id());
        }
        uint64_t na = 0, nu = 0, nc = 0;
        std::vector<Column *> cols(n_columns, nullptr);
        unsigned int nref = n_haps;
        for (uint32_t c = 0; c < n_columns; ++c) {
            std::vector<unsigned int> next = (c + 1 < n_columns) ? ids[c + 1] : std::vector<unsigned int>{};
            cols[c] = new Column(c, &nref, ids[c], next, rs);
        }
        for (uint32_t c = 0; c < n_columns; ++c) {
            active_offset[c] = (uint32_t)na;
            untagged_offset[c] = (uint32_t)nu;
            compat_offset[c] = nc;
            uint32_t u = 0;
            for (unsigned int id : ids[c]) {
                if (na < active_cap) active_ids[na] = id;
                ++na;
                if (!rs->get(id)->hasHaplotag()) {
                    if (nu < untagged_cap) untagged_ids[nu] = id;
                    ++nu;
                    ++u;
                }
            }
            if (c >= 1 && u <= max_enum_bits) {
                for (uint32_t b = 0; b < (1u << u); ++b) {
                    std::vector<unsigned int> cb = cols[c - 1]->get_backward_compatible_bipartitions((int)b, rs);
                    if (nc < compat_cap) compat[nc] = (uint32_t)cb.size();
                    ++nc;
                    for (unsigned int x : cb) {
                        if (nc < compat_cap) compat[nc] = x;
                        ++nc;
                    }
                }
            }
        }
        active_offset[n_columns] = (uint32_t)na;
        untagged_offset[n_columns] = (uint32_t)nu;
        compat_offset[n_columns] = nc;
        needed[0] = na;
        needed[1] = nu;
        needed[2] = nc;
        for (Column *c : cols) delete c;
        delete rs;
        return 0;
    } catch (const std::exception &e) {
        if (err && errlen) {
            std::strncpy(err, e.what(), errlen - 1);
            err[errlen - 1] = 0;
        }
        return 1;
    }
}

// Entry::set_emission_score through the reference's own Entry (src/entry.cpp:50-64).
void ref_entry_emission(const double *scores, uint32_t n, int reg_const, double base_const, double *out) {
    Entry e(0, 0, std::vector<double>(scores, scores + n), 30, reg_const, base_const);
    std::vector<long double> em = e.get_emission_score();
    for (uint32_t i = 0; i < n; ++i) out[i] = (double)em[i];
}

}  // extern "C"
