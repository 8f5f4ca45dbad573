// Drop-in replacement for the reference's src/genotypehmm.cpp — see genotypehmm.h.
//
// Flattens the caller's ReadSet into a gg_problem (include/giggles_b200.h) and calls gg_hmm_run.  What the
// reference constructor does besides arithmetic is kept: the pedigree must hold exactly one individual
// (src/genotypehmm.cpp:33), read ids are reassigned on the caller's ReadSet (:39), unsorted input is rejected with
// the reference's messages (src/columniterator.cpp:31,34), errors surface as std::runtime_error so that Cython's
// `except +` turns them into RuntimeError.
#include "genotypehmm.h"

#include <stdexcept>
#include <string>
#include <unordered_map>

#include "giggles_b200.h"

GenotypeHMM::GenotypeHMM(ReadSet *read_set, const std::vector<float> &recombcost, const Pedigree *pedigree,
                         const unsigned int &n_references, const std::vector<unsigned int> *positions,
                         const std::vector<unsigned int> *n_allele_positions,
                         const std::vector<std::vector<int> > *allele_references)
    : pedigree(pedigree), result(nullptr), column_count(0) {
    if (pedigree->size() != 1) throw std::runtime_error("GenotypeHMM: exactly one individual is supported");
    std::vector<unsigned int> *own_positions = nullptr;
    if (positions == nullptr) {   // ColumnIterator falls back to the read set's own positions (src/columniterator.cpp:15-16)
        own_positions = read_set->get_positions();
        positions = own_positions;
    }
    struct Cleanup {
        std::vector<unsigned int> *p;
        ~Cleanup() { delete p; }
    } cleanup{own_positions};
    const size_t C = positions->size();
    column_count = (unsigned int)C;
    if (n_allele_positions == nullptr || allele_references == nullptr || n_allele_positions->size() != C ||
        allele_references->size() != C)
        throw std::runtime_error("GenotypeHMM: n_allele_positions / allele_references must cover every position");
    read_set->reassignReadIds();

    std::unordered_map<unsigned int, uint32_t> column_of;
    column_of.reserve(C * 2);
    for (size_t c = 0; c < C; ++c) column_of[positions->at(c)] = (uint32_t)c;

    std::vector<uint32_t> n_alleles(n_allele_positions->begin(), n_allele_positions->end());
    std::vector<int32_t> allele_ref(C * (size_t)n_references);
    for (size_t c = 0; c < C; ++c) {
        const std::vector<int> &row = allele_references->at(c);
        if (row.size() != n_references) throw std::runtime_error("GenotypeHMM: allele_references row length != n_references");
        for (unsigned int h = 0; h < n_references; ++h) allele_ref[c * n_references + h] = row[h];
    }

    const unsigned int N = read_set->size();
    std::vector<int8_t> read_tag(N);
    std::vector<uint32_t> read_entry_offset(N + 1, 0), entry_column;
    std::vector<int32_t> entry_allele;
    std::vector<uint64_t> entry_em_offset(1, 0);
    std::vector<double> entry_em;
    for (unsigned int k = 0; k < N; ++k) {
        const Read *read = read_set->get(k);
        if (!read->isSorted()) throw std::runtime_error("ColumnIterator: encountered read with unsorted variants.");
        read_tag[k] = read->hasHaplotag() ? (int8_t)read->getHaplotag() : (int8_t)-1;
        const int n = read->getVariantCount();
        if (n == 0) throw std::runtime_error("Read: firstPosition/lastPosition of a read without variants");
        for (int i = 0; i < n; ++i) {
            auto it = column_of.find((unsigned int)read->getPosition(i));
            if (it == column_of.end()) {
                if (i == 0 || i == n - 1) throw std::runtime_error("GenotypeHMM: read starts or ends at a position that is not a column");
                continue;   // interior entries off the column grid are never visited by the reference's iterators
            }
            entry_column.push_back(it->second);
            entry_allele.push_back(read->getAllele(i));
            const std::vector<long double> em = read->getEmissionProbability(i);   // already normalised by Entry (src/entry.cpp:50-64)
            for (long double v : em) entry_em.push_back((double)v);
            entry_em_offset.push_back(entry_em.size());
        }
        read_entry_offset[k + 1] = (uint32_t)entry_column.size();
    }

    gg_problem p;
    p.n_columns = (uint32_t)C;
    p.n_haps = n_references;
    p.positions = positions->data();
    p.n_alleles = n_alleles.data();
    p.allele_ref = allele_ref.data();
    p.recomb = recombcost.data();
    p.n_reads = N;
    p.read_tag = read_tag.data();
    p.read_entry_offset = read_entry_offset.data();
    p.entry_column = entry_column.data();
    p.entry_allele = entry_allele.data();
    p.entry_em_offset = entry_em_offset.data();
    p.entry_em = entry_em.data();
    if (C > 1 && recombcost.size() < C - 1) throw std::runtime_error("GenotypeHMM: recombcost shorter than the number of column gaps");

    gg_options o;
    o.device = -1;
    o.state_budget_bytes = 0;
    o.precision = 0;
    o.kernel_variant = 0;
    o.flags = 0;
    char err[512] = {0};
    const int rc = gg_hmm_run(&p, &o, &result, err, sizeof(err));
    if (rc != GG_OK) throw std::runtime_error(std::string("GenotypeHMM (giggles_b200): ") + err);
}

GenotypeHMM::~GenotypeHMM() { gg_result_free(result); }

std::vector<long double> GenotypeHMM::get_genotype_likelihoods(unsigned int individual_id, unsigned int position) {
    if (pedigree->id_to_index(individual_id) >= pedigree->size()) throw std::runtime_error("GenotypeHMM: unknown individual");
    uint32_t n = 0;
    const double *l = gg_result_likelihoods(result, position, &n);
    if (l == nullptr) throw std::runtime_error("GenotypeHMM: position out of range");
    return std::vector<long double>(l, l + n);
}
