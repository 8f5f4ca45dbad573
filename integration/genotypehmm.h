// Drop-in replacement for the reference's src/genotypehmm.h.
//
// Same class name, constructor signature and accessor as the reference
// (src/genotypehmm.h:142-146), so giggles/cpp.pxd:100-103 and giggles/core.pyx:474-506 stay untouched and
// `giggles genotype` runs unchanged on its GAF/GFA/VCF front end.  The forward-backward itself runs on the GPU
// through the C ABI of include/giggles_b200.h (libgiggles_b200.so); there is no CPU path.
//
// Everything else of the reference's host core (Read, ReadSet, Entry, Pedigree, Genotype, ...) is used as is:
// this header includes the reference's own readset.h / pedigree.h from the tree it is dropped into.
#ifndef GENOTYPEHMM_H
#define GENOTYPEHMM_H

#include <vector>

#include "pedigree.h"
#include "readset.h"

struct gg_result;

class GenotypeHMM {
public:
    // src/genotypehmm.h:142 — all work happens in the constructor, like the reference (src/genotypehmm.cpp:19-51)
    GenotypeHMM(ReadSet *read_set, const std::vector<float> &recombcost, const Pedigree *pedigree,
                const unsigned int &n_references, const std::vector<unsigned int> *positions = nullptr,
                const std::vector<unsigned int> *n_allele_positions = nullptr,
                const std::vector<std::vector<int> > *allele_references = nullptr);
    ~GenotypeHMM();
    GenotypeHMM(const GenotypeHMM &) = delete;
    GenotypeHMM &operator=(const GenotypeHMM &) = delete;

    // src/genotypehmm.cpp:563-574
    std::vector<long double> get_genotype_likelihoods(unsigned int individual_id, unsigned int position);

private:
    const Pedigree *pedigree;
    gg_result *result;
    unsigned int column_count;
};

#endif
