// Host-side result table handed back through the C ABI (include/giggles_b200.h).
#pragma once
#include <cstdint>
#include <vector>

struct gg_result {
    std::vector<uint64_t> offsets;  // [C+1]
    std::vector<double> values;     // genotype posteriors, all columns back to back
    bool reran_fp64 = false;        // the fp32 range guard replaced this chain's result by an fp64-state rerun
};
