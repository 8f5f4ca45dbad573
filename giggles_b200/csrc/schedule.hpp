// Checkpoint schedule of one forward-backward sweep, sized to an HBM budget.
//
// The reference keeps every floor(sqrt(C))-th backward column and recomputes the
// rest block by block during the forward pass (reference src/genotypehmm.cpp:130,
// 153-156 and 364-383).  Results do not depend on that schedule, so
// this one is driven by bytes instead: a range of columns whose backward columns
// all fit is a LEAF (backward stored, then forward); a range that does not fit is
// swept once backward with two rolling buffers, keeping checkpoints at part
// boundaries, and its parts are processed left to right, recursively.  The
// result is a flat list of kernel launches over offsets into one arena, built on
// the host with no CUDA dependency (unit-tested on CPU).
#pragma once
#include <cstdint>
#include <vector>

namespace gg {

enum OpKind : uint8_t { OP_EMIT = 0, OP_BWD_INIT = 1, OP_BWD = 2, OP_FWD = 3,
                        OP_RUN = 4 /* engine-side fusion of consecutive tiny steps: c = index of the run */ };

struct Op {
    OpKind kind;
    uint32_t c;        // EMIT: first column; BWD_INIT/BWD/FWD: the column PRODUCED
    uint32_t hi;       // EMIT: last column of the range
    uint64_t in_off;   // BWD: beta of column c+1; FWD: alpha of column c-1 (unused for c == 0)
    uint64_t out_off;  // produced column; EMIT: region base (table of column x at + fgoff[x]-fgoff[c])
    uint64_t beta_off; // FWD: beta of column c
    uint64_t fg_out_off, fg_in_off;  // F/G tables of the produced column / of the input column (BWD)
};

struct Schedule {
    std::vector<Op> ops;
    uint64_t arena_bytes = 0;     // high-water mark
    uint32_t levels = 0;          // 1 = everything fitted (no recompute)
    uint64_t n_bwd = 0, n_fwd = 0, n_emit = 0;
    uint64_t bwd_col_bytes = 0;   // bytes of the columns produced by the backward steps (recompute included)
};

// colb[c]: bytes of column c (values + side sums), fgb[c]: bytes of its F/G table.
// Throws std::runtime_error("...budget...") if even the minimal schedule does not fit.
void build_schedule(const std::vector<uint64_t> &colb, const std::vector<uint64_t> &fgb, uint64_t budget,
                    Schedule &out);

// Replays the op list symbolically and throws if any op reads a buffer that does not hold the
// column it expects or two live buffers overlap.  Cheap; run on every schedule.
void validate_schedule(const std::vector<uint64_t> &colb, const std::vector<uint64_t> &fgb, const Schedule &s);

}  // namespace gg
