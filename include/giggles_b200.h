/*
 * giggles_b200.h — C ABI of the B200-native GenotypeHMM forward-backward.
 *
 * This is the drop-in boundary.  The reference binds C++ classes
 * straight into Cython (giggles/cpp.pxd:14-125) and has no C ABI; the entry
 * points below are what a C-ABI rebinding of that one path needs.  Each
 * function names the reference interface it replaces (paths are relative to
 * the reference checkout).
 *
 * Plain pointers and sizes only.  No torch / C++ types.  All functions are
 * synchronous, keep no global state apart from the CUDA context, return 0 on
 * success and a non-zero gg_status plus a message in `err` on failure (the
 * C++ shim `GenotypeHMM` re-throws that as std::runtime_error so Cython's
 * `except +` behaviour of cpp.pxd:100-103 is kept).
 *
 * There is no CPU fallback: without a CUDA device every compute entry point
 * fails with GG_ERR_NO_DEVICE.
 */
#ifndef GIGGLES_B200_H
#define GIGGLES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_MAX_ALLELES 16   /* giggles/vcf.py:406 caps alleles per site      */
#define GG_MAX_UNTAGGED 24  /* 2^u bipartitions per column; u_c <= max_cov   */

typedef enum gg_status {
    GG_OK = 0,
    GG_ERR_INVALID = 1,    /* malformed problem (unsorted reads, bad index…)  */
    GG_ERR_NO_DEVICE = 2,  /* no CUDA device / driver                         */
    GG_ERR_CUDA = 3,       /* a CUDA call failed                              */
    GG_ERR_NOMEM = 4,      /* problem does not fit the device memory budget   */
    GG_ERR_UNSUPPORTED = 5
} gg_status;

/*
 * One chromosome's worth of input to GenotypeHMM::GenotypeHMM
 * (src/genotypehmm.cpp:19-51, ctor declared at src/genotypehmm.h:142), flattened
 * to structure-of-arrays.  Column c is VCF position positions[c]
 * (giggles/cli/genotype.py:265: every position of the chromosome is a column).
 *
 * Reads are in ReadSet order (src/readset.h:41-68: sorted by first position);
 * read k's entries are entry_*[read_entry_offset[k] .. read_entry_offset[k+1]).
 * entry_column must be strictly increasing inside a read (src/read.cpp:74-83)
 * and the first column non-decreasing across reads (src/columniterator.cpp:28-35).
 * Every read needs at least one entry (src/read.cpp:86-88 throws otherwise).
 *
 * entry_em holds the NORMALISED emission probabilities the reference keeps in
 * Entry::emission_score (src/entry.cpp:50-64); use gg_entry_emission() to get
 * them from raw alignment scores.  entry_em_offset[e+1]-entry_em_offset[e]
 * must be >= n_alleles[entry_column[e]].  An entry with entry_allele == -1 is
 * skipped by the emission product but its read still occupies a bipartition
 * bit (src/genotypehmm.cpp:583,612; src/columniterator.cpp:132-137).
 */
typedef struct gg_problem {
    uint32_t n_columns;               /* C                                      */
    uint32_t n_haps;                  /* R = n_references (genotypehmm.h:44)    */
    const uint32_t *positions;        /* [C] strictly increasing; may be NULL   */
    const uint32_t *n_alleles;        /* [C] 1..GG_MAX_ALLELES                  */
    const int32_t *allele_ref;        /* [C*R] panel allele of haplotype h at c, -1 = missing */
    const float *recomb;              /* [>= C-1] recombcost, fp32 like genotypehmm.h:62; step c-1 -> c uses recomb[c-1] */
    uint32_t n_reads;                 /* N                                      */
    const int8_t *read_tag;           /* [N] -1 untagged, 0 = H1, 1 = H2 (src/read.cpp:45-51) */
    const uint32_t *read_entry_offset;/* [N+1] CSR                              */
    const uint32_t *entry_column;     /* [E] column index of the entry          */
    const int32_t *entry_allele;      /* [E] observed allele, -1 = none         */
    const uint64_t *entry_em_offset;  /* [E+1] offsets into entry_em            */
    const double *entry_em;           /* normalised emission probabilities      */
} gg_problem;

typedef struct gg_options {
    int32_t device;                   /* CUDA ordinal; -1 = current device      */
    uint64_t state_budget_bytes;      /* HBM budget for alpha/beta columns; 0 = 94 % of free memory when first asked; later calls
                                         reuse the retired arena of that size (gg_arena_release() makes the next call ask again) */
    int32_t precision;                /* 0 = fp32 state with per-column scaling (production); 1 = fp64 state (validation) */
    int32_t kernel_variant;           /* 0 = auto; for A/B tests: 1 = generic step kernel only (no bulk-copy pipeline), 4 = no chain
                                         kernel, 5 = no warp-per-block kernel, 7 = small panels column by column (no wave runs),
                                         8 = wide panels (R > 128) on the row-tiled pipelined kernel instead of the column-owning one */
    uint32_t flags;                   /* GG_OPT_* */
} gg_options;

/* gg_hmm_create: the handle keeps its tables on the device but no state arena; every gg_hmm_sweep borrows ONE arena
 * shared by all such handles of the device (their sweeps are serialised).  For many chromosomes planned and resident
 * at once (the per-chromosome loop of giggles/cli/genotype.py:196-338) when a single chain already fills HBM. */
#define GG_OPT_SHARED_ARENA 1u

typedef struct gg_result gg_result;   /* host-side result table                 */
typedef struct gg_hmm gg_hmm;         /* device-resident problem + schedule     */

/* Sweep statistics; `algorithmic_bytes` follows SURVEY §8(d): 20 B per state per column. */
typedef struct gg_stats {
    uint64_t n_columns;
    uint64_t state_columns;           /* sum_c 2^{u_c} R^2                      */
    uint64_t algorithmic_bytes;       /* 20 * state_columns                     */
    uint64_t scheduled_bytes;         /* bytes the schedule really moves (incl. checkpoint recompute) */
    uint64_t n_launches;              /* kernels launched by one gg_hmm_sweep   */
    uint64_t n_backward_steps;        /* column steps incl. recompute           */
    uint64_t n_forward_steps;
    uint64_t state_bytes_reserved;    /* arena size                             */
    uint32_t max_untagged;            /* max_c u_c                              */
    uint32_t checkpoint_levels;
    float last_sweep_ms;              /* CUDA-event time of the last sweep      */
    float last_fwdbwd_kernel_ms;      /* summed time of the column-step kernels in the last timed sweep (0 if not timed) */
    float last_fwd_ms;                /* ... of which forward step launches (timed sweeps only) */
    float last_bwd_ms;                /* ... backward step launches                             */
    float last_emit_ms;               /* emission-table launches                                */
    float create_plan_ms;             /* host: column planner (gg_hmm_create)                    */
    float create_total_ms;            /* host + device: planner, schedule, allocations, H2D      */
    float reserved;
} gg_stats;

/* ---- one-shot call: replaces `new GenotypeHMM(...)` (giggles/core.pyx:498 ->
 *      src/genotypehmm.cpp:19-51).  Host buffers in, host result out; the H2D
 *      and D2H copies happen inside. */
int gg_hmm_run(const gg_problem *p, const gg_options *o, gg_result **out,
               char *err, size_t errlen);

/* ---- many chains (chromosomes) on one device: what the per-chromosome loop of giggles/cli/genotype.py:196-338
 *      does one GenotypeHMM at a time.  A few host threads plan and schedule the chains; chains whose
 *      state fits a small arena (all-haplotagged chromosomes: one persistent CTA each) are swept side by side, each
 *      on its own stream; a chain that needs most of HBM has the device to itself while the chains behind it are
 *      being planned.  GG_BATCH_THREADS=1 runs strictly one chain after the other.  results[i] is chain i's result
 *      (the caller frees every non-NULL entry); on failure the status and message of the lowest-numbered failing
 *      chain are returned and results[i] is NULL for chains that failed or were not run. */
int gg_hmm_run_batch(const gg_problem *problems, uint32_t n, const gg_options *o, gg_result **results,
                     char *err, size_t errlen);

/* ---- staged form of the same call, for callers that keep the problem resident
 *      (bench.py times gg_hmm_sweep alone; gg_hmm_run = create+sweep+fetch+destroy). */
int gg_hmm_create(const gg_problem *p, const gg_options *o, gg_hmm **out,
                  char *err, size_t errlen);          /* plan + H2D             */
int gg_hmm_sweep(gg_hmm *h, int timed_kernels, char *err, size_t errlen); /* backward + forward + posterior on device */
int gg_hmm_fetch(gg_hmm *h, gg_result **out, char *err, size_t errlen);    /* D2H posteriors        */
int gg_hmm_stats(const gg_hmm *h, gg_stats *s);
void gg_hmm_destroy(gg_hmm *h);
/* gg_hmm_destroy keeps the state arena of the last handle per device for the next gg_hmm_create / gg_hmm_run
 * (allocating >100 GB costs more than sweeping a small chromosome); this frees it. */
void gg_arena_release(void);
/* test hook: the per-column sums of the stored (unnormalised) alpha / beta columns of the last sweep, [n_columns] each */
int gg_hmm_debug_scales(gg_hmm *h, double *alpha_scale, double *beta_scale);
/* test/profiling hook: per-launch record of the last TIMED sweep (kind: 0 emission, 1 backward init, 2 backward step,
 * 3 forward step, 4 fused backward run, 104 fused forward run; the column it produced; CUDA-event milliseconds) */
uint64_t gg_hmm_debug_op_times(gg_hmm *h, uint32_t *kind, uint32_t *column, float *ms, uint64_t cap);

/* ---- result access: replaces GenotypeHMM::get_genotype_likelihoods
 *      (src/genotypehmm.cpp:563-574, giggles/core.pyx:505-506).  Returns a pointer
 *      to n_genotypes = n_alleles(n_alleles+1)/2 doubles that sum to 1, indexed by
 *      the reference's ordered-pair genotype index a1 + a2(a2+1)/2
 *      (src/genotypehmm.cpp:524-528).  NULL if column is out of range. */
const double *gg_result_likelihoods(const gg_result *r, uint32_t column, uint32_t *n_genotypes);
uint32_t gg_result_n_columns(const gg_result *r);
/* flat view: all columns back to back, offsets[c]..offsets[c+1]              */
const double *gg_result_flat(const gg_result *r, const uint64_t **offsets);
void gg_result_free(gg_result *r);
/* fp32 range guard: 1 if this chain's fp32 posteriors were not finite and the result comes from the one rerun on fp64
 * state that gg_hmm_run / gg_hmm_run_batch then do (the reference computes in 80-bit long double,
 * src/genotypehmm.cpp:203-561); gg_fp64_rerun_count() counts such reruns process-wide. */
int gg_result_reran_fp64(const gg_result *r);
uint64_t gg_fp64_rerun_count(void);
/* wraps caller-provided posteriors (e.g. gathered from other ranks) so gg_result_call can be applied to them */
gg_result *gg_result_from_host(uint32_t n_columns, const uint64_t *offsets, const double *values);

/* ---- column structure, for the bit-identical-structure parity test.  Mirrors
 *      Column::Column (src/column.cpp:10-39) and
 *      Column::get_backward_compatible_bipartitions (src/column.cpp:72-122).
 *      Host only; needs no device.  For column c:
 *        active  = ids of reads spanning c (ColumnIterator::get_next, src/columniterator.cpp:93-141)
 *        untagged= subset without haplotag, bit i of the bipartition index is untagged[i]
 *        shared_mask_prev = bits of column c-1's index whose read is still active at c
 *      The caller owns nothing: pointers stay valid until gg_plan_free. */
typedef struct gg_plan gg_plan;
int gg_plan_build(const gg_problem *p, gg_plan **out, char *err, size_t errlen);
uint32_t gg_plan_n_columns(const gg_plan *pl);
const uint32_t *gg_plan_active(const gg_plan *pl, uint32_t column, uint32_t *n);
const uint32_t *gg_plan_untagged(const gg_plan *pl, uint32_t column, uint32_t *n);
uint32_t gg_plan_shared_mask_prev(const gg_plan *pl, uint32_t column, uint32_t *n_shared);
/* all bipartition indices of column c-1 compatible with index b of column c, ascending */
uint32_t gg_plan_compatible_prev(const gg_plan *pl, uint32_t column, uint32_t b,
                                 uint32_t *out, uint32_t out_cap);
void gg_plan_free(gg_plan *pl);

/* ---- host helpers that define the kernel's inputs bit-for-bit ---- */
/* Entry::set_emission_score (src/entry.cpp:50-64): p_i = 10^-reg + 1/sum_j base^(e_j-e_i), normalised. */
void gg_entry_emission(const double *scores, uint32_t n, int32_t reg_const, double base_const, double *out);
/* the same for many entries at once: entry e has scores[offsets[e] .. offsets[e+1]) (what
 * Read::addVariant does per call, src/read.cpp:69-71) */
void gg_entry_emission_batch(const double *scores, const uint64_t *offsets, uint64_t n_entries, int32_t reg_const,
                             double base_const, double *out);
/* TransitionProbabilityComputer (src/transitionprobabilitycomputer.cpp:10-15). */
void gg_transition_scalars(float recombcost, uint32_t n_haps, double *pr, double *qr);
/* binomial_coefficient (src/binomial.cpp:5-15). */
int32_t gg_binomial_coefficient(int32_t n, int32_t k);

/* ---- derived calls: determine_genotype (giggles/cli/genotype.py:82-95)
 *      and GQ / GL (giggles/vcf.py:850-872) for all columns of a result, in C.
 *      gt_index[c] = genotype index or -1 for "./.", gq[c] integer quality or -1
 *      when undefined, gl_log10 flat like gg_result_flat (floored at -1000). */
int gg_result_call(const gg_result *r, double gt_qual_threshold, int32_t *gt_index,
                   int32_t *gq, double *gl_log10);

/* ---- read selection: readselection(readset, max_cov, preferred_source_ids, bridging)
 *      of giggles/readselect.pyx:231-271 (called by giggles/utils.py:71-76 before every GenotypeHMM).
 *      Host only.  The view is the flat image of what that code reads through cpp.ReadSet / cpp.Read
 *      (getVariantCount, getPosition, getQuality, getSourceID — readselect.pyx:41-46,63-72,127-140).
 *      Returns the selected read indices in increasing order (the reference returns a Python set that
 *      ReadSet.subset turns into an ordered IndexSet, giggles/core.pyx:298-308).  Ties between equally
 *      scored reads are resolved exactly as the reference resolves them under CPython 3.12 + libstdc++
 *      (see giggles_b200/csrc/pyset.hpp).  rc 2 = a read covers fewer than two variants (ValueError there). */
typedef struct gg_readset_view {
    uint32_t n_reads;
    const uint64_t *read_variant_offset; /* [n_reads+1] CSR into the two arrays below */
    const int32_t *variant_position;     /* strictly increasing inside a read */
    const int32_t *variant_quality;
    const int32_t *read_source_id;       /* [n_reads] */
} gg_readset_view;
int gg_readselection(const gg_readset_view *rs, uint32_t max_cov, const int32_t *preferred_source_ids,
                     uint32_t n_preferred, int bridging, uint32_t *selected, uint32_t *n_selected, char *err,
                     size_t errlen);
/* test hook: replays a script of Python-set operations on the order-exact set model (readselect.cpp) */
int64_t gg_pyset_replay(const int64_t *ops, int64_t n, int64_t *out, int64_t cap);

/* ---- batched edit distance: edit_distance(s, t, maxdiff) of giggles/align.pyx:15-107 for n_pairs pairs
 *      at once, on the device.  This is the aligner of the allele detection by realignment in "edit" mode
 *      (giggles/variants.py:50-51,238-239: one call per read x variant x allele, maxdiff = -1).  Pair i compares
 *      pool[s_off[i] .. +s_len[i]) with pool[t_off[i] .. +t_len[i]) (byte strings; the same query may be shared by
 *      several pairs).  maxdiff == NULL or maxdiff[i] == -1: the edit distance.  maxdiff[i] >= 0: the reference's banded
 *      result, bit for bit (the true distance iff it is <= maxdiff, otherwise whatever value > maxdiff the reference's
 *      band bookkeeping yields; abs(len difference) when that already exceeds maxdiff).  H2D / D2H copies inside. */
int gg_edit_distance_batch(const uint8_t *pool, uint64_t pool_bytes, const uint64_t *s_off, const uint32_t *s_len,
                           const uint64_t *t_off, const uint32_t *t_len, const int32_t *maxdiff, uint32_t n_pairs,
                           int32_t *out, int32_t device, char *err, size_t errlen);

/* ---- checkpoint schedule dry run (host only, no device): builds and self-checks the launch
 *      schedule gg_hmm_create would use for columns of the given byte sizes under `budget`.
 *      The reference's fixed sqrt(C) schedule is src/genotypehmm.cpp:130,153-156,364-383. */
int gg_schedule_dry_run(uint32_t n_columns, const uint64_t *col_bytes, const uint64_t *fg_bytes, uint64_t budget,
                        uint64_t *arena_bytes, uint32_t *levels, uint64_t *n_bwd, uint64_t *n_fwd,
                        uint64_t *bwd_col_bytes /* bytes of the columns the backward steps produce, recompute included */,
                        char *err, size_t errlen);

/* the state budget gg_hmm_create uses on `device` when gg_options.state_budget_bytes is 0 (94 % of the free memory,
 * arenas retired by earlier handles counted as free); 0 without a device */
uint64_t gg_default_state_budget(int device);

/* diagnostic: device-to-device copy bandwidth in GB/s (read + write), best of `reps` copies of `bytes` */
double gg_device_copy_gbs(int device, uint64_t bytes, int reps);

const char *gg_version(void);
int gg_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GIGGLES_B200_H */
