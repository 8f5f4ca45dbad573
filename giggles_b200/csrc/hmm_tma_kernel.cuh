// Pipelined column step for sm_100a: bulk-async (TMA) staging of bipartition blocks.
//
// Same arithmetic as step_kernel (hmm_kernels.cuh) for fp32 columns with R % 4 == 0 (R <= 512).  Persistent CTAs:
//
//   warp 0      producer.  A work item is one output bipartition block; it is cut into TILES of TR consecutive
//               rows (TR = R, one tile, while a whole block record fits a stage: R <= 128).  Per tile and per
//               input block the producer issues ONE cp.async.bulk global->shared of the tile's rows into the next
//               free stage of a ring (the first tile also brings the block's row/col/total sums, which sit right
//               behind the values) and, for the forward step, one more for the beta rows of the posterior.  Copies
//               complete on the stage's "full" mbarrier.  The item's F/G emission vectors are prefetched one item
//               ahead and travel in the stage's metadata.
//   warps 1..CW consumers.  Wait on "full", run the row loop out of shared memory (LDS.128), store the produced
//               rows straight from registers (STG.128), hand the stage back through its "empty" mbarrier.
//
// With 3-8 stages of 30-40 KB an SM keeps >100 KB of loads in flight, which is what hides HBM latency; the
// register-staged generic kernel cannot.
#pragma once
#include <type_traits>

#include "hmm_kernels.cuh"

namespace gg {

constexpr int kTmaMaxStages = 8;
constexpr int kTmaMaxConsumerWarps = 8;
constexpr uint32_t kTmaL2Ahead = 2817;   // producer's L2 prefetch: one round ahead, beta blocks only, evict-last; beta copy evict-first (see the producer loop); 0 = off; GG_L2_AHEAD overrides

struct TmaStageMeta {
    uint32_t b_out, b_self, pad0, pad1;
    float fg_out[2 * kMaxClasses + 2];    // F then G of the output block (first load of an item only)
    float fg_self[2 * kMaxClasses + 2];   // backward: F then G of the block held by this stage
};

// Geometry shared by host (sizes) and device (pointers).
struct TmaGeom {
    uint32_t TR, n_tiles, n_iter_t, TRp;          // rows per tile, tiles per block, row iterations per tile, padded rows
    uint32_t side_off, stage_bytes;               // byte offset of the side record inside a stage
    uint32_t off_stage, off_meta, off_acc, off_side, off_colred, off_totred, off_fgo, off_q, off_al_o, off_al_i,
        off_rord, off_rsum, total;
};
__host__ __device__ inline TmaGeom tma_geom(uint32_t R, uint32_t TR, uint32_t n_stages, uint32_t ngrp, uint32_t ncls_out,
                                            bool fwd, bool need_acc) {
    TmaGeom g;
    const uint32_t Rp = (R + 3u) & ~3u;
    g.TR = TR;
    g.n_tiles = (R + TR - 1) / TR;
    g.n_iter_t = (TR + ngrp - 1) / ngrp;
    g.TRp = g.n_iter_t * ngrp;
    g.side_off = TR * R * 4u;
    g.stage_bytes = (g.side_off + (2u * R + 4u) * 4u + 127u) & ~127u;
    uint32_t o = 0;
    g.off_stage = o; o += n_stages * g.stage_bytes;
    g.off_meta = o; o += n_stages * (uint32_t)sizeof(TmaStageMeta);
    o = (o + 15u) & ~15u;
    g.off_acc = o; o += need_acc ? g.side_off : 0;                       // reduce-set accumulator (one tile)
    g.off_side = o; o += (need_acc || g.n_tiles > 1) ? (2u * Rp + 4u) * 4u : 0;   // block side sums kept across tiles
    g.off_colred = o; o += 2 * ngrp * Rp * 4;                            // double-buffered across items
    g.off_totred = o; o += 2 * ngrp * 4;
    g.off_fgo = o; o += 48 * 4;
    g.off_q = o; o += fwd ? ngrp * ncls_out * Rp * 4 : 0;
    g.off_al_o = o; o += Rp;
    g.off_al_i = o; o += Rp;
    o = (o + 7u) & ~7u;
    g.off_rord = o; o += g.n_tiles * g.TRp * 8;                          // per-row metadata (2 words)
    g.off_rsum = o; o += 2 * Rp * 4;                                     // row sums of the produced block
    g.total = (o + 127u) & ~127u;
    return g;
}

// ---- mbarrier / bulk-copy PTX ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void consumer_bar(uint32_t nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// hint: start moving `bytes` (multiple of 16) at `src` from HBM into L2; no completion tracking
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_prefetch_l2_hint(const void *src, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

template <int V>
__device__ __forceinline__ void lds_vec(float (&d)[V], const void *p) {
    if (V == 4) { const float4 t = *reinterpret_cast<const float4 *>(p); d[0] = t.x; d[1] = t.y; d[V > 2 ? 2 : 0] = t.z; d[V > 3 ? 3 : 0] = t.w; }
    else { const float2 t = *reinterpret_cast<const float2 *>(p); d[0] = t.x; d[1] = t.y; }
}
template <int V>
__device__ __forceinline__ void st_vec(void *p, const float (&d)[V]) {
    if (V == 4) *reinterpret_cast<float4 *>(p) = make_float4(d[0], d[1], d[V > 2 ? 2 : 0], d[V > 3 ? 3 : 0]);
    else *reinterpret_cast<float2 *>(p) = make_float2(d[0], d[1]);
}

// row metadata, pre-scaled to byte offsets so the row loop does no address arithmetic:
//   x = byte offset of the row inside its tile ((R1 - row0)*R*4) | valid << 31
//   y = R1*4 (12 bits) | (allele class of R1 in the output column)*4 << 12 (7 bits) | (... input column)*4 << 19
__device__ __forceinline__ uint2 pack_row(uint32_t R1, uint32_t row0, uint32_t R, uint32_t a1o, uint32_t a1i) {
    return make_uint2(((R1 - row0) * R * 4u) | 0x80000000u, (R1 * 4u) | (a1o * 4u) << 12 | (a1i * 4u) << 19);
}

// MT = false: a block is one tile (TR == R), known at compile time so the tile loop and its bookkeeping vanish.
// CWT: consumer warps at compile time (8: two CTAs per SM; 16: one CTA per SM when the shared-memory footprint of a
// column - reduce-set accumulator, many allele classes - does not fit twice); 0 = taken from blockDim.
// TWO: the reduce set has exactly two blocks and both are held in stages (compile-time so that the common
// one-block kernels carry none of its code).
// V: floats per vector access (4: R % 4 == 0; 2: R % 4 == 2, e.g. a 47-sample / 94-haplotype panel).
template <bool FWD, int G, int NS, bool MT, int CWT, bool TWO, int V>
__global__ void __launch_bounds__(32 * (1 + (CWT ? CWT : kTmaMaxConsumerWarps)), (CWT == 8 ? 2 : 1))
step_tma_kernel(const StepArgs<float> a, const uint32_t n_stages, const uint32_t TR_arg, const uint32_t l2_ahead) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kTmaMaxStages];
    __shared__ __align__(8) uint64_t s_empty[kTmaMaxStages];
    __shared__ bool s_last;
    __shared__ uint16_t s_seg[kTmaMaxConsumerWarps * 2 * kMaxClasses];   // forward, single tile: per row group, end of each class segment

    const uint32_t R = a.R, BS = a.BS;
    const uint32_t RR = R * R;
    const uint32_t Rp = (R + 3u) & ~3u;
    const uint32_t TR = MT ? TR_arg : R;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // consumer warps
    const uint32_t CW = CWT ? (uint32_t)CWT : (blockDim.x >> 5) - 1;
    const uint32_t nct = 32 * CW;
    constexpr uint32_t rpi = 32u / G;
    const uint32_t ngrp = CW * rpi;
    const uint32_t NCo = a.nA_out + 1, NCi = a.nA_in + 1;
    const uint32_t fgs_o = 2 * NCo, fgs_i = 2 * NCi;
    const uint32_t nred = FWD ? (1u << a.n_free_left) : (1u << (a.u_in - a.n_sh));
    const uint32_t n_bcast = FWD ? (a.u_out - a.n_sh) : a.n_free_left;
    const uint32_t nload = nred + (FWD ? 1u : 0u);      // loads per tile
    // reduce sets of two (one read starts / ends between the columns: the common non-trivial case) are summed on
    // the fly from two stages; longer ones (and all of them when a block is tiled) are folded into an accumulator
    constexpr bool two = TWO;
    const bool accum = !TWO && nred > 1;
    const TmaGeom L = tma_geom(R, TR, n_stages, ngrp, NCo, FWD, accum);
    const uint32_t n_tiles = MT ? L.n_tiles : 1u;
    TmaStageMeta *s_meta = reinterpret_cast<TmaStageMeta *>(smem_raw + L.off_meta);
    const uint32_t n_items = 1u << a.u_out;

    if (tid == 0) {
        for (uint32_t s = 0; s < n_stages; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // consumer-side tables
    uint8_t *s_al_o = smem_raw + L.off_al_o;
    uint8_t *s_al_i = smem_raw + L.off_al_i;
    uint2 *s_rowmeta = reinterpret_cast<uint2 *>(smem_raw + L.off_rord);
    float *s_q = reinterpret_cast<float *>(smem_raw + L.off_q);
    for (uint32_t t = tid; t < R; t += blockDim.x) {
        s_al_o[t] = a.al_out[t];
        s_al_i[t] = a.al_in[t];
    }
    for (uint32_t t = tid; t < n_tiles * L.TRp; t += blockDim.x) s_rowmeta[t] = make_uint2(0u, 0u);   // padding rows
    if (FWD)
        for (uint32_t t = tid; t < ngrp * NCo * Rp; t += blockDim.x) s_q[t] = 0.f;
    __syncthreads();
    // rows of a tile are visited in allele-class order (forward: the posterior accumulators are parked only when
    // the class changes); one tile: the planner's sorted order; several tiles: a counting sort per tile
    if (n_tiles == 1) {
        for (uint32_t t = tid; t < R; t += blockDim.x) {
            const uint32_t R1 = (FWD && a.row_order) ? a.row_order[t] : t;
            s_rowmeta[t] = pack_row(R1, 0, R, s_al_o[R1], s_al_i[R1]);
        }
        if (FWD && G == 32) {
            // rows are class-sorted and dealt round-robin to the row groups: group g owns sorted rows g, g+ngrp, ...;
            // its rows of class <= c are the first ceil((B_c - g) / ngrp) of them, B_c = #rows of class <= c
            for (uint32_t q = tid; q < ngrp * NCo; q += blockDim.x) {
                const uint32_t g = q / NCo, c = q % NCo;
                uint32_t B = 0;
                for (uint32_t r = 0; r < R; ++r) B += s_al_o[r] <= c ? 1u : 0u;
                s_seg[g * kMaxClasses + c] = (uint16_t)(B > g ? (B - g + ngrp - 1) / ngrp : 0u);
            }
        }
    } else if (tid < n_tiles) {
        const uint32_t row0 = tid * TR, row1 = min(R, row0 + TR);
        uint32_t pos = tid * L.TRp;
        if (FWD) {
            for (uint32_t cls = 0; cls < NCo; ++cls)
                for (uint32_t r = row0; r < row1; ++r)
                    if (s_al_o[r] == cls) s_rowmeta[pos++] = pack_row(r, row0, R, cls, s_al_i[r]);
        } else {
            for (uint32_t r = row0; r < row1; ++r) s_rowmeta[pos++] = pack_row(r, row0, R, s_al_o[r], s_al_i[r]);
        }
    }
    __syncthreads();

    if (warp == 0) {
        // =========================== producer ===========================
        auto decode = [&](uint32_t item, uint32_t &b_out, uint32_t &b_in0) {
            const uint32_t lo = item & ((1u << n_bcast) - 1u), sh = item >> n_bcast;
            if (FWD) {
                b_out = sh | (lo << a.n_sh);
                b_in0 = dev_pdep(sh, a.shared_mask);
            } else {
                b_out = dev_pdep(sh, a.shared_mask) | dev_pdep(lo, a.free_mask);
                b_in0 = sh;
            }
        };
        uint32_t seq = 0;
        uint32_t nb_out = 0, nb_in0 = 0;
        float pf_o0 = 0.f, pf_o1 = 0.f, pf_i0 = 0.f, pf_i1 = 0.f;
        auto prefetch = [&](uint32_t item) {      // F/G vectors one item ahead: their latency never sits between copies
            uint32_t bo = 0, bi = 0;
            if (lane == 0) decode(item, bo, bi);
            nb_out = __shfl_sync(0xffffffffu, bo, 0);
            nb_in0 = __shfl_sync(0xffffffffu, bi, 0);
            if (lane < fgs_o) pf_o0 = a.fg_out[(uint64_t)nb_out * fgs_o + lane];
            if (lane + 32 < fgs_o) pf_o1 = a.fg_out[(uint64_t)nb_out * fgs_o + lane + 32];
            if (!FWD) {
                if (lane < fgs_i) pf_i0 = a.fg_in[(uint64_t)nb_in0 * fgs_i + lane];
                if (lane + 32 < fgs_i) pf_i1 = a.fg_in[(uint64_t)nb_in0 * fgs_i + lane + 32];
            }
        };
        if (blockIdx.x < n_items) prefetch(blockIdx.x);
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const uint32_t b_out = nb_out, b_in0 = nb_in0;
            const float fo0 = pf_o0, fo1 = pf_o1, fi0 = pf_i0, fi1 = pf_i1;
            if (item + gridDim.x < n_items) prefetch(item + gridDim.x);
            if (!MT && (l2_ahead & 255u)) {      // whole-block tiles only: a tiled block (up to 1 MB) would flood L2
                // A forward item holds two stages (alpha, beta) of a three-stage ring, so the beta copy of the next
                // item can only be issued when this item is done and its latency would be exposed.  Ask that block
                // into L2 now, marked evict-last so that it survives until the copy, and let the copy itself mark
                // the lines evict-first (they are read once).  l2_ahead = rounds of look-ahead | 512 (beta blocks
                // only) | 1024 (input blocks only) | 2048 (prefetch evict-last) | 256 (beta copy evict-first); lane j
                // looks after load j of the far item.  Measured on the bench (same-box A/B): 513: +3-5 %
                // on the forward step, 2561: +7 %, 2817 (the default): +8.5 %; input blocks as well: no gain;
                // evict-first on the input copies: -2 %; two or more rounds ahead: 5-25 % slower even with
                // evict-last — the lines do not survive in L2 until they are used.
                const uint64_t far = (uint64_t)item + (uint64_t)(l2_ahead & 255u) * gridDim.x;
                const bool fb = FWD && lane == nred;
                const bool want = (l2_ahead & 512u) ? fb : ((l2_ahead & 1024u) ? !fb : true);
                if (far < n_items && lane < nload && want) {
                    uint32_t bo, bi;
                    decode((uint32_t)far, bo, bi);
                    const uint32_t bk = fb ? bo : (FWD ? (bi | dev_pdep(lane, a.free_mask)) : (bi | (lane << a.n_sh)));
                    if (l2_ahead & 2048u)
                        bulk_prefetch_l2_hint((fb ? a.beta : a.in) + (uint64_t)bk * BS, (fb ? RR : RR + 2u * R + 4u) * 4u, policy_evict_last());
                    else
                        bulk_prefetch_l2((fb ? a.beta : a.in) + (uint64_t)bk * BS, (fb ? RR : RR + 2u * R + 4u) * 4u);
                }
            }
            for (uint32_t t = 0; t < n_tiles; ++t) {
                const uint32_t row0 = t * TR, rows = min(TR, R - row0);
                uint32_t subm = 0;
                for (uint32_t j = 0; j < nload; ++j, ++seq) {
                    const uint32_t s = seq % n_stages;
                    const uint32_t par = ((seq / n_stages) & 1u) ^ 1u;
                    const bool is_beta = FWD && (j == nred);
                    const uint32_t bk = is_beta ? b_out : (FWD ? (b_in0 | subm) : (b_in0 | (j << a.n_sh)));
                    subm = (subm - a.free_mask) & a.free_mask;
                    if (lane == 0) mbar_wait(&s_empty[s], par);
                    __syncwarp();
                    TmaStageMeta *m = &s_meta[s];
                    if (lane == 0) {
                        unsigned char *dst = smem_raw + L.off_stage + s * L.stage_bytes;
                        const float *blk = (is_beta ? a.beta : a.in) + (uint64_t)bk * BS;
                        const uint32_t row_bytes = rows * R * 4u, side_bytes = (2u * R + 4u) * 4u;
                        const bool side = !is_beta && t == 0;
                        if (side && n_tiles == 1) {           // values and side record are contiguous: one copy
                            mbar_expect_tx(&s_full[s], row_bytes + side_bytes);
                            bulk_g2s(dst, blk, row_bytes + side_bytes, &s_full[s]);
                        } else if (is_beta && n_tiles == 1 && (l2_ahead & 256u)) {
                            mbar_expect_tx(&s_full[s], row_bytes);
                            bulk_g2s_hint(dst, blk, row_bytes, &s_full[s], policy_evict_first());
                        } else {
                            mbar_expect_tx(&s_full[s], row_bytes + (side ? side_bytes : 0u));
                            bulk_g2s(dst, blk + (uint64_t)row0 * R, row_bytes, &s_full[s]);
                            if (side) bulk_g2s(dst + L.side_off, blk + RR, side_bytes, &s_full[s]);
                        }
                        m->b_out = b_out;
                        m->b_self = bk;
                    }
                    if (j == 0 && t == 0) {
                        if (lane < fgs_o) m->fg_out[lane] = fo0;
                        if (lane + 32 < fgs_o) m->fg_out[lane + 32] = fo1;
                    }
                    if (!FWD) {
                        if (j == 0) {
                            if (lane < fgs_i) m->fg_self[lane] = fi0;
                            if (lane + 32 < fgs_i) m->fg_self[lane + 32] = fi1;
                        } else {
                            for (uint32_t q = lane; q < fgs_i; q += 32) m->fg_self[q] = a.fg_in[(uint64_t)bk * fgs_i + q];
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_full[s]);
                }
            }
        }
        return;
    }

    // =========================== consumers ===========================
    const uint32_t ctid = tid - 32, cwarp = warp - 1;
    const uint32_t sub = lane / G, gl = lane % G;
    const uint32_t grp = cwarp * rpi + sub;
    float *s_acc = reinterpret_cast<float *>(smem_raw + L.off_acc);
    float *s_side = reinterpret_cast<float *>(smem_raw + L.off_side);
    float *s_colred = reinterpret_cast<float *>(smem_raw + L.off_colred);
    float *s_totred = reinterpret_cast<float *>(smem_raw + L.off_totred);
    float *s_fgo = reinterpret_cast<float *>(smem_raw + L.off_fgo);
    float *s_rsum = reinterpret_cast<float *>(smem_raw + L.off_rsum);
    const bool multi = n_tiles > 1;

    const float inv = (float)(1.0 / *a.in_scale);
    const float cA = (float)a.tA * inv, cB = (float)a.tB * inv, cC = (float)a.tC * inv;
    // this lane's columns: NS groups of V; idle slots shadow the first columns with zero weight and no stores
    bool cok[NS];
    uint32_t ccol[NS];
    uint32_t a2o[NS][V], a2i[NS][V];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        cok[s] = (s * G + gl) * V < R;
        ccol[s] = cok[s] ? (s * G + gl) * V : 0u;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            a2o[s][v] = cok[s] ? s_al_o[ccol[s] + v] : 0u;
            a2i[s][v] = cok[s] ? s_al_i[ccol[s] + v] : 0u;
        }
    }
    uint32_t cokm[NS];                                     // all-ones where the lane's slot holds real columns
#pragma unroll
    for (int s = 0; s < NS; ++s) cokm[s] = cok[s] ? 0xffffffffu : 0u;
    const uint32_t gl0m = gl == 0 ? 0xffffffffu : 0u;
    float qc[NS][V];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int v = 0; v < V; ++v) qc[s][v] = 0.f;
    uint32_t cur_cls = 0xffffffffu;
    float cta_sum = 0.f, sacc = 0.f;
    uint32_t seq = 0, parity_item = 0;

    auto park_q = [&]() {     // posterior accumulators of the finished row class -> this group's private slot
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (cok[s]) {
                float *qd = s_q + (grp * NCo + (cur_cls >> 2)) * Rp + ccol[s];
                float o[V];
                lds_vec<V>(o, qd);
#pragma unroll
                for (int v = 0; v < V; ++v) { o[v] += qc[s][v]; qc[s][v] = 0.f; }
                st_vec<V>(qd, o);
            }
    };

    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, parity_item ^= 1u) {
        float kx[NS][V], kx1[NS][V], kr[NS][V], kc[NS][V], gout[NS][V], colacc[NS][V];
        float totacc = 0.f, totv = 0.f;
        const float *fgo = nullptr, *fgi = nullptr;
        float *outp = nullptr;
        bool weigh = false;
        float *rsum = s_rsum + parity_item * Rp;

        for (uint32_t t = 0; t < n_tiles; ++t) {
            const uint32_t row0 = t * TR, rows = min(TR, R - row0);
            const uint32_t s0 = seq % n_stages;
            mbar_wait(&s_full[s0], (seq / n_stages) & 1u);
            const TmaStageMeta *m0 = &s_meta[s0];
            const uint32_t b_out = m0->b_out;     // read before the stage can go back to the producer
            const float *xs = reinterpret_cast<const float *>(smem_raw + L.off_stage + s0 * L.stage_bytes);
            const float *side = reinterpret_cast<const float *>(smem_raw + L.off_stage + s0 * L.stage_bytes + L.side_off);
            bool early_release = false;
            if (accum) {
                // fold the reduce set into s_acc (and, on the first tile, the side records into s_side); every thread
                // owns the same elements for all k.  Nothing is held across the loop: the set may exceed the ring.
                for (uint32_t k = 0; k < nred; ++k) {
                    const uint32_t sk = (seq + k) % n_stages;
                    if (k > 0) mbar_wait(&s_full[sk], ((seq + k) / n_stages) & 1u);
                    const unsigned char *stb = smem_raw + L.off_stage + sk * L.stage_bytes;
                    const float *st = reinterpret_cast<const float *>(stb);
                    const float *fgk = s_meta[sk].fg_self;
                    float gk[NS][V];
#pragma unroll
                    for (int s = 0; s < NS; ++s)
#pragma unroll
                        for (int v = 0; v < V; ++v) gk[s][v] = FWD ? 1.f : fgk[NCi + a2i[s][v]];
                    for (uint32_t it = 0; it < L.n_iter_t; ++it) {
                        const uint32_t rl = it * ngrp + grp;     // row inside the tile, natural order
                        if (rl < rows) {
                            const float fk = FWD ? 1.f : fgk[s_al_i[row0 + rl]];
#pragma unroll
                            for (int s = 0; s < NS; ++s)
                                if (cok[s]) {
                                    float xv[V];
                                    lds_vec<V>(xv, st + rl * R + ccol[s]);
                                    if (!FWD) {
#pragma unroll
                                        for (int v = 0; v < V; ++v) xv[v] *= fk * gk[s][v];
                                    }
                                    float *dst = s_acc + rl * R + ccol[s];
                                    if (k > 0) {
                                        float o[V];
                                        lds_vec<V>(o, dst);
#pragma unroll
                                        for (int v = 0; v < V; ++v) xv[v] += o[v];
                                    }
                                    st_vec<V>(dst, xv);
                                }
                        }
                    }
                    if (t == 0) {
                        const float *sd = reinterpret_cast<const float *>(stb + L.side_off);
                        for (uint32_t q = ctid; q < 2 * R + 1; q += nct) s_side[q] = (k > 0 ? s_side[q] : 0.f) + sd[q];
                        if (k == 0)
                            for (uint32_t q = ctid; q < fgs_o; q += nct) s_fgo[q] = m0->fg_out[q];
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_empty[sk]);
                }
                consumer_bar(nct);
                xs = s_acc;
                early_release = true;
            } else if (multi && t == 0) {
                // one input block, several tiles: keep its side record and F/G vector for the later tiles
                for (uint32_t q = ctid; q < 2 * R + 1; q += nct) s_side[q] = side[q];
                for (uint32_t q = ctid; q < fgs_o; q += nct) s_fgo[q] = m0->fg_out[q];
                consumer_bar(nct);
            }
            if (accum || multi) side = s_side;
            const float *xs1 = xs, *side1 = side;
            const TmaStageMeta *m1 = m0;
            uint32_t s1 = 0;
            if (two) {
                s1 = (seq + 1) % n_stages;
                mbar_wait(&s_full[s1], ((seq + 1) / n_stages) & 1u);
                m1 = &s_meta[s1];
                xs1 = reinterpret_cast<const float *>(smem_raw + L.off_stage + s1 * L.stage_bytes);
                side1 = reinterpret_cast<const float *>(smem_raw + L.off_stage + s1 * L.stage_bytes + L.side_off);
            }
            const float *bs = xs;
            uint32_t sb = 0;
            if (FWD) {
                sb = (seq + nred) % n_stages;
                mbar_wait(&s_full[sb], ((seq + nred) / n_stages) & 1u);
                bs = reinterpret_cast<const float *>(smem_raw + L.off_stage + sb * L.stage_bytes);
            }
            if (t == 0) {
                // ---- per-lane hoisted column quantities.  With g = G_b[a(R2)] of the output block:
                //   forward : alpha = F[a(R1)] * ( (g*cA) * x + g * rowterm + g*cbase )
                //   backward: beta  = (cA*gin) * (fk*x) + rowterm + cbase, weight of the side sums w = beta * F*g.
                // The reference zeroes beta where an allele of the pair is missing at c-1 (src/genotypehmm.cpp:309);
                // those states have emission 0 (F/G slot 0), so they reach neither the next backward step nor any
                // posterior and the mask is not applied here (the value stays finite).
                fgo = (accum || multi) ? s_fgo : m0->fg_out;
                weigh = !FWD && !accum;
                totv = side[2 * R] + (two ? side1[2 * R] : 0.f);
                const float ctot = cC * totv;
                outp = a.out + (uint64_t)b_out * BS;
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const float cb = cB * (side[R + ccol[s] + v] + (two ? side1[R + ccol[s] + v] : 0.f)) + ctot;
                        gout[s][v] = fgo[NCo + a2o[s][v]];
                        kx1[s][v] = (two && !FWD) ? cA * m1->fg_self[NCi + a2i[s][v]] : 0.f;
                        if (FWD) {
                            kx[s][v] = gout[s][v] * cA; kr[s][v] = gout[s][v]; kc[s][v] = gout[s][v] * cb;
                        } else {
                            kx[s][v] = cA * (weigh ? m0->fg_self[NCi + a2i[s][v]] : 1.f); kr[s][v] = 1.f; kc[s][v] = cb;
                        }
                        colacc[s][v] = 0.f;
                    }
            }
            fgi = s_meta[s0].fg_self;     // backward, nred == 1: F of the (single) input block, valid for this tile's stage
            const char *xrow = reinterpret_cast<const char *>(side);
            const char *fgo_b = reinterpret_cast<const char *>(fgo);
            const char *fgi_b = reinterpret_cast<const char *>(fgi);
            const char *xcolp[NS], *xcolp1[NS], *bcolp[NS];
            char *ocolp[NS];
            const char *xrow1 = reinterpret_cast<const char *>(side1);
            const char *fgi1_b = reinterpret_cast<const char *>(m1->fg_self);
#pragma unroll
            for (int s = 0; s < NS; ++s) {      // per-lane column pointers: the row loop only adds the row offset
                xcolp[s] = reinterpret_cast<const char *>(xs + ccol[s]);
                xcolp1[s] = reinterpret_cast<const char *>(xs1 + ccol[s]);
                bcolp[s] = reinterpret_cast<const char *>(bs + ccol[s]);
                ocolp[s] = reinterpret_cast<char *>(outp + (uint64_t)row0 * R + ccol[s]);
            }
            const uint2 *meta_t = s_rowmeta + t * L.TRp + grp;

            // one row of the block: `seg_cls4` >= 0 means the caller iterates a segment of rows of one allele class
            // (class * 4 given, F of that class in `seg_fo`), so neither is looked up nor checked per row
            auto row_body = [&](auto seg_tag, const uint32_t it, const uint32_t seg_cls4, const float seg_fo) {
                constexpr bool SEG = decltype(seg_tag)::value;      // compile-time: two separate loop bodies
                const uint2 meta = meta_t[it * ngrp];
                const uint32_t rowb = meta.x & 0x7fffffffu;
                const uint32_t r1b = meta.y & 0xfffu, a1ib = (meta.y >> 19) & 0x7fu;
                const uint32_t a1ob = SEG ? seg_cls4 : (meta.y >> 12) & 0x7fu;
                float rowsum_in = *reinterpret_cast<const float *>(xrow + r1b);
                float fk1 = 0.f;
                if (two) {
                    rowsum_in += *reinterpret_cast<const float *>(xrow1 + r1b);
                    if (!FWD) fk1 = *reinterpret_cast<const float *>(fgi1_b + a1ib);
                }
                const float rowterm = cB * rowsum_in;
                const float fo = SEG ? seg_fo : *reinterpret_cast<const float *>(fgo_b + a1ob);
                const float fk = weigh ? *reinterpret_cast<const float *>(fgi_b + a1ib) : 1.f;
                if (FWD && !SEG && a1ob != cur_cls) {   // row class changed: park the posterior accumulators
                    if (cur_cls != 0xffffffffu) park_q();
                    cur_cls = a1ob;
                }
                float rs = 0.f;
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    float x[V], x1[V];
                    lds_vec<V>(x, xcolp[s] + rowb);
#pragma unroll
                    for (int v = 0; v < V; ++v) x1[v] = 0.f;
                    if (two) {
                        float y[V];
                        lds_vec<V>(y, xcolp1[s] + rowb);
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            if (FWD) x[v] += y[v];
                            else x1[v] = fk1 * y[v];
                        }
                    }
                    float o[V];
                    if (FWD) {
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            o[v] = fo * fmaf(kx[s][v], x[v], fmaf(kr[s][v], rowterm, kc[s][v]));
                            colacc[s][v] += o[v];
                        }
                        rs += V == 4 ? (o[0] + o[1]) + (o[V > 2 ? 2 : 0] + o[V > 3 ? 3 : 0]) : o[0] + o[1];
                    } else {
                        float dot = 0.f;
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            const float rest = two ? fmaf(kx1[s][v], x1[v], rowterm + kc[s][v]) : rowterm + kc[s][v];
                            o[v] = fmaf(kx[s][v], fk * x[v], rest);
                            dot = fmaf(o[v], gout[s][v], dot);
                            colacc[s][v] = fmaf(fo, o[v], colacc[s][v]);      // times gout once per block, below
                        }
                        rs = fmaf(fo, dot, rs);
                    }
                    // valid row (sign bit of the metadata word) AND a live column slot, in one test
                    if ((int32_t)(meta.x & cokm[s]) < 0) st_vec<V>(ocolp[s] + rowb, o);
                    if (FWD) {
                        float bv[V];
                        lds_vec<V>(bv, bcolp[s] + rowb);
#pragma unroll
                        for (int v = 0; v < V; ++v) qc[s][v] = fmaf(o[v], bv[v], qc[s][v]);
                    }
                }
#pragma unroll
                for (int sh = G >> 1; sh > 0; sh >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, sh);
                if ((int32_t)(meta.x & gl0m) < 0) *reinterpret_cast<float *>(reinterpret_cast<char *>(rsum) + r1b) = rs;
                totacc += rs;        // identical in every lane of the row group; padding rows contribute 0
            };

            if (FWD && !MT && G == 32) {
                // single tile, one row group per warp: rows are class-sorted, so walk them class by class
                uint32_t it = 0;
                for (uint32_t cls = 0; cls < NCo; ++cls) {
                    const uint32_t end = s_seg[grp * kMaxClasses + cls];
                    if (it < end) {
                        const float seg_fo = fgo[cls];
#pragma unroll 4
                        for (; it < end; ++it) row_body(std::true_type{}, it, cls * 4, seg_fo);
                        cur_cls = cls * 4;
                        park_q();
                    }
                }
                cur_cls = 0xffffffffu;
            } else {
#pragma unroll (NS == 1 ? (CWT == 16 ? 3 : 4) : 2)
                for (uint32_t it = 0; it < L.n_iter_t; ++it) row_body(std::false_type{}, it, 0u, 0.f);
            }
            // ---- hand the stages back
            __syncwarp();
            if (lane == 0) {
                if (!early_release) mbar_arrive(&s_empty[s0]);
                if (two) mbar_arrive(&s_empty[s1]);
                if (FWD) mbar_arrive(&s_empty[sb]);
            }
            seq += nload;
            if (accum && t + 1 < n_tiles) consumer_bar(nct);   // s_acc is rewritten by the next tile
        }
        // backward: the column scale.  Without missing alleles sum(beta_{c-1}[b']) equals the reduced input total
        // exactly (tA + 2R tB + R^2 tC = 1); any positive scalar of that magnitude serves, so use it.
        if (!FWD && ctid == nct - 1) sacc += totv;
        // ---- column sums across the row groups (double-buffered scratch: one barrier per item)
        float *cr = s_colred + parity_item * ngrp * Rp;
        float *tr = s_totred + parity_item * ngrp;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (!FWD) {
#pragma unroll
                for (int v = 0; v < V; ++v) colacc[s][v] *= gout[s][v];
            }
            if (cok[s]) st_vec<V>(cr + grp * Rp + ccol[s], colacc[s]);
        }
        if (gl == 0) tr[grp] = totacc;
        consumer_bar(nct);
        for (uint32_t q = ctid; q < R; q += nct) {
            float cs = 0.f;
            for (uint32_t g = 0; g < ngrp; ++g) cs += cr[g * Rp + q];
            outp[RR + R + q] = cs;
            outp[RR + q] = rsum[q];
        }
        if (ctid == nct - 1) {
            float ts = 0.f;
            for (uint32_t g = 0; g < ngrp; ++g) ts += tr[g];
            outp[RR + 2 * R] = ts;
            cta_sum += ts;
        }
    }

    // ---- kernel epilogue (consumer threads only)
    float *mypart = a.partial + (uint64_t)blockIdx.x * kPartialStride;
    if (FWD) {
        if (cur_cls != 0xffffffffu) park_q();
        consumer_bar(nct);
        for (uint32_t idx = ctid; idx < NCo * R; idx += nct) {
            const uint32_t cls = idx / R, t = idx % R;
            float v = 0.f;
            for (uint32_t g = 0; g < ngrp; ++g) v += s_q[(g * NCo + cls) * Rp + t];
            s_q[cls * Rp + t] = v;
        }
        consumer_bar(nct);
        for (uint32_t p = ctid; p < NCo * NCo; p += nct) {
            const uint32_t i = p / NCo, j = p % NCo;
            float w = 0.f;
            for (uint32_t t = 0; t < R; ++t)
                if (s_al_o[t] == j) w += s_q[i * Rp + t];
            mypart[p] = w;
        }
    } else {
        cta_sum = sacc * inv;     // thread nct-1 holds it; inv puts it on the scale of the stored values
    }
    if (ctid == nct - 1) mypart[kMaxClasses * kMaxClasses] = cta_sum;
    __threadfence();
    consumer_bar(nct);
    if (ctid == 0) {
        const unsigned int ticket = atomicAdd(a.counter, 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    consumer_bar(nct);
    if (!s_last) return;
    __threadfence();
    // the last CTA's final reduction borrows the (now idle) first stage for its fp64 scratch
    double *s_W = reinterpret_cast<double *>(smem_raw + L.off_stage);
    double *s_L = s_W + kPartialStride;
    const uint32_t npair = FWD ? NCo * NCo : 0u;
    for (uint32_t p = cwarp; p < npair + 1; p += CW) {
        const uint32_t slot = (p == npair) ? (uint32_t)(kMaxClasses * kMaxClasses) : p;
        const double acc = warp_sum_partials(a.partial, gridDim.x, slot, lane);
        if (lane == 0) s_W[slot] = acc;
    }
    consumer_bar(nct);
    if (ctid == 0) {
        *a.out_scale = s_W[kMaxClasses * kMaxClasses];
        *a.counter = 0u;
        if (FWD) {
            const uint32_t nA = a.nA_out, ng = nA * (nA + 1) / 2;
            for (uint32_t g = 0; g < ng; ++g) s_L[g] = 0.0;
            double norm = 0.0;
            for (uint32_t i = 1; i <= nA; ++i)
                for (uint32_t j = 1; j <= nA; ++j) {
                    const double w = s_W[i * NCo + j];
                    s_L[(i - 1) + (j - 1) * j / 2] += w;
                    norm += w;
                }
            for (uint32_t g = 0; g < ng; ++g) a.result[g] = s_L[g] / norm;
        }
    }
}

}  // namespace gg
