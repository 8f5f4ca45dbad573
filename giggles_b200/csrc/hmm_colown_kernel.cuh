// Pipelined column step for mid-size panels (32 < R <= 128, fp32, R % 4 == 0; configs[1]'s 88 haplotypes) with the
// column-owning consumer layout of hmm_wide_kernel.cuh: a block is ONE tile (<= 64 KB).
//
// step_tma_kernel gives a warp one row per iteration: at R = 88 that is 22 of 32 lanes at work, a 5-step shuffle
// reduction and ~30 instructions of per-row bookkeeping for every row (ncu: the backward step is bound by issue slots
// and LDS / SHFL latency, not by HBM).  Here a warp's four 8-lane groups work on FOUR rows at a time, each lane owning
// one float4 of columns for all rows of the block: 32 columns per warp, ceil(R / 32) column groups (R = 88: 3 groups,
// 22 of 24 lanes busy), RH row interleaves.  The per-row bookkeeping is shared by four rows, a row sum is a 3-step
// butterfly plus one fold over the column groups per block, column sums are private to a lane (one fold over the row
// slots by shuffle), and the posterior is accumulated per (row interleave, row slot, allele class of the row, column)
// in shared memory with rows in natural order (no class sorting, no parking).  One consumer barrier per block as
// before (row / column partials are double-buffered; the block total is one warp's job).
// Producer warp, stage ring, mbarrier protocol, F/G prefetch one item ahead and the forward step's L2 look-ahead for
// the beta block are those of step_tma_kernel.  Reduce sets of one or two blocks (TWO); larger ones stay there.
#pragma once
#include "hmm_tma_kernel.cuh"

namespace gg {

constexpr uint32_t kColownMaxWarps = 9;        // consumer warps: column groups x row interleaves (two CTAs per SM at 96 registers; 12 warps at 72 registers measured slower)

struct ColownGeom {
    uint32_t side_off, stage_bytes;
    uint32_t off_stage, off_meta, off_q, off_rowpart, off_colpart, off_al_o, off_al_i, total;
};
__host__ __device__ inline ColownGeom colown_geom(uint32_t R, uint32_t n_stages, uint32_t CW, uint32_t RH, uint32_t ncls_out, bool fwd) {
    ColownGeom g;
    const uint32_t Rp = (R + 3u) & ~3u;
    g.side_off = R * R * 4u;
    g.stage_bytes = (g.side_off + (2u * R + 4u) * 4u + 127u) & ~127u;
    uint32_t o = 0;
    g.off_stage = o; o += n_stages * g.stage_bytes;
    g.off_meta = o; o += n_stages * (uint32_t)sizeof(TmaStageMeta);
    o = (o + 15u) & ~15u;
    g.off_q = o; o += fwd ? RH * 4u * ncls_out * Rp * 4u : 0u;      // posterior: [row interleave][row slot][class][column]
    g.off_rowpart = o; o += 2u * CW * Rp * 4u;                      // row sums per column group, by item parity
    g.off_colpart = o; o += 2u * RH * Rp * 4u;                      // column sums per row interleave, by item parity
    g.off_al_o = o; o += Rp;
    g.off_al_i = o; o += Rp;
    g.total = (o + 127u) & ~127u;
    return g;
}

template <bool FWD, bool TWO>
__global__ void __launch_bounds__(32 * (1 + kColownMaxWarps), 2)
colown_step_kernel(const StepArgs<float> a, const uint32_t n_stages, const uint32_t CW, const uint32_t RH, const uint32_t l2_ahead) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kTmaMaxStages];
    __shared__ __align__(8) uint64_t s_empty[kTmaMaxStages];
    __shared__ bool s_last;

    const uint32_t R = a.R, BS = a.BS, RR = R * R, Rp = (R + 3u) & ~3u;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t NW = CW * RH, nct = 32 * NW;
    const uint32_t NCo = a.nA_out + 1, NCi = a.nA_in + 1;
    const uint32_t fgs_o = 2 * NCo, fgs_i = 2 * NCi;
    constexpr uint32_t nred = TWO ? 2u : 1u;
    const uint32_t n_bcast = FWD ? (a.u_out - a.n_sh) : a.n_free_left;
    constexpr uint32_t nload = nred + (FWD ? 1u : 0u);
    const ColownGeom L = colown_geom(R, n_stages, CW, RH, NCo, FWD);
    TmaStageMeta *s_meta = reinterpret_cast<TmaStageMeta *>(smem_raw + L.off_meta);
    const uint32_t n_items = 1u << a.u_out;

    if (tid == 0) {
        for (uint32_t s = 0; s < n_stages; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint8_t *s_al_o = smem_raw + L.off_al_o;
    uint8_t *s_al_i = smem_raw + L.off_al_i;
    float *s_q = reinterpret_cast<float *>(smem_raw + L.off_q);
    for (uint32_t t = tid; t < R; t += blockDim.x) {
        s_al_o[t] = a.al_out[t];
        s_al_i[t] = a.al_in[t];
    }
    if (FWD)
        for (uint32_t t = tid; t < RH * 4u * NCo * Rp; t += blockDim.x) s_q[t] = 0.f;
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
        const uint32_t blk_bytes = RR * 4u, side_bytes = (2u * R + 4u) * 4u;
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const uint32_t b_out = nb_out, b_in0 = nb_in0;
            const float fo0 = pf_o0, fo1 = pf_o1, fi0 = pf_i0, fi1 = pf_i1;
            if (item + gridDim.x < n_items) prefetch(item + gridDim.x);
            if (FWD && (l2_ahead & 255u)) {
                // the forward item holds two stages of a short ring: ask the NEXT round's beta block into L2 now
                // (evict-last), copy it evict-first when its turn comes — see step_tma_kernel
                const uint64_t far = (uint64_t)item + (uint64_t)(l2_ahead & 255u) * gridDim.x;
                if (far < n_items && lane == 0) {
                    uint32_t bo, bi;
                    decode((uint32_t)far, bo, bi);
                    bulk_prefetch_l2_hint(a.beta + (uint64_t)bo * BS, blk_bytes, policy_evict_last());
                }
            }
            uint32_t subm = 0;
#pragma unroll
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
                    if (is_beta) {
                        mbar_expect_tx(&s_full[s], blk_bytes);
                        if (l2_ahead & 256u) bulk_g2s_hint(dst, blk, blk_bytes, &s_full[s], policy_evict_first());
                        else bulk_g2s(dst, blk, blk_bytes, &s_full[s]);
                    } else {                               // values and side record are contiguous: one copy
                        mbar_expect_tx(&s_full[s], blk_bytes + side_bytes);
                        bulk_g2s(dst, blk, blk_bytes + side_bytes, &s_full[s]);
                    }
                    m->b_out = b_out;
                    m->b_self = bk;
                }
                if (j == 0) {
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
        return;
    }

    // =========================== consumers ===========================
    const uint32_t ctid = tid - 32, cwarp = warp - 1;
    const uint32_t cg = cwarp % CW, rh = cwarp / CW;        // column group, row interleave
    const uint32_t slot = lane >> 3, j8 = lane & 7u;        // row slot of the warp, lane inside the 8-lane group
    const uint32_t cv = cg * 8u + j8;
    const bool cok = cv * 4u < R;
    const uint32_t col = cok ? cv * 4u : 0u;
    float *s_rowpart = reinterpret_cast<float *>(smem_raw + L.off_rowpart);
    float *s_colpart = reinterpret_cast<float *>(smem_raw + L.off_colpart);

    const float inv = (float)(1.0 / *a.in_scale);
    const float cA = (float)a.tA * inv, cB = (float)a.tB * inv, cC = (float)a.tC * inv;
    uint32_t a2o[4], a2i[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        a2o[v] = s_al_o[col + v];
        a2i[v] = s_al_i[col + v];
    }
    float *q_base = s_q + (size_t)((rh * 4u + slot) * NCo) * Rp + col;     // + class * Rp
    float cta_sum = 0.f, sacc = 0.f;
    uint32_t seq = 0, par_item = 0;
    const uint32_t n_quad = (R + 3u) >> 2;

    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, par_item ^= 1u) {
        const uint32_t s0 = seq % n_stages;
        mbar_wait(&s_full[s0], (seq / n_stages) & 1u);
        const float *st0 = reinterpret_cast<const float *>(smem_raw + L.off_stage + s0 * L.stage_bytes);
        const float *st1 = st0;
        const TmaStageMeta *m0 = &s_meta[s0], *m1 = m0;
        uint32_t s1 = 0, sb = 0;
        if (TWO) {
            s1 = (seq + 1) % n_stages;
            mbar_wait(&s_full[s1], ((seq + 1) / n_stages) & 1u);
            st1 = reinterpret_cast<const float *>(smem_raw + L.off_stage + s1 * L.stage_bytes);
            m1 = &s_meta[s1];
        }
        const float *stb = st0;
        if (FWD) {
            sb = (seq + nred) % n_stages;
            mbar_wait(&s_full[sb], ((seq + nred) / n_stages) & 1u);
            stb = reinterpret_cast<const float *>(smem_raw + L.off_stage + sb * L.stage_bytes);
        }
        // side record (row sums, column sums, total of the input block) sits right behind the values in the stage
        const float *sd0 = st0 + RR, *sd1 = st1 + RR;
        const float *fgo = m0->fg_out, *fgi0 = m0->fg_self, *fgi1 = m1->fg_self;
        float *outp = a.out + (uint64_t)m0->b_out * BS;
        float *rowpart = s_rowpart + par_item * CW * Rp;
        float *colpart = s_colpart + par_item * RH * Rp;
        const float totv = sd0[2 * R] + (TWO ? sd1[2 * R] : 0.f);
        const float ctot = cC * totv;
        float kx[4], kx1[4], kr[4], kc[4], gout[4], colacc[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float cb = cB * (sd0[R + col + v] + (TWO ? sd1[R + col + v] : 0.f)) + ctot;
            const float g = cok ? fgo[NCo + a2o[v]] : 0.f;
            gout[v] = g;
            if (FWD) {
                kx[v] = g * cA; kr[v] = g; kc[v] = g * cb; kx1[v] = 0.f;
            } else {
                kx[v] = cok ? cA * fgi0[NCi + a2i[v]] : 0.f;
                kx1[v] = (TWO && cok) ? cA * fgi1[NCi + a2i[v]] : 0.f;
                kr[v] = 1.f; kc[v] = cb;
            }
            colacc[v] = 0.f;
        }
#pragma unroll 2
        for (uint32_t it = rh; it < n_quad; it += RH) {
            const uint32_t r = 4u * it + slot;
            const bool valid = r < R;
            const uint32_t R1 = valid ? r : 0u;
            const uint32_t a1o = s_al_o[R1], a1i = s_al_i[R1];
            const float rowterm = cB * (sd0[R1] + (TWO ? sd1[R1] : 0.f));
            const float fo = valid ? fgo[a1o] : 0.f;
            const uint32_t eoff = R1 * R + col;
            const float4 x4 = *reinterpret_cast<const float4 *>(st0 + eoff);
            float x[4] = {x4.x, x4.y, x4.z, x4.w};
            float o[4];
            float rs;
            if (FWD) {
                if (TWO) {
                    const float4 y4 = *reinterpret_cast<const float4 *>(st1 + eoff);
                    x[0] += y4.x; x[1] += y4.y; x[2] += y4.z; x[3] += y4.w;
                }
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    o[v] = fo * fmaf(kx[v], x[v], fmaf(kr[v], rowterm, kc[v]));
                    colacc[v] += o[v];
                }
                rs = (o[0] + o[1]) + (o[2] + o[3]);
            } else {
                const float fk = fgi0[a1i];
                float y[4] = {0.f, 0.f, 0.f, 0.f};
                float fk1 = 0.f;
                if (TWO) {
                    const float4 y4 = *reinterpret_cast<const float4 *>(st1 + eoff);
                    y[0] = y4.x; y[1] = y4.y; y[2] = y4.z; y[3] = y4.w;
                    fk1 = fgi1[a1i];
                }
                float dot = 0.f;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float rest = TWO ? fmaf(kx1[v], fk1 * y[v], rowterm + kc[v]) : rowterm + kc[v];
                    o[v] = fmaf(kx[v], fk * x[v], rest);
                    dot = fmaf(o[v], gout[v], dot);
                    colacc[v] = fmaf(fo, o[v], colacc[v]);       // times gout once per block, below
                }
                rs = fo * dot;
            }
            if (valid && cok) {
                *reinterpret_cast<float4 *>(outp + eoff) = make_float4(o[0], o[1], o[2], o[3]);
                if (FWD) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(stb + eoff);
                    float4 *qp = reinterpret_cast<float4 *>(q_base + (size_t)a1o * Rp);
                    float4 q4 = *qp;
                    q4.x = fmaf(o[0], b4.x, q4.x); q4.y = fmaf(o[1], b4.y, q4.y);
                    q4.z = fmaf(o[2], b4.z, q4.z); q4.w = fmaf(o[3], b4.w, q4.w);
                    *qp = q4;
                }
            }
#pragma unroll
            for (int sh = 4; sh > 0; sh >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, sh);
            if (j8 == 0 && valid) rowpart[cg * Rp + R1] = rs;
        }
        // ---- hand the stages back
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&s_empty[s0]);
            if (TWO) mbar_arrive(&s_empty[s1]);
            if (FWD) mbar_arrive(&s_empty[sb]);
        }
        seq += nload;
        // ---- column sums over the warp's four row slots, then (after the barrier) over the row interleaves; row sums
        //      over the column groups; block total by the first consumer warp.  Fixed orders throughout.
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (!FWD) colacc[v] *= gout[v];
            colacc[v] += __shfl_xor_sync(0xffffffffu, colacc[v], 8);
            colacc[v] += __shfl_xor_sync(0xffffffffu, colacc[v], 16);
        }
        if (slot == 0 && cok) *reinterpret_cast<float4 *>(colpart + rh * Rp + col) = make_float4(colacc[0], colacc[1], colacc[2], colacc[3]);
        consumer_bar(nct);
        for (uint32_t q = ctid; q < R; q += nct) {
            float rsum = 0.f, cs = 0.f;
            for (uint32_t g = 0; g < CW; ++g) rsum += rowpart[g * Rp + q];
            for (uint32_t g = 0; g < RH; ++g) cs += colpart[g * Rp + q];
            outp[RR + q] = rsum;
            outp[RR + R + q] = cs;
        }
        if (cwarp == 0) {
            float tp = 0.f;
            for (uint32_t q = lane; q < R; q += 32) {
                float rsum = 0.f;
                for (uint32_t g = 0; g < CW; ++g) rsum += rowpart[g * Rp + q];
                tp += rsum;
            }
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) tp += __shfl_xor_sync(0xffffffffu, tp, sh);
            if (lane == 0) {
                outp[RR + 2 * R] = tp;
                cta_sum += tp;
                sacc += totv;
            }
        }
    }

    // ---- kernel epilogue (consumer threads only): as step_tma_kernel; thread 0 of the first consumer warp holds the sums
    float *mypart = a.partial + (uint64_t)blockIdx.x * kPartialStride;
    if (FWD) {
        consumer_bar(nct);
        for (uint32_t idx = ctid; idx < NCo * R; idx += nct) {
            const uint32_t cls = idx / R, t = idx % R;
            float v = 0.f;
            for (uint32_t g = 0; g < RH * 4u; ++g) v += s_q[(size_t)(g * NCo + cls) * Rp + t];
            s_q[(size_t)cls * Rp + t] = v;
        }
        consumer_bar(nct);
        for (uint32_t p = ctid; p < NCo * NCo; p += nct) {
            const uint32_t i = p / NCo, jc = p % NCo;
            float w = 0.f;
            for (uint32_t t = 0; t < R; ++t)
                if (s_al_o[t] == jc) w += s_q[(size_t)i * Rp + t];
            mypart[p] = w;
        }
    } else {
        cta_sum = sacc * inv;
    }
    if (ctid == 0) mypart[kMaxClasses * kMaxClasses] = cta_sum;
    __threadfence();
    consumer_bar(nct);
    if (ctid == 0) {
        const unsigned int ticket = atomicAdd(a.counter, 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    consumer_bar(nct);
    if (!s_last) return;
    __threadfence();
    double *s_W = reinterpret_cast<double *>(smem_raw + L.off_stage);
    double *s_L = s_W + kPartialStride;
    const uint32_t npair = FWD ? NCo * NCo : 0u;
    for (uint32_t p = cwarp; p < npair + 1; p += NW) {
        const uint32_t slt = (p == npair) ? (uint32_t)(kMaxClasses * kMaxClasses) : p;
        const double acc = warp_sum_partials(a.partial, gridDim.x, slt, lane);
        if (lane == 0) s_W[slt] = acc;
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
                for (uint32_t jj = 1; jj <= nA; ++jj) {
                    const double w = s_W[i * NCo + jj];
                    s_L[(i - 1) + (jj - 1) * jj / 2] += w;
                    norm += w;
                }
            for (uint32_t g = 0; g < ng; ++g) a.result[g] = s_L[g] / norm;
        }
    }
}

}  // namespace gg
