// Pipelined column step for WIDE panels (128 < R <= 512, fp32, R % 4 == 0): warps own COLUMNS.
//
// step_tma_kernel gives every consumer warp whole rows (32 lanes x NS float4).  At R = 400 that is four vectors per
// lane and 168 registers, and the forward step's posterior scratch — one accumulator per (row group, allele class of
// the row, column) — grows to 8 x 17 x 400 floats on 16-allele columns, so that only two to four consumer warps fit
// beside the stages (configs[4]: forward step at 0.22 of the measured HBM peak).  Here a consumer warp owns 64
// columns (16 float4; its two half-warps work on two rows at a time) for ALL rows of a block:
//   * column sums and the posterior accumulators are private to one lane: the posterior scratch is one
//     accumulator per (row interleave, allele class of the row, column) whatever the number of column groups, and
//     rows are visited in natural order (no class sorting, no parking);
//   * row sums become a 16-lane shuffle plus one fold over the column-owning warps per block;
//   * one float4 of state per lane: ~60 registers, no spills.
// The producer warp, the stage ring and the mbarrier protocol are those of step_tma_kernel with row tiles: a work
// item is one output block, cut into tiles of TR rows; per tile one cp.async.bulk per input block of the reduce set
// (the first tile also brings the block's side record) and, forward, one for the beta rows.  Reduce sets of one or
// two blocks (TWO) are handled; larger ones stay with step_tma_kernel.
#pragma once
#include "hmm_tma_kernel.cuh"

namespace gg {

constexpr uint32_t kWideMaxWarps = 16;     // consumer warps: CW column groups x RH row interleaves

struct WideGeom {
    uint32_t TR, n_tiles, side_off, stage_bytes;
    uint32_t off_stage, off_meta, off_side, off_fgo, off_q, off_rowpart, off_colpart, off_totred, off_al_o, off_al_i, total;
};
__host__ __device__ inline WideGeom wide_geom(uint32_t R, uint32_t TR, uint32_t n_stages, uint32_t CW, uint32_t RH, uint32_t ncls_out, bool fwd) {
    WideGeom g;
    const uint32_t Rp = (R + 3u) & ~3u;
    g.TR = TR;
    g.n_tiles = (R + TR - 1) / TR;
    g.side_off = TR * R * 4u;
    g.stage_bytes = (g.side_off + (2u * R + 4u) * 4u + 127u) & ~127u;
    uint32_t o = 0;
    g.off_stage = o; o += n_stages * g.stage_bytes;
    g.off_meta = o; o += n_stages * (uint32_t)sizeof(TmaStageMeta);
    o = (o + 15u) & ~15u;
    g.off_side = o; o += (2u * Rp + 4u) * 4u;                   // side sums of the (reduced) input block
    g.off_fgo = o; o += 48u * 4u;                               // F/G of the output block
    g.off_q = o; o += fwd ? RH * ncls_out * Rp * 4u : 0u;       // posterior: [row interleave][class of the row][column]
    g.off_rowpart = o; o += CW * Rp * 4u;                       // row sums per column group
    g.off_colpart = o; o += RH * Rp * 4u;
    g.off_totred = o; o += 2u * kWideMaxWarps * 4u;
    g.off_al_o = o; o += Rp;
    g.off_al_i = o; o += Rp;
    g.total = (o + 127u) & ~127u;
    return g;
}

template <bool FWD, bool TWO>
__global__ void __launch_bounds__(32 * (1 + kWideMaxWarps), 1)
wide_step_kernel(const StepArgs<float> a, const uint32_t n_stages, const uint32_t TR, const uint32_t CW, const uint32_t RH) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kTmaMaxStages];
    __shared__ __align__(8) uint64_t s_empty[kTmaMaxStages];
    __shared__ bool s_last;

    const uint32_t R = a.R, BS = a.BS, RR = R * R, Rp = (R + 3u) & ~3u;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t NW = CW * RH, nct = 32 * NW;
    const uint32_t NCo = a.nA_out + 1, NCi = a.nA_in + 1;
    const uint32_t fgs_o = 2 * NCo, fgs_i = 2 * NCi;
    const uint32_t nred = TWO ? 2u : 1u;
    const uint32_t n_bcast = FWD ? (a.u_out - a.n_sh) : a.n_free_left;
    const uint32_t nload = nred + (FWD ? 1u : 0u);
    const WideGeom L = wide_geom(R, TR, n_stages, CW, RH, NCo, FWD);
    const uint32_t n_tiles = L.n_tiles;
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
        for (uint32_t t = tid; t < RH * NCo * Rp; t += blockDim.x) s_q[t] = 0.f;
    __syncthreads();

    if (warp == 0) {
        // =========================== producer (as step_tma_kernel, tiled blocks) ===========================
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
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            uint32_t bo = 0, bi = 0;
            if (lane == 0) decode(item, bo, bi);
            const uint32_t b_out = __shfl_sync(0xffffffffu, bo, 0), b_in0 = __shfl_sync(0xffffffffu, bi, 0);
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
                        mbar_expect_tx(&s_full[s], row_bytes + (side ? side_bytes : 0u));
                        bulk_g2s(dst, blk + (uint64_t)row0 * R, row_bytes, &s_full[s]);
                        if (side) bulk_g2s(dst + L.side_off, blk + RR, side_bytes, &s_full[s]);
                        m->b_out = b_out;
                        m->b_self = bk;
                    }
                    if (j == 0 && t == 0)
                        for (uint32_t q = lane; q < fgs_o; q += 32) m->fg_out[q] = a.fg_out[(uint64_t)b_out * fgs_o + q];
                    if (!FWD)
                        for (uint32_t q = lane; q < fgs_i; q += 32) m->fg_self[q] = a.fg_in[(uint64_t)bk * fgs_i + q];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_full[s]);
                }
            }
        }
        return;
    }

    // =========================== consumers ===========================
    const uint32_t ctid = tid - 32, cwarp = warp - 1;
    const uint32_t cg = cwarp % CW, rh = cwarp / CW;        // column group, row interleave
    const uint32_t h = lane >> 4, j16 = lane & 15u;
    const uint32_t cv = cg * 16u + j16;
    const bool cok = cv * 4u < R;
    const uint32_t col = cok ? cv * 4u : 0u;
    float *s_rowpart = reinterpret_cast<float *>(smem_raw + L.off_rowpart);
    float *s_colpart = reinterpret_cast<float *>(smem_raw + L.off_colpart);
    float *s_totred = reinterpret_cast<float *>(smem_raw + L.off_totred);

    const float inv = (float)(1.0 / *a.in_scale);
    const float cA = (float)a.tA * inv, cB = (float)a.tB * inv, cC = (float)a.tC * inv;
    uint32_t a2o[4], a2i[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        a2o[v] = s_al_o[col + v];
        a2i[v] = s_al_i[col + v];
    }
    float *q_base = s_q + (size_t)(rh * NCo) * Rp + col;     // + class * Rp
    float cta_sum = 0.f, sacc = 0.f;
    uint32_t seq = 0, par_item = 0;

    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, par_item ^= 1u) {
        float kx[4], kx1[4], kr[4], kc[4], gout[4], colacc[4];
        float totv = 0.f;
        float *outp = nullptr;
        // single buffers: the two barriers of the item epilogue separate their last readers from the next item's writers
        float *s_side = reinterpret_cast<float *>(smem_raw + L.off_side);
        float *s_fgo = reinterpret_cast<float *>(smem_raw + L.off_fgo);
        float *rowpart = s_rowpart;

        for (uint32_t t = 0; t < n_tiles; ++t) {
            const uint32_t row0 = t * TR, rows = min(TR, R - row0);
            const uint32_t s0 = seq % n_stages;
            mbar_wait(&s_full[s0], (seq / n_stages) & 1u);
            const unsigned char *st0 = smem_raw + L.off_stage + s0 * L.stage_bytes;
            const unsigned char *st1 = st0;
            const TmaStageMeta *m0 = &s_meta[s0], *m1 = m0;
            uint32_t s1 = 0, sb = 0;
            if (TWO) {
                s1 = (seq + 1) % n_stages;
                mbar_wait(&s_full[s1], ((seq + 1) / n_stages) & 1u);
                st1 = smem_raw + L.off_stage + s1 * L.stage_bytes;
                m1 = &s_meta[s1];
            }
            const unsigned char *stb = st0;
            if (FWD) {
                sb = (seq + nred) % n_stages;
                mbar_wait(&s_full[sb], ((seq + nred) / n_stages) & 1u);
                stb = smem_raw + L.off_stage + sb * L.stage_bytes;
            }
            if (t == 0) {
                // the block's side sums (row, column, total of the reduced input) and F/G of the output block are kept
                // for the later tiles
                const float *sd0 = reinterpret_cast<const float *>(st0 + L.side_off);
                const float *sd1 = reinterpret_cast<const float *>(st1 + L.side_off);
                for (uint32_t q = ctid; q < 2 * R + 1; q += nct) s_side[q] = sd0[q] + (TWO ? sd1[q] : 0.f);
                for (uint32_t q = ctid; q < fgs_o; q += nct) s_fgo[q] = m0->fg_out[q];
                outp = a.out + (uint64_t)m0->b_out * BS;
                consumer_bar(nct);
                totv = s_side[2 * R];
                const float ctot = cC * totv;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float cb = cB * s_side[R + col + v] + ctot;
                    const float g = cok ? s_fgo[NCo + a2o[v]] : 0.f;
                    gout[v] = g;
                    if (FWD) {
                        kx[v] = g * cA; kr[v] = g; kc[v] = g * cb; kx1[v] = 0.f;
                    } else {
                        kx[v] = cok ? cA * m0->fg_self[NCi + a2i[v]] : 0.f;
                        kx1[v] = (TWO && cok) ? cA * m1->fg_self[NCi + a2i[v]] : 0.f;
                        kr[v] = 1.f; kc[v] = cb;
                    }
                    colacc[v] = 0.f;
                }
            }
            const float *fgi0 = m0->fg_self, *fgi1 = m1->fg_self;
            const uint32_t npair = (rows + 1u) >> 1;
#pragma unroll 2
            for (uint32_t it = rh; it < npair; it += RH) {
                const uint32_t r = 2u * it + h;
                const bool valid = r < rows;
                const uint32_t rr = valid ? r : 0u;
                const uint32_t R1 = row0 + rr;
                const uint32_t a1o = s_al_o[R1], a1i = s_al_i[R1];
                const float rowterm = cB * s_side[R1];
                const float fo = valid ? s_fgo[a1o] : 0.f;
                const uint32_t eoff = rr * R + col;
                const float4 x4 = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(st0) + eoff);
                float x[4] = {x4.x, x4.y, x4.z, x4.w};
                float o[4];
                float rs;
                if (FWD) {
                    if (TWO) {
                        const float4 y4 = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(st1) + eoff);
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
                        const float4 y4 = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(st1) + eoff);
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
                if (valid && cok) *reinterpret_cast<float4 *>(outp + (uint64_t)R1 * R + col) = make_float4(o[0], o[1], o[2], o[3]);
                if (FWD) {
                    // posterior: q[class of the row][column] += alpha * beta.  The two half-warps hold two rows of the
                    // same columns: when the rows share a class the lower half adds both products, otherwise each half
                    // adds its own — one accumulator array per row interleave, no two lanes on one address
                    const float4 b4 = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(stb) + eoff);
                    const bool live = valid && cok;
                    float p0 = live ? o[0] * b4.x : 0.f, p1 = live ? o[1] * b4.y : 0.f;
                    float p2 = live ? o[2] * b4.z : 0.f, p3 = live ? o[3] * b4.w : 0.f;
                    const uint32_t cls_me = valid ? a1o : 0xffffffffu;
                    const uint32_t cls_ot = __shfl_xor_sync(0xffffffffu, cls_me, 16);
                    const float r0 = __shfl_xor_sync(0xffffffffu, p0, 16), r1 = __shfl_xor_sync(0xffffffffu, p1, 16);
                    const float r2 = __shfl_xor_sync(0xffffffffu, p2, 16), r3 = __shfl_xor_sync(0xffffffffu, p3, 16);
                    const bool same = cls_ot == cls_me;
                    if (h == 0 && same) { p0 += r0; p1 += r1; p2 += r2; p3 += r3; }
                    if (live && (h == 0 || !same)) {
                        float4 *qp = reinterpret_cast<float4 *>(q_base + (size_t)a1o * Rp);
                        float4 q4 = *qp;
                        q4.x += p0; q4.y += p1; q4.z += p2; q4.w += p3;
                        *qp = q4;
                    }
                }
#pragma unroll
                for (int sh = 8; sh > 0; sh >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, sh);
                if (j16 == 0 && valid) rowpart[cg * Rp + R1] = rs;
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&s_empty[s0]);
                if (TWO) mbar_arrive(&s_empty[s1]);
                if (FWD) mbar_arrive(&s_empty[sb]);
            }
            seq += nload;
        }
        // ---- item epilogue: column sums (private to a lane pair; RH > 1: folded over the row interleaves), row sums
        //      folded over the column groups in fixed order, block total
        if (!FWD && ctid == nct - 1) sacc += totv;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (!FWD) colacc[v] *= gout[v];
            colacc[v] += __shfl_xor_sync(0xffffffffu, colacc[v], 16);
        }
        if (RH == 1) {
            if (h == 0 && cok) *reinterpret_cast<float4 *>(outp + RR + R + col) = make_float4(colacc[0], colacc[1], colacc[2], colacc[3]);
        } else if (h == 0 && cok) {
            *reinterpret_cast<float4 *>(s_colpart + rh * Rp + col) = make_float4(colacc[0], colacc[1], colacc[2], colacc[3]);
        }
        consumer_bar(nct);
        float tpart = 0.f;
        for (uint32_t q = ctid; q < R; q += nct) {
            float rsum = 0.f;
            for (uint32_t g = 0; g < CW; ++g) rsum += rowpart[g * Rp + q];
            outp[RR + q] = rsum;
            tpart += rsum;
            if (RH > 1) {
                float cs = 0.f;
                for (uint32_t g = 0; g < RH; ++g) cs += s_colpart[g * Rp + q];
                outp[RR + R + q] = cs;
            }
        }
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) tpart += __shfl_xor_sync(0xffffffffu, tpart, sh);
        float *tr = s_totred + par_item * kWideMaxWarps;
        if (lane == 0) tr[cwarp] = tpart;
        consumer_bar(nct);
        if (ctid == nct - 1) {
            float ts = 0.f;
            for (uint32_t g = 0; g < NW; ++g) ts += tr[g];
            outp[RR + 2 * R] = ts;
            cta_sum += ts;
        }
    }

    // ---- kernel epilogue (consumer threads only): as step_tma_kernel
    float *mypart = a.partial + (uint64_t)blockIdx.x * kPartialStride;
    if (FWD) {
        consumer_bar(nct);
        for (uint32_t idx = ctid; idx < NCo * R; idx += nct) {
            const uint32_t cls = idx / R, t = idx % R;
            float v = 0.f;
            for (uint32_t g = 0; g < RH; ++g) v += s_q[(size_t)(g * NCo + cls) * Rp + t];
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
    double *s_W = reinterpret_cast<double *>(smem_raw + L.off_stage);
    double *s_L = s_W + kPartialStride;
    const uint32_t npair = FWD ? NCo * NCo : 0u;
    for (uint32_t p = cwarp; p < npair + 1; p += NW) {
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
