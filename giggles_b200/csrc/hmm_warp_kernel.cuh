// Small-panel column step: one WARP per bipartition block.
//
// For small panels (R <= 32, e.g. the 10-haplotype panel of BASELINE.json configs[0]) a block is only R*R <= 1024
// values: a CTA-per-block kernel spends its time in block barriers and per-item bookkeeping.  Here every warp owns
// whole blocks (no block-level synchronisation at all): a group of G lanes owns a row, 32/G rows are processed per
// iteration, row sums are shuffles inside the group, column sums shuffles across the groups.  The genotype posterior
// uses the fact that the allele class of a (row, column) POSITION is the same for every block of a column: each lane
// keeps one accumulator per position it owns for the whole kernel and bins them by class once, at the end.
// Same arithmetic as step_kernel (hmm_kernels.cuh); fp32.
#pragma once
#include "hmm_kernels.cuh"

namespace gg {

constexpr int kWarpKernelWarps = 8;

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// One column step of one warp: the lane's column / row identities for this column pair, then its share of the
// output blocks (items first_item, first_item + item_stride, ...).  Shared by the per-column kernel below and by the
// multi-column wave kernel (hmm_wave_kernel.cuh).  qa: per-position posterior accumulators (forward), warp_sum: sum
// of the produced column over this warp's items.
template <int V, int G, bool FWD>
__device__ __forceinline__ void warp_step(const StepArgs<float> &a, const uint8_t *al_out, const uint8_t *al_in, const float inv,
                                          const uint32_t first_item, const uint32_t item_stride,
                                          float (&qa)[(V * G * G / 32) > 0 ? (V * G * G / 32) : 1][V], float &warp_sum) {
    constexpr int NI = (V * G * G / 32) > 0 ? (V * G * G / 32) : 1;
    const uint32_t lane = threadIdx.x & 31u;
    constexpr uint32_t rpi = 32u / G;
    const uint32_t R = a.R;
    const uint32_t sub = lane / G, gl = lane % G;
    const uint32_t NCo = a.nA_out + 1, NCi = a.nA_in + 1;
    const uint32_t fgs_o = 2 * NCo, fgs_i = 2 * NCi;
    const bool has_in = a.in != nullptr;
    const uint64_t RR = (uint64_t)R * R, BS = a.BS;
    const uint64_t o_row = RR, o_col = RR + R, o_tot = RR + 2 * (uint64_t)R;
    const uint32_t nred = !has_in ? 0u : (FWD ? (1u << a.n_free_left) : (1u << (a.u_in - a.n_sh)));
    const uint32_t n_bcast = FWD ? (a.u_out - a.n_sh) : a.n_free_left;
    const uint32_t n_iter = (R + rpi - 1) / rpi;
    // index maps that are plain shifts (no read started or ended out of order): skip the bit-deposit loops
    const bool sh_id = a.shared_mask == ((a.n_sh >= 32 ? 0u : (1u << a.n_sh)) - 1u);
    const bool fr_id = a.free_mask == ((((a.n_free_left >= 32 ? 0u : (1u << a.n_free_left)) - 1u)) << a.n_sh);

    const bool cok = gl * V < R;
    const uint32_t ccol = cok ? gl * V : 0u;
    uint32_t a2o[V], a2i[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        a2o[v] = cok ? al_out[ccol + v] : 0u;
        a2i[v] = (cok && has_in) ? al_in[ccol + v] : 0u;
    }
    // per row slot: row index and its allele classes (fixed for the whole kernel)
    uint32_t rowi[NI], rowoff[NI], a1o[NI], a1i[NI];
    bool rok[NI];
#pragma unroll
    for (int it = 0; it < NI; ++it) {
        const uint32_t r = it * rpi + sub;
        rok[it] = (uint32_t)it < n_iter && r < R;
        rowi[it] = rok[it] ? r : 0u;
        rowoff[it] = rowi[it] * R;
        a1o[it] = rok[it] ? al_out[rowi[it]] : 0u;
        a1i[it] = (rok[it] && has_in) ? al_in[rowi[it]] : 0u;
    }
    const float cA = (float)a.tA * inv, cB = (float)a.tB * inv, cC = (float)a.tC * inv;
    const uint32_t n_items = 1u << a.u_out;
    for (uint32_t item = first_item; item < n_items; item += item_stride) {
        const uint32_t lo = item & ((1u << n_bcast) - 1u), sh = item >> n_bcast;
        uint32_t b_out, b_in0;
        if (FWD) {
            b_out = sh | (lo << a.n_sh);
            b_in0 = sh_id ? sh : dev_pdep(sh, a.shared_mask);
        } else {
            b_out = (sh_id ? sh : dev_pdep(sh, a.shared_mask)) | (fr_id ? (lo << a.n_sh) : dev_pdep(lo, a.free_mask));
            b_in0 = sh;
        }
        if (nred == 1 && item + item_stride < n_items) {
            // the next item's input block, beta block and F/G vectors on their way into L1 while this one is computed
            // (columns of small panels sit in L2; a warp's items are otherwise a chain of exposed L2 round trips)
            const uint32_t nx = item + item_stride;
            const uint32_t nlo = nx & ((1u << n_bcast) - 1u), nsh = nx >> n_bcast;
            uint32_t nb_out, nb_in;
            if (FWD) {
                nb_out = nsh | (nlo << a.n_sh);
                nb_in = sh_id ? nsh : dev_pdep(nsh, a.shared_mask);
            } else {
                nb_out = (sh_id ? nsh : dev_pdep(nsh, a.shared_mask)) | (fr_id ? (nlo << a.n_sh) : dev_pdep(nlo, a.free_mask));
                nb_in = nsh;
            }
            const uint32_t lines = (uint32_t)((BS * 4 + 127) / 128);
            const char *pa = reinterpret_cast<const char *>(a.in + (uint64_t)nb_in * BS);
            const char *pb = FWD ? reinterpret_cast<const char *>(a.beta + (uint64_t)nb_out * BS)
                                 : reinterpret_cast<const char *>(a.fg_in + (uint64_t)nb_in * fgs_i);
            if (lane < lines) prefetch_l1(pa + lane * 128);
            else if (lane < 2 * lines && FWD) prefetch_l1(pb + (lane - lines) * 128);
            else if (lane == 30 && !FWD) prefetch_l1(pb);
            else if (lane == 31) prefetch_l1(a.fg_out + (uint64_t)nb_out * fgs_o);
        }
        const float *fgo = a.fg_out + (uint64_t)b_out * fgs_o;
        float colv[V], gout[V], gin0[V], colacc[V];
        float totv = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            colv[v] = 0.f;
            colacc[v] = 0.f;
            gout[v] = fgo[NCo + a2o[v]];
            gin0[v] = (!FWD && has_in) ? a.fg_in[(uint64_t)b_in0 * fgs_i + NCi + a2i[v]] : 1.f;
        }
        float *outp = a.out + (uint64_t)b_out * BS;
        float totacc = 0.f;
        if (nred == 1) {
            // ---- fast path (no read starts or ends between the columns): one input block, straight-line code
            const float *blk = a.in + (uint64_t)b_in0 * BS;
            const float *bblk = FWD ? a.beta + (uint64_t)b_out * BS : blk;
            const float *fgi = FWD ? fgo : a.fg_in + (uint64_t)b_in0 * fgs_i;
            load_vec<float, V>(colv, blk + (uint32_t)o_col + ccol);
            totv = blk[(uint32_t)o_tot];
            float kc[V], kx[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                kc[v] = fmaf(cB, colv[v], cC * totv);
                kx[v] = FWD ? cA : cA * gin0[v];
            }
            constexpr int CH = NI < 4 ? NI : 4;        // rows whose loads are issued together
#pragma unroll
            for (int c0 = 0; c0 < NI; c0 += CH) {
                float xin[CH][V], bin[CH][V], rowin[CH], fo_r[CH], fk_r[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const int it = c0 + j;
                    load_vec<float, V>(xin[j], blk + rowoff[it] + ccol);
                    if (FWD) load_vec<float, V>(bin[j], bblk + rowoff[it] + ccol);
                    rowin[j] = blk[(uint32_t)o_row + rowi[it]];
                    fo_r[j] = fgo[a1o[it]];
                    fk_r[j] = FWD ? 1.f : fgi[a1i[it]];
                }
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const int it = c0 + j;
                    const bool live = rok[it] && cok;
                    const float rowterm = cB * rowin[j];
                    float o[V], rs = 0.f;
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const float base = fmaf(kx[v], fk_r[j] * xin[j][v], rowterm + kc[v]);
                        float w;
                        if (FWD) {
                            o[v] = (fo_r[j] * gout[v]) * base;
                            w = o[v];
                            qa[it][v] = fmaf(live ? o[v] : 0.f, bin[j][v], qa[it][v]);
                        } else {
                            o[v] = (a1o[it] != 0 && a2o[v] != 0) ? base : 0.f;
                            w = o[v] * (fo_r[j] * gout[v]);
                        }
                        w = live ? w : 0.f;
                        rs += w;
                        colacc[v] += w;
                    }
                    if (live) store_vec<float, V>(outp + rowoff[it] + ccol, o);
#pragma unroll
                    for (int s = G >> 1; s > 0; s >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, s);
                    if (gl == 0 && rok[it]) outp[(uint32_t)o_row + rowi[it]] = rs;
                    totacc += rs;
                }
            }
        } else {
        if (has_in) {
            uint32_t subm = 0;
            for (uint32_t k = 0; k < nred; ++k) {
                const uint64_t bk = FWD ? (uint64_t)(b_in0 | subm) : (uint64_t)(b_in0 | (k << a.n_sh));
                subm = (subm - a.free_mask) & a.free_mask;
                const float *blk = a.in + bk * BS;
#pragma unroll
                for (int v = 0; v < V; ++v) colv[v] += blk[o_col + ccol + v];
                totv += blk[o_tot];
            }
        }
        const float ctot = cC * totv;
#pragma unroll
        for (int it = 0; it < NI; ++it) {
            if ((uint32_t)it < n_iter) {       // warp-uniform
                const uint32_t R1 = rowi[it];
                float x[V];
#pragma unroll
                for (int v = 0; v < V; ++v) x[v] = 0.f;
                float rowv = 0.f;
                if (has_in) {
                    uint32_t subm = 0;
                    for (uint32_t k = 0; k < nred; ++k) {
                        const uint64_t bk = FWD ? (uint64_t)(b_in0 | subm) : (uint64_t)(b_in0 | (k << a.n_sh));
                        subm = (subm - a.free_mask) & a.free_mask;
                        const float *blk = a.in + bk * BS;
                        rowv += blk[o_row + R1];
                        float ld[V];
                        load_vec<float, V>(ld, blk + (uint64_t)R1 * R + ccol);
                        if (FWD) {
#pragma unroll
                            for (int v = 0; v < V; ++v) x[v] += ld[v];
                        } else {
                            const float *fgk = a.fg_in + bk * fgs_i;
                            const float fk = fgk[a1i[it]];
#pragma unroll
                            for (int v = 0; v < V; ++v) x[v] += ld[v] * (fk * (k == 0 ? gin0[v] : fgk[NCi + a2i[v]]));
                        }
                    }
                }
                const float fo = fgo[a1o[it]];
                const float rowterm = cB * rowv + ctot;
                float o[V], rs = 0.f;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float base = has_in ? fmaf(cA, x[v], fmaf(cB, colv[v], rowterm)) : 1.f;
                    float w;
                    if (FWD) {
                        o[v] = (fo * gout[v]) * base;
                        w = o[v];
                    } else {
                        o[v] = (a1o[it] != 0 && a2o[v] != 0) ? base : 0.f;
                        w = o[v] * (fo * gout[v]);
                    }
                    w = (rok[it] && cok) ? w : 0.f;
                    rs += w;
                    colacc[v] += w;
                }
                if (rok[it] && cok) {
                    store_vec<float, V>(outp + (uint64_t)R1 * R + ccol, o);
                    if (FWD) {
                        float bv[V];
                        load_vec<float, V>(bv, a.beta + (uint64_t)b_out * BS + (uint64_t)R1 * R + ccol);
#pragma unroll
                        for (int v = 0; v < V; ++v) qa[it][v] = fmaf(o[v], bv[v], qa[it][v]);
                    }
                }
#pragma unroll
                for (int s = G >> 1; s > 0; s >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, s);
                if (gl == 0 && rok[it]) outp[o_row + R1] = rs;
                totacc += rs;
            }
        }
        }
        // column sums / totals across the row groups of the warp
#pragma unroll
        for (int s = G; s < 32; s <<= 1) {
#pragma unroll
            for (int v = 0; v < V; ++v) colacc[v] += __shfl_xor_sync(0xffffffffu, colacc[v], s);
            totacc += __shfl_xor_sync(0xffffffffu, totacc, s);
        }
        if (sub == 0 && cok) store_vec<float, V>(outp + o_col + ccol, colacc);
        if (lane == 0) outp[o_tot] = totacc;
        // column scale: forward = sum of the produced alpha; backward = the reduced input total (without missing
        // alleles sum(beta_{c-1}[b']) equals it exactly since tA + 2R tB + R^2 tC = 1; any positive scalar serves)
        warp_sum += FWD ? totacc : totv * inv;
    }

}

template <int V, int G, bool FWD>
__global__ void __launch_bounds__(32 * kWarpKernelWarps)
step_warp_kernel(const StepArgs<float> a) {
    typedef float T;
    constexpr int NI = (V * G * G / 32) > 0 ? (V * G * G / 32) : 1;     // row iterations: ceil(R / (32/G)) for R <= V*G
    extern __shared__ __align__(16) float s_pos[];      // forward: [warps][R*R] per-position posterior accumulators
    __shared__ float s_sum[kWarpKernelWarps];
    __shared__ double s_W[kPartialStride];
    __shared__ double s_L[GG_MAX_ALLELES * (GG_MAX_ALLELES + 1) / 2];
    __shared__ bool s_last;

    const uint32_t R = a.R;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr uint32_t rpi = 32u / G;
    const uint32_t sub = lane / G, gl = lane % G;
    const uint32_t NCo = a.nA_out + 1;
    const uint64_t RR = (uint64_t)R * R;
    const bool cok = gl * V < R;
    const uint32_t ccol = cok ? gl * V : 0u;
    const uint32_t n_iter = (R + rpi - 1) / rpi;
    uint32_t rowi[NI];
    bool rok[NI];
#pragma unroll
    for (int it = 0; it < NI; ++it) {
        const uint32_t r = it * rpi + sub;
        rok[it] = (uint32_t)it < n_iter && r < R;
        rowi[it] = rok[it] ? r : 0u;
    }
    float qa[NI][V];
#pragma unroll
    for (int it = 0; it < NI; ++it)
#pragma unroll
        for (int v = 0; v < V; ++v) qa[it][v] = 0.f;

    float inv = 1.f;
    if (a.in != nullptr) inv = (float)(1.0 / *a.in_scale);
    float warp_sum = 0.f;     // FWD: sum of produced alpha; BWD: sum of produced beta
    __syncthreads();
    warp_step<V, G, FWD>(a, a.al_out, a.al_in, inv, blockIdx.x * kWarpKernelWarps + warp, gridDim.x * kWarpKernelWarps, qa, warp_sum);

    // ---- kernel epilogue: per-position posterior accumulators -> (fixed-order fold over the warps) -> allele-class
    //      bins, one warp per class pair; CTA partials; last CTA.  No atomics: results are bit-reproducible.
    float *mypart = a.partial + (uint64_t)blockIdx.x * kPartialStride;
    if (FWD) {
        float *mine = s_pos + warp * (uint32_t)RR;
        for (uint32_t t = lane; t < (uint32_t)RR; t += 32) mine[t] = 0.f;
        __syncwarp();
#pragma unroll
        for (int it = 0; it < NI; ++it)
            if (rok[it] && cok) store_vec<float, V>(mine + rowi[it] * R + ccol, qa[it]);
    }
    if (lane == 0) s_sum[warp] = warp_sum;
    __syncthreads();
    if (FWD) {
        for (uint32_t t = tid; t < (uint32_t)RR; t += blockDim.x) {
            float q = 0.f;
#pragma unroll
            for (int w = 0; w < kWarpKernelWarps; ++w) q += s_pos[w * (uint32_t)RR + t];
            s_pos[t] = q;      // warp 0's slot now holds the fold (each thread rewrites only what it read)
        }
        __syncthreads();
        for (uint32_t p = warp; p < NCo * NCo; p += kWarpKernelWarps) {
            const uint32_t ci = p / NCo, cj = p % NCo;
            float w = 0.f;
            for (uint32_t t = lane; t < (uint32_t)RR; t += 32) {
                const uint32_t r = t / R, c = t % R;
                if (a.al_out[r] == ci && a.al_out[c] == cj) w += s_pos[t];
            }
            for (uint32_t s = 16; s; s >>= 1) w += __shfl_xor_sync(0xffffffffu, w, s);
            if (lane == 0) mypart[p] = w;
        }
    }
    if (tid == 0) {
        float ts = 0.f;
        for (int w = 0; w < kWarpKernelWarps; ++w) ts += s_sum[w];
        mypart[kMaxClasses * kMaxClasses] = ts;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int ticket = atomicAdd(a.counter, 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const uint32_t npair = FWD ? NCo * NCo : 0u;
    for (uint32_t p = warp; p < npair + 1; p += kWarpKernelWarps) {
        const uint32_t slot = (p == npair) ? (uint32_t)(kMaxClasses * kMaxClasses) : p;
        const double acc = warp_sum_partials(a.partial, gridDim.x, slot, lane);
        if (lane == 0) s_W[slot] = acc;
    }
    __syncthreads();
    if (tid == 0) {
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
