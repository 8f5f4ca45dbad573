// Chain kernel: a run of consecutive columns WITHOUT untagged reads (u_c = 0, one bipartition block per column)
// swept by one persistent CTA.
//
// This is the regime real `giggles genotype` runs live in: without --keep-untagged only haplotagged reads survive
// (reference giggles/cli/__init__.py:167-172), every column is a single R x R block (31 KB at R = 88) and the sweep
// is a strictly sequential chain of tiny steps.  Launching one kernel per column costs ~10 us of launch + prologue
// + last-CTA latency per step; here the column lives in REGISTERS (each thread keeps its <= 32 values of the R x R
// block), the per-column parameters are prefetched one step ahead, the reductions are two block barriers per step,
// and only what later kernels need goes to HBM: every backward column (the forward pass reads them) and the last
// forward column of the run.  Same arithmetic as step_kernel / step_tma_kernel.
#pragma once
#include "hmm_kernels.cuh"

namespace gg {

constexpr int kChainWarps = 16;
constexpr int kChainThreads = 32 * kChainWarps;
constexpr int kChainChunk = 32;    // step descriptors staged in shared memory per chunk
constexpr int kChainMaxNIT = 8;   // row iterations per warp (template parameter NIT = ceil(R / 16)): R <= 128

struct ChainStep {
    uint64_t out_off;      // produced column record (backward: every step; forward: used for the last step only)
    uint64_t beta_off;     // forward: beta record of this column
    uint64_t fg_out_off;   // F/G table of the produced column
    uint64_t fg_in_off;    // backward: F/G table of the input column
    uint64_t gl_off;       // forward: offset of this column's genotype table
    uint32_t c_out, c_in;  // produced / input column (c_in == 0xffffffff: none)
    uint32_t nA_out, nA_in;
    float tA, tB, tC, pad;
    uint64_t pad2;         // 80 bytes: descriptors are staged with 16-byte copies
};
static_assert(sizeof(ChainStep) % 16 == 0, "ChainStep must be a multiple of 16 bytes");

struct ChainArgs {
    unsigned char *arena;
    const ChainStep *steps;
    uint32_t n_steps;
    const float *in;          // input column record of the first step (nullptr: the chain starts here)
    const double *in_scale;   // its column sum
    double *scale;            // per-column sums (alpha or beta array), indexed by column
    double *result;
    const uint8_t *allele1;   // [C][R] panel alleles + 1
    uint32_t R, BS, max_classes;
    uint32_t no_fast_posterior;   // A/B switch: always use the shared-memory posterior path
};

struct ChainParams {          // per-step parameters staged in shared memory one step ahead
    float F_out[kMaxClasses + 3], G_out[kMaxClasses + 3], F_in[kMaxClasses + 3], G_in[kMaxClasses + 3];
    uint8_t al_out[128], al_in[128];
};

__host__ __device__ inline uint32_t chain_smem_bytes(uint32_t R, uint32_t ncls, bool fwd) {
    const uint32_t Rp = (R + 3u) & ~3u;
    uint32_t o = 2 * (uint32_t)sizeof(ChainParams);
    o += 2 * kChainChunk * (uint32_t)sizeof(ChainStep);   // step descriptors, two chunks
    o += kChainWarps * Rp * 4;          // column partials per warp
    o += 2 * Rp * 4;                    // folded column sums (double-buffered by step parity)
    o += 2 * kChainWarps * 4 + 16;      // total partials, folded totals
    if (fwd) {
        o += kChainWarps * ncls * Rp * 4;   // posterior partials per warp and row class
        o += ncls * Rp * 4;                 // folded
        o += kMaxClasses * kMaxClasses * 4; // W[i][j]
        o += kChainWarps * 4 * 4;           // bi-allelic fast path: per-warp W partials
    }
    return (o + 127u) & ~127u;
}

template <bool FWD, int NIT>
__global__ void __launch_bounds__(kChainThreads, 1)
chain_kernel(const ChainArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t R = a.R, RR = R * R, Rp = (R + 3u) & ~3u;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t NCmax = a.max_classes;
    ChainParams *s_par = reinterpret_cast<ChainParams *>(smem_raw);
    ChainStep *s_steps = reinterpret_cast<ChainStep *>(smem_raw + 2 * sizeof(ChainParams));
    float *s_colp = reinterpret_cast<float *>(smem_raw + 2 * sizeof(ChainParams) + 2 * kChainChunk * sizeof(ChainStep));
    float *s_col = s_colp + kChainWarps * Rp;
    float *s_totp = s_col + 2 * Rp;
    float *s_tot = s_totp + 2 * kChainWarps;            // [4]
    float *s_Q = s_tot + 4;
    float *s_Qf = s_Q + (FWD ? kChainWarps * NCmax * Rp : 0);
    float *s_W = s_Qf + (FWD ? NCmax * Rp : 0);
    float *s_Wp = s_W + kMaxClasses * kMaxClasses;

    const bool cok = lane * 4 < R;
    const uint32_t ccol = cok ? lane * 4 : 0u;
    uint32_t rows[NIT];
    bool rok[NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
        rows[i] = warp + kChainWarps * i;
        rok[i] = rows[i] < R;
        if (!rok[i]) rows[i] = 0;
    }

    // step descriptors: chunk q (steps q*32 ..) lives in s_steps[q & 1]; a chunk is fetched a whole chunk ahead
    auto load_chunk = [&](uint32_t q) {
        const uint32_t first = q * kChainChunk;
        if (first >= a.n_steps) return;
        const uint32_t n16 = min((uint32_t)kChainChunk, a.n_steps - first) * (uint32_t)(sizeof(ChainStep) / 16);
        if (tid < n16)
            reinterpret_cast<uint4 *>(s_steps + (q & 1u) * kChainChunk)[tid] = reinterpret_cast<const uint4 *>(a.steps + first)[tid];
    };
    auto step_at = [&](uint32_t k) -> const ChainStep & { return s_steps[((k / kChainChunk) & 1u) * kChainChunk + (k % kChainChunk)]; };
    // per-step parameters (F/G vectors, panel alleles): loaded into registers TWO steps ahead (issue), stored to
    // shared memory one step ahead (commit) - their global-load latency never sits on the step's critical path
    float pf_f = 0.f, pf_g = 0.f;
    uint32_t pf_al = 0;
    auto issue_params = [&](uint32_t k) {
        if (k >= a.n_steps) return;
        const ChainStep &st = step_at(k);
        const uint32_t NCo = st.nA_out + 1, NCi = st.nA_in + 1;
        if (tid < NCo) {
            const float *fgo = reinterpret_cast<const float *>(a.arena + st.fg_out_off);
            pf_f = fgo[tid]; pf_g = fgo[NCo + tid];
        } else if (!FWD && st.c_in != 0xffffffffu && tid >= 32 && tid - 32 < NCi) {
            const float *fgi = reinterpret_cast<const float *>(a.arena + st.fg_in_off);
            pf_f = fgi[tid - 32]; pf_g = fgi[NCi + tid - 32];
        } else if (tid >= 64 && tid - 64 < R) {
            pf_al = a.allele1[(uint64_t)st.c_out * R + (tid - 64)];
        } else if (!FWD && st.c_in != 0xffffffffu && tid >= 256 && tid - 256 < R) {
            pf_al = a.allele1[(uint64_t)st.c_in * R + (tid - 256)];
        }
    };
    auto commit_params = [&](uint32_t k) {
        if (k >= a.n_steps) return;
        const ChainStep &st = step_at(k);
        ChainParams *p = &s_par[k & 1u];
        const uint32_t NCo = st.nA_out + 1, NCi = st.nA_in + 1;
        if (tid < NCo) { p->F_out[tid] = pf_f; p->G_out[tid] = pf_g; }
        else if (!FWD && st.c_in != 0xffffffffu && tid >= 32 && tid - 32 < NCi) { p->F_in[tid - 32] = pf_f; p->G_in[tid - 32] = pf_g; }
        else if (tid >= 64 && tid - 64 < R) p->al_out[tid - 64] = (uint8_t)pf_al;
        else if (!FWD && st.c_in != 0xffffffffu && tid >= 256 && tid - 256 < R) p->al_in[tid - 256] = (uint8_t)pf_al;
    };
    load_chunk(0);
    load_chunk(1);
    __syncthreads();

    // ---- state of the input column: values in registers, its side sums in registers / shared memory
    float v[NIT][4];       // forward: alpha_{c-1};  backward: beta_c
    float rowM[NIT];       // forward: row sums of alpha_{c-1}
    float colM[4] = {0.f, 0.f, 0.f, 0.f};
    float totM = 0.f, inv = 1.f;
    const bool has_in = a.in != nullptr;
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
        rowM[i] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) v[i][q] = 0.f;
    }
    if (has_in) {
        inv = (float)(1.0 / *a.in_scale);
#pragma unroll
        for (int i = 0; i < NIT; ++i)
            if (rok[i] && cok) {
                const float4 x = *reinterpret_cast<const float4 *>(a.in + rows[i] * R + ccol);
                v[i][0] = x.x; v[i][1] = x.y; v[i][2] = x.z; v[i][3] = x.w;
            }
        if (FWD) {
#pragma unroll
            for (int i = 0; i < NIT; ++i) rowM[i] = rok[i] ? a.in[RR + rows[i]] : 0.f;
            if (cok) {
                const float4 x = *reinterpret_cast<const float4 *>(a.in + RR + R + ccol);
                colM[0] = x.x; colM[1] = x.y; colM[2] = x.z; colM[3] = x.w;
            }
            totM = a.in[RR + 2 * R];
        }
    }
    issue_params(0);
    commit_params(0);
    issue_params(1);
    float bnext[NIT][4];   // forward: beta of the column being produced, prefetched one step ahead
    auto prefetch_beta = [&](uint32_t k) {
        if (!FWD || k >= a.n_steps) return;
        const float *b = reinterpret_cast<const float *>(a.arena + step_at(k).beta_off);
#pragma unroll
        for (int i = 0; i < NIT; ++i)
            if (rok[i] && cok) {
                const float4 x = *reinterpret_cast<const float4 *>(b + rows[i] * R + ccol);
                bnext[i][0] = x.x; bnext[i][1] = x.y; bnext[i][2] = x.z; bnext[i][3] = x.w;
            }
    };
#pragma unroll
    for (int i = 0; i < NIT; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) bnext[i][q] = 0.f;
    prefetch_beta(0);
    __syncthreads();

    uint64_t prev_out_off = 0;
    for (uint32_t k = 0; k < a.n_steps; ++k) {
        const ChainStep st = step_at(k);
        const ChainParams *p = &s_par[k & 1u];
        const uint32_t par = k & 1u;
        const uint32_t NCo = st.nA_out + 1;
        commit_params(k + 1);
        issue_params(k + 2);
        if (k % kChainChunk == 0 && k > 0) load_chunk(k / kChainChunk + 1);   // the chunk after next; its buffer was last read in step k-1
        float *outp = reinterpret_cast<float *>(a.arena + st.out_off);
        const bool have_prev = FWD ? (k > 0 || has_in) : true;
        const bool fast = FWD && NCo == 3 && !a.no_fast_posterior;
        uint32_t a2o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) a2o[q] = cok ? p->al_out[ccol + q] : 0u;

        float cpart[4] = {0.f, 0.f, 0.f, 0.f}, tpart = 0.f;
        float rsum[NIT];
        if (FWD) {
            // alpha_c = E_c * T(alpha_{c-1}); posterior with beta_c
            const float cA = st.tA * inv, cB = st.tB * inv, cC = st.tC * inv;
            float g[4], cb[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                g[q] = p->G_out[a2o[q]];
                cb[q] = cB * colM[q] + cC * totM;
            }
            // posterior partials.  Bi-allelic columns (the common case): two register accumulators per lane and row
            // class, binned by column class with four warp reductions.  Otherwise: per-warp shared-memory partials.
            float acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
            if (!fast) {
                for (uint32_t z = lane; z < NCo * (Rp >> 2); z += 32)
                    reinterpret_cast<float4 *>(s_Q + warp * NCmax * Rp)[z] = make_float4(0.f, 0.f, 0.f, 0.f);
                __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const uint32_t a1 = p->al_out[rows[i]];
                const float f = rok[i] ? p->F_out[a1] : 0.f;
                const float rt = cB * rowM[i];
                float rs = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float base = have_prev ? fmaf(cA, v[i][q], rt + cb[q]) : 1.f;
                    v[i][q] = (f * g[q]) * base;          // alpha_c replaces alpha_{c-1} in place
                    rs += v[i][q];
                    cpart[q] += v[i][q];
                }
                for (int sh = 16; sh > 0; sh >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, sh);
                rsum[i] = rs;
                tpart += rs;
                if (fast) {
                    if (a1 == 1) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc1[q] = fmaf(v[i][q], bnext[i][q], acc1[q]);
                    } else {           // class 2, or a missing allele / idle row slot whose alpha is 0
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc2[q] = fmaf(v[i][q], bnext[i][q], acc2[q]);
                    }
                } else if (rok[i] && cok) {
                    float4 *qd = reinterpret_cast<float4 *>(s_Q + (warp * NCmax + a1) * Rp + ccol);
                    float4 acc = *qd;
                    acc.x = fmaf(v[i][0], bnext[i][0], acc.x); acc.y = fmaf(v[i][1], bnext[i][1], acc.y);
                    acc.z = fmaf(v[i][2], bnext[i][2], acc.z); acc.w = fmaf(v[i][3], bnext[i][3], acc.w);
                    *qd = acc;
                }
            }
            if (fast) {
                float w4[4] = {0.f, 0.f, 0.f, 0.f};       // (row class, column class) = (1,1) (1,2) (2,1) (2,2)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    w4[0] += a2o[q] == 1 ? acc1[q] : 0.f; w4[1] += a2o[q] == 2 ? acc1[q] : 0.f;
                    w4[2] += a2o[q] == 1 ? acc2[q] : 0.f; w4[3] += a2o[q] == 2 ? acc2[q] : 0.f;
                }
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                    for (int sh = 16; sh > 0; sh >>= 1) w4[z] += __shfl_xor_sync(0xffffffffu, w4[z], sh);
                    if (lane == 0) s_Wp[warp * 4 + z] = w4[z];
                }
            }
            prefetch_beta(k + 1);      // beta of the next column: in flight across the barriers below
        } else {
            // X = beta_c * E_c and its sums; beta_{c-1} = T(X) comes after the fold
            float gi[4];
            const bool has_col_in = st.c_in != 0xffffffffu;
#pragma unroll
            for (int q = 0; q < 4; ++q) gi[q] = (has_col_in && cok) ? p->G_in[p->al_in[ccol + q]] : 0.f;
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const float f = (has_col_in && rok[i]) ? p->F_in[p->al_in[rows[i]]] : 0.f;
                float rs = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    v[i][q] = v[i][q] * (f * gi[q]);      // X = beta_c * E_c, in place
                    rs += v[i][q];
                    cpart[q] += v[i][q];
                }
                for (int sh = 16; sh > 0; sh >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, sh);
                rsum[i] = rs;
                tpart += rs;
            }
        }
        if (cok) *reinterpret_cast<float4 *>(s_colp + warp * Rp + ccol) = make_float4(cpart[0], cpart[1], cpart[2], cpart[3]);
        if (lane == 0) s_totp[warp] = tpart;
        __syncthreads();                                   // B1
        // ---- folds (fixed order: deterministic)
        for (uint32_t t = tid; t < R; t += kChainThreads) {
            float cs = 0.f;
#pragma unroll
            for (int w = 0; w < kChainWarps; ++w) cs += s_colp[w * Rp + t];
            s_col[par * Rp + t] = cs;
        }
        if (tid == kChainThreads - 1) {
            float ts = 0.f;
#pragma unroll
            for (int w = 0; w < kChainWarps; ++w) ts += s_totp[w];
            s_tot[par] = ts;
        }
        if (FWD && fast) {
            if (tid < 4) {
                float q = 0.f;
#pragma unroll
                for (int w = 0; w < kChainWarps; ++w) q += s_Wp[w * 4 + tid];
                s_W[(1 + (tid >> 1)) * 3 + 1 + (tid & 1)] = q;
            }
        } else if (FWD) {
            for (uint32_t idx = tid; idx < NCo * R; idx += kChainThreads) {
                const uint32_t cls = idx / R, t = idx % R;
                float q = 0.f;
#pragma unroll
                for (int w = 0; w < kChainWarps; ++w) q += s_Q[(w * NCmax + cls) * Rp + t];
                s_Qf[cls * Rp + t] = q;
            }
        }
        __syncthreads();                                   // B2
        const float tot = s_tot[par];
        float colN[4] = {0.f, 0.f, 0.f, 0.f};
        if (cok) {
            const float4 x = *reinterpret_cast<const float4 *>(s_col + par * Rp + ccol);
            colN[0] = x.x; colN[1] = x.y; colN[2] = x.z; colN[3] = x.w;
        }
        if (FWD) {
            // the produced column becomes the next step's input; only the last one of the run goes to HBM
#pragma unroll
            for (int i = 0; i < NIT; ++i) rowM[i] = rsum[i];
#pragma unroll
            for (int q = 0; q < 4; ++q) colM[q] = colN[q];
            totM = tot;
            inv = 1.f / tot;
            if (tid == 0) a.scale[st.c_out] = (double)tot;
            if (k + 1 == a.n_steps) {
#pragma unroll
                for (int i = 0; i < NIT; ++i)
                    if (rok[i] && cok) {
                        *reinterpret_cast<float4 *>(outp + rows[i] * R + ccol) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
                        if (lane == 0) outp[RR + rows[i]] = rsum[i];
                    }
                if (warp == 0 && cok) *reinterpret_cast<float4 *>(outp + RR + R + ccol) = make_float4(colN[0], colN[1], colN[2], colN[3]);
                if (tid == 0) outp[RR + 2 * R] = tot;
            }
            // W[i][j] = sum over columns of class j of the folded posterior partials; one warp per pair
            if (!fast) {
            for (uint32_t pr = warp; pr < NCo * NCo; pr += kChainWarps) {
                const uint32_t ci = pr / NCo, cj = pr % NCo;
                float w = 0.f;
                for (uint32_t t = lane; t < R; t += 32)
                    if (p->al_out[t] == cj) w += s_Qf[ci * Rp + t];
                for (int sh = 16; sh > 0; sh >>= 1) w += __shfl_xor_sync(0xffffffffu, w, sh);
                if (lane == 0) s_W[pr] = w;
            }
            __syncthreads();                               // B3
            }
            if (warp == 0) {
                const uint32_t nA = st.nA_out, ng = nA * (nA + 1) / 2;
                double norm = 0.0;
                for (uint32_t i2 = 1; i2 <= nA; ++i2)
                    for (uint32_t j2 = 1; j2 <= nA; ++j2) norm += (double)s_W[i2 * NCo + j2];
                for (uint32_t gidx = lane; gidx < ng; gidx += 32) {
                    // genotype gidx collects the ordered pairs (i, j) with i + j(j+1)/2 == gidx  (src/genotypehmm.cpp:524-528)
                    double acc = 0.0;
                    for (uint32_t j2 = 0; j2 < nA; ++j2) {
                        const uint32_t base = j2 * (j2 + 1) / 2;
                        if (gidx >= base && gidx - base < nA) acc += (double)s_W[(gidx - base + 1) * NCo + (j2 + 1)];
                    }
                    a.result[st.gl_off + gidx] = acc / norm;
                }
            }
        } else {
            // side sums of the INPUT column's record are what was just folded (row/col/total of beta_c * E_c): the
            // first column of the run already has them; the columns this kernel produced get them now
            if (k > 0) {
                float *inrec = reinterpret_cast<float *>(a.arena + prev_out_off);
#pragma unroll
                for (int i = 0; i < NIT; ++i)
                    if (rok[i] && lane == 0) inrec[RR + rows[i]] = rsum[i];
                if (warp == 0 && cok) *reinterpret_cast<float4 *>(inrec + RR + R + ccol) = make_float4(colN[0], colN[1], colN[2], colN[3]);
                if (tid == 0) inrec[RR + 2 * R] = tot;
            }
            // beta_{c-1} = T(X) / tot  (T's coefficients sum to one: the produced column sums to ~1)
            const bool has_col_in = st.c_in != 0xffffffffu;
            const float s = has_col_in ? 1.f / tot : 1.f;
            const float cA = st.tA * s, cB = st.tB * s, cC = st.tC * s;
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const float rt = cB * rsum[i] + cC * tot;
#pragma unroll
                for (int q = 0; q < 4; ++q) v[i][q] = has_col_in ? fmaf(cA, v[i][q], fmaf(cB, colN[q], rt)) : 1.f;
                if (rok[i] && cok) *reinterpret_cast<float4 *>(outp + rows[i] * R + ccol) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
            }
            if (tid == 0) a.scale[st.c_out] = has_col_in ? 1.0 : (double)RR;
        }
        prev_out_off = st.out_off;
    }
    if (!FWD && a.n_steps > 0) {
        // side sums of the last produced column (row/col/total of beta * E of that column): one more fold
        const ChainStep st = step_at(a.n_steps - 1);
        __syncthreads();
        ChainParams *p = &s_par[0];
        const uint32_t NCo = st.nA_out + 1;
        const float *fgo = reinterpret_cast<const float *>(a.arena + st.fg_out_off);
        if (tid < NCo) { p->F_out[tid] = fgo[tid]; p->G_out[tid] = fgo[NCo + tid]; }
        if (tid >= 64 && tid - 64 < R) p->al_out[tid - 64] = a.allele1[(uint64_t)st.c_out * R + (tid - 64)];
        __syncthreads();
        float *outp = reinterpret_cast<float *>(a.arena + st.out_off);
        float g[4], cpart[4] = {0.f, 0.f, 0.f, 0.f}, tpart = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) g[q] = cok ? p->G_out[p->al_out[ccol + q]] : 0.f;
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const float f = rok[i] ? p->F_out[p->al_out[rows[i]]] : 0.f;
            float rs = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float x = v[i][q] * (f * g[q]);
                rs += x;
                cpart[q] += x;
            }
            for (int sh = 16; sh > 0; sh >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, sh);
            if (rok[i] && lane == 0) outp[RR + rows[i]] = rs;
            tpart += rs;
        }
        if (cok) *reinterpret_cast<float4 *>(s_colp + warp * Rp + ccol) = make_float4(cpart[0], cpart[1], cpart[2], cpart[3]);
        if (lane == 0) s_totp[warp] = tpart;
        __syncthreads();
        for (uint32_t t = tid; t < R; t += kChainThreads) {
            float cs = 0.f;
#pragma unroll
            for (int w = 0; w < kChainWarps; ++w) cs += s_colp[w * Rp + t];
            outp[RR + R + t] = cs;
        }
        if (tid == kChainThreads - 1) {
            float ts = 0.f;
#pragma unroll
            for (int w = 0; w < kChainWarps; ++w) ts += s_totp[w];
            outp[RR + 2 * R] = ts;
        }
    }
}

}  // namespace gg
