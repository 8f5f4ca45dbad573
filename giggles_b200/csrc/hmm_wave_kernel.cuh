// Small panels (R <= 32), many columns per launch: the wave kernel.
//
// A column of configs[0] (10 haplotypes, <= 2^15 blocks of 100 values) is a few MB that sit in L2, and one column step
// is a few microseconds of work: launched column by column, a sweep is bound by launch latency, the per-launch
// prologue and the last-CTA epilogue (10,000 launches for 5,000 columns).  Here ONE cooperative launch walks a whole
// run of consecutive steps of one direction: every warp does its share of a step's output blocks with the same code
// as the per-column kernel (warp_step, hmm_warp_kernel.cuh), the CTAs meet at a grid barrier (one atomic arrival
// counter, acquire / release), and the next step starts.  What a launch boundary used to provide is done in place:
//
//   column scale   every CTA arrives at the barrier with its share of the produced column's sum in fixed point, in the
//                  same 64-bit atomic add that counts the arrival; integer addition is order-independent, so all CTAs
//                  read the same total and fold the same 1/sum into the next step (CTA 0 also stores it where later
//                  launches expect it);
//   posterior      forward steps bin their per-position accumulators by allele-class pair in registers (the class of a
//                  (row, column) position is the same for every block of a column), one partial vector per CTA and
//                  step; wave_posterior_kernel adds them up in fixed order after the run.
//
// No atomics on data, fixed summation orders: results are bit-reproducible (not bit-identical to the per-column
// kernels: the partial sums are grouped differently).
#pragma once
#include "hmm_warp_kernel.cuh"

#ifndef GG_WAVE_MIN_CTAS
#define GG_WAVE_MIN_CTAS 1
#endif

namespace gg {

struct WaveStep {
    StepArgs<float> a;
    uint64_t w_off;       // forward: offset (floats) of this step's [grid][nA_out^2] posterior partials
    uint64_t pad;
    uint8_t al_out[32], al_in[32];   // allele + 1 of the produced / input column (R <= 32), so that one record is all a step needs
};
static_assert(sizeof(WaveStep) % 16 == 0, "WaveStep is staged with 32-bit copies and kept 16-byte aligned");

struct WaveArgs {
    const WaveStep *steps;
    uint32_t n_steps;
    float *w_partial;             // forward: posterior partials of every step
    unsigned long long *barrier;  // [3] arrivals << 52 | fixed-point column sum, zero at launch; step s uses slot s % 3
};

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Grid barrier of the (co-resident, cooperatively launched) grid that also adds up the column sum: every CTA arrives
// with ONE 64-bit atomic add of (1 << 52 | its share of the sum in 2^-32 fixed point).  Integer addition does not depend
// on the order of arrival, so every CTA reads the same total whatever the timing - the scale of the next step costs no
// memory round trip of its own and stays bit-reproducible.  (The sum only has to be a consistent positive scale; 2^-32
// resolution against column sums of order one is ample.)  Thread 0 polls relaxed and fences once: an acquire load per
// poll would drop the SM's L1 contents on every iteration.  Three slots in rotation; CTA 0 clears the slot of the step
// before, which every CTA has left by then and nobody reaches again before CTA 0's next arrival.  Returns the total.
__device__ __forceinline__ double wave_barrier_sum(unsigned long long *bar, const uint32_t step, const float cta_sum, double *s_total) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const double clamped = fmin(fmax((double)cta_sum, 0.0), 1024.0);
        const unsigned long long mine = (1ull << 52) | (unsigned long long)__double2ll_rn(clamped * 4294967296.0);
        unsigned long long *slot = bar + step % 3u;
        __threadfence();
        atomicAdd(slot, mine);
        unsigned long long v;
        do { v = ld_relaxed_u64(slot); } while ((v >> 52) < gridDim.x);
        __threadfence();
        if (blockIdx.x == 0) bar[(step + 2u) % 3u] = 0ull;
        *s_total = (double)(v & ((1ull << 52) - 1ull)) * (1.0 / 4294967296.0);
    }
    __syncthreads();
    return *s_total;
}

constexpr int kLaneMaxAlleles = 4;     // forward lane path: posterior bins per thread are kLaneMaxAlleles^2 registers

// Thread-per-block column step for very small panels (R even, <= 16; configs[0] has R = 10).  With one warp per block
// (warp_step) a 100-value block costs ~500 warp instructions, most of them index arithmetic, predicates and shuffles
// for 39 % busy lanes.  Here ONE THREAD owns a whole bipartition block: row sums, column sums and the block total are
// plain register arithmetic (no shuffles), two rows at a time move as R/2 128-bit accesses, and a warp advances 32
// blocks at once.  Neighbouring lanes work on neighbouring blocks (496 bytes apart for R = 10), so a warp's accesses
// walk 32 lines that L1 holds across the row loop.  Same arithmetic as warp_step; summation orders differ.
//   W: this thread's posterior sums per (row allele, column allele), alleles < kLaneMaxAlleles (forward only);
//   sum_acc: this thread's share of the produced column's sum.
template <int R, bool FWD>
__device__ __forceinline__ void lane_step(const StepArgs<float> &a, const uint8_t *al_o, const uint8_t *al_i, const float inv,
                                          const uint32_t first_item, const uint32_t item_stride,
                                          float (&W)[kLaneMaxAlleles][kLaneMaxAlleles], float &sum_acc) {
    static_assert(R % 2 == 0 && R >= 2 && R <= 16, "lane_step: even panels up to 16 haplotypes");
    constexpr uint32_t RR = R * R, o_row = RR, o_col = RR + R, o_tot = RR + 2 * R;
    constexpr int NV = R / 2;                      // float4 per pair of rows
    const uint32_t BS = a.BS;
    const uint32_t NCo = a.nA_out + 1, NCi = a.nA_in + 1, fgs_o = 2 * NCo, fgs_i = 2 * NCi, nA = a.nA_out;
    const uint32_t nred = FWD ? (1u << a.n_free_left) : (1u << (a.u_in - a.n_sh));
    const uint32_t n_bcast = FWD ? (a.u_out - a.n_sh) : a.n_free_left;
    const bool sh_id = a.shared_mask == ((a.n_sh >= 32 ? 0u : (1u << a.n_sh)) - 1u);
    const bool fr_id = a.free_mask == ((((a.n_free_left >= 32 ? 0u : (1u << a.n_free_left)) - 1u)) << a.n_sh);
    const float cA = (float)a.tA * inv, cB = (float)a.tB * inv, cC = (float)a.tC * inv;
    const uint32_t n_items = 1u << a.u_out;
    uint32_t co[R], ci[R];                         // allele classes of the columns (the same in every thread)
#pragma unroll
    for (int j = 0; j < R; ++j) {
        co[j] = al_o[j];
        ci[j] = al_i[j];
    }
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
        const float *fgo = a.fg_out + (uint64_t)b_out * fgs_o;
        float g[R], cb[R], colacc[R], gk0[R];
        float totv = 0.f;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            g[j] = fgo[NCo + co[j]];
            cb[j] = 0.f;
            colacc[j] = 0.f;
            gk0[j] = 1.f;
        }
        {   // column sums and total of the reduced input
            uint32_t subm = 0;
            for (uint32_t k = 0; k < nred; ++k) {
                const uint64_t bk = FWD ? (uint64_t)(b_in0 | subm) : (uint64_t)(b_in0 | (k << a.n_sh));
                subm = (subm - a.free_mask) & a.free_mask;
                const float *blk = a.in + bk * BS;
#pragma unroll
                for (int j = 0; j < R; j += 2) {
                    const float2 t = *reinterpret_cast<const float2 *>(blk + o_col + j);
                    cb[j] += t.x;
                    cb[j + 1] += t.y;
                }
                totv += blk[o_tot];
            }
        }
        const float ctot = cC * totv;
#pragma unroll
        for (int j = 0; j < R; ++j) cb[j] = fmaf(cB, cb[j], ctot);
        if (!FWD && nred == 1) {
            const float *fgi = a.fg_in + (uint64_t)b_in0 * fgs_i;
#pragma unroll
            for (int j = 0; j < R; ++j) gk0[j] = fgi[NCi + ci[j]];
        }
        float *outp = a.out + (uint64_t)b_out * BS;
        const float *bblk = FWD ? a.beta + (uint64_t)b_out * BS : nullptr;
        if (nred == 1) {
            // every line of the input block (and of beta, which comes from HBM) on its way to L1 before the row loop
            // walks them: one exposed round trip per block instead of one per pair of rows
            constexpr int kLines = (int)((RR + 2 * R + 4 + 3) / 4 * 16 + 127) / 128 + 1;
            const char *pi = reinterpret_cast<const char *>(a.in + (uint64_t)b_in0 * BS);
#pragma unroll
            for (int l = 0; l < kLines; ++l) prefetch_l1(pi + l * 128);
            if (FWD) {
#pragma unroll
                for (int l = 0; l < kLines; ++l) prefetch_l1(reinterpret_cast<const char *>(bblk) + l * 128);
            }
        }
        float tot = 0.f;
#pragma unroll
        for (uint32_t r = 0; r < R; r += 2) {
            float x[2 * R], rowv0 = 0.f, rowv1 = 0.f;
#pragma unroll
            for (int e = 0; e < 2 * R; ++e) x[e] = 0.f;
            const uint32_t ai0 = al_i[r], ai1 = al_i[r + 1];
            uint32_t subm = 0;
            for (uint32_t k = 0; k < nred; ++k) {
                const uint64_t bk = FWD ? (uint64_t)(b_in0 | subm) : (uint64_t)(b_in0 | (k << a.n_sh));
                subm = (subm - a.free_mask) & a.free_mask;
                const float *blk = a.in + bk * BS;
                const float2 rv = *reinterpret_cast<const float2 *>(blk + o_row + r);
                rowv0 += rv.x;
                rowv1 += rv.y;
                float fk0 = 1.f, fk1 = 1.f;
                const float *fgk = nullptr;
                if (!FWD) {
                    fgk = a.fg_in + bk * fgs_i;
                    fk0 = fgk[ai0];
                    fk1 = fgk[ai1];
                }
                const float4 *src = reinterpret_cast<const float4 *>(blk + r * R);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const float4 t = src[v];
                    const float ld[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        constexpr int dummy = 0; (void)dummy;
                        const int e = 4 * v + q, j = e % R;
                        if (FWD) {
                            x[e] += ld[q];
                        } else {
                            const float gk = nred == 1 ? gk0[j] : fgk[NCi + ci[j]];
                            x[e] = fmaf(ld[q], (e < R ? fk0 : fk1) * gk, x[e]);
                        }
                    }
                }
            }
            float xo[2 * R], rs[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint32_t ao = al_o[r + q];
                const float fo = fgo[ao];
                const float rowterm = cB * (q == 0 ? rowv0 : rowv1);
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const float base = fmaf(cA, x[q * R + j], rowterm + cb[j]);
                    float o, w;
                    if (FWD) {
                        o = (fo * g[j]) * base;
                        w = o;
                    } else {
                        o = (ao != 0 && co[j] != 0) ? base : 0.f;
                        w = o * (fo * g[j]);
                    }
                    acc += w;
                    colacc[j] += w;
                    xo[q * R + j] = o;
                }
                rs[q] = acc;
                tot += acc;
            }
            float4 *dst = reinterpret_cast<float4 *>(outp + r * R);
#pragma unroll
            for (int v = 0; v < NV; ++v) dst[v] = make_float4(xo[4 * v], xo[4 * v + 1], xo[4 * v + 2], xo[4 * v + 3]);
            *reinterpret_cast<float2 *>(outp + o_row + r) = make_float2(rs[0], rs[1]);
            if (FWD) {
                const float4 *bsrc = reinterpret_cast<const float4 *>(bblk + r * R);
                float rowc[2][kLaneMaxAlleles];
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int c = 0; c < kLaneMaxAlleles; ++c) rowc[q][c] = 0.f;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const float4 t = bsrc[v];
                    const float bv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int e = 4 * v + q, j = e % R;
                        const float fb = xo[e] * bv[q];
#pragma unroll
                        for (int c = 0; c < kLaneMaxAlleles; ++c)
                            if ((uint32_t)c < nA) rowc[e < R ? 0 : 1][c] += (co[j] == (uint32_t)c + 1) ? fb : 0.f;
                    }
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint32_t ao = al_o[r + q];
#pragma unroll
                    for (int i = 0; i < kLaneMaxAlleles; ++i)
#pragma unroll
                        for (int c = 0; c < kLaneMaxAlleles; ++c)
                            if ((uint32_t)i < nA && (uint32_t)c < nA) W[i][c] += (ao == (uint32_t)i + 1) ? rowc[q][c] : 0.f;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < R; j += 2) *reinterpret_cast<float2 *>(outp + o_col + j) = make_float2(colacc[j], colacc[j + 1]);
        outp[o_tot] = tot;
        sum_acc += FWD ? tot : totv * inv;
    }
}

// RL: the panel size when the thread-per-block path (lane_step) applies to it, else 0
template <int V, int G, bool FWD, int RL>
__global__ void __launch_bounds__(32 * kWarpKernelWarps, GG_WAVE_MIN_CTAS)
wave_kernel(const WaveArgs w) {
    constexpr int NI = (V * G * G / 32) > 0 ? (V * G * G / 32) : 1;
    __shared__ float s_sum[kWarpKernelWarps];
    __shared__ float s_w[FWD ? kWarpKernelWarps : 1][FWD ? GG_MAX_ALLELES * GG_MAX_ALLELES : 1];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr uint32_t rpi = 32u / G;
    const uint32_t sub = lane / G, gl = lane % G;
    const uint32_t first_item = blockIdx.x * kWarpKernelWarps + warp, item_stride = gridDim.x * kWarpKernelWarps;

    // the step record (arguments and the two allele vectors) is staged in shared memory one step ahead: the barrier's
    // acquire drops the L1 contents, so without this every step would begin with dependent L2 round trips
    __shared__ __align__(16) WaveStep s_step[2];
    __shared__ double s_total;
    auto stage_args = [&](uint32_t s, uint32_t slot) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(w.steps + s);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_step[slot]);
        if (tid < sizeof(WaveStep) / 4) dst[tid] = __ldcg(src + tid);
    };
    stage_args(0, 0);
    __syncthreads();

    float inv = 1.f;
    for (uint32_t s = 0; s < w.n_steps; ++s) {
        const uint32_t slot = s & 1u;
        const StepArgs<float> &a = s_step[slot].a;
        const uint8_t *al_out = s_step[slot].al_out, *al_in = s_step[slot].al_in;
        if (s + 1 < w.n_steps) stage_args(s + 1, slot ^ 1u);      // visible after the __syncthreads before the barrier
        if (s == 0) inv = (float)(1.0 / *a.in_scale);
        float warp_sum = 0.f;
        const bool lane_path = RL > 0 && (!FWD || a.nA_out <= (uint32_t)kLaneMaxAlleles);
        if (lane_path) {
            // thread-per-block: warp tasks of 32 consecutive blocks, dealt round-robin over the CTAs first so that a
            // column of a few thousand blocks still spreads over all SMs
            constexpr int RLL = RL > 0 ? RL : 2;
            float W[kLaneMaxAlleles][kLaneMaxAlleles];
#pragma unroll
            for (int i = 0; i < kLaneMaxAlleles; ++i)
#pragma unroll
                for (int j = 0; j < kLaneMaxAlleles; ++j) W[i][j] = 0.f;
            float sum_acc = 0.f;
            lane_step<RLL, FWD>(a, al_out, al_in, inv, (blockIdx.x + gridDim.x * warp) * 32u + lane, gridDim.x * kWarpKernelWarps * 32u, W, sum_acc);
#pragma unroll
            for (int o = 16; o; o >>= 1) sum_acc += __shfl_xor_sync(0xffffffffu, sum_acc, o);
            warp_sum = sum_acc;
            if (FWD) {
                const uint32_t nA = a.nA_out;
#pragma unroll
                for (int i = 0; i < kLaneMaxAlleles; ++i)
#pragma unroll
                    for (int j = 0; j < kLaneMaxAlleles; ++j)
                        if ((uint32_t)i < nA && (uint32_t)j < nA) {
                            float q = W[i][j];
#pragma unroll
                            for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                            if (lane == 0) s_w[warp][i * nA + j] = q;
                        }
            }
        } else {
        float qa[NI][V];
#pragma unroll
        for (int it = 0; it < NI; ++it)
#pragma unroll
            for (int v = 0; v < V; ++v) qa[it][v] = 0.f;
        warp_step<V, G, FWD>(a, al_out, al_in, inv, first_item, item_stride, qa, warp_sum);

        if (FWD) {
            // posterior partials of this warp: sum of the accumulators per (row allele, column allele), alleles 1-based
            // classes (0 = missing never carries mass)
            const uint32_t R = a.R, nA = a.nA_out;
            const bool cok = gl * V < R;
            const uint32_t ccol = cok ? gl * V : 0u;
            const uint32_t n_iter = (R + rpi - 1) / rpi;
            uint32_t cls_c[V], cls_r[NI];
#pragma unroll
            for (int v = 0; v < V; ++v) cls_c[v] = cok ? al_out[ccol + v] : 0u;
#pragma unroll
            for (int it = 0; it < NI; ++it) {
                const uint32_t r = it * rpi + sub;
                cls_r[it] = ((uint32_t)it < n_iter && r < R && cok) ? al_out[r] : 0u;
            }
            for (uint32_t i = 1; i <= nA; ++i)
                for (uint32_t j = 1; j <= nA; ++j) {
                    float q = 0.f;
#pragma unroll
                    for (int it = 0; it < NI; ++it)
#pragma unroll
                        for (int v = 0; v < V; ++v) q += (cls_r[it] == i && cls_c[v] == j) ? qa[it][v] : 0.f;
#pragma unroll
                    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                    if (lane == 0) s_w[warp][(i - 1) * nA + (j - 1)] = q;
                }
        }
        }
        if (lane == 0) s_sum[warp] = warp_sum;
        __syncthreads();
        if (FWD) {
            const uint32_t nA = a.nA_out;
            float *dst = w.w_partial + s_step[slot].w_off + (uint64_t)blockIdx.x * nA * nA;
            for (uint32_t p = tid; p < nA * nA; p += blockDim.x) {
                float q = 0.f;
#pragma unroll
                for (int k = 0; k < kWarpKernelWarps; ++k) q += s_w[k][p];
                dst[p] = q;
            }
        }
        float ts = 0.f;
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < kWarpKernelWarps; ++k) ts += s_sum[k];
        }
        const double total = wave_barrier_sum(w.barrier, s, ts, &s_total);
        if (blockIdx.x == 0 && tid == 0) *a.out_scale = total;
        inv = (float)(1.0 / total);
    }
}

#ifndef GG_PAIR_CTAS
#define GG_PAIR_CTAS 1      // measured: 1 CTA per SM 88.9 ms, 2 CTAs 94.1 ms per configs[0] sweep (the grid barrier grows with the CTA count)
#endif

// Row-pair-per-lane column step for very small even panels (R <= 16; configs[0] has R = 10): a bipartition block is
// worked on by R/2 lanes of an 8-lane group, each lane owning two rows (2R values, R/2 128-bit accesses), four blocks
// per warp.  ncu on the thread-per-block mapping (lane_step) showed what bounds a small-panel step: not bytes and not
// issue slots (6-9 % busy) but the latency of ONE thread walking a whole block (~1,500 dependent instructions at 8
// warps per SM, 190-210 registers) plus the load round trips in front of it.  Splitting a block over five lanes cuts
// that chain to a fifth; columns with fewer blocks than the grid has threads (every column up to 2^13 blocks) now occupy
// five times as many lanes instead of leaving them idle.
// Row sums stay register arithmetic; column sums and the block total are a 3-step butterfly inside the group.  The
// posterior is kept per owned (row, column) position for the whole step (P) and binned by allele class once per step.
template <int R, bool FWD>
__device__ __forceinline__ void pair_step(const StepArgs<float> &a, const uint8_t *al_o, const uint8_t *al_i, const float inv,
                                          const uint32_t first_task, const uint32_t task_stride, float (&P)[2 * R], float &sum_acc) {
    static_assert(R % 2 == 0 && R >= 2 && R <= 16, "pair_step: even panels up to 16 haplotypes");
    constexpr uint32_t RR = R * R, o_row = RR, o_col = RR + R, o_tot = RR + 2 * R;
    constexpr int NP = R / 2;                      // row pairs per block = float4 per row pair
    const uint32_t lane = threadIdx.x & 31u, g8 = lane & 7u, sub = lane >> 3;
    const bool act = g8 < (uint32_t)NP;
    const uint32_t r = act ? 2u * g8 : 0u;
    const uint32_t BS = a.BS;
    // the step record sits in shared memory, which hides the address space of the pointers in it from the compiler
    // (generic LD/ST); say that they are global
    const float *const p_in = a.in, *const p_beta = a.beta, *const p_fgo = a.fg_out, *const p_fgi = a.fg_in;
    float *const p_out = a.out;
    __builtin_assume(__isGlobal(p_in));
    __builtin_assume(__isGlobal(p_out));
    __builtin_assume(__isGlobal(p_fgo));
    if (FWD) __builtin_assume(__isGlobal(p_beta));
    if (!FWD) __builtin_assume(__isGlobal(p_fgi));
    const uint32_t NCo = a.nA_out + 1, NCi = a.nA_in + 1, fgs_o = 2 * NCo, fgs_i = 2 * NCi;
    const uint32_t nred = FWD ? (1u << a.n_free_left) : (1u << (a.u_in - a.n_sh));
    const uint32_t n_bcast = FWD ? (a.u_out - a.n_sh) : a.n_free_left;
    const bool sh_id = a.shared_mask == ((a.n_sh >= 32 ? 0u : (1u << a.n_sh)) - 1u);
    const bool fr_id = a.free_mask == ((((a.n_free_left >= 32 ? 0u : (1u << a.n_free_left)) - 1u)) << a.n_sh);
    const float cA = (float)a.tA * inv, cB = (float)a.tB * inv, cC = (float)a.tC * inv;
    const uint32_t n_items = 1u << a.u_out, n_tasks = (n_items + 3u) >> 2;
    uint32_t co[R], ci[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        co[j] = al_o[j];
        ci[j] = al_i[j];
    }
    const uint32_t ao[2] = {al_o[r], al_o[r + 1]}, ai[2] = {al_i[r], al_i[r + 1]};
    // the first block of every item's reduce set (and its beta block) is asked into L1 one task ahead: lane l looks
    // after line l % 4 of block l / 4 % 4 of the next task, lanes 16.. after the beta lines — a warp's next round then
    // starts on L1 hits instead of an L2 round trip behind the previous round's stores
    constexpr uint32_t kLines = ((RR + 2 * R + 4 + 3) / 4 * 16 + 127) / 128;      // 128-byte lines per block record
    auto prefetch_task = [&](uint32_t task) {
        const uint32_t pi = (lane >> 2) & 3u, pl = lane & 3u;
        const uint32_t it = task * 4u + pi;
        if (it >= n_items) return;
        const uint32_t plo = it & ((1u << n_bcast) - 1u), psh = it >> n_bcast;
        const bool want_beta = lane >= 16;
        if (want_beta && !FWD) return;
        uint32_t blk_idx;
        if (FWD) blk_idx = want_beta ? (psh | (plo << a.n_sh)) : (sh_id ? psh : dev_pdep(psh, a.shared_mask));
        else blk_idx = psh;
        const char *base = reinterpret_cast<const char *>((want_beta ? p_beta : p_in) + (uint64_t)blk_idx * BS);
        for (uint32_t l = pl; l < kLines; l += 4) prefetch_l1(base + l * 128);
    };
    for (uint32_t task = first_task; task < n_tasks; task += task_stride) {
        if (task + task_stride < n_tasks) prefetch_task(task + task_stride);
        const uint32_t item_raw = task * 4u + sub;
        const bool valid = item_raw < n_items, live = valid && act;
        const uint32_t item = valid ? item_raw : 0u;
        const uint32_t lo = item & ((1u << n_bcast) - 1u), sh = item >> n_bcast;
        uint32_t b_out, b_in0;
        if (FWD) {
            b_out = sh | (lo << a.n_sh);
            b_in0 = sh_id ? sh : dev_pdep(sh, a.shared_mask);
        } else {
            b_out = (sh_id ? sh : dev_pdep(sh, a.shared_mask)) | (fr_id ? (lo << a.n_sh) : dev_pdep(lo, a.free_mask));
            b_in0 = sh;
        }
        const float *fgo = p_fgo + (uint64_t)b_out * fgs_o;
        float *outp = p_out + (uint64_t)b_out * BS;
        float gq[R], cb[R], colacc[R], gk0[R], x[2 * R];
        float totv = 0.f, rowv0 = 0.f, rowv1 = 0.f;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            gq[j] = fgo[NCo + co[j]];
            cb[j] = 0.f;
            colacc[j] = 0.f;
            gk0[j] = 1.f;
        }
#pragma unroll
        for (int e = 0; e < 2 * R; ++e) x[e] = 0.f;
        if (!FWD && nred == 1) {
            const float *fgi = p_fgi + (uint64_t)b_in0 * fgs_i;
#pragma unroll
            for (int j = 0; j < R; ++j) gk0[j] = fgi[NCi + ci[j]];
        }
        {
            uint32_t subm = 0;
            for (uint32_t k = 0; k < nred; ++k) {
                const uint64_t bk = FWD ? (uint64_t)(b_in0 | subm) : (uint64_t)(b_in0 | (k << a.n_sh));
                subm = (subm - a.free_mask) & a.free_mask;
                const float *blk = p_in + bk * BS;
                const float4 *src = reinterpret_cast<const float4 *>(blk + r * R);
                float4 t4[NP];
#pragma unroll
                for (int v = 0; v < NP; ++v) t4[v] = src[v];
#pragma unroll
                for (int j = 0; j < R; j += 2) {
                    const float2 t = *reinterpret_cast<const float2 *>(blk + o_col + j);
                    cb[j] += t.x;
                    cb[j + 1] += t.y;
                }
                const float2 rv = *reinterpret_cast<const float2 *>(blk + o_row + r);
                rowv0 += rv.x;
                rowv1 += rv.y;
                totv += blk[o_tot];
                float fk0 = 1.f, fk1 = 1.f;
                const float *fgk = nullptr;
                if (!FWD) {
                    fgk = p_fgi + bk * fgs_i;
                    fk0 = fgk[ai[0]];
                    fk1 = fgk[ai[1]];
                }
#pragma unroll
                for (int v = 0; v < NP; ++v) {
                    const float ld[4] = {t4[v].x, t4[v].y, t4[v].z, t4[v].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int e = 4 * v + q, j = e % R;
                        if (FWD) {
                            x[e] += ld[q];
                        } else {
                            const float gk = nred == 1 ? gk0[j] : fgk[NCi + ci[j]];
                            x[e] = fmaf(ld[q], (e < R ? fk0 : fk1) * gk, x[e]);
                        }
                    }
                }
            }
        }
        const float ctot = cC * totv;
#pragma unroll
        for (int j = 0; j < R; ++j) cb[j] = fmaf(cB, cb[j], ctot);
        float xo[2 * R], rs[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float fo = live ? fgo[ao[q]] : 0.f;
            const float rowterm = cB * (q == 0 ? rowv0 : rowv1);
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const float base = fmaf(cA, x[q * R + j], rowterm + cb[j]);
                float o, w;
                if (FWD) {
                    o = (fo * gq[j]) * base;
                    w = o;
                } else {
                    o = (ao[q] != 0 && co[j] != 0) ? base : 0.f;
                    w = o * (fo * gq[j]);
                }
                acc += w;
                colacc[j] += w;
                xo[q * R + j] = o;
            }
            rs[q] = acc;
        }
        if (live) {
            float4 *dst = reinterpret_cast<float4 *>(outp + r * R);
#pragma unroll
            for (int v = 0; v < NP; ++v) dst[v] = make_float4(xo[4 * v], xo[4 * v + 1], xo[4 * v + 2], xo[4 * v + 3]);
            *reinterpret_cast<float2 *>(outp + o_row + r) = make_float2(rs[0], rs[1]);
        }
        if (FWD) {
            const float4 *bsrc = reinterpret_cast<const float4 *>(p_beta + (uint64_t)b_out * BS + r * R);
#pragma unroll
            for (int v = 0; v < NP; ++v) {
                const float4 t = bsrc[v];
                P[4 * v] = fmaf(xo[4 * v], t.x, P[4 * v]);
                P[4 * v + 1] = fmaf(xo[4 * v + 1], t.y, P[4 * v + 1]);
                P[4 * v + 2] = fmaf(xo[4 * v + 2], t.z, P[4 * v + 2]);
                P[4 * v + 3] = fmaf(xo[4 * v + 3], t.w, P[4 * v + 3]);
            }
        }
        // column sums and block total over the lanes of the group (idle lanes carry zeros)
        float tot = rs[0] + rs[1];
#pragma unroll
        for (int shf = 1; shf < 8; shf <<= 1) {
#pragma unroll
            for (int j = 0; j < R; ++j) colacc[j] += __shfl_xor_sync(0xffffffffu, colacc[j], shf);
            tot += __shfl_xor_sync(0xffffffffu, tot, shf);
        }
        if (live) {
#pragma unroll
            for (int jj = 0; jj < NP; ++jj)
                if (g8 == (uint32_t)jj) *reinterpret_cast<float2 *>(outp + o_col + 2 * jj) = make_float2(colacc[2 * jj], colacc[2 * jj + 1]);
            if (g8 == 0) {
                outp[o_tot] = tot;
                sum_acc += FWD ? tot : totv * inv;
            }
        }
    }
}

template <int R, bool FWD>
__global__ void __launch_bounds__(32 * kWarpKernelWarps, GG_PAIR_CTAS)
pair_wave_kernel(const WaveArgs w) {
    __shared__ float s_sum[kWarpKernelWarps];
    __shared__ float s_w[FWD ? kWarpKernelWarps : 1][FWD ? GG_MAX_ALLELES * GG_MAX_ALLELES : 1];
    __shared__ __align__(16) WaveStep s_step[2];
    __shared__ double s_total;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    auto stage_args = [&](uint32_t s, uint32_t slot) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(w.steps + s);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_step[slot]);
        if (tid < sizeof(WaveStep) / 4) dst[tid] = __ldcg(src + tid);
    };
    stage_args(0, 0);
    __syncthreads();
    float inv = 1.f;
    for (uint32_t s = 0; s < w.n_steps; ++s) {
        const uint32_t slot = s & 1u;
        const StepArgs<float> &a = s_step[slot].a;
        const uint8_t *al_out = s_step[slot].al_out, *al_in = s_step[slot].al_in;
        if (s + 1 < w.n_steps) stage_args(s + 1, slot ^ 1u);
        if (s == 0) inv = (float)(1.0 / *a.in_scale);
        float P[2 * R];
#pragma unroll
        for (int e = 0; e < 2 * R; ++e) P[e] = 0.f;
        float sum_acc = 0.f;
        // warp tasks of four consecutive blocks, dealt round-robin over the CTAs first
        pair_step<R, FWD>(a, al_out, al_in, inv, blockIdx.x + gridDim.x * warp, gridDim.x * kWarpKernelWarps, P, sum_acc);
#pragma unroll
        for (int o = 16; o; o >>= 1) sum_acc += __shfl_xor_sync(0xffffffffu, sum_acc, o);
        if (FWD) {
            // bin the per-position accumulators by (allele of the row, allele of the column); classes are 1-based, 0 =
            // missing never carries mass
            const uint32_t nA = a.nA_out, g8 = lane & 7u;
            const bool act = g8 < (uint32_t)(R / 2);
            const uint32_t r = act ? 2u * g8 : 0u;
            const uint32_t cr0 = act ? al_out[r] : 0u, cr1 = act ? al_out[r + 1] : 0u;
            uint32_t cc[R];
#pragma unroll
            for (int j = 0; j < R; ++j) cc[j] = al_out[j];
            for (uint32_t i = 1; i <= nA; ++i)
                for (uint32_t c = 1; c <= nA; ++c) {
                    float q = 0.f;
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        q += (cr0 == i && cc[j] == c) ? P[j] : 0.f;
                        q += (cr1 == i && cc[j] == c) ? P[R + j] : 0.f;
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                    if (lane == 0) s_w[warp][(i - 1) * nA + (c - 1)] = q;
                }
        }
        if (lane == 0) s_sum[warp] = sum_acc;
        __syncthreads();
        if (FWD) {
            const uint32_t nA = a.nA_out;
            float *dst = w.w_partial + s_step[slot].w_off + (uint64_t)blockIdx.x * nA * nA;
            for (uint32_t p = tid; p < nA * nA; p += blockDim.x) {
                float q = 0.f;
#pragma unroll
                for (int k = 0; k < kWarpKernelWarps; ++k) q += s_w[k][p];
                dst[p] = q;
            }
        }
        float ts = 0.f;
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < kWarpKernelWarps; ++k) ts += s_sum[k];
        }
        const double total = wave_barrier_sum(w.barrier, s, ts, &s_total);
        if (blockIdx.x == 0 && tid == 0) *a.out_scale = total;
        inv = (float)(1.0 / total);
    }
}

// genotype tables of the forward steps of a wave run: one warp per step adds the per-CTA partials up in fixed order
__global__ void __launch_bounds__(32)
wave_posterior_kernel(const WaveStep *steps, uint32_t n_steps, const float *w_partial, uint32_t grid) {
    __shared__ double s_L[GG_MAX_ALLELES * (GG_MAX_ALLELES + 1) / 2];
    const uint32_t s = blockIdx.x, lane = threadIdx.x;
    if (s >= n_steps) return;
    const StepArgs<float> &a = steps[s].a;
    const uint32_t nA = a.nA_out, ng = nA * (nA + 1) / 2, np = nA * nA;
    const float *src = w_partial + steps[s].w_off;
    for (uint32_t g = lane; g < ng; g += 32) s_L[g] = 0.0;
    __syncwarp();
    double norm = 0.0;
    for (uint32_t p = 0; p < np; ++p) {
        double acc = 0.0;
        for (uint32_t k = lane; k < grid; k += 32) acc += (double)src[(uint64_t)k * np + p];
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            // ordered allele pair -> genotype index, reference src/genotypehmm.cpp:524-528
            const uint32_t i = p / nA, j = p % nA;
            s_L[i + j * (j + 1) / 2] += acc;
            norm += acc;
        }
    }
    norm = __shfl_sync(0xffffffffu, norm, 0);
    __syncwarp();
    for (uint32_t g = lane; g < ng; g += 32) a.result[g] = s_L[g] / norm;
}

}  // namespace gg
