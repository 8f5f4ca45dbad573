// Small panels (R <= 32), many columns per launch: the wave kernel.
//
// A column of configs[0] (10 haplotypes, <= 2^15 blocks of 100 values) is a few MB that sit in L2, and one column step
// is a few microseconds of work: launched column by column, a sweep is bound by launch latency, the per-launch
// prologue and the last-CTA epilogue (10,000 launches for 5,000 columns).  Here ONE cooperative launch walks a whole
// run of consecutive steps of one direction: every warp does its share of a step's output blocks with the same code
// as the per-column kernel (warp_step, hmm_warp_kernel.cuh), the CTAs meet at a grid barrier (one atomic arrival
// counter, acquire / release), and the next step starts.  What a launch boundary used to provide is done in place:
//
//   column scale   every CTA leaves the sum of what it produced in a double-buffered slot; after the barrier each CTA
//                  adds the slots up in the same fixed order, so all of them fold the same 1/sum into the next step
//                  (CTA 0 also stores it where later launches expect it);
//   posterior      forward steps bin their per-position accumulators by allele-class pair in registers (the class of a
//                  (row, column) position is the same for every block of a column), one partial vector per CTA and
//                  step; wave_posterior_kernel adds them up in fixed order after the run.
//
// No atomics on data, fixed summation orders: results are bit-reproducible (not bit-identical to the per-column
// kernels: the partial sums are grouped differently).
#pragma once
#include "hmm_warp_kernel.cuh"

#ifndef GG_WAVE_MIN_CTAS
#define GG_WAVE_MIN_CTAS 2
#endif

namespace gg {

struct WaveStep {
    StepArgs<float> a;
    uint64_t w_off;       // forward: offset (floats) of this step's [grid][nA_out^2] posterior partials
    uint64_t pad;
};

struct WaveArgs {
    const WaveStep *steps;
    uint32_t n_steps;
    float *sum_partial;           // [2][gridDim.x]
    float *w_partial;             // forward: posterior partials of every step
    unsigned int *barrier;        // arrival counter, zero at launch
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all CTAs of the (co-resident, cooperatively launched) grid; `target` = arrivals expected so far
__device__ __forceinline__ void wave_barrier(unsigned int *bar, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (ld_acquire_u32(bar) < target) { }
    }
    __syncthreads();
}

template <int V, int G, bool FWD>
__global__ void __launch_bounds__(32 * kWarpKernelWarps, GG_WAVE_MIN_CTAS)
wave_kernel(const WaveArgs w) {
    constexpr int NI = (V * G * G / 32) > 0 ? (V * G * G / 32) : 1;
    __shared__ float s_sum[kWarpKernelWarps];
    __shared__ float s_inv;
    __shared__ float s_w[FWD ? kWarpKernelWarps : 1][FWD ? GG_MAX_ALLELES * GG_MAX_ALLELES : 1];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr uint32_t rpi = 32u / G;
    const uint32_t sub = lane / G, gl = lane % G;
    const uint32_t first_item = blockIdx.x * kWarpKernelWarps + warp, item_stride = gridDim.x * kWarpKernelWarps;

    // step arguments and the two allele vectors are staged in shared memory one step ahead (the grid barrier's acquire
    // drops the L1 contents, so without this every step would begin with two dependent L2 round trips)
    __shared__ __align__(16) WaveStep s_step[2];
    __shared__ uint8_t s_al[2][2][32];
    auto stage_args = [&](uint32_t s, uint32_t slot) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(w.steps + s);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_step[slot]);
        if (tid < sizeof(WaveStep) / 4) dst[tid] = __ldcg(src + tid);
    };
    auto stage_alleles = [&](uint32_t slot) {       // after a __syncthreads that follows stage_args(slot)
        const StepArgs<float> &n = s_step[slot].a;
        if (tid < 32) s_al[slot][0][tid] = tid < n.R ? n.al_out[tid] : (uint8_t)0;
        else if (tid < 64) s_al[slot][1][tid - 32] = tid - 32 < n.R ? n.al_in[tid - 32] : (uint8_t)0;
    };
    stage_args(0, 0);
    __syncthreads();
    stage_alleles(0);
    __syncthreads();

    float inv = 1.f;
    for (uint32_t s = 0; s < w.n_steps; ++s) {
        const uint32_t slot = s & 1u;
        const StepArgs<float> &a = s_step[slot].a;
        const uint8_t *al_out = s_al[slot][0], *al_in = s_al[slot][1];
        if (s + 1 < w.n_steps) stage_args(s + 1, slot ^ 1u);      // visible after the __syncthreads before the barrier
        if (s == 0) inv = (float)(1.0 / *a.in_scale);
        float qa[NI][V];
#pragma unroll
        for (int it = 0; it < NI; ++it)
#pragma unroll
            for (int v = 0; v < V; ++v) qa[it][v] = 0.f;
        float warp_sum = 0.f;
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
        if (lane == 0) s_sum[warp] = warp_sum;
        __syncthreads();
        if (s + 1 < w.n_steps) stage_alleles(slot ^ 1u);          // next step's arguments are in place since the barrier above
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
        float *slots = w.sum_partial + (s & 1u) * gridDim.x;
        if (tid == 0) {
            float ts = 0.f;
#pragma unroll
            for (int k = 0; k < kWarpKernelWarps; ++k) ts += s_sum[k];
            slots[blockIdx.x] = ts;
        }
        wave_barrier(w.barrier, (s + 1) * gridDim.x);
        // the column sum, identically in every CTA (warp 0: lane l takes slots l, l+32, ..., then a butterfly)
        if (warp == 0) {
            double acc = 0.0;
            for (uint32_t k = lane; k < gridDim.x; k += 32) acc += (double)__ldcg(slots + k);
#pragma unroll
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                s_inv = (float)(1.0 / acc);
                if (blockIdx.x == 0) *a.out_scale = acc;
            }
        }
        __syncthreads();
        inv = s_inv;
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
