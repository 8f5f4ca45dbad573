// sm_100a kernels of the GenotypeHMM column sweep.
//
// One column of the HMM holds S = 2^u * R^2 values of type T (fp32 in
// production, fp64 for the precise mode) in the reference's order b, R1, R2
// (src/column.cpp:41-45).  Each bipartition block b is one contiguous record of
// BS = R^2 + 2R + 4 elements (rounded up to a multiple of 4) that carries its
// "side sums" behind the values, so that a whole block moves with ONE bulk copy:
//
//     block b at b*BS:  [ values R*R ][ row sums R ][ col sums R ][ total, 0, 0, 0 ]
//
// The side sums are what makes the Li-Stephens transition elementwise.  The
// reference evaluates (src/genotypehmm.cpp:319 backward, :505 forward)
//     T(M)[R1,R2] = qr^2 M + qr pr (row[R1] + col[R2] - 2M) + pr^2 (tot - row[R1] - col[R2] + M)
//                 = tA*M + tB*(row[R1] + col[R2]) + tC*tot,   tA=(qr-pr)^2, tB=pr(qr-pr), tC=pr^2
// with row/col/tot the sums of M over R2 / R1 / both.  Every step kernel
// writes, next to the column it produces, the row/col/tot sums of the
// quantity the NEXT step will push through T (forward: alpha itself;
// backward: beta*emission), so a step reads each input value exactly once
// and needs no dense R^2 x R^2 contraction and no second pass.
//
//   forward  (reference compute_forward_column, src/genotypehmm.cpp:340-561)
//     alpha_c[b,r]  = E_c(b,r) * T_c( sum_{terminating bits} alpha_{c-1} )[shared bits of b, r]
//     L_c[g]       += alpha_c[b,r] * beta_c[b,r],  g = a(R1) + a(R2)(a(R2)+1)/2     (:524-537)
//   backward (reference compute_backward_column, src/genotypehmm.cpp:203-337)
//     beta_{c-1}[b',r] = [alleles of r defined at c-1] * T_c( sum_{new bits} beta_c * E_c )[shared bits of b', r]
//
// E_c(b,r) = F_b[a_c(R1)] * G_b[a_c(R2)] (reference update_emission_probability,
// src/genotypehmm.cpp:576-629): F_b / G_b are the products of the per-read
// emission vectors over the reads on side 0 / side 1 of bipartition b; the
// emission kernel tabulates them per column (fp64 products, rounded once).
//
// Scaling: a column is stored unnormalised and its sum goes to a per-column
// device scalar (written by the last CTA to finish, fixed summation order, so
// results are deterministic); the consumer folds 1/sum into tA,tB,tC.  All
// outputs are invariant to these scales.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gg {

constexpr int kMaxClasses = GG_MAX_ALLELES + 1;           // allele classes incl. "missing" (index 0)
constexpr int kPartialStride = kMaxClasses * kMaxClasses + 3;  // W[i][j], column sum, spare

struct ColDesc {              // device copy of the per-column plan (emission kernel only)
    uint32_t u, n_alleles, em_count, present;   // present: bit i set = allele i occurs in the panel at this column
    uint64_t em_off, em_prob_off;
    uint64_t fg_prefix;       // prefix sum (elements) of the F/G table sizes
};

template <typename T>
struct StepArgs {
    const T *in;              // input column (values + side sums); nullptr for the first forward / last backward column
    T *out;                   // output column
    const T *beta;            // forward: beta of the output column (posterior)
    const T *fg_out;          // F/G table of the output column
    const T *fg_in;           // backward: F/G table of the input column
    const double *in_scale;   // sum of the input column (device scalar)
    double *out_scale;        // sum of the output column
    double *result;           // forward: genotype table of this column
    T *partial;               // [gridDim.x][kPartialStride]
    unsigned int *counter;    // ticket for the last-CTA reduction (self-resetting)
    const uint8_t *al_out;    // [R] allele+1 of the output column
    const uint8_t *al_in;     // [R] allele+1 of the input column
    const uint16_t *row_order;// [R] haplotypes of the output column sorted by allele (forward), may be nullptr
    uint32_t R, G;            // haplotypes; lanes per row (power of two <= 32)
    uint32_t BS;              // block stride in elements: R*R + 2R + 4 rounded up to 4
    uint32_t u_out, u_in;
    uint32_t nA_out, nA_in;
    uint32_t n_sh;            // bits shared between the two columns
    uint32_t shared_mask;     // their positions in the LEFT column's index
    uint32_t free_mask;       // positions of the left column's terminating bits
    uint32_t n_free_left;     // popcount(free_mask)
    double tA, tB, tC;
};

template <typename T, int V> struct VecOf;
template <> struct VecOf<float, 1> { typedef float type; };
template <> struct VecOf<float, 2> { typedef float2 type; };
template <> struct VecOf<float, 4> { typedef float4 type; };
template <> struct VecOf<double, 1> { typedef double type; };
template <> struct VecOf<double, 2> { typedef double2 type; };

template <typename T, int V>
__device__ __forceinline__ void load_vec(T (&d)[V], const T *p) {
    typedef typename VecOf<T, V>::type VT;
    union { VT v; T e[V]; } x;
    x.v = *reinterpret_cast<const VT *>(p);
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] = x.e[i];
}
template <typename T, int V>
__device__ __forceinline__ void store_vec(T *p, const T (&d)[V]) {
    typedef typename VecOf<T, V>::type VT;
    union { VT v; T e[V]; } x;
#pragma unroll
    for (int i = 0; i < V; ++i) x.e[i] = d[i];
    *reinterpret_cast<VT *>(p) = x.v;
}

__device__ __forceinline__ uint32_t dev_pdep(uint32_t x, uint32_t mask) {
    uint32_t r = 0;
    for (uint32_t m = mask; m; m &= m - 1) {
        if (x & 1u) r |= m & (0u - m);
        x >>= 1;
    }
    return r;
}

// Last-CTA reduction: fixed-order sum of one slot over the per-CTA partial records, by ONE WARP (lane l takes records
// l, l+32, ... in fp64, then a butterfly).  The order depends on the grid size only, so results stay bit-reproducible;
// a single thread walking all records serially cost several microseconds per launch (every load an L2 round trip).
template <typename T>
__device__ __forceinline__ double warp_sum_partials(const T *partial, uint32_t nblk, uint32_t slot, uint32_t lane) {
    double acc = 0.0;
    for (uint32_t blk = lane; blk < nblk; blk += 32) acc += (double)__ldcg(partial + (uint64_t)blk * kPartialStride + slot);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

// Shared-memory carve-up, identical on host (size) and device (pointers).
struct SmemLayout {
    uint32_t off_rowv, off_colv, off_fo, off_fi, off_colred, off_totred, off_q, off_al_o, off_al_i, off_rord, total;
};
template <typename T>
__host__ __device__ inline SmemLayout smem_layout(uint32_t R, uint32_t ngrp, uint32_t ncls_out, bool fwd) {
    SmemLayout L;
    uint32_t Rp = (R + 3u) & ~3u;
    uint32_t o = 0;
    L.off_rowv = o; o += Rp * sizeof(T);
    L.off_colv = o; o += Rp * sizeof(T);
    L.off_fo = o; o += 20 * sizeof(T);
    L.off_fi = o; o += 20 * sizeof(T);
    L.off_colred = o; o += ngrp * Rp * sizeof(T);
    L.off_totred = o; o += 2 * ngrp * sizeof(T);
    L.off_q = o; o += fwd ? ngrp * ncls_out * Rp * sizeof(T) : 0;
    L.off_al_o = o; o += Rp;
    L.off_al_i = o; o += Rp;
    o = (o + 3u) & ~3u;
    L.off_rord = o; o += Rp * 2;
    L.total = (o + 15u) & ~15u;
    return L;
}

// ---------------------------------------------------------------------------------------------
// Emission tables: fg[b][side][slot], slot 0 = missing allele (0), slot i+1 = allele i.
//
// One thread per (bipartition b, side): the products over the reads on that side for every allele, in fp64
// (reference update_emission_probability, src/genotypehmm.cpp:576-629, without its Gray-code iteration order).
// Range (SURVEY §7 hard part 1): a column's emissions may be scaled by any positive constant (A.6), and inside a
// block F may be traded against G.  PASS 0 finds the column's largest product m_c = max_b max F_b * max G_b
// (emax[c], atomic max on the bit pattern of a non-negative double); PASS 1 writes F_b * 2^-eF(b) and
// G_b * 2^(eF(b) - eM(c)) with eF = exponent of max F_b and eM = exponent of m_c.  The best state of every column
// then has emission in [1, 4) whatever the reads say, and because the factors are powers of two the fp32 roundings
// - hence the posteriors - are bit-identical to the unshifted tables wherever those did not underflow.
// ---------------------------------------------------------------------------------------------
// NAM: compile-time bound on the alleles of the column (2, 4, 8 or GG_MAX_ALLELES): the products live in registers,
// so the loop over alleles must be unrolled, and unrolled to 16 a biallelic column spent 14 of 16 issue slots on
// predicated-off multiplies (configs[0]: the two passes were 15 % of a sweep).
template <typename T, int PASS, int NAM>
__device__ __forceinline__ void emission_body(const ColDesc &d, const uint32_t c, const uint8_t *__restrict__ em_side,
                                              const double *__restrict__ em_prob, T *__restrict__ fg, double *__restrict__ emax) {
    const uint32_t nA = d.n_alleles, NC = nA + 1;
    const uint64_t n2 = (uint64_t)2 << d.u;                      // (b, side) pairs; even, so lanes 2k / 2k+1 share b
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    int eM = 0;
    if (PASS == 1) {
        const double m = emax[c];
        eM = m > 0.0 ? ilogb(m) : 0;
    }
    double wmax = 0.0;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n2; base += stride) {
        const uint64_t idx = base + lane;
        const bool active = idx < n2;
        const uint32_t side = (uint32_t)(idx & 1), b = (uint32_t)(idx >> 1);
        double p[NAM];
#pragma unroll
        for (int i = 0; i < NAM; ++i) p[i] = 1.0;
        if (active)
            for (uint32_t e = 0; e < d.em_count; ++e) {
                const uint32_t sc = em_side[d.em_off + e];
                const uint32_t s = sc >= 64 ? sc - 64 : ((b >> sc) & 1u);
                if (s == side) {
                    const double *q = em_prob + d.em_prob_off + (uint64_t)e * nA;
#pragma unroll
                    for (int i = 0; i < NAM; ++i)
                        if ((uint32_t)i < nA) p[i] *= q[i];
                }
            }
        // only alleles some panel haplotype carries here can be the allele of a state (the others never reach the sweep)
        double mx = 0.0;
#pragma unroll
        for (int i = 0; i < NAM; ++i)
            if ((uint32_t)i < nA && ((d.present >> i) & 1u)) mx = fmax(mx, p[i]);
        const double other = __shfl_xor_sync(0xffffffffu, mx, 1);
        if (PASS == 0) {
            if (active) wmax = fmax(wmax, mx * other);
        } else if (active) {
            const double mF = side == 0 ? mx : other;
            const int eF = mF > 0.0 ? ilogb(mF) : 0;
            const int sh = side == 0 ? -eF : eF - eM;
            T *dst = fg + idx * NC;
            dst[0] = T(0);
#pragma unroll
            for (int i = 0; i < NAM; ++i)
                if ((uint32_t)i < nA) dst[1 + i] = (T)ldexp(p[i], sh);
        }
    }
    if (PASS == 0) {
#pragma unroll
        for (int o = 16; o; o >>= 1) wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        if (lane == 0 && wmax > 0.0)
            atomicMax(reinterpret_cast<unsigned long long *>(emax + c), (unsigned long long)__double_as_longlong(wmax));
    }
}

template <typename T, int PASS>
__global__ void __launch_bounds__(256)
emission_kernel(const ColDesc *__restrict__ cols, const uint8_t *__restrict__ em_side,
                const double *__restrict__ em_prob, T *__restrict__ region, uint64_t prefix_lo, uint32_t lo,
                double *__restrict__ emax) {
    const uint32_t c = lo + blockIdx.y;
    const ColDesc d = cols[c];
    T *fg = PASS == 1 ? region + (d.fg_prefix - prefix_lo) : nullptr;
    if (d.n_alleles <= 2) emission_body<T, PASS, 2>(d, c, em_side, em_prob, fg, emax);
    else if (d.n_alleles <= 4) emission_body<T, PASS, 4>(d, c, em_side, em_prob, fg, emax);
    else if (d.n_alleles <= 8) emission_body<T, PASS, (GG_MAX_ALLELES < 8 ? GG_MAX_ALLELES : 8)>(d, c, em_side, em_prob, fg, emax);
    else emission_body<T, PASS, GG_MAX_ALLELES>(d, c, em_side, em_prob, fg, emax);
}

// ---------------------------------------------------------------------------------------------
// Column step.  FWD: produce alpha of column `out` from alpha of the column to its left and
// accumulate the genotype posterior with beta.  !FWD: produce beta of column `out` from beta of
// the column to its right.  One work item = one output bipartition block (R*R values); a CTA
// walks items blockIdx.x, +gridDim.x, ...; inside an item each group of G lanes owns one row R1
// and each lane NS*V fixed columns R2.
// ---------------------------------------------------------------------------------------------
template <typename T, int V, int NS, bool FWD>
__global__ void __launch_bounds__(256)
step_kernel(const StepArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t R = a.R, G = a.G;
    const uint32_t Rp = (R + 3u) & ~3u;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const uint32_t rpi = 32u / G;                 // rows per warp iteration
    const uint32_t sub = lane / G, gl = lane % G;
    const uint32_t grp = warp * rpi + sub, ngrp = nwarps * rpi;
    const uint32_t NCo = a.nA_out + 1, NCi = a.nA_in + 1;
    const uint32_t fgs_o = 2 * NCo, fgs_i = 2 * NCi;
    const SmemLayout L = smem_layout<T>(R, ngrp, NCo, FWD);
    T *s_rowv = reinterpret_cast<T *>(smem_raw + L.off_rowv);
    T *s_colv = reinterpret_cast<T *>(smem_raw + L.off_colv);
    T *s_fo = reinterpret_cast<T *>(smem_raw + L.off_fo);
    T *s_fi = reinterpret_cast<T *>(smem_raw + L.off_fi);
    T *s_colred = reinterpret_cast<T *>(smem_raw + L.off_colred);
    T *s_totred = reinterpret_cast<T *>(smem_raw + L.off_totred);
    T *s_q = reinterpret_cast<T *>(smem_raw + L.off_q);
    uint8_t *s_al_o = smem_raw + L.off_al_o;
    uint8_t *s_al_i = smem_raw + L.off_al_i;
    uint16_t *s_rord = reinterpret_cast<uint16_t *>(smem_raw + L.off_rord);
    __shared__ uint32_t s_item[4];
    __shared__ double s_W[kPartialStride];
    __shared__ double s_L[GG_MAX_ALLELES * (GG_MAX_ALLELES + 1) / 2];
    __shared__ T s_totv;
    __shared__ bool s_last;

    const bool has_in = a.in != nullptr;
    const uint64_t RR = (uint64_t)R * R;
    const uint64_t nb_out = (uint64_t)1 << a.u_out;
    const uint64_t BS = a.BS;
    const uint64_t o_row = RR, o_col = RR + R, o_tot = RR + 2 * (uint64_t)R;   // side sums inside a block

    // reduce set of an item: FWD sums the left column over its terminating bits (submasks of
    // free_mask); BWD sums the right column over its new bits (the bits above n_sh).
    const uint32_t nred = !has_in ? 0u : (FWD ? (1u << a.n_free_left) : (1u << (a.u_in - a.n_sh)));
    // FWD items: new bits fastest; BWD items: terminating bits fastest (neighbouring CTAs share inputs)
    const uint32_t n_bcast = FWD ? (a.u_out - a.n_sh) : a.n_free_left;

    for (uint32_t t = tid; t < R; t += blockDim.x) {
        s_al_o[t] = a.al_out[t];
        s_al_i[t] = has_in ? a.al_in[t] : (uint8_t)0;
        s_rord[t] = (FWD && a.row_order) ? a.row_order[t] : (uint16_t)t;
    }
    if (FWD)
        for (uint32_t t = tid; t < ngrp * NCo * Rp; t += blockDim.x) s_q[t] = T(0);

    T inv = T(1);
    if (has_in) inv = (T)(1.0 / *a.in_scale);
    const T cA = (T)a.tA * inv, cB = (T)a.tB * inv, cC = (T)a.tC * inv;
    __syncthreads();

    // per-lane column identities (fixed for the whole kernel)
    uint32_t ccol[NS];
    uint8_t a2o[NS][V], a2i[NS][V];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        ccol[s] = (s * G + gl) * V;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const bool ok = ccol[s] < R;
            a2o[s][v] = ok ? s_al_o[ccol[s] + v] : (uint8_t)0;
            a2i[s][v] = ok ? s_al_i[ccol[s] + v] : (uint8_t)0;
        }
    }

    T qc[NS][V];               // FWD: posterior accumulators of the current row class
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int v = 0; v < V; ++v) qc[s][v] = T(0);
    uint32_t cur_cls = 0xffffffffu;
    T cta_sum = T(0);          // thread 0: sum of the produced column over this CTA's items (FWD: alpha)
    T sacc = T(0);             // BWD: per-thread sum of the produced beta values

    const uint32_t n_items = (uint32_t)nb_out;
    const uint32_t n_iter = (R + ngrp - 1) / ngrp;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        if (tid == 0) {
            const uint32_t lo = item & ((1u << n_bcast) - 1u), sh = item >> n_bcast;
            uint32_t b_out, b_in0;
            if (FWD) {
                b_out = sh | (lo << a.n_sh);
                b_in0 = dev_pdep(sh, a.shared_mask);
            } else {
                b_out = dev_pdep(sh, a.shared_mask) | dev_pdep(lo, a.free_mask);
                b_in0 = sh;
            }
            s_item[0] = b_out;
            s_item[1] = b_in0;
        }
        __syncthreads();
        const uint32_t b_out = s_item[0], b_in0 = s_item[1];
        // ---- item prologue: side sums of the reduced input, F vectors
        if (has_in) {
            for (uint32_t t = tid; t < R; t += blockDim.x) {
                T rv = T(0), cv = T(0);
                uint32_t subm = 0;
                for (uint32_t k = 0; k < nred; ++k) {
                    const uint64_t bk = FWD ? (uint64_t)(b_in0 | subm) : (uint64_t)(b_in0 | (k << a.n_sh));
                    rv += a.in[bk * BS + o_row + t];
                    cv += a.in[bk * BS + o_col + t];
                    subm = (subm - a.free_mask) & a.free_mask;
                }
                s_rowv[t] = rv;
                s_colv[t] = cv;
            }
            if (tid == blockDim.x - 1) {
                T tv = T(0);
                uint32_t subm = 0;
                for (uint32_t k = 0; k < nred; ++k) {
                    const uint64_t bk = FWD ? (uint64_t)(b_in0 | subm) : (uint64_t)(b_in0 | (k << a.n_sh));
                    tv += a.in[bk * BS + o_tot];
                    subm = (subm - a.free_mask) & a.free_mask;
                }
                s_totv = tv;
            }
        }
        if (tid < NCo) s_fo[tid] = a.fg_out[(uint64_t)b_out * fgs_o + tid];
        if (!FWD && has_in && tid < NCi) s_fi[tid] = a.fg_in[(uint64_t)b_in0 * fgs_i + tid];
        __syncthreads();

        // ---- per-lane hoisted column quantities
        T cbase[NS][V], gout[NS][V], gin0[NS][V], colacc[NS][V];
        const T ctot = has_in ? cC * s_totv : T(0);
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const bool ok = ccol[s] < R;
                cbase[s][v] = (has_in && ok) ? cB * s_colv[ccol[s] + v] + ctot : T(0);
                gout[s][v] = a.fg_out[(uint64_t)b_out * fgs_o + NCo + a2o[s][v]];
                gin0[s][v] = (!FWD && has_in) ? a.fg_in[(uint64_t)b_in0 * fgs_i + NCi + a2i[s][v]] : T(1);
                colacc[s][v] = T(0);
            }
        T totacc = T(0);
        const uint64_t out_base = (uint64_t)b_out * BS;

#pragma unroll 2
        for (uint32_t it = 0; it < n_iter; ++it) {
            const uint32_t ridx = it * ngrp + grp;
            const bool rvalid = ridx < R;
            const uint32_t R1 = rvalid ? s_rord[ridx] : 0u;
            const uint32_t a1o = s_al_o[R1];
            T x[NS][V];
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int v = 0; v < V; ++v) x[s][v] = T(0);
            T rowterm = T(0);
            if (has_in) {
                rowterm = cB * s_rowv[R1];
                const uint32_t a1i = s_al_i[R1];
                uint32_t subm = 0;
                for (uint32_t k = 0; k < nred; ++k) {
                    const uint64_t bk = FWD ? (uint64_t)(b_in0 | subm) : (uint64_t)(b_in0 | (k << a.n_sh));
                    subm = (subm - a.free_mask) & a.free_mask;
                    const T *src = a.in + bk * BS + (uint64_t)R1 * R;
                    T fk = T(1);
                    if (!FWD) fk = (k == 0) ? s_fi[a1i] : a.fg_in[bk * fgs_i + a1i];
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        if (rvalid && ccol[s] < R) {
                            T ld[V];
                            load_vec<T, V>(ld, src + ccol[s]);
#pragma unroll
                            for (int v = 0; v < V; ++v) {
                                if (FWD) {
                                    x[s][v] += ld[v];
                                } else {
                                    const T gk = (k == 0) ? gin0[s][v] : a.fg_in[bk * fgs_i + NCi + a2i[s][v]];
                                    x[s][v] += ld[v] * (fk * gk);
                                }
                            }
                        }
                    }
                }
            }
            // ---- produce the output row
            T rs = T(0);
            const T fo = s_fo[a1o];
            if (FWD && a1o != cur_cls) {   // row class changed: park the posterior accumulators
                if (cur_cls != 0xffffffffu) {
#pragma unroll
                    for (int s = 0; s < NS; ++s)
                        if (ccol[s] < R) {
#pragma unroll
                            for (int v = 0; v < V; ++v) {
                                s_q[(grp * NCo + cur_cls) * Rp + ccol[s] + v] += qc[s][v];
                                qc[s][v] = T(0);
                            }
                        }
                }
                cur_cls = a1o;
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (rvalid && ccol[s] < R) {
                    T o[V], w[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const T base = has_in ? (cA * x[s][v] + rowterm + cbase[s][v]) : T(1);
                        if (FWD) {
                            o[v] = (fo * gout[s][v]) * base;
                            w[v] = o[v];
                        } else {
                            o[v] = (a1o != 0 && a2o[s][v] != 0) ? base : T(0);
                            w[v] = o[v] * (fo * gout[s][v]);
                            sacc += o[v];
                        }
                        rs += w[v];
                        colacc[s][v] += w[v];
                    }
                    store_vec<T, V>(a.out + out_base + (uint64_t)R1 * R + ccol[s], o);
                    if (FWD) {
                        T bv[V];
                        load_vec<T, V>(bv, a.beta + out_base + (uint64_t)R1 * R + ccol[s]);
#pragma unroll
                        for (int v = 0; v < V; ++v) qc[s][v] += o[v] * bv[v];
                    }
                }
            }
            for (uint32_t o = G >> 1; o; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
            if (gl == 0 && rvalid) {
                a.out[out_base + o_row + R1] = rs;
                totacc += rs;
            }
        }
        // ---- item epilogue: column sums across the row groups, block total
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (ccol[s] < R) {
#pragma unroll
                for (int v = 0; v < V; ++v) s_colred[grp * Rp + ccol[s] + v] = colacc[s][v];
            }
        if (gl == 0) s_totred[grp] = totacc;
        __syncthreads();
        for (uint32_t t = tid; t < R; t += blockDim.x) {
            T cs = T(0);
            for (uint32_t g = 0; g < ngrp; ++g) cs += s_colred[g * Rp + t];
            a.out[out_base + o_col + t] = cs;
        }
        if (tid == 0) {
            T ts = T(0);
            for (uint32_t g = 0; g < ngrp; ++g) ts += s_totred[g];
            a.out[out_base + o_tot] = ts;
            cta_sum += ts;
        }
        // the next item's first __syncthreads orders these reads before the smem is rewritten
    }

    // ---- kernel epilogue: per-CTA partials, then the last CTA reduces them in a fixed order
    T *mypart = a.partial + (uint64_t)blockIdx.x * kPartialStride;
    if (FWD) {
        if (cur_cls != 0xffffffffu) {
#pragma unroll
            for (int s = 0; s < NS; ++s)
                if (ccol[s] < R) {
#pragma unroll
                    for (int v = 0; v < V; ++v) s_q[(grp * NCo + cur_cls) * Rp + ccol[s] + v] += qc[s][v];
                }
        }
        __syncthreads();
        for (uint32_t idx = tid; idx < NCo * R; idx += blockDim.x) {   // fold the row groups into group 0
            const uint32_t cls = idx / R, t = idx % R;
            T v = T(0);
            for (uint32_t g = 0; g < ngrp; ++g) v += s_q[(g * NCo + cls) * Rp + t];
            s_q[cls * Rp + t] = v;   // group 0 slot (cls*Rp+t == (0*NCo+cls)*Rp+t)
        }
        __syncthreads();
        for (uint32_t p = tid; p < NCo * NCo; p += blockDim.x) {       // W[i][j] = sum over columns of class j
            const uint32_t i = p / NCo, j = p % NCo;
            T w = T(0);
            for (uint32_t t = 0; t < R; ++t)
                if (s_al_o[t] == j) w += s_q[i * Rp + t];
            mypart[p] = w;
        }
        if (tid == 0) mypart[kMaxClasses * kMaxClasses] = cta_sum;
    } else {
        // block reduction of the beta sum
        for (uint32_t o = 16; o; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
        __syncthreads();
        if (lane == 0) s_totred[warp] = sacc;
        __syncthreads();
        if (tid == 0) {
            T ts = T(0);
            for (uint32_t w = 0; w < nwarps; ++w) ts += s_totred[w];
            mypart[kMaxClasses * kMaxClasses] = ts;
        }
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
    for (uint32_t p = warp; p < npair + 1; p += nwarps) {
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
            // ordered allele pair -> genotype index, reference src/genotypehmm.cpp:524-528
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

// The chain's last backward column: beta = 1 on every state whose two alleles are present (reference
// src/genotypehmm.cpp:115-128 fills it with ones; the states with a missing allele carry no emission and are kept at
// 0 as everywhere else), side sums of beta * E.  The values of every block are the same 0/1 pattern, only the side
// record depends on b, so this is a pure fill: each thread keeps its column pattern in registers and streams rows.
// A chain whose columns are tens of GB re-creates this column several times per sweep (schedule.cpp, exact planner)
// — the generic step kernel took 16 ms for 33 GB (per-item latency), this one is write-bound.
// row[R1] = F[a(R1)] * SG, col[R2] = G[a(R2)] * SF, total = SF * SG with SF = sum_j n_j F[j], SG = sum_j n_j G[j]
// (n_j = haplotypes carrying allele j), summed in allele order: the same bits whatever the grid.
template <typename T, int V>
__global__ void __launch_bounds__(256) init_kernel(const StepArgs<T> a) {
    __shared__ uint8_t s_al[1024 + 8];
    __shared__ uint32_t s_cnt[kMaxClasses];
    const uint32_t R = a.R, NC = a.nA_out + 1, fgs = 2 * NC, RR = R * R;
    const uint32_t tid = threadIdx.x;
    for (uint32_t t = tid; t < R; t += blockDim.x) s_al[t] = a.al_out[t];
    __syncthreads();
    if (tid < NC) {
        uint32_t n = 0;
        for (uint32_t r = 0; r < R; ++r) n += s_al[r] == tid ? 1u : 0u;
        s_cnt[tid] = n;
    }
    __syncthreads();
    const uint32_t nvec = R / V;
    const bool fast = nvec <= blockDim.x;
    const uint32_t cchunk = fast ? tid % nvec : 0u, r0 = fast ? tid / nvec : 0u, rstep = fast ? blockDim.x / nvec : 1u;
    T pat[V], zero[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        pat[v] = s_al[cchunk * V + v] != 0 ? T(1) : T(0);
        zero[v] = T(0);
    }
    const uint64_t nb = (uint64_t)1 << a.u_out;
    for (uint64_t b = blockIdx.x; b < nb; b += gridDim.x) {
        T *blk = a.out + b * a.BS;
        if (fast) {
            if (r0 < rstep)
                for (uint32_t R1 = r0; R1 < R; R1 += rstep)
                    store_vec<T, V>(blk + (uint64_t)R1 * R + cchunk * V, s_al[R1] != 0 ? pat : zero);
        } else {
            for (uint32_t q = tid; q < R * nvec; q += blockDim.x) {
                const uint32_t R1 = q / nvec, cc = q % nvec;
                T o[V];
#pragma unroll
                for (int v = 0; v < V; ++v) o[v] = (s_al[R1] != 0 && s_al[cc * V + v] != 0) ? T(1) : T(0);
                store_vec<T, V>(blk + (uint64_t)R1 * R + cc * V, o);
            }
        }
        const T *fg = a.fg_out + b * fgs;
        T SF = T(0), SG = T(0);
        for (uint32_t j = 1; j < NC; ++j) {
            SF += (T)s_cnt[j] * fg[j];
            SG += (T)s_cnt[j] * fg[NC + j];
        }
        for (uint32_t t = tid; t < R; t += blockDim.x) {
            blk[RR + t] = fg[s_al[t]] * SG;
            blk[RR + R + t] = fg[NC + s_al[t]] * SF;
        }
        if (tid == 0) blk[RR + 2 * R] = SF * SG;
    }
    if (blockIdx.x == 0 && tid == 0) {
        const double nv = (double)(R - s_cnt[0]);
        *a.out_scale = nv * nv * (double)nb;
    }
}


}  // namespace gg
