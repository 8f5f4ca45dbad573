// Batched edit distance on the device: gg_edit_distance_batch (include/giggles_b200.h).
//
// Replaces edit_distance(s, t, maxdiff) of the reference (giggles/align.pyx:15-107), which the allele detection by
// realignment calls once per (read, variant, allele) in "edit" mode (giggles/variants.py:50-51,238-239: the aligner IS
// edit_distance, called as aligner(query, allele), i.e. maxdiff = -1) — SURVEY §8 f4.
//
//   maxdiff = -1  unit-cost Levenshtein distance.  Bit-parallel (Myers 1999, in Hyyro's block formulation): the DP
//                 column is kept as vertical +1/-1 deltas in 64-bit words, one text character costs ~17 integer
//                 instructions per 64 pattern positions instead of 64 min/add cells.
//                   short_kernel   pattern <= 64 after stripping: one THREAD per pair, one word.
//                   long_kernel    longer patterns: one WARP per pair, lane k owns pattern word k; the horizontal
//                                  delta that word k hands to word k+1 is passed by shuffle while the lanes work on
//                                  staggered text columns (lane k on column step-k), so all 32 words advance every
//                                  step; patterns of more than 2048 characters are strip-mined in groups of 32 words
//                                  with the inter-group deltas parked in a per-pair byte array.
//   maxdiff >= 0  the reference's banded loop, literally (dp_kernel): its result above maxdiff is "some value greater
//                 than maxdiff" that depends on the band bookkeeping (stale cells outside the band, the early break),
//                 so bit-exactness needs the same loop, one thread per pair.  Not on the genotyping path.
//
// Integer work throughout; results are bit-exact against the reference's Cython module (tests/test_editdist.py).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/giggles_b200.h"

namespace gg_ed {

constexpr int kCodes = 16;            // distinct byte values a batch may contain on the bit-parallel path (+ "other")
constexpr int kShortThreads = 128;
constexpr int kLongWarps = 4;

struct Pair {
    uint64_t s_off, t_off;
    uint32_t s_len, t_len;
    int32_t maxdiff;
    uint32_t out_index;
    uint64_t scratch_off;             // long_kernel: inter-group deltas; dp_kernel: the costs[] array (ints)
};

__constant__ uint8_t c_code[256];     // byte -> code 1..kCodes-1 (0 = does not occur in the batch)

// identical prefixes and suffixes do not change the distance (align.pyx:48-58)
__device__ __forceinline__ void strip(const uint8_t *&s, uint32_t &m, const uint8_t *&t, uint32_t &n) {
    while (m > 0 && n > 0 && s[0] == t[0]) { ++s; ++t; --m; --n; }
    while (m > 0 && n > 0 && s[m - 1] == t[n - 1]) { --m; --n; }
}

// One 64-position block of the bit-parallel recurrence.  Pv / Mv: vertical +1 / -1 deltas; hin / hout: horizontal
// delta entering at the top / leaving at bit `last` of the block.
__device__ __forceinline__ int advance_block(uint64_t &Pv, uint64_t &Mv, uint64_t Eq, int hin, uint64_t last) {
    const uint64_t Xv = Eq | Mv;
    if (hin < 0) Eq |= 1ull;
    const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    uint64_t Ph = Mv | ~(Xh | Pv);
    uint64_t Mh = Pv & Xh;
    int hout = 0;
    if (Ph & last) hout = 1;
    else if (Mh & last) hout = -1;
    Ph <<= 1;
    Mh <<= 1;
    if (hin < 0) Mh |= 1ull;
    else if (hin > 0) Ph |= 1ull;
    Pv = Mh | ~(Xv | Ph);
    Mv = Ph & Xv;
    return hout;
}

// ---- pattern of at most 64 characters: one thread per pair ----------------------------------------------------
__global__ void __launch_bounds__(kShortThreads)
short_kernel(const uint8_t *__restrict__ pool, const Pair *__restrict__ pairs, uint32_t n_pairs, int32_t *__restrict__ out) {
    __shared__ uint64_t s_peq[kCodes][kShortThreads];
    const uint32_t i = blockIdx.x * kShortThreads + threadIdx.x;
    if (i >= n_pairs) return;
    const Pair p = pairs[i];
    const uint8_t *s = pool + p.s_off, *t = pool + p.t_off;
    uint32_t m = p.s_len, n = p.t_len;
    strip(s, m, t, n);
    if (m > n) {                      // the distance is symmetric: the shorter string is the pattern
        const uint8_t *x = s; s = t; t = x;
        const uint32_t y = m; m = n; n = y;
    }
    if (m == 0) { out[p.out_index] = (int32_t)n; return; }
#pragma unroll
    for (int c = 0; c < kCodes; ++c) s_peq[c][threadIdx.x] = 0;
    for (uint32_t k = 0; k < m; ++k) s_peq[c_code[s[k]]][threadIdx.x] |= 1ull << k;
    s_peq[0][threadIdx.x] = 0;        // code 0 = "not in the batch's alphabet" never matches
    uint64_t Pv = ~0ull, Mv = 0;
    const uint64_t last = 1ull << (m - 1);
    int32_t score = (int32_t)m;
    for (uint32_t j = 0; j < n; ++j)
        score += advance_block(Pv, Mv, s_peq[c_code[t[j]]][threadIdx.x], 1, last);     // top row D[0][j] = j: hin = +1
    out[p.out_index] = score;
}

// ---- longer patterns: one warp per pair, 32 words per group, staggered columns ----------------------------------
__global__ void __launch_bounds__(32 * kLongWarps)
long_kernel(const uint8_t *__restrict__ pool, const Pair *__restrict__ pairs, uint32_t n_pairs, int8_t *__restrict__ scratch,
            int32_t *__restrict__ out) {
    __shared__ uint64_t s_peq[kLongWarps][kCodes][32];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t i = blockIdx.x * kLongWarps + warp;
    if (i >= n_pairs) return;
    const Pair p = pairs[i];
    const uint8_t *s = pool + p.s_off, *t = pool + p.t_off;
    uint32_t m = p.s_len, n = p.t_len;
    {   // identical prefix / suffix, 32 characters per ballot (flanking overhangs are identical by construction)
        const uint32_t lim = min(m, n);
        uint32_t pre = 0;
        for (;;) {
            const uint32_t k = pre + lane;
            const unsigned eq = __ballot_sync(0xffffffffu, k < lim && s[k] == t[k]);
            if (eq == 0xffffffffu) { pre += 32; continue; }
            pre += __ffs(~eq) - 1;
            break;
        }
        s += pre; t += pre; m -= pre; n -= pre;
        const uint32_t lim2 = min(m, n);
        uint32_t suf = 0;
        for (;;) {
            const uint32_t k = suf + lane;
            const unsigned eq = __ballot_sync(0xffffffffu, k < lim2 && s[m - 1 - k] == t[n - 1 - k]);
            if (eq == 0xffffffffu) { suf += 32; continue; }
            suf += __ffs(~eq) - 1;
            break;
        }
        m -= suf; n -= suf;
    }
    if (m > n) {
        const uint8_t *x = s; s = t; t = x;
        const uint32_t y = m; m = n; n = y;
    }
    if (m == 0) {
        if (lane == 0) out[p.out_index] = (int32_t)n;
        return;
    }
    int8_t *hbuf = scratch + p.scratch_off;          // [n] horizontal deltas between word groups
    const uint32_t W = (m + 63) / 64, n_groups = (W + 31) / 32;
    int32_t score = (int32_t)m;
    for (uint32_t g = 0; g < n_groups; ++g) {
        const uint32_t word = g * 32 + lane;
        const bool live = word < W;
        const uint32_t w_in_group = min(32u, W - g * 32);
        // match masks of this lane's word
        __syncwarp();
#pragma unroll
        for (int c = 0; c < kCodes; ++c) s_peq[warp][c][lane] = 0;
        if (live) {
            const uint32_t base = word * 64, cnt = min(64u, m - base);
            for (uint32_t k = 0; k < cnt; ++k) s_peq[warp][c_code[s[base + k]]][lane] |= 1ull << k;
            s_peq[warp][0][lane] = 0;
        }
        __syncwarp();
        const bool is_last_word = word == W - 1;
        const uint64_t last = 1ull << (is_last_word ? (m - 1) % 64 : 63);
        uint64_t Pv = ~0ull, Mv = 0;
        int hout = 0;
        const uint32_t steps = n + w_in_group - 1;
        for (uint32_t step = 0; step < steps; ++step) {
            const int from_below = __shfl_up_sync(0xffffffffu, hout, 1);      // word k-1's delta for the column this lane takes now
            const int64_t j = (int64_t)step - lane;
            if (live && j >= 0 && j < (int64_t)n) {
                const int hin = lane == 0 ? (g == 0 ? 1 : (int)hbuf[j]) : from_below;
                hout = advance_block(Pv, Mv, s_peq[warp][c_code[t[j]]][lane], hin, last);
                if (is_last_word) score += hout;
                else if (lane == 31) hbuf[j] = (int8_t)hout;                  // read by the next group's lane 0
            }
        }
    }
    // the lane that owns the last word holds the score
    const uint32_t owner = (W - 1) % 32;
    score = __shfl_sync(0xffffffffu, score, owner);
    if (lane == 0) out[p.out_index] = score;
}

// ---- the reference's loops, literally (align.pyx:60-107): banded mode and the fallback for wide alphabets ---------
__global__ void __launch_bounds__(64)
dp_kernel(const uint8_t *__restrict__ pool, const Pair *__restrict__ pairs, uint32_t n_pairs, int32_t *__restrict__ scratch,
          int32_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const Pair p = pairs[i];
    const uint8_t *sv = pool + p.s_off, *tv = pool + p.t_off;
    int m = (int)p.s_len, n = (int)p.t_len;
    const int e = p.maxdiff;
    {
        uint32_t mu = (uint32_t)m, nu = (uint32_t)n;
        strip(sv, mu, tv, nu);
        m = (int)mu; n = (int)nu;
    }
    int *costs = scratch + p.scratch_off;            // [m + 1] (sized for the unstripped m)
    for (int k = 0; k <= m; ++k) costs[k] = k;
    if (e < 0) {
        for (int j = 1; j <= n; ++j) {
            int prev = costs[0];
            costs[0] += 1;
            for (int k = 1; k <= m; ++k) {
                const int match = sv[k - 1] == tv[j - 1];
                const int c = min(min(prev + 1 - match, costs[k] + 1), costs[k - 1] + 1);
                prev = costs[k];
                costs[k] = c;
            }
        }
        out[p.out_index] = costs[m];
        return;
    }
    int smallest = 0;
    for (int j = 1; j <= n; ++j) {
        const int stop = min(j + e + 1, m + 1);
        int start, prev;
        if (j <= e) {
            prev = costs[0];
            costs[0] += 1;
            smallest = costs[0];
            start = 1;
        } else {
            start = j - e;
            prev = costs[start - 1];
            smallest = e + 1;
        }
        for (int k = start; k < stop; ++k) {
            const int match = sv[k - 1] == tv[j - 1];
            const int c = min(min(prev + 1 - match, costs[k] + 1), costs[k - 1] + 1);
            prev = costs[k];
            costs[k] = c;
            smallest = min(smallest, c);
        }
        if (smallest > e) break;
    }
    out[p.out_index] = smallest > e ? smallest : costs[m];
}

struct DevMem {
    void *p = nullptr;
    ~DevMem() { if (p) cudaFree(p); }
    void alloc(size_t n) {
        if (cudaMalloc(&p, std::max<size_t>(n, 16)) != cudaSuccess) {
            cudaGetLastError();
            throw std::runtime_error("gg_edit_distance_batch: out of device memory");
        }
    }
};

#define ED_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

}  // namespace gg_ed

extern "C" int gg_edit_distance_batch(const uint8_t *pool, uint64_t pool_bytes, const uint64_t *s_off, const uint32_t *s_len,
                                      const uint64_t *t_off, const uint32_t *t_len, const int32_t *maxdiff, uint32_t n_pairs,
                                      int32_t *out, int32_t device, char *err, size_t errlen) {
    using namespace gg_ed;
    auto fail = [&](int rc, const std::string &m) {
        if (err && errlen) {
            std::strncpy(err, m.c_str(), errlen - 1);
            err[errlen - 1] = 0;
        }
        return rc;
    };
    try {
        if (n_pairs == 0) return GG_OK;
        if (!s_off || !s_len || !t_off || !t_len || !out || (!pool && pool_bytes)) return fail(GG_ERR_INVALID, "gg_edit_distance_batch: null argument");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            return fail(GG_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
        }
        if (device >= 0) ED_CUDA(cudaSetDevice(device));
        // alphabet of the batch: the bit-parallel kernels index their match masks by a small code per byte value
        bool seen[256] = {false};
        std::vector<Pair> shortp, longp, dpp;
        uint64_t long_scratch = 0, dp_scratch = 0;
        for (uint32_t i = 0; i < n_pairs; ++i) {
            if (s_off[i] + s_len[i] > pool_bytes || t_off[i] + t_len[i] > pool_bytes)
                return fail(GG_ERR_INVALID, "gg_edit_distance_batch: pair " + std::to_string(i) + " reaches outside the pool");
            const int e = maxdiff ? maxdiff[i] : -1;
            const int64_t m = s_len[i], n = t_len[i];
            // align.pyx:38-40: lengths too different for the band -> the difference, nothing else is looked at
            if (e != -1 && std::llabs(m - n) > e) { out[i] = (int32_t)std::llabs(m - n); continue; }
            if (e < -1) return fail(GG_ERR_INVALID, "gg_edit_distance_batch: maxdiff must be -1 or >= 0");
            Pair p{s_off[i], t_off[i], s_len[i], t_len[i], e, i, 0};
            if (e >= 0) {
                p.scratch_off = dp_scratch;
                dp_scratch += (uint64_t)s_len[i] + 1;
                dpp.push_back(p);
            } else if (std::min(m, n) <= 64) {
                shortp.push_back(p);
            } else {
                p.scratch_off = long_scratch;
                long_scratch += std::max(s_len[i], t_len[i]);
                longp.push_back(p);
            }
        }
        for (const auto *v : {&shortp, &longp})
            for (const Pair &p : *v) {
                for (uint32_t k = 0; k < p.s_len; ++k) seen[pool[p.s_off + k]] = true;
                for (uint32_t k = 0; k < p.t_len; ++k) seen[pool[p.t_off + k]] = true;
            }
        uint8_t code[256] = {0};
        int n_codes = 1;
        bool wide = false;
        for (int b = 0; b < 256; ++b)
            if (seen[b]) {
                if (n_codes >= kCodes) { wide = true; break; }
                code[b] = (uint8_t)n_codes++;
            }
        if (wide) {                       // more than 15 distinct byte values: everything through the plain DP
            for (auto *v : {&shortp, &longp})
                for (Pair p : *v) {
                    p.scratch_off = dp_scratch;
                    dp_scratch += (uint64_t)p.s_len + 1;
                    dpp.push_back(p);
                }
            shortp.clear();
            longp.clear();
        }
        if (shortp.empty() && longp.empty() && dpp.empty()) return GG_OK;
        cudaStream_t st;
        ED_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
        DevMem d_pool, d_pairs, d_out, d_scr8, d_scr32;
        std::vector<Pair> all;
        all.insert(all.end(), shortp.begin(), shortp.end());
        all.insert(all.end(), longp.begin(), longp.end());
        all.insert(all.end(), dpp.begin(), dpp.end());
        d_pool.alloc(pool_bytes);
        d_pairs.alloc(all.size() * sizeof(Pair));
        d_out.alloc((size_t)n_pairs * sizeof(int32_t));
        d_scr8.alloc(long_scratch);
        d_scr32.alloc(dp_scratch * sizeof(int32_t));
        ED_CUDA(cudaMemcpyAsync(d_pool.p, pool, pool_bytes, cudaMemcpyHostToDevice, st));
        ED_CUDA(cudaMemcpyAsync(d_pairs.p, all.data(), all.size() * sizeof(Pair), cudaMemcpyHostToDevice, st));
        ED_CUDA(cudaMemcpyToSymbolAsync(c_code, code, 256, 0, cudaMemcpyHostToDevice, st));
        // pairs settled on the host (length difference beyond the band) keep their value: start from the caller's array
        ED_CUDA(cudaMemcpyAsync(d_out.p, out, (size_t)n_pairs * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        const Pair *dp = static_cast<const Pair *>(d_pairs.p);
        const uint8_t *dpool = static_cast<const uint8_t *>(d_pool.p);
        int32_t *dout = static_cast<int32_t *>(d_out.p);
        if (!shortp.empty()) {
            const uint32_t n = (uint32_t)shortp.size();
            short_kernel<<<(n + kShortThreads - 1) / kShortThreads, kShortThreads, 0, st>>>(dpool, dp, n, dout);
            ED_CUDA(cudaGetLastError());
        }
        if (!longp.empty()) {
            const uint32_t n = (uint32_t)longp.size();
            long_kernel<<<(n + kLongWarps - 1) / kLongWarps, 32 * kLongWarps, 0, st>>>(dpool, dp + shortp.size(), n,
                                                                                     static_cast<int8_t *>(d_scr8.p), dout);
            ED_CUDA(cudaGetLastError());
        }
        if (!dpp.empty()) {
            const uint32_t n = (uint32_t)dpp.size();
            dp_kernel<<<(n + 63) / 64, 64, 0, st>>>(dpool, dp + shortp.size() + longp.size(), n, static_cast<int32_t *>(d_scr32.p), dout);
            ED_CUDA(cudaGetLastError());
        }
        ED_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)n_pairs * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        ED_CUDA(cudaStreamSynchronize(st));
        return GG_OK;
    } catch (const std::exception &e) {
        return fail(GG_ERR_CUDA, e.what());
    }
}
