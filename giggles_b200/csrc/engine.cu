// Device engine + compute half of the C ABI (include/giggles_b200.h).
//
// gg_hmm_create  = plan on the host (plan.cpp), checkpoint schedule (schedule.cpp), H2D of the tables
// gg_hmm_sweep   = replay the schedule: emission tables, backward steps, forward steps (+ posterior)
// gg_hmm_fetch   = D2H of the genotype tables
// There is no CPU path: without a CUDA device every entry point fails with GG_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/giggles_b200.h"
#include "hmm_kernels.cuh"
#include "hmm_tma_kernel.cuh"
#include "hmm_wide_kernel.cuh"
#include "hmm_colown_kernel.cuh"
#include "hmm_chain_kernel.cuh"
#include "hmm_warp_kernel.cuh"
#include "hmm_wave_kernel.cuh"
#include "plan.hpp"
#include "result.hpp"
#include "schedule.hpp"

namespace gg {

struct CudaError : std::runtime_error {
    int status;
    CudaError(int st, const std::string &m) : std::runtime_error(m), status(st) {}
};

#define GG_CUDA(call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            int st__ = (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver) ? GG_ERR_NO_DEVICE    \
                       : (e__ == cudaErrorMemoryAllocation ? GG_ERR_NOMEM : GG_ERR_CUDA);                     \
            throw CudaError(st__, std::string(#call) + ": " + cudaGetErrorString(e__));                       \
        }                                                                                                     \
    } while (0)

// A typed view into the handle's table arena (one allocation per handle, cached across handles: a dozen
// cudaMalloc/cudaFree pairs per call cost more than the sweep of a small chromosome).
template <typename U>
struct DevBuf {
    U *p = nullptr;
    size_t n = 0;
    void bind(void *base, size_t count) {
        p = count ? static_cast<U *>(base) : nullptr;
        n = count;
    }
    void upload(const std::vector<U> &h, cudaStream_t st) {
        if (h.size() != n) throw std::runtime_error("internal: table size mismatch");
        if (!h.empty()) GG_CUDA(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(U), cudaMemcpyHostToDevice, st));
    }
    void release() {
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

// Retired arenas are kept for the next gg_hmm_create (cudaMalloc/cudaFree of a >100 GB arena costs more than a sweep
// of a small window, and cudaFree synchronises the device, which would serialise chains that run concurrently from
// several host threads).  Per key (2*device = state arena, 2*device+1 = table arena): at most ONE large arena — a
// request it does not fit frees it first, there is no room for two — and a small pool of arenas of up to 4 GB for
// concurrent small chains (all-haplotagged chromosomes).  gg_arena_release() frees everything.
struct ArenaCache {
    static constexpr size_t kSmallArena = (size_t)4 << 30, kSmallPoolBytes = (size_t)24 << 30, kSmallPoolCount = 64;
    struct Ent { unsigned char *p; size_t n; };
    std::mutex mu;
    std::map<int, std::vector<Ent>> pool;
    static bool fits(size_t n, size_t need) { return n >= need && n <= need + need / 2 + (64u << 20); }
    unsigned char *take(int key, size_t need, size_t &got) {
        std::lock_guard<std::mutex> g(mu);
        std::vector<Ent> &v = pool[key];
        size_t best = v.size();
        for (size_t i = 0; i < v.size(); ++i)
            if (fits(v[i].n, need) && (best == v.size() || v[i].n < v[best].n)) best = i;
        if (best != v.size()) {
            Ent e = v[best];
            v.erase(v.begin() + best);
            got = e.n;
            return e.p;
        }
        if (need > kSmallArena) {          // a large request needs the room: drop whatever is cached under this key
            for (Ent &e : v) cudaFree(e.p);
            v.clear();
        }
        return nullptr;
    }
    // what default_budget may count as free again: the large cached arena (or the largest cached one)
    size_t cached(int key) {
        std::lock_guard<std::mutex> g(mu);
        size_t m = 0;
        for (const Ent &e : pool[key]) m = std::max(m, e.n);
        return m;
    }
    void give(int key, unsigned char *p, size_t n) {
        std::lock_guard<std::mutex> g(mu);
        std::vector<Ent> &v = pool[key];
        if (n > kSmallArena) {             // keep one large arena only
            for (size_t i = 0; i < v.size();)
                if (v[i].n > kSmallArena) { cudaFree(v[i].p); v.erase(v.begin() + i); } else ++i;
        }
        v.push_back({p, n});
        size_t small_bytes = 0, small_count = 0;
        for (const Ent &e : v)
            if (e.n <= kSmallArena) { small_bytes += e.n; ++small_count; }
        for (size_t i = 0; i < v.size() && (small_bytes > kSmallPoolBytes || small_count > kSmallPoolCount);) {
            if (v[i].n <= kSmallArena) {   // oldest first
                small_bytes -= v[i].n; --small_count;
                cudaFree(v[i].p);
                v.erase(v.begin() + i);
            } else ++i;
        }
    }
    void clear() {
        std::lock_guard<std::mutex> g(mu);
        for (auto &kv : pool)
            for (Ent &e : kv.second) cudaFree(e.p);
        pool.clear();
    }
};
static ArenaCache g_arena_cache;

// One state arena per device shared by every handle created with GG_OPT_SHARED_ARENA: such a handle holds its tables
// only and borrows the arena for the duration of a sweep (sweeps of these handles are serialised by the lock).  This
// is how many chromosomes stay planned and resident on one device while the state memory - which a single chain of
// HBM-filling columns needs whole - exists once.  Grows on demand, never shrinks until gg_arena_release().
struct SharedArena {
    std::mutex mu;
    unsigned char *p = nullptr;
    size_t n = 0;
};
static std::mutex g_shared_mu;
static std::map<int, std::unique_ptr<SharedArena>> g_shared_arenas;
static SharedArena &shared_arena_of(int device) {
    std::lock_guard<std::mutex> g(g_shared_mu);
    std::unique_ptr<SharedArena> &sa = g_shared_arenas[device];
    if (!sa) sa.reset(new SharedArena());
    return *sa;
}
static std::atomic<uint64_t> g_fp64_reruns{0};      // fp32 range guard: chains rerun on fp64 state (process-wide)
static bool all_finite(const gg_result *r) {
    if (r)
        for (double v : r->values)
            if (!std::isfinite(v)) return false;
    return true;
}
static std::mutex g_budget_mu;
static std::map<int, uint64_t> g_last_queried_budget;   // per device: what the last cudaMemGetInfo-based budget was

// cudaMalloc that gives the cached arenas back to the driver and tries once more before reporting out-of-memory
static cudaError_t device_alloc(unsigned char **p, size_t n) {
    cudaError_t e = cudaMalloc(p, n);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        g_arena_cache.clear();
        e = cudaMalloc(p, n);
    }
    return e;
}

// cudaFuncSetAttribute is per device.  Runs `configure` exactly once per (device, kernel family) and holds every other
// caller back until it has finished: a thread that merely saw "already started" could otherwise launch (or ask for an
// occupancy figure) with more than 48 KB of dynamic shared memory before the attribute is set (cold process, several
// host threads: gg_hmm_run_batch workers, capi.hmm_run_many).  A failed configuration is retried by the next caller.
template <typename F>
static void configure_once(int device, int family, F &&configure) {
    static std::mutex mu;
    static std::set<std::pair<int, int>> done;
    std::lock_guard<std::mutex> g(mu);
    if (done.count({device, family})) return;
    configure();
    done.insert({device, family});
}

struct LaunchShape {
    int V, NS;
    uint32_t G, nwarps_fwd, nwarps_bwd;
};

}  // namespace gg

using namespace gg;

struct gg_hmm {
    Plan plan;
    Schedule sched;
    bool fp64 = false;
    bool shared_arena = false;   // GG_OPT_SHARED_ARENA: the state arena is borrowed per sweep
    int kernel_variant = 0;   // 0 = auto, 1 = generic step kernel only, 2 = pipelined, never two CTAs per SM, 3 = pipelined, never 16 consumer warps, 4 = no chain kernel, 5 = no warp-per-block kernel (nor wave runs), 7 = no wave runs (small panels column by column), 8 = no wide-panel kernel (R > 128 on the row-tiled pipelined kernel), 9 = no column-owning kernel for R <= 128
    int device = 0;
    int num_sms = 0;
    size_t elem = 4;
    cudaStream_t stream = nullptr;
    std::vector<uint64_t> colb, fgb, fg_prefix_elems;
    std::vector<ColDesc> host_cols;
    DevBuf<ColDesc> d_cols;
    DevBuf<uint8_t> d_em_side, d_allele1;
    DevBuf<double> d_em_prob, d_scale_a, d_scale_b, d_result, d_emax;
    DevBuf<uint16_t> d_row_order;
    DevBuf<unsigned char> d_partial;
    unsigned char *arena = nullptr;      // state arena (columns, F/G tables)
    size_t arena_bytes = 0;
    unsigned char *tables = nullptr;     // table arena (everything else on the device)
    size_t tables_bytes = 0;
    DevBuf<unsigned int> d_counter;
    // runs of consecutive single-block columns fused into one chain-kernel launch each
    struct Run { bool fwd; uint32_t first, n, max_classes; uint64_t in_off; int64_t in_col; bool wave; };
    // wave runs (hmm_wave_kernel.cuh): consecutive small-panel steps of one direction in one cooperative launch
    std::vector<Op> wave_ops;                 // the schedule ops behind the wave steps, in order (runs[].first indexes it)
    std::vector<uint64_t> wave_w_off;         // per wave op: offset of its posterior partials (floats)
    uint64_t wave_w_floats = 0;
    uint32_t wave_grid = 0;
    const void *wave_built_for = nullptr;     // state arena the device copy of the step arguments points into
    DevBuf<unsigned char> d_wave_steps, d_wave_w, d_wave_sum;
    DevBuf<unsigned int> d_wave_bar;
    std::vector<Run> runs;
    std::vector<ChainStep> chain_steps;
    DevBuf<ChainStep> d_chain_steps;
    std::vector<Op> exec_ops;   // the schedule after fusion (what gg_hmm_sweep replays)
    uint32_t max_grid = 0;
    LaunchShape shape{};
    std::map<std::tuple<const void *, int, size_t>, int> occ_cache;
    gg_stats stats{};
    std::vector<cudaEvent_t> ev;   // timing events (timed sweeps only)
    std::vector<float> op_ms;      // per-op durations of the last timed sweep
    ~gg_hmm() {
        // work queued on the stream (a sweep that failed half-way, a caller that destroys after an error) may still
        // use the arenas: drain it before another handle can take them from the cache; after a device error they are
        // freed instead of being cached
        const bool clean = !stream || cudaStreamSynchronize(stream) == cudaSuccess;
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
        if (arena && !shared_arena) { if (clean) g_arena_cache.give(2 * device, arena, arena_bytes); else cudaFree(arena); }
        if (tables) { if (clean) g_arena_cache.give(2 * device + 1, tables, tables_bytes); else cudaFree(tables); }
        if (!clean) cudaGetLastError();
    }
};

namespace gg {

static void set_err(char *err, size_t errlen, const std::string &m) {
    if (err && errlen) {
        std::strncpy(err, m.c_str(), errlen - 1);
        err[errlen - 1] = 0;
    }
}

static uint32_t pow2ceil(uint32_t x) {
    uint32_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

template <typename T>
static LaunchShape choose_shape(uint32_t R) {
    LaunchShape s{};
    const int maxV = sizeof(T) == 4 ? 4 : 2;
    s.V = (maxV >= 4 && R % 4 == 0) ? 4 : (R % 2 == 0 ? 2 : 1);
    const uint32_t vecs = R / s.V;
    s.G = std::min<uint32_t>(32, pow2ceil(vecs));
    uint32_t ns = (vecs + s.G - 1) / s.G;
    s.NS = (int)pow2ceil(ns);
    if (s.NS > 8) throw CudaError(GG_ERR_UNSUPPORTED, "n_haps too large for the vector width its parity allows (use a multiple of 4)");
    const uint32_t rpi = 32 / s.G;
    s.nwarps_fwd = s.nwarps_bwd = std::min<uint32_t>(8, pow2ceil((R + rpi - 1) / rpi));
    return s;
}

template <typename T, int V, int NS>
static const void *kernel_ptr(bool fwd) {
    return fwd ? (const void *)step_kernel<T, V, NS, true> : (const void *)step_kernel<T, V, NS, false>;
}

template <typename T>
static const void *pick_kernel(int V, int NS, bool fwd) {
#define GG_PICK(VV, NN) \
    if (V == VV && NS == NN) return kernel_ptr<T, VV, NN>(fwd);
    GG_PICK(1, 1) GG_PICK(1, 2) GG_PICK(1, 4) GG_PICK(1, 8)
    GG_PICK(2, 1) GG_PICK(2, 2) GG_PICK(2, 4) GG_PICK(2, 8)
    if (sizeof(T) == 4) {
        if (V == 4 && NS == 1) return kernel_ptr<float, 4, 1>(fwd);
        if (V == 4 && NS == 2) return kernel_ptr<float, 4, 2>(fwd);
        if (V == 4 && NS == 4) return kernel_ptr<float, 4, 4>(fwd);
        if (V == 4 && NS == 8) return kernel_ptr<float, 4, 8>(fwd);
    }
#undef GG_PICK
    throw CudaError(GG_ERR_UNSUPPORTED, "no kernel for this vector shape");
}

static constexpr size_t kMaxDynSmem = 200 * 1024;
static constexpr size_t kMaxDynSmemTma = 222 * 1024;   // 227 KB per CTA minus the kernel's static shared memory
static constexpr size_t kMaxDynSmemTma2 = 112 * 1024;  // two CTAs per SM: (228 KB - 2 * (1 KB reserved + <1 KB static)) / 2

template <typename T>
struct Engine {
    static uint64_t block_stride(uint32_t R) { return ((uint64_t)R * R + 2 * R + 4 + 3) & ~(uint64_t)3; }
    static uint64_t col_elems(uint32_t u, uint32_t R) { return block_stride(R) << u; }
    static uint64_t fg_elems(uint32_t u, uint32_t nA) { return ((uint64_t)2 * (nA + 1)) << u; }

    // the kernel arguments of one column step of the schedule (pointers into the current state arena)
    static StepArgs<T> make_args(gg_hmm *h, const Op &op, bool fwd, bool init) {
        const Plan &pl = h->plan;
        const uint32_t R = pl.n_haps;
        const uint32_t co = op.c;
        const ColumnPlan &pco = pl.cols[co];
        unsigned char *arena = h->arena;
        StepArgs<T> a{};
        a.R = R;
        a.G = h->shape.G;
        a.BS = (uint32_t)block_stride(R);
        a.out = reinterpret_cast<T *>(arena + op.out_off);
        a.fg_out = reinterpret_cast<const T *>(arena + op.fg_out_off);
        a.u_out = pco.u;
        a.nA_out = pco.n_alleles;
        a.al_out = h->d_allele1.p + (size_t)co * R;
        a.partial = reinterpret_cast<T *>(h->d_partial.p);
        a.counter = h->d_counter.p;
        a.result = h->d_result.p + pco.gl_off;
        a.tA = 1.0; a.tB = 0.0; a.tC = 0.0;
        if (fwd) {
            a.beta = reinterpret_cast<const T *>(arena + op.beta_off);
            a.row_order = h->d_row_order.p + (size_t)co * R;
            a.out_scale = h->d_scale_a.p + co;
            if (co > 0) {
                const ColumnPlan &pci = pl.cols[co - 1];
                a.in = reinterpret_cast<const T *>(arena + op.in_off);
                a.in_scale = h->d_scale_a.p + (co - 1);
                a.al_in = h->d_allele1.p + (size_t)(co - 1) * R;
                a.u_in = pci.u;
                a.nA_in = pci.n_alleles;
                a.n_sh = pco.n_sh;
                a.shared_mask = pco.shared_mask_prev;
                a.free_mask = pco.term_mask_prev;
                a.n_free_left = pco.n_term_prev;
                a.tA = pco.tA; a.tB = pco.tB; a.tC = pco.tC;
            } else {
                a.n_sh = pco.u;   // no bits to broadcast: item index == bipartition index
                a.al_in = a.al_out;
            }
        } else {
            a.out_scale = h->d_scale_b.p + co;
            if (!init) {
                const uint32_t ci = co + 1;
                const ColumnPlan &pci = pl.cols[ci];
                a.in = reinterpret_cast<const T *>(arena + op.in_off);
                a.fg_in = reinterpret_cast<const T *>(arena + op.fg_in_off);
                a.in_scale = h->d_scale_b.p + ci;
                a.al_in = h->d_allele1.p + (size_t)ci * R;
                a.u_in = pci.u;
                a.nA_in = pci.n_alleles;
                a.n_sh = pci.n_sh;
                a.shared_mask = pci.shared_mask_prev;
                a.free_mask = pci.term_mask_prev;
                a.n_free_left = pci.n_term_prev;
                a.tA = pci.tA; a.tB = pci.tB; a.tC = pci.tC;
            } else {
                a.n_sh = pco.u;
                a.shared_mask = pco.u >= 32 ? 0xffffffffu : ((1u << pco.u) - 1u);
                a.free_mask = 0;
                a.n_free_left = 0;
                a.al_in = a.al_out;
            }
        }
        return a;
    }

    static void launch_step(gg_hmm *h, const Op &op, bool fwd, bool init) {
        const uint32_t R = h->plan.n_haps;
        const ColumnPlan &pco = h->plan.cols[op.c];
        const StepArgs<T> a = make_args(h, op, fwd, init);
        if (init) {
            // the all-ones last backward column: a fill (init_kernel, hmm_kernels.cuh)
            const uint64_t nb = (uint64_t)1 << pco.u;
            const uint32_t grid = (uint32_t)std::min<uint64_t>(nb, (uint64_t)h->num_sms * 8);
            const int V = h->shape.V;
            if (V == 4) { if constexpr (sizeof(T) == 4) init_kernel<T, 4><<<grid, 256, 0, h->stream>>>(a); }
            else if (V == 2) init_kernel<T, 2><<<grid, 256, 0, h->stream>>>(a);
            else init_kernel<T, 1><<<grid, 256, 0, h->stream>>>(a);
            GG_CUDA(cudaGetLastError());
            ++h->stats.n_launches;
            return;
        }
        if (try_launch_warp(h, a, pco, fwd)) return;
        if (try_launch_tma(h, a, pco, fwd)) return;
        const uint32_t rpi = 32 / h->shape.G;
        uint32_t nwarps = fwd ? h->shape.nwarps_fwd : h->shape.nwarps_bwd;
        SmemLayout L = smem_layout<T>(R, nwarps * rpi, pco.n_alleles + 1, fwd);
        while (L.total > kMaxDynSmem && nwarps > 1) {
            nwarps >>= 1;
            L = smem_layout<T>(R, nwarps * rpi, pco.n_alleles + 1, fwd);
        }
        if (L.total > kMaxDynSmem) throw CudaError(GG_ERR_UNSUPPORTED, "column needs more shared memory than one SM has");
        const void *k = pick_kernel<T>(h->shape.V, h->shape.NS, fwd);
        const int threads = (int)nwarps * 32;
        auto key = std::make_tuple(k, threads, (size_t)L.total);
        auto it = h->occ_cache.find(key);
        int occ;
        if (it == h->occ_cache.end()) {
            GG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, L.total));
            if (occ < 1) throw CudaError(GG_ERR_CUDA, "step kernel does not fit on an SM");
            h->occ_cache[key] = occ;
        } else {
            occ = it->second;
        }
        const uint64_t n_items = (uint64_t)1 << pco.u;
        uint32_t grid = (uint32_t)std::min<uint64_t>(n_items, (uint64_t)h->num_sms * occ);
        grid = std::min(grid, h->max_grid);
        void *params[] = {(void *)&a};
        GG_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(threads), params, L.total, h->stream));
        ++h->stats.n_launches;
    }

    // Warp-per-block kernel for small panels (hmm_warp_kernel.cuh): fp32, R <= 32.
    static const void *warp_kernel(bool fwd, int V, uint32_t G) {
#define GG_WK(VV, GG) if (V == VV && G == GG) return fwd ? (const void *)step_warp_kernel<VV, GG, true> : (const void *)step_warp_kernel<VV, GG, false>;
        GG_WK(2, 1) GG_WK(2, 2) GG_WK(2, 4) GG_WK(2, 8) GG_WK(2, 16) GG_WK(4, 1) GG_WK(4, 2) GG_WK(4, 4) GG_WK(4, 8)
#undef GG_WK
        return nullptr;
    }

    static bool try_launch_warp(gg_hmm *h, const StepArgs<T> &a, const ColumnPlan &pco, bool fwd) {
        if (sizeof(T) != 4 || h->kernel_variant == 1 || h->kernel_variant == 5 || a.in == nullptr || a.R > 32) return false;
        const void *k = warp_kernel(fwd, h->shape.V, a.G);
        if (!k) return false;
        const size_t smem = fwd ? (size_t)kWarpKernelWarps * a.R * a.R * sizeof(float) : 16;
        const int threads = 32 * kWarpKernelWarps;
        auto key = std::make_tuple(k, threads, smem);
        auto it = h->occ_cache.find(key);
        int occ;
        if (it == h->occ_cache.end()) {
            GG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, smem));
            h->occ_cache[key] = occ;
        } else {
            occ = it->second;
        }
        if (occ < 1) return false;
        const uint64_t n_items = (uint64_t)1 << pco.u;
        uint32_t grid = (uint32_t)std::min<uint64_t>((n_items + kWarpKernelWarps - 1) / kWarpKernelWarps, (uint64_t)h->num_sms * occ);
        grid = std::min(grid, h->max_grid);
        StepArgs<float> af;
        std::memcpy(&af, &a, sizeof(af));
        void *params[] = {(void *)&af};
        GG_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(threads), params, smem, h->stream));
        ++h->stats.n_launches;
        return true;
    }

    // Wide panels (hmm_wide_kernel.cuh): fp32, R % 4 == 0, 128 < R <= 512, reduce sets of one or two blocks.
    static bool try_launch_wide(gg_hmm *h, const StepArgs<T> &a, const ColumnPlan &pco, bool fwd, uint32_t nred) {
        const bool two = nred == 2;
        const void *k = fwd ? (two ? (const void *)wide_step_kernel<true, true> : (const void *)wide_step_kernel<true, false>)
                            : (two ? (const void *)wide_step_kernel<false, true> : (const void *)wide_step_kernel<false, false>);
        configure_once(h->device, 5, [] {
            for (const void *kk : {(const void *)wide_step_kernel<true, true>, (const void *)wide_step_kernel<true, false>,
                                   (const void *)wide_step_kernel<false, true>, (const void *)wide_step_kernel<false, false>})
                GG_CUDA(cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmemTma));
        });
        const uint32_t ncls = pco.n_alleles + 1;
        const uint32_t cw = (a.R + 63u) / 64u;
        const uint32_t nload = nred + (fwd ? 1u : 0u);
        // two row interleaves (twice the consumer warps) first; tiles of ~32 KB, then 3/4 and 1/2 of that; a ring that
        // holds two tiles' worth of loads (2 x nload stages) or more, so that the next tile is in flight while this one
        // is consumed; only then shallower rings
        struct { uint32_t rh, tr, stages; WideGeom L; bool ok; } pick{};
        const uint32_t tr_max = std::max<uint32_t>(2u, std::min<uint32_t>(a.R, (32u * 1024u) / (a.R * 4u)) & ~1u);
        for (uint32_t min_st : {2u * nload, nload + 1u}) {
            for (uint32_t rh : {2u, 1u}) {
                if (cw * rh > kWideMaxWarps) continue;
                for (uint32_t tr : {tr_max, std::max<uint32_t>(2u, (tr_max * 3u / 4u) & ~1u), std::max<uint32_t>(2u, (tr_max / 2u) & ~1u)}) {
                    for (uint32_t st = std::min<uint32_t>(kTmaMaxStages, 3u * nload); st >= min_st && !pick.ok; --st) {
                        const WideGeom L = wide_geom(a.R, tr, st, cw, rh, ncls, fwd);
                        if (L.total <= kMaxDynSmemTma) pick = {rh, tr, st, L, true};
                    }
                    if (pick.ok) break;
                }
                if (pick.ok) break;
            }
            if (pick.ok) break;
        }
        if (!pick.ok) return false;
        const uint64_t n_items = (uint64_t)1 << pco.u;
        const uint32_t grid = (uint32_t)std::min<uint64_t>(n_items, (uint64_t)h->num_sms);
        StepArgs<float> af;
        std::memcpy(&af, &a, sizeof(af));
        uint32_t ns = pick.stages, tr = pick.tr, cwv = cw, rhv = pick.rh;
        void *params[] = {(void *)&af, (void *)&ns, (void *)&tr, (void *)&cwv, (void *)&rhv};
        GG_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(32 * (1 + cw * pick.rh)), params, pick.L.total, h->stream));
        ++h->stats.n_launches;
        return true;
    }

    // Mid-size panels (hmm_colown_kernel.cuh): fp32, R % 4 == 0, 32 < R <= 128, reduce sets of one or two blocks, two
    // CTAs per SM.
    static bool try_launch_colown(gg_hmm *h, const StepArgs<T> &a, const ColumnPlan &pco, bool fwd, uint32_t nred, uint32_t l2a) {
        const bool two = nred == 2;
        const void *k = fwd ? (two ? (const void *)colown_step_kernel<true, true> : (const void *)colown_step_kernel<true, false>)
                            : (two ? (const void *)colown_step_kernel<false, true> : (const void *)colown_step_kernel<false, false>);
        configure_once(h->device, 6, [] {
            for (const void *kk : {(const void *)colown_step_kernel<true, true>, (const void *)colown_step_kernel<true, false>,
                                   (const void *)colown_step_kernel<false, true>, (const void *)colown_step_kernel<false, false>})
                GG_CUDA(cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmemTma2));
        });
        const uint32_t ncls = pco.n_alleles + 1;
        const uint32_t cw = (a.R + 31u) / 32u;
        static const uint32_t rh_env = [] { const char *e = std::getenv("GG_COLOWN_RH"); return e ? (uint32_t)std::atoi(e) : 0u; }();
        uint32_t rh = rh_env ? rh_env : std::max<uint32_t>(1u, std::min<uint32_t>(4u, kColownMaxWarps / cw));
        while (rh > 1 && cw * rh > kColownMaxWarps) --rh;
        const uint32_t nload = nred + (fwd ? 1u : 0u);
        ColownGeom L{};
        uint32_t stages = 0;
        for (uint32_t st = std::min<uint32_t>(kTmaMaxStages, nload + 2u); st >= nload + 1u; --st) {
            L = colown_geom(a.R, st, cw, rh, ncls, fwd);
            if (L.total <= kMaxDynSmemTma2) { stages = st; break; }
        }
        if (!stages) return false;
        const int threads = 32 * (1 + (int)(cw * rh));
        auto key = std::make_tuple(k, threads, (size_t)L.total);
        auto it = h->occ_cache.find(key);
        int occ;
        if (it == h->occ_cache.end()) {
            GG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, L.total));
            h->occ_cache[key] = occ;
        } else {
            occ = it->second;
        }
        if (occ < 2) return false;
        const uint64_t n_items = (uint64_t)1 << pco.u;
        const uint32_t grid = (uint32_t)std::min<uint64_t>(n_items, (uint64_t)h->num_sms * 2);
        StepArgs<float> af;
        std::memcpy(&af, &a, sizeof(af));
        uint32_t ns = stages, cwv = cw, rhv = rh, l2v = l2a;
        void *params[] = {(void *)&af, (void *)&ns, (void *)&cwv, (void *)&rhv, (void *)&l2v};
        GG_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(threads), params, L.total, h->stream));
        ++h->stats.n_launches;
        return true;
    }

    // Pipelined bulk-copy kernel (hmm_tma_kernel.cuh): fp32, R % 4 == 0, R <= 512.
    static const void *tma_kernel(bool fwd, int V, uint32_t G, int NS, bool mt, int cwt, bool two) {
#define GG_TMA(VV, GG, NN, MM, CC, TT) if (V == VV && G == GG && NS == NN && mt == MM && cwt == CC && two == TT) return fwd ? (const void *)step_tma_kernel<true, GG, NN, MM, CC, TT, VV> : (const void *)step_tma_kernel<false, GG, NN, MM, CC, TT, VV>;
        // float4 rows (R % 4 == 0)
        GG_TMA(4, 1, 1, false, 8, false) GG_TMA(4, 2, 1, false, 8, false) GG_TMA(4, 4, 1, false, 8, false) GG_TMA(4, 8, 1, false, 8, false)
        GG_TMA(4, 16, 1, false, 8, false) GG_TMA(4, 32, 1, false, 8, false) GG_TMA(4, 32, 1, false, 16, false)
        GG_TMA(4, 16, 1, false, 8, true) GG_TMA(4, 32, 1, false, 8, true) GG_TMA(4, 32, 1, false, 16, true)
        GG_TMA(4, 16, 2, false, 8, false) GG_TMA(4, 16, 2, false, 8, true)
        GG_TMA(4, 32, 1, true, 0, false) GG_TMA(4, 32, 2, true, 0, false) GG_TMA(4, 32, 4, true, 0, false)
        // float2 rows (R % 4 == 2, e.g. 94 haplotypes): panels of 34..256
        GG_TMA(2, 32, 1, false, 8, false) GG_TMA(2, 32, 1, false, 16, false) GG_TMA(2, 32, 1, false, 8, true) GG_TMA(2, 32, 1, false, 16, true)
        GG_TMA(2, 32, 2, false, 8, false) GG_TMA(2, 32, 2, false, 16, false) GG_TMA(2, 32, 2, false, 8, true) GG_TMA(2, 32, 2, false, 16, true)
        GG_TMA(2, 32, 2, true, 0, false) GG_TMA(2, 32, 4, true, 0, false)
#undef GG_TMA
        return nullptr;
    }

    static bool try_launch_tma(gg_hmm *h, const StepArgs<T> &a, const ColumnPlan &pco, bool fwd) {
        if (sizeof(T) != 4 || h->kernel_variant == 1 || a.in == nullptr) return false;
        const int V = h->shape.V;
        int NS = h->shape.NS;
        uint32_t G = a.G;
        if (V < 2 || a.R > 512) return false;
        // rows of 17..32 vectors (R = 68..128 for float4): two rows per warp (16 lanes x 2 vectors each) instead of one
        // row on 17..32 of 32 lanes: the per-row bookkeeping and the shuffle reduction are shared by two rows
        static const int half_rows_env = [] { const char *e = std::getenv("GG_TMA_HALFROWS"); return e ? std::atoi(e) : 0; }();
        if (V == 4 && G == 32 && NS == 1 && a.R <= 128 && (half_rows_env & (fwd ? 2 : 1))) { G = 16; NS = 2; }
        configure_once(h->device, 1, [] {
            for (int vv : {2, 4})
                for (uint32_t g = 1; g <= 32; g <<= 1)
                    for (bool f : {false, true})
                        for (int ns : {1, 2, 4})
                            for (bool mt : {false, true})
                                for (int cwt : {0, 8, 16})
                                    for (bool tw : {false, true})
                                        if (const void *kk = tma_kernel(f, vv, g, ns, mt, cwt, tw))
                                            GG_CUDA(cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmemTma));
        });
        const uint32_t rpi = 32 / G;
        const uint32_t nred = fwd ? (1u << a.n_free_left) : (1u << (a.u_in - a.n_sh));
        const uint32_t ncls = pco.n_alleles + 1;
        if (a.R > 128 && V == 4 && nred <= 2 && h->kernel_variant != 8 && try_launch_wide(h, a, pco, fwd, nred)) return true;
        static const uint32_t l2_ahead_env0 = [] {
            const char *e = std::getenv("GG_L2_AHEAD");
            return e ? (uint32_t)std::atoi(e) : kTmaL2Ahead;
        }();
        // measured on the R = 88 bench columns (same box, bytes actually moved): backward one-block steps 5.62 -> 5.77
        // TB/s, two-block reduce sets 4.18 -> 4.76 TB/s; forward 6.03 -> 5.93 TB/s — so the backward step only.
        // GG_COLOWN: bit 0 = backward, bit 1 = forward
        static const int colown_env = [] { const char *e = std::getenv("GG_COLOWN"); return e ? std::atoi(e) : 1; }();
        const int kv = h->kernel_variant;
        if ((colown_env & (fwd ? 2 : 1)) && a.R > 32 && a.R <= 128 && V == 4 && nred <= 2 && kv != 2 && kv != 3 && kv != 9 &&
            try_launch_colown(h, a, pco, fwd, nred, l2_ahead_env0))
            return true;
        // geometry search.  Whole block = one tile (R <= 128): two CTAs of 8 consumer warps per SM when the column's
        // footprint fits twice, else one CTA of 16 (same warp count, deeper ring), else one of 8.  Larger R: row tiles
        // of <= ~40 KB, most consumer warps first, >= 3 stages.
        // both blocks of a two-block reduce set held in stages (where that kernel variant is compiled)
        const bool two_ok = nred == 2 && a.R <= 128 && tma_kernel(fwd, V, G, NS, false, 8, true) != nullptr;
        struct Pick { uint32_t cw, tr, stages, ctas; int cwt; TmaGeom L; bool ok; } pick{};
        auto fit = [&](uint32_t cw, uint32_t tr, uint32_t lo, uint32_t hi, size_t cap, uint32_t &stages, TmaGeom &L) {
            for (stages = hi; stages >= lo; --stages) {
                L = tma_geom(a.R, tr, stages, cw * rpi, ncls, fwd, nred > 1 && !(two_ok && tr == a.R));
                if (L.total <= cap) return true;
            }
            return false;
        };
        if (a.R <= 128 && tma_kernel(fwd, V, G, NS, false, 8, false)) {
            uint32_t st; TmaGeom L;
            if (h->kernel_variant != 2 && fit(8, a.R, 3, 4, kMaxDynSmemTma2, st, L)) pick = {8, a.R, st, 2, 8, L, true};
            else if (tma_kernel(fwd, V, G, NS, false, 16, false) && h->kernel_variant != 3 &&
                     fit(16, a.R, 3, kTmaMaxStages, kMaxDynSmemTma, st, L)) pick = {16, a.R, st, 1, 16, L, true};
            else if (fit(8, a.R, 3, kTmaMaxStages, kMaxDynSmemTma, st, L)) pick = {8, a.R, st, 1, 8, L, true};
        }
        if (!pick.ok && G == 32) {
            for (uint32_t cw : {8u, 4u, 2u}) {
                const uint32_t ngrp = cw * rpi;
                uint32_t tr_max = std::min<uint32_t>(a.R, (40u * 1024u) / (a.R * 4u));
                tr_max = std::max<uint32_t>(ngrp, tr_max / ngrp * ngrp);
                for (uint32_t tr = tr_max; tr >= ngrp && !pick.ok; tr -= ngrp) {
                    uint32_t st; TmaGeom L;
                    const uint32_t trc = std::min<uint32_t>(tr, a.R);
                    if (trc < a.R && fit(cw, trc, 3, 6, kMaxDynSmemTma, st, L)) pick = {cw, trc, st, 1, 0, L, true};
                    if (tr == ngrp || tr < 2 * ngrp) break;
                }
                if (pick.ok) break;
            }
        }
        if (!pick.ok) return false;
        const void *k = tma_kernel(fwd, V, G, NS, pick.cwt == 0, pick.cwt, two_ok && pick.cwt != 0);
        if (!k) return false;
        const int threads = 32 * (1 + (int)pick.cw);
        auto key = std::make_tuple(k, threads, (size_t)pick.L.total);
        auto it = h->occ_cache.find(key);
        int occ;
        if (it == h->occ_cache.end()) {
            GG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, pick.L.total));
            h->occ_cache[key] = occ;
        } else {
            occ = it->second;
        }
        if (occ < 1) return false;
        const uint32_t ctas = std::min<uint32_t>(pick.ctas, (uint32_t)occ);
        const uint64_t n_items = (uint64_t)1 << pco.u;
        const uint32_t grid = (uint32_t)std::min<uint64_t>(n_items, (uint64_t)h->num_sms * ctas);
        StepArgs<float> af;
        std::memcpy(&af, &a, sizeof(af));
        uint32_t ns = pick.stages, tr = pick.tr;
        static const uint32_t l2_ahead_env = [] {
            const char *e = std::getenv("GG_L2_AHEAD");
            return e ? (uint32_t)std::atoi(e) : kTmaL2Ahead;
        }();
        uint32_t l2a = l2_ahead_env;
        void *params[] = {(void *)&af, (void *)&ns, (void *)&tr, (void *)&l2a};
        GG_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(threads), params, pick.L.total, h->stream));
        ++h->stats.n_launches;
        return true;
    }

    // PASS 1 of the emission kernel for the columns of one EMIT op (tables scaled by the column maxima of PASS 0)
    static void launch_emit(gg_hmm *h, const Op &op) {
        T *region = reinterpret_cast<T *>(h->arena + op.out_off);
        launch_emission<1>(h, op.c, op.hi, region, h->fg_prefix_elems[op.c]);
    }

    // PASS 0 for every column: the per-column emission maxima (once per sweep, ~0.1 % of its bytes)
    static void launch_emax(gg_hmm *h) {
        const uint32_t C = h->plan.n_columns;
        GG_CUDA(cudaMemsetAsync(h->d_emax.p, 0, (size_t)C * sizeof(double), h->stream));
        launch_emission<0>(h, 0, C - 1, nullptr, 0);
    }

    template <int PASS>
    static void launch_emission(gg_hmm *h, uint32_t lo, uint32_t hi, T *region, uint64_t prefix_lo) {
        const Plan &pl = h->plan;
        while (lo <= hi) {
            const uint32_t n = std::min<uint32_t>(hi - lo + 1, 32768);
            uint32_t umax = 0;
            for (uint32_t c = lo; c < lo + n; ++c) umax = std::max(umax, pl.cols[c].u);
            const uint64_t work = (uint64_t)2 << umax;
            const uint32_t bx = (uint32_t)std::min<uint64_t>((work + 255) / 256, 2048);
            emission_kernel<T, PASS><<<dim3(bx, n), 256, 0, h->stream>>>(h->d_cols.p, h->d_em_side.p, h->d_em_prob.p, region,
                                                                         prefix_lo, lo, h->d_emax.p);
            GG_CUDA(cudaGetLastError());
            ++h->stats.n_launches;
            lo += n;
        }
    }

    // ---- chain kernel (hmm_chain_kernel.cuh): fuse consecutive steps over columns without untagged reads
    static const void *chain_kernel_ptr(bool fwd, uint32_t nit) {
        switch (nit) {
#define GG_CH(NN) case NN: return fwd ? (const void *)chain_kernel<true, NN> : (const void *)chain_kernel<false, NN>;
            GG_CH(1) GG_CH(2) GG_CH(3) GG_CH(4) GG_CH(5) GG_CH(6) GG_CH(7) GG_CH(8)
#undef GG_CH
        }
        return nullptr;
    }

    static void fuse_runs(gg_hmm *h) {
        const Plan &pl = h->plan;
        const uint32_t R = pl.n_haps;
        const std::vector<Op> &ops = h->sched.ops;
        h->exec_ops.clear();
        h->runs.clear();
        h->chain_steps.clear();
        const bool eligible = sizeof(T) == 4 && h->kernel_variant != 1 && h->kernel_variant != 4 && R % 4 == 0 && R <= 128;
        auto single = [&](uint32_t c) { return pl.cols[c].u == 0; };
        auto op_ok = [&](const Op &op) {
            if (!eligible) return false;
            if (op.kind == OP_BWD) return single(op.c) && single(op.c + 1);
            if (op.kind == OP_BWD_INIT) return single(op.c);
            if (op.kind == OP_FWD) return single(op.c) && (op.c == 0 || single(op.c - 1));
            return false;
        };
        size_t i = 0;
        while (i < ops.size()) {
            if (!op_ok(ops[i])) { h->exec_ops.push_back(ops[i++]); continue; }
            const bool fwd = ops[i].kind == OP_FWD;
            size_t j = i;
            while (j < ops.size() && op_ok(ops[j]) && (ops[j].kind == OP_FWD) == fwd &&
                   (j == i || (fwd ? ops[j].c == ops[j - 1].c + 1 : (ops[j].c + 1 == ops[j - 1].c && ops[j].in_off == ops[j - 1].out_off))))
                ++j;
            if (j - i < 4) { for (; i < j; ++i) h->exec_ops.push_back(ops[i]); continue; }
            gg_hmm::Run run{};
            run.fwd = fwd;
            run.first = (uint32_t)h->chain_steps.size();
            run.n = (uint32_t)(j - i);
            run.in_col = -1;
            for (size_t q = i; q < j; ++q) {
                const Op &op = ops[q];
                const ColumnPlan &pco = pl.cols[op.c];
                ChainStep st{};
                st.out_off = op.out_off;
                st.beta_off = op.beta_off;
                st.fg_out_off = op.fg_out_off;
                st.fg_in_off = op.fg_in_off;
                st.gl_off = pco.gl_off;
                st.c_out = op.c;
                st.nA_out = pco.n_alleles;
                st.c_in = 0xffffffffu;
                st.nA_in = 1;
                const ColumnPlan *pt = nullptr;      // whose transition
                if (fwd && op.c > 0) { st.c_in = op.c - 1; pt = &pco; }
                if (op.kind == OP_BWD) { st.c_in = op.c + 1; pt = &pl.cols[op.c + 1]; }
                if (st.c_in != 0xffffffffu) st.nA_in = pl.cols[st.c_in].n_alleles;
                if (pt) { st.tA = (float)pt->tA; st.tB = (float)pt->tB; st.tC = (float)pt->tC; }
                run.max_classes = std::max(run.max_classes, pco.n_alleles + 1);
                if (q == i && st.c_in != 0xffffffffu) { run.in_off = op.in_off; run.in_col = st.c_in; }
                h->chain_steps.push_back(st);
            }
            Op fused{};
            fused.kind = OP_RUN;
            fused.c = (uint32_t)h->runs.size();
            h->runs.push_back(run);
            h->exec_ops.push_back(fused);
            i = j;
        }
    }

    // ---- wave kernel (hmm_wave_kernel.cuh): fuse consecutive steps of one direction over small panels (R <= 32)
    static bool wave_eligible(const gg_hmm *h) {
        return sizeof(T) == 4 && h->kernel_variant != 1 && h->kernel_variant != 5 && h->kernel_variant != 7 && h->plan.n_haps <= 32 &&
               warp_kernel(true, h->shape.V, h->shape.G) != nullptr;
    }
    // panels of an even number of haplotypes up to 16 get the thread-per-block path inside the wave kernel
    static const void *wave_kernel_ptr(bool fwd, int V, uint32_t G, uint32_t R) {
#define GG_WV(VV, GG, RL) if (V == VV && G == GG && rl == RL) return fwd ? (const void *)wave_kernel<VV, GG, true, RL> : (const void *)wave_kernel<VV, GG, false, RL>;
        const uint32_t rl = (R % 2 == 0 && R <= 16) ? R : 0;
        static const bool no_pair = std::getenv("GG_NO_PAIR_WAVE") != nullptr;      // A/B: thread-per-block path instead
        if (rl && !no_pair) {
#define GG_PW(RL) if (rl == RL) return fwd ? (const void *)pair_wave_kernel<RL, true> : (const void *)pair_wave_kernel<RL, false>;
            GG_PW(2) GG_PW(4) GG_PW(6) GG_PW(8) GG_PW(10) GG_PW(12) GG_PW(14) GG_PW(16)
#undef GG_PW
        }
        GG_WV(2, 1, 2) GG_WV(4, 1, 4) GG_WV(2, 4, 6) GG_WV(4, 2, 8) GG_WV(2, 8, 10) GG_WV(4, 4, 12) GG_WV(2, 8, 14) GG_WV(4, 4, 16)
        GG_WV(2, 1, 0) GG_WV(2, 2, 0) GG_WV(2, 4, 0) GG_WV(2, 8, 0) GG_WV(2, 16, 0) GG_WV(4, 1, 0) GG_WV(4, 2, 0) GG_WV(4, 4, 0) GG_WV(4, 8, 0)
#undef GG_WV
        return nullptr;
    }

    static void fuse_waves(gg_hmm *h) {
        h->wave_ops.clear();
        h->wave_w_off.clear();
        h->wave_w_floats = 0;
        if (!wave_eligible(h)) return;
        const std::vector<Op> ops = h->exec_ops;
        h->exec_ops.clear();
        auto op_ok = [&](const Op &op) { return op.kind == OP_BWD || (op.kind == OP_FWD && op.c > 0); };
        size_t i = 0;
        while (i < ops.size()) {
            if (!op_ok(ops[i])) { h->exec_ops.push_back(ops[i++]); continue; }
            const bool fwd = ops[i].kind == OP_FWD;
            size_t j = i;
            while (j < ops.size() && op_ok(ops[j]) && (ops[j].kind == OP_FWD) == fwd &&
                   (j == i || (fwd ? ops[j].c == ops[j - 1].c + 1 : (ops[j].c + 1 == ops[j - 1].c && ops[j].in_off == ops[j - 1].out_off))))
                ++j;
            // every eligible step goes through the wave kernel, single ones too: its summation orders depend on the
            // wave grid only (problem and device), so the posteriors stay bit-identical across checkpoint schedules
            gg_hmm::Run run{};
            run.fwd = fwd;
            run.wave = true;
            run.first = (uint32_t)h->wave_ops.size();
            run.n = (uint32_t)(j - i);
            for (size_t q = i; q < j; ++q) {
                h->wave_ops.push_back(ops[q]);
                h->wave_w_off.push_back(h->wave_w_floats);
                if (fwd) {
                    const uint64_t nA = h->plan.cols[ops[q].c].n_alleles;
                    h->wave_w_floats += nA * nA;          // per CTA; scaled by the grid at allocation
                }
            }
            Op fused{};
            fused.kind = OP_RUN;
            fused.c = (uint32_t)h->runs.size();
            h->runs.push_back(run);
            h->exec_ops.push_back(fused);
            i = j;
        }
    }

    // device copy of the wave steps' arguments: they hold pointers into the state arena, so it is (re)built whenever
    // the handle sweeps in another arena (shared-arena handles, a handle recreated on a cached arena)
    static void upload_wave_steps(gg_hmm *h) {
        if (h->wave_ops.empty() || h->wave_built_for == h->arena) return;
        std::vector<WaveStep> steps(h->wave_ops.size());
        for (size_t q = 0; q < steps.size(); ++q) {
            const Op &op = h->wave_ops[q];
            StepArgs<T> a = make_args(h, op, op.kind == OP_FWD, false);
            std::memcpy(&steps[q].a, &a, sizeof(StepArgs<float>));
            steps[q].w_off = h->wave_w_off[q] * h->wave_grid;
            steps[q].pad = 0;
            const uint32_t Rl = h->plan.n_haps;
            const uint32_t c_in = op.kind == OP_FWD ? op.c - 1 : op.c + 1;
            std::memset(steps[q].al_out, 0, 64);
            std::memcpy(steps[q].al_out, h->plan.allele1.data() + (size_t)op.c * Rl, Rl);
            std::memcpy(steps[q].al_in, h->plan.allele1.data() + (size_t)c_in * Rl, Rl);
        }
        GG_CUDA(cudaMemcpyAsync(h->d_wave_steps.p, steps.data(), steps.size() * sizeof(WaveStep), cudaMemcpyHostToDevice, h->stream));
        GG_CUDA(cudaStreamSynchronize(h->stream));      // `steps` is pageable and about to go away
        h->wave_built_for = h->arena;
    }

    static void launch_wave(gg_hmm *h, const gg_hmm::Run &run) {
        const int V = h->shape.V;
        const uint32_t G = h->shape.G;
        const void *k = wave_kernel_ptr(run.fwd, V, G, h->plan.n_haps);
        WaveArgs w{};
        w.steps = reinterpret_cast<const WaveStep *>(h->d_wave_steps.p) + run.first;
        w.n_steps = run.n;
        w.w_partial = reinterpret_cast<float *>(h->d_wave_w.p);
        w.barrier = reinterpret_cast<unsigned long long *>(h->d_wave_bar.p);
        GG_CUDA(cudaMemsetAsync(h->d_wave_bar.p, 0, 3 * sizeof(unsigned long long), h->stream));
        void *params[] = {(void *)&w};
        GG_CUDA(cudaLaunchCooperativeKernel(k, dim3(h->wave_grid), dim3(32 * kWarpKernelWarps), params, 0, h->stream));
        ++h->stats.n_launches;
        if (run.fwd) {
            wave_posterior_kernel<<<run.n, 32, 0, h->stream>>>(w.steps, run.n, w.w_partial, h->wave_grid);
            GG_CUDA(cudaGetLastError());
            ++h->stats.n_launches;
        }
    }

    static void launch_run(gg_hmm *h, const Op &op) {
        if (h->runs[op.c].wave) {
            launch_wave(h, h->runs[op.c]);
            return;
        }
        const gg_hmm::Run &run = h->runs[op.c];
        const uint32_t R = h->plan.n_haps;
        ChainArgs a{};
        a.arena = h->arena;
        a.steps = h->d_chain_steps.p + run.first;
        a.n_steps = run.n;
        a.in = run.in_col >= 0 ? reinterpret_cast<const float *>(h->arena + run.in_off) : nullptr;
        a.in_scale = run.in_col >= 0 ? (run.fwd ? h->d_scale_a.p : h->d_scale_b.p) + run.in_col : nullptr;
        a.scale = run.fwd ? h->d_scale_a.p : h->d_scale_b.p;
        a.result = h->d_result.p;
        a.allele1 = h->d_allele1.p;
        a.R = R;
        a.BS = (uint32_t)block_stride(R);
        a.max_classes = run.max_classes;
        a.no_fast_posterior = h->kernel_variant == 6;
        const void *k = chain_kernel_ptr(run.fwd, (R + kChainWarps - 1) / kChainWarps);
        const uint32_t smem = chain_smem_bytes(R, run.max_classes, run.fwd);
        configure_once(h->device, 2, [] {
            for (uint32_t n = 1; n <= (uint32_t)kChainMaxNIT; ++n)
                for (bool f : {false, true})
                    GG_CUDA(cudaFuncSetAttribute(chain_kernel_ptr(f, n), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmemTma));
        });
        void *params[] = {(void *)&a};
        GG_CUDA(cudaLaunchKernel(k, dim3(1), dim3(kChainThreads), params, smem, h->stream));
        ++h->stats.n_launches;
    }

    static void set_attrs(int device) {
        configure_once(device, sizeof(T) == 4 ? 3 : 4, [] {
            for (int V : {1, 2, 4})
                for (int NS : {1, 2, 4, 8})
                    for (bool fwd : {false, true}) {
                        if (V == 4 && sizeof(T) != 4) continue;
                        const void *k = pick_kernel<T>(V, NS, fwd);
                        GG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
                    }
        });
    }

    // ---- host half of gg_hmm_create: sizes, checkpoint schedule, run fusion.  No CUDA call: a worker thread can do it
    //      for chain i+1 while the device sweeps chain i (gg_hmm_run_batch).
    static void prepare_host(gg_hmm *h, uint64_t budget) {
        Plan &pl = h->plan;
        const uint32_t C = pl.n_columns, R = pl.n_haps;
        h->elem = sizeof(T);
        h->shape = choose_shape<T>(R);
        h->colb.resize(C);
        h->fgb.resize(C);
        h->fg_prefix_elems.assign(C + 1, 0);
        h->host_cols.assign(C, ColDesc{});
        for (uint32_t c = 0; c < C; ++c) {
            const ColumnPlan &cp = pl.cols[c];
            h->colb[c] = col_elems(cp.u, R) * sizeof(T);
            h->fgb[c] = fg_elems(cp.u, cp.n_alleles) * sizeof(T);
            h->fg_prefix_elems[c + 1] = h->fg_prefix_elems[c] + fg_elems(cp.u, cp.n_alleles);
            ColDesc &cd = h->host_cols[c];
            cd.u = cp.u;
            cd.n_alleles = cp.n_alleles;
            cd.em_count = cp.em_count;
            cd.em_off = cp.em_off;
            cd.em_prob_off = cp.em_prob_off;
            cd.fg_prefix = h->fg_prefix_elems[c];
            uint32_t present = 0;
            const uint8_t *al = pl.allele1.data() + (size_t)c * R;
            for (uint32_t hh = 0; hh < R; ++hh)
                if (al[hh]) present |= 1u << (al[hh] - 1);
            cd.present = present;
        }
        // if everything fits there is no reason to reserve more than that
        uint64_t all = 0, maxcol = 0;
        for (uint32_t c = 0; c < C; ++c) {
            all += h->colb[c] + h->fgb[c] + 512;
            maxcol = std::max(maxcol, h->colb[c]);
        }
        budget = std::min<uint64_t>(budget, all + 3 * maxcol + 4096);
        try {
            build_schedule(h->colb, h->fgb, budget, h->sched);
            validate_schedule(h->colb, h->fgb, h->sched);
        } catch (const CudaError &) {
            throw;
        } catch (const std::runtime_error &e) {
            const std::string m = e.what();
            throw CudaError(m.find("budget") != std::string::npos ? GG_ERR_NOMEM : GG_ERR_INVALID, m);
        }
        fuse_runs(h);
        fuse_waves(h);
    }

    // ---- device half: static tables H2D, arena
    static void create_device(gg_hmm *h) {
        Plan &pl = h->plan;
        const uint32_t C = pl.n_columns;
        set_attrs(h->device);
        h->max_grid = (uint32_t)h->num_sms * 16;
        // one table arena: sub-allocations at 256-byte granularity
        size_t off = 0;
        auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        const size_t o_cols = carve(h->host_cols.size() * sizeof(ColDesc));
        const size_t o_side = carve(pl.em_side.size());
        const size_t o_prob = carve(pl.em_prob.size() * sizeof(double));
        const size_t o_al = carve(pl.allele1.size());
        const size_t o_ro = carve(pl.row_order.size() * sizeof(uint16_t));
        const size_t o_ch = carve(h->chain_steps.size() * sizeof(ChainStep));
        const size_t o_sa = carve((size_t)C * sizeof(double));
        const size_t o_sb = carve((size_t)C * sizeof(double));
        const size_t o_emax = carve((size_t)C * sizeof(double));
        const size_t o_res = carve((size_t)pl.gl_offsets[C] * sizeof(double));
        const size_t o_cnt = carve(sizeof(unsigned int));
        const size_t o_part = carve((size_t)h->max_grid * kPartialStride * sizeof(T));
        // wave runs: a grid that is co-resident (cooperative launch), as many CTAs per SM as the kernel is compiled for
        h->wave_grid = 0;
        if (!h->wave_ops.empty()) {
            int occ = 0;
            GG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wave_kernel_ptr(true, h->shape.V, h->shape.G, pl.n_haps), 32 * kWarpKernelWarps, 0));
            int occ_b = 0;
            GG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, wave_kernel_ptr(false, h->shape.V, h->shape.G, pl.n_haps), 32 * kWarpKernelWarps, 0));
            occ = std::min(occ, occ_b);
            if (occ < 1) throw CudaError(GG_ERR_CUDA, "wave kernel does not fit on an SM");
            const uint64_t max_items = (uint64_t)1 << pl.max_u;
            const bool pair = pl.n_haps % 2 == 0 && pl.n_haps <= 16 && std::getenv("GG_NO_PAIR_WAVE") == nullptr;
            h->wave_grid = (uint32_t)std::min<uint64_t>((uint64_t)h->num_sms * std::min(occ, pair ? GG_PAIR_CTAS : GG_WAVE_MIN_CTAS),
                                                       std::max<uint64_t>(1, (max_items + kWarpKernelWarps - 1) / kWarpKernelWarps));
        }
        const size_t o_wsteps = carve(h->wave_ops.size() * sizeof(WaveStep));
        const size_t o_wsum = carve((size_t)2 * h->wave_grid * sizeof(float));
        const size_t o_wbar = carve(3 * sizeof(unsigned long long));
        const size_t o_ww = carve((size_t)h->wave_w_floats * h->wave_grid * sizeof(float));
        const size_t need = off + 256;
        h->tables = g_arena_cache.take(2 * h->device + 1, need, h->tables_bytes);
        if (!h->tables) {
            GG_CUDA(device_alloc(&h->tables, need));
            h->tables_bytes = need;
        }
        unsigned char *tb = h->tables;
        h->d_cols.bind(tb + o_cols, h->host_cols.size());
        h->d_em_side.bind(tb + o_side, pl.em_side.size());
        h->d_em_prob.bind(tb + o_prob, pl.em_prob.size());
        h->d_allele1.bind(tb + o_al, pl.allele1.size());
        h->d_row_order.bind(tb + o_ro, pl.row_order.size());
        h->d_chain_steps.bind(tb + o_ch, h->chain_steps.size());
        h->d_scale_a.bind(tb + o_sa, C);
        h->d_scale_b.bind(tb + o_sb, C);
        h->d_emax.bind(tb + o_emax, C);
        h->d_result.bind(tb + o_res, pl.gl_offsets[C]);
        h->d_counter.bind(tb + o_cnt, 1);
        h->d_partial.bind(tb + o_part, (size_t)h->max_grid * kPartialStride * sizeof(T));
        h->d_wave_steps.bind(tb + o_wsteps, h->wave_ops.size() * sizeof(WaveStep));
        h->d_wave_sum.bind(tb + o_wsum, (size_t)2 * h->wave_grid * sizeof(float));
        h->d_wave_bar.bind(tb + o_wbar, 6);
        h->d_wave_w.bind(tb + o_ww, (size_t)h->wave_w_floats * h->wave_grid * sizeof(float));
        h->wave_built_for = nullptr;
        h->d_cols.upload(h->host_cols, h->stream);
        h->d_em_side.upload(pl.em_side, h->stream);
        h->d_em_prob.upload(pl.em_prob, h->stream);
        h->d_allele1.upload(pl.allele1, h->stream);
        h->d_row_order.upload(pl.row_order, h->stream);
        h->d_chain_steps.upload(h->chain_steps, h->stream);
        GG_CUDA(cudaMemsetAsync(h->d_counter.p, 0, sizeof(unsigned int), h->stream));
        if (!h->shared_arena) {
            h->arena = g_arena_cache.take(2 * h->device, h->sched.arena_bytes, h->arena_bytes);
            if (!h->arena && h->sched.arena_bytes) {
                GG_CUDA(device_alloc(&h->arena, h->sched.arena_bytes));
                h->arena_bytes = h->sched.arena_bytes;
            }
            poison_arena(h);
        }
        GG_CUDA(cudaMemsetAsync(h->d_scale_a.p, 0, C * sizeof(double), h->stream));
        GG_CUDA(cudaMemsetAsync(h->d_scale_b.p, 0, C * sizeof(double), h->stream));
        if (pl.gl_offsets[C]) GG_CUDA(cudaMemsetAsync(h->d_result.p, 0, pl.gl_offsets[C] * sizeof(double), h->stream));
        GG_CUDA(cudaStreamSynchronize(h->stream));
        // stats that do not depend on a sweep
        const uint32_t R = pl.n_haps;
        gg_stats &st = h->stats;
        st.n_columns = C;
        st.state_columns = pl.state_columns;
        st.algorithmic_bytes = 5 * (uint64_t)sizeof(T) * pl.state_columns;
        st.state_bytes_reserved = h->sched.arena_bytes;
        st.max_untagged = pl.max_u;
        st.checkpoint_levels = h->sched.levels;
        st.n_backward_steps = h->sched.n_bwd;
        st.n_forward_steps = h->sched.n_fwd;
        uint64_t sb = 0;
        auto S = [&](uint32_t c) { return ((uint64_t)R * R) << pl.cols[c].u; };
        for (const Op &op : h->sched.ops) {
            if (op.kind == OP_FWD) sb += (2 * S(op.c) + (op.c > 0 ? S(op.c - 1) : 0)) * sizeof(T);
            else if (op.kind == OP_BWD) sb += (S(op.c) + S(op.c + 1)) * sizeof(T);
            else if (op.kind == OP_BWD_INIT) sb += S(op.c) * sizeof(T);
        }
        st.scheduled_bytes = sb;
    }

    // test hook: GG_POISON_ARENA=1 fills the state arena with NaN bit patterns so that any read of a value the
    // schedule has not produced yet shows up in the posteriors
    static void poison_arena(gg_hmm *h) {
        if (const char *poison = std::getenv("GG_POISON_ARENA"))
            if (poison[0] == '1' && h->arena) GG_CUDA(cudaMemsetAsync(h->arena, 0xFF, h->arena_bytes, h->stream));
    }

    static void sweep(gg_hmm *h, int timed) {
        if (!h->shared_arena) {
            sweep_with_arena(h, timed);
            return;
        }
        SharedArena &sa = shared_arena_of(h->device);
        std::lock_guard<std::mutex> g(sa.mu);
        if (sa.n < h->sched.arena_bytes) {
            if (sa.p) GG_CUDA(cudaFree(sa.p));
            sa.p = nullptr;
            sa.n = 0;
            GG_CUDA(device_alloc(&sa.p, h->sched.arena_bytes));
            sa.n = h->sched.arena_bytes;
        }
        h->arena = sa.p;
        h->arena_bytes = h->sched.arena_bytes;
        try {
            poison_arena(h);               // whatever another chain left behind must not matter either
            sweep_with_arena(h, timed);    // returns after the stream has drained
        } catch (...) {
            cudaStreamSynchronize(h->stream);
            h->arena = nullptr;
            throw;
        }
        h->arena = nullptr;
    }

    static void sweep_with_arena(gg_hmm *h, int timed) {
        gg_stats &st = h->stats;
        st.n_launches = 0;
        const std::vector<Op> &ops = h->exec_ops;
        const size_t nops = ops.size();
        const bool per_kernel = timed != 0;
        if (h->ev.size() < 2) {
            h->ev.resize(2);
            GG_CUDA(cudaEventCreate(&h->ev[0]));
            GG_CUDA(cudaEventCreate(&h->ev[1]));
        }
        if (per_kernel && h->ev.size() < 2 + 2 * nops) {
            size_t old = h->ev.size();
            h->ev.resize(2 + 2 * nops);
            for (size_t i = old; i < h->ev.size(); ++i) GG_CUDA(cudaEventCreate(&h->ev[i]));
        }
        upload_wave_steps(h);
        GG_CUDA(cudaEventRecord(h->ev[0], h->stream));
        launch_emax(h);
        for (size_t i = 0; i < nops; ++i) {
            const Op &op = ops[i];
            if (per_kernel) GG_CUDA(cudaEventRecord(h->ev[2 + 2 * i], h->stream));
            switch (op.kind) {
            case OP_EMIT: launch_emit(h, op); break;
            case OP_BWD_INIT: launch_step(h, op, false, true); break;
            case OP_BWD: launch_step(h, op, false, false); break;
            case OP_FWD: launch_step(h, op, true, false); break;
            case OP_RUN: launch_run(h, op); break;
            }
            if (per_kernel) GG_CUDA(cudaEventRecord(h->ev[3 + 2 * i], h->stream));
        }
        GG_CUDA(cudaEventRecord(h->ev[1], h->stream));
        GG_CUDA(cudaStreamSynchronize(h->stream));
        GG_CUDA(cudaEventElapsedTime(&st.last_sweep_ms, h->ev[0], h->ev[1]));
        st.last_fwdbwd_kernel_ms = 0.f;
        st.last_fwd_ms = st.last_bwd_ms = st.last_emit_ms = 0.f;
        if (per_kernel) {
            h->op_ms.assign(nops, 0.f);
            for (size_t i = 0; i < nops; ++i) {
                float ms = 0.f;
                GG_CUDA(cudaEventElapsedTime(&ms, h->ev[2 + 2 * i], h->ev[3 + 2 * i]));
                h->op_ms[i] = ms;
                switch (ops[i].kind) {
                case OP_EMIT: st.last_emit_ms += ms; break;
                case OP_FWD: st.last_fwd_ms += ms; break;
                case OP_RUN: (h->runs[ops[i].c].fwd ? st.last_fwd_ms : st.last_bwd_ms) += ms; break;
                default: st.last_bwd_ms += ms; break;
                }
            }
            st.last_fwdbwd_kernel_ms = st.last_fwd_ms + st.last_bwd_ms;
        }
    }
};

}  // namespace gg

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
template <typename F>
static int guarded(char *err, size_t errlen, F &&f) {
    try {
        f();
        return GG_OK;
    } catch (const CudaError &e) {
        set_err(err, errlen, e.what());
        return e.status;
    } catch (const std::bad_alloc &) {
        set_err(err, errlen, "out of host memory");
        return GG_ERR_NOMEM;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return GG_ERR_INVALID;
    }
}

extern "C" {

int gg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// budget for the state arena: the caller's, or 94 % of what is free (the retired arena of the last handle counts as free);
// the rest (>= 10 GB on a B200) is left for the static tables, the result, the CUDA context and NCCL
static uint64_t default_budget(const gg_options *o, int device) {
    uint64_t budget = o ? o->state_budget_bytes : 0;
    if (budget == 0) {
        size_t free_b = 0, total_b = 0;
        GG_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t cached = g_arena_cache.cached(2 * device);
        budget = (uint64_t)((free_b + cached) * 0.94);
        // free memory jitters by a few MB between calls; plan for the retired arena when it is about the size the
        // budget asks for, so that it is reused instead of being freed and allocated again (hundreds of ms at >100 GB)
        if (cached && budget > cached && cached >= budget - budget / 8) budget = cached;
        std::lock_guard<std::mutex> g(g_budget_mu);
        g_last_queried_budget[device] = budget;
    }
    return budget;
}

uint64_t gg_default_state_budget(int device) {
    try {
        if (device >= 0) GG_CUDA(cudaSetDevice(device));
        int dev = 0;
        GG_CUDA(cudaGetDevice(&dev));
        return default_budget(nullptr, dev);
    } catch (const std::exception &) {
        return 0;
    }
}

static void hmm_schedule(gg_hmm *h, uint64_t budget);

// Schedule `h` under the caller's budget or the default one.  cudaMemGetInfo takes a driver-wide lock and was seen to
// cost 5-75 ms now and then, more than the rest of gg_hmm_create together, so it is asked only when it can change the
// outcome: not when the retired arena already holds the whole problem without recomputation, and not when that arena
// IS the last default budget (there is nothing more to get).
static void schedule_default(gg_hmm *h, const gg_options *o, int device) {
    if (o && o->state_budget_bytes) {
        hmm_schedule(h, o->state_budget_bytes);
        return;
    }
    const uint64_t cached = g_arena_cache.cached(2 * device);
    if (cached) {
        uint64_t last = 0;
        {
            std::lock_guard<std::mutex> g(g_budget_mu);
            auto it = g_last_queried_budget.find(device);
            if (it != g_last_queried_budget.end()) last = it->second;
        }
        bool ok = true;
        try {
            hmm_schedule(h, cached);
        } catch (const std::exception &) {
            ok = false;                      // does not fit the retired arena at all: ask the driver
        }
        if (ok && (h->sched.levels <= 1 || (last && cached >= last - last / 8))) return;
    }
    hmm_schedule(h, default_budget(o, device));
}

static void need_device(const gg_options *o, int *device, int *num_sms) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        throw CudaError(GG_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
    }
    const int dev = o ? o->device : -1;
    if (dev >= 0) GG_CUDA(cudaSetDevice(dev));
    GG_CUDA(cudaGetDevice(device));
    GG_CUDA(cudaDeviceGetAttribute(num_sms, cudaDevAttrMultiProcessorCount, *device));
}

// host-only halves of create (no CUDA calls): validate + plan, then schedule + run fusion
static gg_hmm *hmm_plan(const gg_problem *p, const gg_options *o) {
    gg_hmm *h = new gg_hmm();
    try {
        const auto t_begin = std::chrono::steady_clock::now();
        build_plan(*p, h->plan);
        h->stats.create_plan_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        h->stats.create_total_ms = h->stats.create_plan_ms;
        h->fp64 = o && o->precision == 1;
        h->kernel_variant = o ? o->kernel_variant : 0;
        h->shared_arena = o && (o->flags & GG_OPT_SHARED_ARENA);
    } catch (...) {
        delete h;
        throw;
    }
    return h;
}

static void hmm_schedule(gg_hmm *h, uint64_t budget) {
    const auto t_begin = std::chrono::steady_clock::now();
    if (h->fp64) Engine<double>::prepare_host(h, budget);
    else Engine<float>::prepare_host(h, budget);
    h->stats.create_total_ms += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
}

static void hmm_create_device(gg_hmm *h, int device, int num_sms) {
    const auto t_begin = std::chrono::steady_clock::now();
    h->device = device;
    h->num_sms = num_sms;
    GG_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    if (h->fp64) Engine<double>::create_device(h);
    else Engine<float>::create_device(h);
    h->stats.create_total_ms += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
}

int gg_hmm_create(const gg_problem *p, const gg_options *o, gg_hmm **out, char *err, size_t errlen) {
    if (out) *out = nullptr;
    gg_hmm *h = nullptr;
    int rc = guarded(err, errlen, [&] {
        if (!p || !out) throw std::runtime_error("gg_hmm_create: null argument");
        const auto t0 = std::chrono::steady_clock::now();
        auto ms = [&] { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
        h = hmm_plan(p, o);          // validates the problem before any device work
        const float t1 = ms();
        int device = 0, num_sms = 0;
        need_device(o, &device, &num_sms);
        const float t2 = ms();
        const float t3 = ms();
        schedule_default(h, o, device);
        const float t4 = ms();
        hmm_create_device(h, device, num_sms);
        if (std::getenv("GG_TRACE"))
            std::fprintf(stderr, "[gg create] plan %.2f ms, device query %.2f, (%.2f) budget + schedule %.2f, device tables/arena/uploads %.2f\n",
                         t1, t2 - t1, t3 - t2, t4 - t3, ms() - t4);
    });
    if (rc != GG_OK) {
        delete h;
        return rc;
    }
    *out = h;
    return GG_OK;
}

int gg_hmm_sweep(gg_hmm *h, int timed_kernels, char *err, size_t errlen) {
    return guarded(err, errlen, [&] {
        if (!h) throw std::runtime_error("gg_hmm_sweep: null handle");
        GG_CUDA(cudaSetDevice(h->device));
        if (h->plan.n_columns == 0) return;
        if (h->fp64) Engine<double>::sweep(h, timed_kernels);
        else Engine<float>::sweep(h, timed_kernels);
    });
}

int gg_hmm_fetch(gg_hmm *h, gg_result **out, char *err, size_t errlen) {
    if (out) *out = nullptr;
    return guarded(err, errlen, [&] {
        if (!h || !out) throw std::runtime_error("gg_hmm_fetch: null argument");
        GG_CUDA(cudaSetDevice(h->device));
        gg_result *r = new gg_result();
        r->offsets = h->plan.gl_offsets;
        r->values.resize(h->plan.gl_offsets.empty() ? 0 : h->plan.gl_offsets.back());
        if (!r->values.empty()) {
            cudaError_t e = cudaMemcpyAsync(r->values.data(), h->d_result.p, r->values.size() * sizeof(double),
                                            cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) {
                delete r;
                throw CudaError(GG_ERR_CUDA, std::string("D2H of the genotype tables: ") + cudaGetErrorString(e));
            }
        }
        *out = r;
    });
}

int gg_hmm_stats(const gg_hmm *h, gg_stats *s) {
    if (!h || !s) return GG_ERR_INVALID;
    *s = h->stats;
    return GG_OK;
}

void gg_hmm_destroy(gg_hmm *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    delete h;
}

void gg_arena_release(void) {
    g_arena_cache.clear();
    std::lock_guard<std::mutex> g(g_shared_mu);
    for (auto &kv : g_shared_arenas) {
        std::lock_guard<std::mutex> g2(kv.second->mu);
        if (kv.second->p) cudaFree(kv.second->p);
        kv.second->p = nullptr;
        kv.second->n = 0;
    }
}

// STREAM-style device copy bandwidth (read + write bytes per second, best of `reps`), the same measurement that
// defines the roofline denominator; lets a benchmark report what THIS device does on the day.
double gg_device_copy_gbs(int device, uint64_t bytes, int reps) {
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return 0.0;
    unsigned char *a = nullptr, *b = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double best = 0.0;
    if (cudaMalloc(&a, bytes) == cudaSuccess && cudaMalloc(&b, bytes) == cudaSuccess && cudaEventCreate(&e0) == cudaSuccess &&
        cudaEventCreate(&e1) == cudaSuccess) {
        cudaMemset(a, 1, bytes);
        for (int i = 0; i < reps + 1; ++i) {
            cudaEventRecord(e0);
            cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice);
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) break;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (i > 0 && ms > 0.f) best = std::max(best, 2.0 * (double)bytes / (ms * 1e6));
        }
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (a) cudaFree(a);
    if (b) cudaFree(b);
    cudaGetLastError();
    return best;
}

// Independent chains back to back on one device; chain i+1 is planned and scheduled on a host thread while the device
// sweeps chain i (the reference plans and sweeps one chromosome after the other, giggles/cli/genotype.py:196-338).
int gg_hmm_run_batch(const gg_problem *problems, uint32_t n, const gg_options *o, gg_result **results, char *err, size_t errlen) {
    if (results)
        for (uint32_t i = 0; i < n; ++i) results[i] = nullptr;
    return guarded(err, errlen, [&] {
        if ((!problems || !results) && n) throw std::runtime_error("gg_hmm_run_batch: null argument");
        if (n == 0) return;
        std::unique_ptr<gg_hmm> first(hmm_plan(&problems[0], o));   // a malformed first chain is reported before any device work
        int device = 0, num_sms = 0;
        need_device(o, &device, &num_sms);
        const uint64_t budget = default_budget(o, device);
        // A few host threads take chains off a counter; each plans and schedules its chain (host only), then does the
        // device part.  Chains whose state arena is small (all-haplotagged chromosomes: one persistent CTA each) hold a
        // SHARED lock and therefore run side by side, each on its own stream; a chain that needs a large arena takes
        // the lock exclusively — there is room for one such arena — while the other threads go on planning the chains
        // behind it, so their host work stays hidden behind its sweep as before.  GG_BATCH_THREADS overrides the
        // thread count (1 = strictly one chain after the other).
        unsigned workers = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        if (const char *e = std::getenv("GG_BATCH_THREADS")) workers = (unsigned)std::max(1, std::atoi(e));
        workers = std::min<unsigned>(workers, n);
        const bool trace = std::getenv("GG_TRACE") != nullptr;
        std::atomic<uint32_t> next{0};
        std::atomic<bool> failed{false};
        std::shared_mutex device_gate;
        std::mutex err_mu;
        int first_rc = GG_OK;
        uint32_t first_chain = 0;
        std::string first_msg;
        auto record = [&](uint32_t i, int rc, const std::string &m) {
            std::lock_guard<std::mutex> g(err_mu);
            if (!failed.load() || i < first_chain) {
                first_rc = rc;
                first_chain = i;
                first_msg = m;
            }
            failed.store(true);
        };
        // one chain on the device at the given precision; returns its status (results[i] set on success)
        auto run_chain = [&](uint32_t i, std::unique_ptr<gg_hmm> h, const gg_options *oo, std::string &msg) -> int {
            const auto t0 = std::chrono::steady_clock::now();
            auto ms_since = [&] { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
            if (!h) h.reset(hmm_plan(&problems[i], oo));
            hmm_schedule(h.get(), budget);
            const float t_plan = ms_since();
            bool large = h->sched.arena_bytes > ArenaCache::kSmallArena;
            int rc = GG_OK;
            float t_gate = 0.f;
            for (int attempt = 0; attempt < 2; ++attempt) {
                std::unique_lock<std::shared_mutex> excl(device_gate, std::defer_lock);
                std::shared_lock<std::shared_mutex> shar(device_gate, std::defer_lock);
                if (large) excl.lock(); else shar.lock();
                t_gate = ms_since();
                char e2[512] = {0};
                try {
                    if (!h->stream) hmm_create_device(h.get(), device, num_sms);
                    rc = gg_hmm_sweep(h.get(), 0, e2, sizeof(e2));
                    if (rc == GG_OK) rc = gg_hmm_fetch(h.get(), &results[i], e2, sizeof(e2));
                    msg = e2;
                } catch (const CudaError &e) {
                    rc = e.status;
                    msg = e.what();
                }
                if (rc == GG_ERR_NOMEM && !large && attempt == 0) {
                    // small chains share the device and their arenas are not counted against each other: when one of
                    // them runs out of memory, let it try again alone instead of failing the batch
                    std::unique_ptr<gg_hmm> again(hmm_plan(&problems[i], oo));
                    cudaSetDevice(device);
                    h = std::move(again);
                    hmm_schedule(h.get(), budget);
                    large = true;
                    continue;
                }
                cudaSetDevice(device);
                h.reset();                 // the arena goes back to the cache before the gate opens
                break;
            }
            if (trace)
                std::fprintf(stderr, "[gg batch] chain %u: plan+schedule %.1f ms, waited %.1f ms for the device, device part %.1f ms (%s arena)\n",
                             i, t_plan, t_gate - t_plan, ms_since() - t_gate, large ? "large" : "small");
            return rc;
        };
        auto work = [&] {
            cudaSetDevice(device);
            for (;;) {
                const uint32_t i = next.fetch_add(1);
                if (i >= n || failed.load()) break;
                try {
                    std::string msg;
                    std::unique_ptr<gg_hmm> h;
                    if (i == 0) h = std::move(first);          // only one thread draws chain 0
                    int rc = run_chain(i, std::move(h), o, msg);
                    // the fp32 range guard of gg_hmm_run, per chain (one rerun on fp64 state, under the gate its
                    // larger arena calls for)
                    if (rc == GG_OK && !(o && o->precision == 1) && !all_finite(results[i])) {
                        gg_options o64{};
                        if (o) o64 = *o; else o64.device = -1;
                        o64.precision = 1;
                        gg_result *r32 = results[i];
                        results[i] = nullptr;
                        std::string msg64;
                        if (run_chain(i, nullptr, &o64, msg64) == GG_OK && results[i]) {
                            results[i]->reran_fp64 = true;
                            ++g_fp64_reruns;
                            gg_result_free(r32);
                        } else {
                            if (results[i]) gg_result_free(results[i]);
                            results[i] = r32;      // keep the fp32 result, as gg_hmm_run does
                        }
                    }
                    if (rc != GG_OK) record(i, rc, msg);
                } catch (const CudaError &e) {
                    record(i, e.status, e.what());
                } catch (const std::exception &e) {
                    record(i, GG_ERR_INVALID, e.what());
                }
            }
        };
        std::vector<std::thread> pool;
        for (unsigned w = 1; w < workers; ++w) pool.emplace_back(work);
        work();
        for (std::thread &t : pool) t.join();
        if (failed.load()) throw CudaError(first_rc, std::string("chain ") + std::to_string(first_chain) + ": " + first_msg);
    });
}

// per-op record of the last timed sweep: kind (OpKind), produced column (EMIT: first column; RUN: first produced column),
// duration in ms.  Returns the number of ops (may exceed cap).
uint64_t gg_hmm_debug_op_times(gg_hmm *h, uint32_t *kind, uint32_t *column, float *ms, uint64_t cap) {
    if (!h) return 0;
    const size_t n = std::min(h->exec_ops.size(), h->op_ms.size());
    for (size_t i = 0; i < n && i < cap; ++i) {
        const Op &op = h->exec_ops[i];
        kind[i] = (uint32_t)op.kind + (op.kind == OP_RUN && h->runs[op.c].fwd ? 100u : 0u);
        column[i] = op.kind != OP_RUN ? op.c : (h->runs[op.c].wave ? h->wave_ops[h->runs[op.c].first].c : h->chain_steps[h->runs[op.c].first].c_out);
        ms[i] = h->op_ms[i];
    }
    return n;
}

int gg_hmm_debug_scales(gg_hmm *h, double *alpha_scale, double *beta_scale) {
    if (!h) return GG_ERR_INVALID;
    const size_t n = h->plan.n_columns * sizeof(double);
    if (n == 0) return GG_OK;
    if (cudaMemcpy(alpha_scale, h->d_scale_a.p, n, cudaMemcpyDeviceToHost) != cudaSuccess) return GG_ERR_CUDA;
    if (cudaMemcpy(beta_scale, h->d_scale_b.p, n, cudaMemcpyDeviceToHost) != cudaSuccess) return GG_ERR_CUDA;
    return GG_OK;
}

int gg_hmm_run(const gg_problem *p, const gg_options *o, gg_result **out, char *err, size_t errlen) {
    if (out) *out = nullptr;
    gg_hmm *h = nullptr;
    const bool trace = std::getenv("GG_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_since = [&] { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    int rc = gg_hmm_create(p, o, &h, err, errlen);
    if (rc != GG_OK) return rc;
    const float t_create = ms_since();
    rc = gg_hmm_sweep(h, 0, err, errlen);
    const float t_sweep = ms_since();
    if (rc == GG_OK) rc = gg_hmm_fetch(h, out, err, errlen);
    const float t_fetch = ms_since();
    const bool was_fp32 = !h->fp64;
    const float dev_ms = h->stats.last_sweep_ms;
    const unsigned long long arena_b = h->arena_bytes;
    gg_hmm_destroy(h);
    if (trace)
        std::fprintf(stderr, "[gg run] create %.1f ms, sweep %.1f (device %.1f), fetch %.1f, destroy %.1f, arena %llu B\n", t_create,
                     t_sweep - t_create, dev_ms, t_fetch - t_sweep, ms_since() - t_fetch, arena_b);
    if (rc != GG_OK || !was_fp32) return rc;
    // Range guard of the fp32 path.  Every column's emission tables are shifted so that its best state has emission
    // ~1 (emission_kernel) and every stored column is rescaled by its sum, which keeps chains in range whatever the
    // reads say; this backstop remains for whatever that does not cover (the reference computes in 80-bit long
    // double).  A non-finite posterior triggers ONE rerun on fp64 state, reported by gg_result_reran_fp64();
    // if that is non-finite too (e.g. a column whose panel alleles are all missing, where the reference itself
    // returns NaN, SURVEY B-6) the fp64 result is returned as is.
    if (all_finite(*out)) return GG_OK;
    gg_options o64{};
    if (o) o64 = *o; else o64.device = -1;
    o64.precision = 1;
    gg_result *r64 = nullptr;
    gg_hmm *h64 = nullptr;
    char err2[256];
    if (gg_hmm_create(p, &o64, &h64, err2, sizeof(err2)) != GG_OK) return GG_OK;      // keep the fp32 result
    int rc2 = gg_hmm_sweep(h64, 0, err2, sizeof(err2));
    if (rc2 == GG_OK) rc2 = gg_hmm_fetch(h64, &r64, err2, sizeof(err2));
    gg_hmm_destroy(h64);
    if (rc2 == GG_OK && r64) {
        gg_result_free(*out);
        r64->reran_fp64 = true;
        ++g_fp64_reruns;
        *out = r64;
    }
    return GG_OK;
}

uint64_t gg_fp64_rerun_count(void) { return g_fp64_reruns.load(); }

}  // extern "C"
