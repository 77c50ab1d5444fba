// Device/host helpers shared by the sampled (sampled.cu) and closed-form (closed.cu) steps.
#pragma once
#include "common.cuh"
#include "internal.h"

namespace vfmb {

constexpr int kGridCap = 2048;      // most blocks a step kernel is launched with (block partials are sized for it)

struct DevCfg {
    int B, F, d, S, n_classes;
    int row_stride, row_offset;                          // global id = row*stride + offset (sharding)
    int class_bound[kMaxFields];
    float class_size[kMaxFields];
    float n_train;
    uint64_t seed;
};

static DevCfg make_dev(const vfmb_config* c) {
    DevCfg r{};
    r.B = c->B; r.F = c->F; r.d = c->d; r.S = c->S; r.n_classes = c->n_classes;
    for (int i = 0; i < kMaxFields; ++i) { r.class_bound[i] = c->class_bound[i]; r.class_size[i] = c->class_size[i]; }
    r.n_train = c->n_train; r.seed = c->seed;
    r.row_stride = c->row_stride > 0 ? c->row_stride : 1;
    r.row_offset = c->row_stride > 0 ? c->row_offset : 0;
    return r;
}

__device__ __forceinline__ int class_of(const DevCfg& c, int row) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < kMaxFields - 1; ++i) k += (i < c.n_classes - 1 && row >= c.class_bound[i]) ? 1 : 0;
    return k;
}

// eps for the VEC elements starting at k of row `rowid` (unique rank u)

// block-wide deterministic reduction of K per-thread doubles into partials[block][K]; returns
// true in every thread of the block that arrived last (which then owns the final reduction)
template <int K>
__device__ __forceinline__ bool block_partials(const double (&acc)[K], double* __restrict__ partials,
                                               int32_t* __restrict__ counter) {
    __shared__ double s_red[K][8];
    __shared__ bool s_last;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double s = warp_sum(acc[i]);
        if (lane == 0) s_red[i][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += s_red[threadIdx.x][w];
        partials[(size_t)blockIdx.x * K + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

// Final reduction by the last block: thread t adds the partials of blocks t, t+256, ... in that
// order, then the 256 per-thread sums are combined by a fixed shuffle/shared-memory tree.  The
// association order depends only on gridDim, so the result is bitwise reproducible.
// Must be called by every thread of the (256-thread) block; result valid in thread 0.
template <int K>
__device__ __forceinline__ void final_sums(const double* __restrict__ partials, double (&out)[K]) {
    __shared__ double s_fin[K][8];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double s = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) s += __ldcg(partials + (size_t)b * K + i);
        s = warp_sum(s);
        if (lane == 0) s_fin[i][warp] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += s_fin[i][w];
        out[i] = s;
    }
}

// Work distribution of the step kernels ("hybrid"): a warp takes CH = 4*GPW consecutive units
// (unique rows / samples / work items; GPW = 32/LPR rows fit in a warp).  Everything that is one
// scalar per unit -- record decode, bias row, train count, bias noise (one Philox block),
// likelihood, bias Adam -- is done lane-parallel by the first CH lanes with coalesced loads;
// the warp then walks the CH units in 4 rounds, GPW rows per round with LPR lanes per row for the
// wide (d-element) work, fetching each unit's scalars by shuffle.  Rows of the later rounds are
// prefetched into L2 up front.  This keeps ~50 warps per SM busy while taking the per-row scalar
// code out of the wide path, where it ran at 1/LPR lane efficiency.
#ifndef VFMB_ROUNDS
#define VFMB_ROUNDS 4
#endif
constexpr int kRounds = VFMB_ROUNDS;
#define GPW_OF(LPR) (32 / (LPR))

template <typename T>
__device__ __forceinline__ T bcast(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// hand a per-group value back to the lane that owns unit (it*GPW + g)
template <int LPR, typename T>
__device__ __forceinline__ void hand_back(T& mine, T val, int it, int lane) {
    constexpr int GPW = kWarp / LPR;
#pragma unroll
    for (int g = 0; g < GPW; ++g) {
        T v = __shfl_sync(0xffffffffu, val, g * LPR);
        if (lane == it * GPW + g) mine = v;
    }
}

__device__ __forceinline__ void prefetch_row(const float* p, int bytes) {
    for (int off = 0; off < bytes; off += 128) prefetch_l2(reinterpret_cast<const char*>(p) + off);
}


// hyper-parameters as torch rounds them: every derived constant is formed in double first
struct AdamDev {
    double lr, beta1, beta2;
    float b2, omb1, omb2, eps;      // beta2, 1-beta1, 1-beta2, eps as fp32
};
static AdamDev make_adam(const vfmb_adam* a) {
    AdamDev h{};
    double lr = a ? a->lr : 1e-3, b1 = a ? a->beta1 : 0.9, b2 = a ? a->beta2 : 0.999, e = a ? a->eps : 1e-8;
    h.lr = lr; h.beta1 = b1; h.beta2 = b2;
    h.b2 = (float)b2; h.omb1 = (float)(1.0 - b1); h.omb2 = (float)(1.0 - b2); h.eps = (float)e;
    return h;
}

// step_size = lr / (1 - beta1^t), inv_bc2_sqrt = 1 / sqrt(1 - beta2^t)
__device__ __forceinline__ void adam_coeffs(const AdamDev& h, int t, float* step_size, float* inv_bc2_sqrt) {
    double bc1 = 1.0 - pow(h.beta1, (double)t);
    double bc2 = 1.0 - pow(h.beta2, (double)t);
    *step_size = (float)(h.lr / bc1);
    *inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
}

// torch _single_tensor_adam: m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2);
// p.addcdiv_(m, sqrt(v)/bc2_sqrt + eps, -step_size)
__device__ __forceinline__ void adam_elem(float& p, float& m, float& v, float g, const AdamDev& h,
                                          float step_size, float inv_bc2_sqrt) {
    m = fmaf(g - m, h.omb1, m);
    v = fmaf(v, h.b2, h.omb2 * g * g);
    float denom = fmaf(fast_sqrt(v), inv_bc2_sqrt, h.eps);
    p = fmaf(-step_size, m * fast_rcp(denom), p);
}

// Backward, phase B: per unique row, chain rule from (g_v, g_w) to (mean, raw scale) plus the KL
// gradient, then Adam on the row (or the dense-gradient store).  This is the HBM-bound kernel of
// the step: parameters and both Adam moments of every touched row are read and written once.

// ------------------------------------------------------------------------------- k_combine
// Rows cut by tile boundaries: add the tile partials in tile order (tail slot of the first tile,
// then the head slots of the following tiles).  A block scans 256 rows.  Rows with few partials
// are handled per warp (GPW lane groups take contiguous ranges, group sums added in group order);
// the rare hot rows (> kHotPartials tiles: a Zipf head row with thousands of occurrences) are
// handled by the whole block, 8*GPW groups over contiguous ranges and a fixed-order shared-memory
// reduction.  Every association is fixed by the plan, so the result is bitwise reproducible.

template <int VEC, int LPR, int NV, int UNR>
__device__ __forceinline__ void sum_head_slots(const float* __restrict__ gslot, int dp, int d, int gl,
                                               int lo, int hi, Vec<VEC> (&acc)[NV], float& gw) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[i].v[j] = 0.f;
    gw = 0.f;
    for (int t = lo; t < hi; t += UNR) {
        Vec<VEC> part[UNR][NV]; float pw[UNR];
#pragma unroll
        for (int q = 0; q < UNR; ++q) {
            const int tt = min(t + q, hi - 1);
            const float* sp = gslot + ((size_t)tt * 2) * dp;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                int k = (gl + i * LPR) * VEC;
                if (k < d) part[q][i] = ld_vec_nc<VEC>(sp + k);
            }
            pw[q] = __ldg(sp + d);
        }
#pragma unroll
        for (int q = 0; q < UNR; ++q)
            if (t + q < hi) {
#pragma unroll
                for (int i = 0; i < NV; ++i)
#pragma unroll
                    for (int j = 0; j < VEC; ++j) acc[i].v[j] += part[q][i].v[j];
                gw += pw[q];
            }
    }
}

// one row cut by few tile boundaries, summed by a warp: its GPW lane groups take contiguous slot
// ranges; total = tail(tA) + group 0 + group 1 + ... (fixed order)
template <int VEC, int LPR, int NV>
__device__ __forceinline__ void combine_warp_row(int u, int tA, int tB, int d, int F,
                                                 const float* __restrict__ gslot, const float* __restrict__ vs,
                                                 float* __restrict__ grow, float* __restrict__ gws) {
    constexpr int GPW = kWarp / LPR;
    const int dp = d + 4;
    const int lane = threadIdx.x & 31, gl = lane % LPR, gidx = lane / LPR;
    const int per = (tB - tA + GPW - 1) / GPW;
    const int lo = tA + 1 + gidx * per, hi = min(tB + 1, lo + per);
    Vec<VEC> acc[NV]; float gw;
    sum_head_slots<VEC, LPR, NV, 4>(gslot, dp, d, gl, lo, hi, acc, gw);
    // total = tail(tA) + group 0 + group 1 + ...
    const float* tp = gslot + ((size_t)tA * 2 + 1) * dp;
    float gw_tot = __ldg(tp + d);
#pragma unroll
    for (int g = 0; g < GPW; ++g) gw_tot += __shfl_sync(0xffffffffu, gw, g * LPR);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        int k = (gl + i * LPR) * VEC;
        Vec<VEC> tot;
        if (k < d) tot = ld_vec_nc<VEC>(tp + k);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
#pragma unroll
            for (int g = 0; g < GPW; ++g) {
                float v = __shfl_sync(0xffffffffu, acc[i].v[j], g * LPR + gl);
                if (k < d) tot.v[j] += v;
            }
        }
        if (k < d && gidx == 0) {
            if (F > 2) {                           // pairwise: sum r_n (S_n - v_u)
                Vec<VEC> own = ld_vec_nc<VEC>(vs + (size_t)u * d + k);
#pragma unroll
                for (int j = 0; j < VEC; ++j) tot.v[j] = fmaf(-gw_tot, own.v[j], tot.v[j]);
            }
            st_vec<VEC>(grow + (size_t)u * d + k, tot);
        }
    }
    if (lane == 0) gws[u] = gw_tot;
}

// one hot row (> kHotPartials tile partials), summed by the whole block: 8*GPW lane groups over
// contiguous slot ranges, then a fixed-order shared-memory reduction by one group
template <int VEC, int LPR, int NV>
__device__ __forceinline__ void combine_hot_row(int u, int d, int F, const int32_t* __restrict__ urec,
                                                const float* __restrict__ gslot, const float* __restrict__ vs,
                                                float* __restrict__ grow, float* __restrict__ gws, float* s_part) {
    constexpr int GPW = kWarp / LPR, NG = 8 * GPW;
    const int dp = d + 4;
    const int lane = threadIdx.x & 31, gl = lane % LPR, gidx = lane / LPR, warp = threadIdx.x >> 5;
    const int4 rec = __ldg(reinterpret_cast<const int4*>(urec) + u);
    const int tA = rec.z / kTile, tB = (rec.z + rec.y - 1) / kTile;
    const int g = warp * GPW + gidx;                        // group id 0 .. NG-1
    const int per = (tB - tA + NG - 1) / NG;
    const int lo = min(tB + 1, tA + 1 + g * per), hi = min(tB + 1, lo + per);
    Vec<VEC> acc[NV]; float gw;
    sum_head_slots<VEC, LPR, NV, 8>(gslot, dp, d, gl, lo, hi, acc, gw);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        int k = (gl + i * LPR) * VEC;
        if (k < d) st_vec<VEC>(s_part + (size_t)g * dp + k, acc[i]);
    }
    if (gl == 0) s_part[(size_t)g * dp + d] = gw;
    __syncthreads();
    if (warp == 0 && gidx == 0) {                          // one lane group adds the NG sums in order
        const float* tp = gslot + ((size_t)tA * 2 + 1) * dp;
        float gw_tot = __ldg(tp + d);
        for (int q = 0; q < NG; ++q) gw_tot += s_part[(size_t)q * dp + d];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            int k = (gl + i * LPR) * VEC;
            if (k < d) {
                Vec<VEC> tot = ld_vec_nc<VEC>(tp + k);
                for (int q = 0; q < NG; ++q)
#pragma unroll
                    for (int j = 0; j < VEC; ++j) tot.v[j] += s_part[(size_t)q * dp + k + j];
                if (F > 2) {
                    Vec<VEC> own = ld_vec_nc<VEC>(vs + (size_t)u * d + k);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) tot.v[j] = fmaf(-gw_tot, own.v[j], tot.v[j]);
                }
                st_vec<VEC>(grow + (size_t)u * d + k, tot);
            }
        }
        if (gl == 0) gws[u] = gw_tot;
    }
    __syncthreads();
}

// every row cut by a tile boundary, from the plan's lists (no scan over the unique rows): the
// `cut` list holds the rows with <= kHotPartials partials from its front (meta[5] entries, one
// warp each), the hot rows from its back (meta[3] entries, one block each).
template <int VEC, int LPR, int NV>
__global__ void __launch_bounds__(256)
k_combine_cut(int d, int F, const int32_t* __restrict__ urec, const int32_t* __restrict__ meta,
              const int32_t* __restrict__ cut, int cut_cap, const float* __restrict__ gslot,
              const float* __restrict__ vs, float* __restrict__ grow, float* __restrict__ gws) {
    extern __shared__ float s_part[];                    // [8*GPW][d+4] block-level partial sums
    const int n_hot = meta[3], n_cut = meta[5];
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = gridDim.x * (blockDim.x >> 5);
    for (int i = gwarp; i < n_cut; i += nwarps) {
        const int u = __ldg(cut + i);
        const int4 rec = __ldg(reinterpret_cast<const int4*>(urec) + u);
        combine_warp_row<VEC, LPR, NV>(u, rec.z / kTile, (rec.z + rec.y - 1) / kTile, d, F, gslot, vs, grow, gws);
    }
    __syncthreads();
    for (int i = blockIdx.x; i < n_hot; i += gridDim.x)
        combine_hot_row<VEC, LPR, NV>(__ldg(cut + cut_cap - 1 - i), d, F, urec, gslot, vs, grow, gws, s_part);
}

template <int VEC, int LPR, int NV, int HOT_ONLY = 0>
__global__ void __launch_bounds__(256)
k_combine(int d, int F, const int32_t* __restrict__ urec, const int32_t* __restrict__ meta,
          const float* __restrict__ gslot, const float* __restrict__ vs,
          float* __restrict__ grow, float* __restrict__ gws) {
    constexpr int GPW = kWarp / LPR, NG = 8 * GPW;
    extern __shared__ float s_part[];                    // [NG][dp] block-level partial sums
    __shared__ int s_hot[256];
    __shared__ int s_nhot;
    const int U = meta[0];
    const int dp = d + 4;
    const int lane = threadIdx.x & 31, gl = lane % LPR, gidx = lane / LPR, warp = threadIdx.x >> 5;
    for (int bbase = blockIdx.x * 256; bbase < U; bbase += gridDim.x * 256) {
        if (threadIdx.x == 0) s_nhot = 0;
        __syncthreads();
        const int base = bbase + warp * 32;
        const int ul = base + lane;
        int tA_l = 0, tB_l = 0;
        if (ul < U) {
            const int4 rec = __ldg(reinterpret_cast<const int4*>(urec) + ul);
            tA_l = rec.z / kTile;                          // first / last tile of the row's segment
            tB_l = (rec.z + rec.y - 1) / kTile;
        }
        const int P_l = tB_l - tA_l;
        if (P_l > kHotPartials) s_hot[atomicAdd(&s_nhot, 1)] = ul;   // order only affects scheduling
        // HOT_ONLY: the rows with few partials are summed by the consumer (k_adam_rows<FLAVOR >= 1>)
        unsigned todo = HOT_ONLY ? 0u : __ballot_sync(0xffffffffu, P_l > 0 && P_l <= kHotPartials);
        // ---- per-warp: rows with few partials
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int u = base + src;
            const int tA = __shfl_sync(0xffffffffu, tA_l, src), tB = __shfl_sync(0xffffffffu, tB_l, src);
            combine_warp_row<VEC, LPR, NV>(u, tA, tB, d, F, gslot, vs, grow, gws);
        }
        __syncthreads();
        // ---- whole block: hot rows
        const int nhot = s_nhot;
        for (int hi_ = 0; hi_ < nhot; ++hi_)
            combine_hot_row<VEC, LPR, NV>(s_hot[hi_], d, F, urec, gslot, vs, grow, gws, s_part);
        __syncthreads();                                   // s_nhot / s_hot are reused by the next pass
    }
}


#define VFMB_LAYOUT_SWITCH(L, ...)                                                     \
    do {                                                                               \
        if (L.vec == 4 && L.lpr == 4 && L.nv == 1) { constexpr int VEC = 4, LPR = 4, NV = 1; __VA_ARGS__; }        \
        else if (L.vec == 4 && L.lpr == 8 && L.nv == 1) { constexpr int VEC = 4, LPR = 8, NV = 1; __VA_ARGS__; }   \
        else if (L.vec == 4 && L.lpr == 16 && L.nv == 1) { constexpr int VEC = 4, LPR = 16, NV = 1; __VA_ARGS__; } \
        else if (L.vec == 4 && L.lpr == 32 && L.nv == 1) { constexpr int VEC = 4, LPR = 32, NV = 1; __VA_ARGS__; } \
        else if (L.vec == 4 && L.lpr == 32 && L.nv == 2) { constexpr int VEC = 4, LPR = 32, NV = 2; __VA_ARGS__; } \
        else if (L.vec == 1 && L.lpr == 4 && L.nv == 1) { constexpr int VEC = 1, LPR = 4, NV = 1; __VA_ARGS__; }   \
        else if (L.vec == 1 && L.lpr == 8 && L.nv == 1) { constexpr int VEC = 1, LPR = 8, NV = 1; __VA_ARGS__; }   \
        else if (L.vec == 1 && L.lpr == 16 && L.nv == 1) { constexpr int VEC = 1, LPR = 16, NV = 1; __VA_ARGS__; } \
        else if (L.vec == 1 && L.lpr == 32 && L.nv == 1) { constexpr int VEC = 1, LPR = 32, NV = 1; __VA_ARGS__; } \
        else if (L.vec == 1 && L.lpr == 32 && L.nv == 2) { constexpr int VEC = 1, LPR = 32, NV = 2; __VA_ARGS__; } \
        else return set_error(VFMB_ESHAPE, "unsupported embedding size %d", cfg->d);   \
    } while (0)


// Carving of vfmb_step_io.partials (doubles): block partials of the scalar reductions, the tile
// head/tail slots of the gather kernels, and (closed form) the prior-gradient partials.
struct ScratchMap {
    size_t gslot_off, pg_part_off, pg_sum_off, blk_class_off, total_doubles;   // offsets in floats
    int nblk_max, pg_width;
};
static inline ScratchMap scratch_map(int B, int F, int d, int64_t u_cap) {
    ScratchMap m{};
    const int64_t n = (int64_t)B * F;
    const int64_t slots = 2 * ((n + kTile - 1) / kTile + 1);
    m.nblk_max = (int)(u_cap / 32 + kMaxFields + 1);
    const size_t nred = (size_t)(m.nblk_max > kGridCap ? m.nblk_max : kGridCap);
    size_t off = nred * 16 * 2;                             // block partials (doubles), in floats
    m.gslot_off = off;        off += (size_t)slots * (3 * d + 4);
    m.pg_width = 2 * d + 4;
    m.pg_part_off = off;      off += (size_t)m.nblk_max * m.pg_width;
    m.pg_sum_off = off;       off += (size_t)kMaxFields * m.pg_width;
    m.blk_class_off = off;    off += (size_t)m.nblk_max + 16;
    m.total_doubles = (off + 1) / 2;
    return m;
}

// Grid of a (grid-stride) step kernel: one persistent resident wave (blocks/SM from the occupancy
// calculator).  (One warp per `per_warp` units -- as many blocks as the work needs -- measured 4 %
// slower on the ml20m step: more block prologues, no gain from hardware load balancing.)
template <typename K>
static inline int resident_blocks(K kernel, int block, size_t smem) {
    static int cached = 0;                                  // one instance per kernel type
    if (cached == 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        cached = per_sm * kNumSMs;
    }
    return cached;
}
template <typename K>
static inline int grid_resident(K kernel, int64_t units, int per_warp, size_t smem = 0, bool may_reserve = true) {
    int64_t warps = (units + per_warp - 1) / per_warp;
    int64_t g = (warps + 7) / 8;
    // leave `reserve` block slots per SM free: room for the plan kernels running concurrently
    const int reserve = may_reserve ? grid_reserve() : 0;
    int cap = resident_blocks(kernel, 256, smem);
    if (reserve > 0 && cap / kNumSMs > reserve + 1) cap -= reserve * kNumSMs;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

// grid for a kernel whose warps each take `per_warp` units out of at most `units`
static inline int grid_warps(int64_t units, int per_warp) {
    int64_t warps = (units + per_warp - 1) / per_warp;
    int64_t g = (warps + 7) / 8;
    if (g < 1) g = 1;
    if (g > kMaxGrid) g = kMaxGrid;
    return (int)g;
}

}  // namespace vfmb
