// Device/host helpers shared by the sampled (sampled.cu) and closed-form (closed.cu) steps.
#pragma once
#include "common.cuh"
#include "internal.h"

namespace vfmb {

constexpr int kGridCap = 2048;      // most blocks a step kernel is launched with (block partials are sized for it)

struct DevCfg {
    int B, F, d, S, n_classes;
    int pairwise;                                        // row gradients hold sum r_n S_n: the row kernel removes (sum r_n) v_u
    int row_stride, row_offset;                          // global id = row*stride + offset (sharding)
    int class_bound[kMaxFields];
    float class_size[kMaxFields];
    float n_train;
    uint64_t seed;
};

static DevCfg make_dev(const vfmb_config* c) {
    DevCfg r{};
    r.B = c->B; r.F = c->F; r.d = c->d; r.S = c->S; r.n_classes = c->n_classes;
    // F > 2 fields are scored pairwise; F == 1 is the owner side of a row-sharded step, which
    // is told through `interaction` what its requesters' model is
    r.pairwise = (c->F > 2 || (c->F == 1 && c->interaction == VFMB_INTER_PAIRWISE)) ? 1 : 0;
    for (int i = 0; i < kMaxFields; ++i) { r.class_bound[i] = c->class_bound[i]; r.class_size[i] = c->class_size[i]; }
    r.n_train = c->n_train; r.seed = c->seed;
    r.row_stride = c->row_stride > 0 ? c->row_stride : 1;
    r.row_offset = c->row_stride > 0 ? c->row_offset : 0;
    return r;
}

// ---- mode B (row-sharded tables): destinations in other ranks' exchange buffers (NVLink peer memory)
// n == 0: not in use.  n == P: base[q] = the exchange region of rank q mapped into this process; this
// rank owns chunk `rank` ([CAP] slots) of every region, slot (q, j) -> base[q] + (rank * CAP + j) * width.
struct Peers {
    void* base[kMaxFields];
    int n, rank;
};
// k_stage, owner side: the sampled row of unique rank u goes straight into the slot of every rank that
// asked for it (sorted occurrences of the owner's plan = request slots s; requester = s / CAP)
struct RowPut {
    Peers pe;
    const int32_t* occ;
    int CAP, SP, n_real;         // slot pitch in floats (d + 4: [row | bias, pad]); rows >= n_real are padding
};
// k_score, requester side: the rank's additive scalars {NLL sum, residual sum, squared-error sum, KL of
// its owned rows, overflow flag} into slot `rank` of every rank's tail region (summed in rank order there)
struct TailPut {
    Peers pe;
    const float* stats_owner;    // stats of this rank's owner-side stage (KL of the rows it owns)
    const int32_t* overflow;     // sticky bucket-overflow flag of this rank
    float n_local;
    int pitch;
};

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

__device__ __forceinline__ void prefetch_row(const float* p, int bytes, bool keep = false) {
    if (keep) for (int off = 0; off < bytes; off += 128) prefetch_l2_keep(reinterpret_cast<const char*>(p) + off);
    else for (int off = 0; off < bytes; off += 128) prefetch_l2(reinterpret_cast<const char*>(p) + off);
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

// ------------------------------------------------------------------------------- cut rows
// The backward's segmented reduction walks the sorted occurrence list in tiles: k_gather in block tiles
// of 512 positions (rows cut inside a block never get here: shared memory), k_cgather in tiles of kTile
// positions (tile_span below: rows of up to `keep` occurrences are kept whole).  A row cut by the tile
// boundaries leaves one partial per
// tile it touches: in the tail slot of its first tile, in the head slots of the following ones.
// The lane group that stores the LAST partial adds them -- no separate combine launch (12 us of
// pure latency on the ml20m step in round 1).  To keep that serial sum short for a Zipf head row
// (thousands of occurrences = hundreds of tiles) it is a two-level tree with a fixed shape:
//   level 1  the row's partials in runs of kFan consecutive tiles; the group that completes a run
//            (arrival counter of the run, indexed by its first tile: a tile boundary cuts at most
//            one row, so (tile, head/tail) names a run) adds the run's slots in tile order;
//   level 2  a row with more than kFan partials: the run sums are written back to the first slot
//            of each run, and the group that completes the last run adds them in run order.
// The association is fixed by the plan (tile order inside a run, run order across), so the result
// is bitwise reproducible whichever groups happen to finish.  NW = d-wide sections per slot
// (1: sampled step; 3: closed form [A | Bq | C]).  arrive: int32 [3 * n_tiles1], zero between steps
// (every counter is reset by the group that completes it).
constexpr int kFan = 32;

template <int VEC, int LPR, int NV, int NW>
__device__ __forceinline__ void sum_slots(const float* gslot, int dp, int d, int gl, int t_first, bool first_is_tail,
                                          int count, int stride, Vec<VEC> (&tot)[NW][NV], float& gw) {
    constexpr int UNR = (NW == 1) ? (NV == 1 ? 4 : 2) : 1;
    const float* tp = gslot + ((size_t)t_first * 2 + (first_is_tail ? 1 : 0)) * dp;
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int k = (gl + i * LPR) * VEC;
            if (k < d) tot[w][i] = ld_vec_cg<VEC>(tp + w * d + k);
        }
    gw = __ldcg(tp + NW * d);
    for (int q0 = 1; q0 < count; q0 += UNR) {
        Vec<VEC> part[UNR][NW][NV]; float pw[UNR];
#pragma unroll
        for (int q = 0; q < UNR; ++q) {
            const float* sp = gslot + ((size_t)(t_first + min(q0 + q, count - 1) * stride) * 2) * dp;   // head slots
#pragma unroll
            for (int w = 0; w < NW; ++w)
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int k = (gl + i * LPR) * VEC;
                    if (k < d) part[q][w][i] = ld_vec_cg<VEC>(sp + w * d + k);
                }
            pw[q] = __ldcg(sp + NW * d);
        }
#pragma unroll
        for (int q = 0; q < UNR; ++q)
            if (q0 + q < count) {
#pragma unroll
                for (int w = 0; w < NW; ++w)
#pragma unroll
                    for (int i = 0; i < NV; ++i)
#pragma unroll
                        for (int j = 0; j < VEC; ++j) tot[w][i].v[j] += part[q][w][i].v[j];
                gw += pw[q];
            }
    }
}

// Tile bounds, snapped to row boundaries: a row of at most kTile occurrences is never cut -- it
// belongs whole to the tile in which it starts -- so only longer rows (the Zipf head) leave partials and
// pay for the finisher's fences and atomics.  Tile t nominally covers positions [t kTile, (t+1) kTile);
// a boundary that falls inside a short row moves to the end of that row (the tile on its left takes the
// whole row, at most kTile - 1 extra positions; a tile can come out empty).  head_u / tail_u: the long
// row cut at the tile's start / end, or -1.  Every lane computes the same values (broadcast loads).
struct TileSpan { int t0, t1, head_u, tail_u; };
// launch knobs of the gather kernels (host side: Tuning); none of them changes a result bit EXCEPT keep,
// which moves the cuts (and with them the association of a long row's sum)
struct GatherKnobs { int keep; int dyn; int light_fence; };
__device__ __forceinline__ int snap_boundary(int b, int N, int keep, const int32_t* __restrict__ pos_rank,
                                             const int32_t* __restrict__ urec, int* cut_u) {
    *cut_u = -1;
    if (b <= 0) return 0;
    if (b >= N) return N;
    const int u = __ldg(pos_rank + b);
    if (__ldg(pos_rank + b - 1) != u) return b;
    const int4 rec = __ldg(reinterpret_cast<const int4*>(urec) + u);
    if (rec.y <= keep) return rec.z + rec.y;
    *cut_u = u;
    return b;
}
__device__ __forceinline__ TileSpan tile_span(int tile, int N, int keep, const int32_t* __restrict__ pos_rank,
                                              const int32_t* __restrict__ urec) {
    TileSpan s;
    s.t0 = snap_boundary(tile * kTile, N, keep, pos_rank, urec, &s.head_u);
    s.t1 = snap_boundary((tile + 1) * kTile, N, keep, pos_rank, urec, &s.tail_u);
    return s;
}
// release / acquire around the arrival counters: fence.acq_rel is enough for the last-arriver pattern (the
// partials are stored before the fence, counted after it; read after the counter and a second fence);
// __threadfence() is fence.sc
__device__ __forceinline__ void fence_gpu(bool light) {
    if (light) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    else __threadfence();
}

// called by every lane of a group after it stored the partial of cut row u for `tile`
template <int VEC, int LPR, int NV, int NW>
__device__ __noinline__ void finish_cut_row(int u, int tile, int d, const float* __restrict__ own_row,
                                            const int32_t* __restrict__ urec, float* gslot, float* __restrict__ out_row,
                                            float* __restrict__ out_w, int32_t* arrive, int n_tiles1,
                                            bool light_fence, int tile_size = kTile) {
    // own_row (F > 2, pairwise): the row's own sampled vector, removed from the sum; out_row / out_w:
    // where the finished gradient goes (the per-rank scratch, or the owner's slot over NVLink)
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const unsigned gmask = group_mask<LPR>();
    const int dp = NW * d + 4;
    fence_gpu(light_fence);                                        // this group's partial is visible ...
    __syncwarp(gmask);
    const int4 rec = __ldg(reinterpret_cast<const int4*>(urec) + u);
    const int tA = rec.z / tile_size, tB = (rec.z + rec.y - 1) / tile_size;     // cut rows are cut at the nominal boundaries
    const int run = (tile - tA) / kFan;
    const int first = tA + run * kFan, last = min(tB, first + kFan - 1);
    int32_t* c1 = arrive + 2 * first + (run == 0 ? 1 : 0);
    int old = 0;
    if (gl == 0) old = atomicAdd(c1, 1);                    // ... before it is counted
    old = __shfl_sync(gmask, old, 0, LPR);
    if (old != last - first) return;                        // the group that completes the run goes on
    fence_gpu(light_fence);
    Vec<VEC> tot[NW][NV]; float gw;
    sum_slots<VEC, LPR, NV, NW>(gslot, dp, d, gl, first, run == 0, last - first + 1, 1, tot, gw);
    if (gl == 0) *c1 = 0;                                   // counter ready for the next step
    if (tB - tA >= kFan) {                                  // more than one run: second level
        float* sp = gslot + ((size_t)first * 2 + (run == 0 ? 1 : 0)) * dp;
#pragma unroll
        for (int w = 0; w < NW; ++w)
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int k = (gl + i * LPR) * VEC;
                if (k < d) st_vec<VEC>(sp + w * d + k, tot[w][i]);
            }
        if (gl == 0) sp[NW * d] = gw;
        fence_gpu(light_fence);
        __syncwarp(gmask);
        int32_t* c2 = arrive + 2 * n_tiles1 + tA;
        const int n_runs = (tB - tA) / kFan + 1;
        if (gl == 0) old = atomicAdd(c2, 1);
        old = __shfl_sync(gmask, old, 0, LPR);
        if (old != n_runs - 1) return;
        fence_gpu(light_fence);
        sum_slots<VEC, LPR, NV, NW>(gslot, dp, d, gl, tA, true, n_runs, kFan, tot, gw);
        if (gl == 0) *c2 = 0;
    }
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int k = (gl + i * LPR) * VEC;
            if (k < d) {
                if (NW == 1 && own_row) {                   // pairwise (F > 2): sum r_n (S_n - v_u)
                    const Vec<VEC> own = ld_vec_nc<VEC>(own_row + k);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) tot[w][i].v[j] = fmaf(-gw, own.v[j], tot[w][i].v[j]);
                }
                st_vec<VEC>(out_row + w * d + k, tot[w][i]);
            }
        }
    if (gl == 0) *out_w = gw;
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
    size_t gslot_off, pg_part_off, pg_sum_off, blk_class_off, arrive_off, total_doubles;   // offsets in floats
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
    m.arrive_off = off;       off += 3 * ((size_t)slots / 2) + 16;   // arrival counters of the cut rows (int32 [3][n_tiles
                                                                    // + 1], zero-initialised with the buffer, self-resetting)
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
