// Batch plan: sorted unique rows, inverse map, segments, work items.
//
// Replaces the three torch.unique calls of vfm-torch.py:190-192 and the
// per-group ones of vfm-tomasrch.py:537-545.  Integer outputs are bit-identical
// to torch.unique(sorted=True): the (row id, occurrence) pairs are radix-sorted
// by row id with a stable sort, so inside a segment the occurrences stay in
// ascending batch order -- that fixed order is what makes the segmented
// gradient reduction deterministic.
//
//
// The sort is an own tiled LSD radix sort (no library calls on the plan path): with B*F of a few
// hundred thousand keys the problem is latency-bound, so it is built from few, wide-grid kernels
// with plain dependencies instead of a chained look-back --
//   k_sort_hist     per tile: digit histogram (pass 1: + range check and the normalisers Z_f)
//   k_sort_scan     one CTA per bin: global offset of (bin, tile)
//   k_sort_scatter  per tile: stable rank of every key among its digit (warp match + per-warp
//                   counters), scatter
// (per-key global atomics for the next pass's histogram were tried and cost 20 us: a Zipf head
// row sends thousands of atomics to one address)
// digits of <= 10 bits: 2 passes up to 2^20 rows, 3 passes up to 2^30.  Then
//   k_plan_heads / k_plan_scatter  unique ranks (head-flag scan), uniq / inverse / segments
//   k_plan_finish   flat records the step kernels read
#include "common.cuh"
#include "internal.h"

namespace vfmb {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kMaxBins = 1024;
constexpr int kMaxTiles = 512;      // tiles per pass (the tile grows with B*F beyond 512K keys)

// ---- S1: digit histogram per tile; the first pass also range-checks the ids and accumulates
// the per-column normalisers Z_f = sum_n 1 / cnt_train(x[n,f])
// (Z_f equals the sum over uniq(x[:,f]) of cnt_f(u)/cnt_train(u), vfm-torch.py:305-306)
template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads)
k_sort_hist(const int64_t* __restrict__ x, const int32_t* __restrict__ kin,
            const float* __restrict__ train_counts, int N, int F, int R,
            int tile, int n_tiles, int shift, int bins, int32_t* __restrict__ hist, int32_t* __restrict__ gtot,
            int32_t* __restrict__ meta, double* __restrict__ partials, float* __restrict__ z) {
    chain_wait();                                           // launched with launch_chained()
    __shared__ int s_hist[kMaxBins];
    __shared__ double s_part[kMaxFields][kSortWarps];
    __shared__ bool s_last;
    const int tid = threadIdx.x, t = blockIdx.x;
    for (int b = tid; b < bins; b += kSortThreads) s_hist[b] = 0;
    __syncthreads();
    double acc[kMaxFields];
#pragma unroll
    for (int f = 0; f < kMaxFields; ++f) acc[f] = 0.0;
    const int lo = t * tile, hi = min(N, lo + tile);
    if (FIRST) {
#pragma unroll 4
        for (int o = lo + tid; o < hi; o += kSortThreads) {
            int64_t id = x[o];
            if (id < 0 || id >= R) { atomicOr(&meta[2], 1); id = 0; }
            atomicAdd(&s_hist[(int)id & (bins - 1)], 1);
            const float tc = 1.0f / __ldg(train_counts + id);
            const int f = o % F;
#pragma unroll
            for (int g = 0; g < kMaxFields; ++g) if (g == f) acc[g] += (double)tc;
        }
    } else {
#pragma unroll 4
        for (int o = lo + tid; o < hi; o += kSortThreads) atomicAdd(&s_hist[(kin[o] >> shift) & (bins - 1)], 1);
    }
    __syncthreads();
    for (int b = tid; b < bins; b += kSortThreads) {
        const int c = s_hist[b];
        hist[(size_t)b * n_tiles + t] = c;
        if (c) atomicAdd(&gtot[b], c);
    }
    if (!FIRST) return;
    // deterministic Z_f: block partials, the block that arrives last adds them in block order
    const int lane = tid & 31, warp = tid >> 5;
    for (int f = 0; f < F; ++f) {
        double sfw = warp_sum(acc[f]);
        if (lane == 0) s_part[f][warp] = sfw;
    }
    __syncthreads();
    if (tid < F) {
        double sf = 0.0;
        for (int w = 0; w < kSortWarps; ++w) sf += s_part[tid][w];
        partials[(size_t)t * kMaxFields + tid] = sf;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&meta[4], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (warp < F) {      // warp f reduces column f: lane l adds blocks l, l+32, ... then a fixed tree
            double sf = 0.0;
            for (unsigned b = lane; b < gridDim.x; b += 32) sf += __ldcg(partials + (size_t)b * kMaxFields + warp);
            sf = warp_sum(sf);
            if (lane == 0) z[warp] = (float)sf;
        }
    }
}

// ---- S2: hist[bin][tile] -> global offset of the first key of (bin, tile) ----------------------
__global__ void __launch_bounds__(128)
k_sort_scan(int32_t* __restrict__ hist, const int32_t* __restrict__ gtot, int n_tiles) {
    chain_wait();                                           // launched with launch_chained()
    __shared__ int s_w[4];
    __shared__ int s_carry;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int below = 0;                                         // keys in bins < b
    for (int i = tid; i < b; i += 128) below += gtot[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if (lane == 0) s_w[warp] = below;
    __syncthreads();
    if (tid == 0) s_carry = s_w[0] + s_w[1] + s_w[2] + s_w[3];
    __syncthreads();
    int32_t* row = hist + (size_t)b * n_tiles;
    for (int t0 = 0; t0 < n_tiles; t0 += 128) {
        const int t = t0 + tid;
        const int c = t < n_tiles ? row[t] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        __syncthreads();                                   // s_w / s_carry of the previous chunk are consumed
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += s_w[w];
        const int carry = s_carry;
        if (t < n_tiles) row[t] = carry + woff + incl - c;
        __syncthreads();
        if (tid == 127) s_carry = carry + woff + incl;
    }
}

// ---- S3: stable scatter of one pass ------------------------------------------------------------
// A tile is walked in chunks of 256 keys (the keys of 4 chunks are fetched up front, so the
// memory latency is paid once per 1024 keys).  Inside a chunk a key's rank among equal digits is
// (keys of earlier warps) + (earlier lanes of its warp, by match_any); s_base[digit] carries the
// offset of the digit's next free slot for this tile.  FIRST reads the int64 ids.
constexpr int kPre = 4;
template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads)
k_sort_scatter(const int64_t* __restrict__ x, int R, const int32_t* __restrict__ kin,
               const int32_t* __restrict__ vin, int32_t* __restrict__ kout, int32_t* __restrict__ vout,
               int N, int tile_shift, int n_tiles, int shift, int bins, const int32_t* __restrict__ base) {
    chain_wait();                                           // launched with launch_chained()
    __shared__ int s_base[kMaxBins];
    __shared__ int s_cnt[kSortWarps][kMaxBins];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, t = blockIdx.x;
    for (int b = tid; b < bins; b += kSortThreads) {
        s_base[b] = base[(size_t)b * n_tiles + t];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) s_cnt[w][b] = 0;
    }
    __syncthreads();
    const int lo = t << tile_shift, hi = min(N, lo + (1 << tile_shift));
    for (int g0 = lo; g0 < hi; g0 += kPre * kSortThreads) {
        int keys[kPre], vals[kPre];
#pragma unroll
        for (int c = 0; c < kPre; ++c) {
            const int i = g0 + c * kSortThreads + tid;
            keys[c] = 0; vals[c] = 0;
            if (i < hi) {
                if (FIRST) { const int64_t id = x[i]; keys[c] = (id < 0 || id >= R) ? 0 : (int)id; vals[c] = i; }
                else { keys[c] = kin[i]; vals[c] = vin[i]; }
            }
        }
#pragma unroll
        for (int c = 0; c < kPre; ++c) {
            if (g0 + c * kSortThreads >= hi) break;            // block-uniform
            const bool valid = g0 + c * kSortThreads + tid < hi;
            const int key = keys[c];
            const int digit = (key >> shift) & (bins - 1);
            const unsigned dm = valid ? (unsigned)digit : (0x40000000u | (unsigned)lane);
            const unsigned peers = __match_any_sync(0xffffffffu, dm);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            const int cnt = __popc(peers);
            const bool leader = valid && rank == 0;
            if (leader) s_cnt[warp][digit] = cnt;
            __syncthreads();
            int pos = 0;
            if (valid) {
                int off = 0;
                for (int w = 0; w < warp; ++w) off += s_cnt[w][digit];
                pos = s_base[digit] + off + rank;
            }
            __syncthreads();
            if (leader) { atomicAdd(&s_base[digit], cnt); s_cnt[warp][digit] = 0; }
            if (valid) { kout[pos] = key; vout[pos] = vals[c]; }
        }
    }
}

// ---- P2: number of segment heads per tile of the sorted keys -----------------------------------
__global__ void __launch_bounds__(kSortThreads)
k_plan_heads(const int32_t* __restrict__ keys_s, int N, int tile_shift, int32_t* __restrict__ tile_heads) {
    chain_wait();                                           // launched with launch_chained()
    __shared__ int s_w[kSortWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, t = blockIdx.x;
    const int lo = t << tile_shift, hi = min(N, lo + (1 << tile_shift));
    int c = 0;
    for (int i = lo + tid; i < hi; i += kSortThreads) c += (i == 0 || keys_s[i] != keys_s[i - 1]) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) s_w[warp] = c;
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < kSortWarps; ++w) tot += s_w[w];
        tile_heads[t] = tot;
    }
}

// ---- P3: unique ranks (scan of the head flags) and scatter of uniq / seg_off / inverse --------
__global__ void __launch_bounds__(kSortThreads)
k_plan_scatter(const int32_t* __restrict__ keys_s, const int32_t* __restrict__ vals_s,
               const int32_t* __restrict__ tile_heads, int N, int tile_shift, int32_t* __restrict__ uniq,
               int32_t* __restrict__ seg_off, int32_t* __restrict__ inverse, int32_t* __restrict__ occ,
               int32_t* __restrict__ pos_of, int32_t* __restrict__ pos_rank, int32_t* __restrict__ meta) {
    chain_wait();                                           // launched with launch_chained()
    __shared__ int s_w[kSortWarps];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, t = blockIdx.x;
    int below = 0;                                         // heads in the tiles before this one
    for (int i = tid; i < t; i += kSortThreads) below += tile_heads[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if (lane == 0) s_w[warp] = below;
    __syncthreads();
    if (tid == 0) { int tot = 0; for (int w = 0; w < kSortWarps; ++w) tot += s_w[w]; s_carry = tot; }
    __syncthreads();
    const int lo = t << tile_shift, hi = min(N, lo + (1 << tile_shift));
    for (int g0 = lo; g0 < hi; g0 += kPre * kSortThreads) {
        int ks[kPre], os[kPre], hs[kPre];
#pragma unroll
        for (int c = 0; c < kPre; ++c) {
            const int i = g0 + c * kSortThreads + tid;
            ks[c] = 0; os[c] = 0; hs[c] = 0;
            if (i < hi) {
                ks[c] = keys_s[i];
                os[c] = vals_s[i];
                hs[c] = (i == 0 || keys_s[i - 1] != ks[c]) ? 1 : 0;
            }
        }
#pragma unroll
        for (int c = 0; c < kPre; ++c) {
            if (g0 + c * kSortThreads >= hi) break;            // block-uniform
            const int i = g0 + c * kSortThreads + tid;
            const bool valid = i < hi;
            const int k = ks[c], o = os[c], head = hs[c];
            int incl = head;
#pragma unroll
            for (int sft = 1; sft < 32; sft <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, sft); if (lane >= sft) incl += v; }
            __syncthreads();                               // s_w / s_carry of the previous chunk are consumed
            if (lane == 31) s_w[warp] = incl;
            __syncthreads();
            int woff = 0;
            for (int w = 0; w < warp; ++w) woff += s_w[w];
            const int carry = s_carry;
            const int r = carry + woff + incl - 1;         // unique rank of position i
            if (valid) {
                if (head) { uniq[r] = k; seg_off[r] = i; }
                inverse[o] = r;
                occ[i] = o;
                pos_of[o] = i;
                pos_rank[i] = r;
                if (i == N - 1) { meta[0] = r + 1; seg_off[r + 1] = N; }
            }
            __syncthreads();
            if (tid == kSortThreads - 1) s_carry = carry + woff + incl;
        }
    }
}

// ---- P4: flat records the step kernels read with one load each -----------------------------
//   partner[i]  per sorted position: F==2 the rank of the sample's other field, else sample n
//   urec[u]     {row id, segment length, segment offset, batch count (= length; the owner of a
//               sharded row overwrites it with the count summed over ranks)}
struct ClassBounds { int n; int bound[kMaxFields]; };
__device__ __forceinline__ int plan_class_of(const ClassBounds& cb, int row) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < kMaxFields - 1; ++i) k += (i < cb.n - 1 && row >= cb.bound[i]) ? 1 : 0;
    return k;
}

__global__ void __launch_bounds__(256)
k_plan_finish(const int32_t* __restrict__ uniq, const int32_t* __restrict__ seg_off,
              const int32_t* __restrict__ occ, const int32_t* __restrict__ inverse, int N, int F,
              ClassBounds cb, int32_t* __restrict__ partner, int32_t* __restrict__ urec,
              int32_t* __restrict__ class_off, int32_t* __restrict__ meta, int32_t* __restrict__ hot, int hot_cap) {
    chain_wait();                                           // launched with launch_chained()
    const int U = meta[0];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = tid; i < N; i += nth) {
        int o = occ[i];
        partner[i] = (F == 2) ? inverse[o ^ 1] : o / F;
    }
    for (int u = tid; u < U; u += nth) {
        const int seg0 = seg_off[u];
        const int rowid = uniq[u];
        const int len = seg_off[u + 1] - seg0;
        reinterpret_cast<int4*>(urec)[u] = make_int4(rowid, len, seg0, len);
        // rows cut by backward-tile boundaries, as lists (diagnostics: the step kernels finish cut rows
        // in place, see finish_cut_row): <= kHotPartials tiles from the front, more from the back
        if (hot) {
            const int span = (seg0 + len - 1) / kTile - seg0 / kTile;
            if (span > kHotPartials) hot[hot_cap - 1 - atomicAdd(&meta[3], 1)] = u;
            else if (span > 0) hot[atomicAdd(&meta[5], 1)] = u;
        }
        // class_off[g] = first rank whose class is >= g (uniq is sorted, classes are id ranges)
        const int cls = plan_class_of(cb, rowid);
        const int prev = (u == 0) ? -1 : plan_class_of(cb, uniq[u - 1]);
        for (int g = prev + 1; g <= cls; ++g) class_off[g] = u;
        if (u == U - 1) for (int g = cls + 1; g <= kMaxFields; ++g) class_off[g] = U;
    }
}

static inline int bits_for(int R) {
    int b = 1;
    while (b < 31 && (1LL << b) < (long long)R) ++b;
    return b;
}

// geometry of the sort for N keys of `bits` bits
struct SortGeom {
    int npass, w[3], shift[3], bins[3];
    int tile_shift, n_tiles;
};
static SortGeom sort_geom(int64_t N, int R) {
    SortGeom g{};
    const int bits = bits_for(R);
    g.npass = (bits + 9) / 10;
    int done = 0;
    for (int p = 0; p < g.npass; ++p) {
        const int w = (bits - done + (g.npass - p) - 1) / (g.npass - p);
        g.w[p] = w; g.shift[p] = done; g.bins[p] = 1 << w;
        done += w;
    }
    g.tile_shift = 10;
    while (((N + (1LL << g.tile_shift) - 1) >> g.tile_shift) > kMaxTiles) ++g.tile_shift;
    g.n_tiles = (int)((N + (1LL << g.tile_shift) - 1) >> g.tile_shift);
    return g;
}

struct PlanWs {
    int32_t *kA, *vA, *kB, *vB;
    int32_t* counters;         // [3][kMaxBins] bin totals of the passes (accumulated: zeroed per build)
    int32_t* hist;             // [bins][n_tiles] histogram of the current pass
    int32_t* tile_heads;
    double* partials;
    size_t total;
};

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static PlanWs carve(void* base, int64_t N, const SortGeom& g) {
    PlanWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return (char*)base + o; };
    w.kA = (int32_t*)take((size_t)N * 4);
    w.vA = (int32_t*)take((size_t)N * 4);
    w.kB = (int32_t*)take((size_t)N * 4);
    w.vB = (int32_t*)take((size_t)N * 4);
    w.counters = (int32_t*)take(3 * (size_t)kMaxBins * 4);
    w.hist = (int32_t*)take((size_t)kMaxBins * g.n_tiles * 4);
    w.tile_heads = (int32_t*)take((size_t)g.n_tiles * 4);
    w.partials = (double*)take((size_t)kMaxTiles * kMaxFields * 8);
    w.total = off;
    return w;
}

}  // namespace vfmb

using namespace vfmb;

extern "C" int vfmb_plan_capacity(int32_t B, int32_t F, int32_t R, vfmb_plan_capacity_t* out) {
    if (!out || B <= 0 || F < 1 || F > VFMB_MAX_FIELDS || R <= 0)
        return set_error(VFMB_EINVAL, "vfmb_plan_capacity: bad B/F/R");
    int64_t N = (int64_t)B * F;
    if (N >= (1LL << 30)) return set_error(VFMB_ESHAPE, "vfmb_plan_capacity: B*F must be < 2^30");
    out->u_cap = N < R ? N : R;
    out->n_tiles = (N + kTile - 1) / kTile;
    out->tile = kTile;
    out->cut_rows_cap = (int32_t)cut_list_capacity(out->n_tiles);
    PlanWs w = carve(nullptr, N, sort_geom(N, R));
    out->workspace_bytes = (int64_t)w.total;
    return 0;
}

extern "C" int vfmb_plan_build(const vfmb_config* cfg, const int64_t* x, const float* train_counts,
                               const vfmb_plan* plan, void* workspace, size_t workspace_bytes,
                               vfmb_stream stream_) {
    if (!cfg || !x || !plan || !workspace || !train_counts)
        return set_error(VFMB_EINVAL, "vfmb_plan_build: null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    vfmb_plan_capacity_t cap;
    int rc = vfmb_plan_capacity(cfg->B, cfg->F, cfg->R, &cap);
    if (rc) return rc;
    if (workspace_bytes < (size_t)cap.workspace_bytes)
        return set_error(VFMB_ESPACE, "vfmb_plan_build: workspace too small");
    const int N = cfg->B * cfg->F;
    const SortGeom g = sort_geom(N, cfg->R);
    PlanWs w = carve(workspace, N, g);
    int32_t* gtot[3] = {w.counters, w.counters + kMaxBins, w.counters + 2 * kMaxBins};

    CUDA_TRY(cudaMemsetAsync(plan->meta, 0, 8 * sizeof(int32_t), stream));
    CUDA_TRY(cudaMemsetAsync(w.counters, 0, 3 * kMaxBins * sizeof(int32_t), stream));
    const int32_t* kin = nullptr; const int32_t* vin = nullptr;
    int32_t* kout = w.kA; int32_t* vout = w.vA;
    for (int p = 0; p < g.npass; ++p) {
        if (p == 0)
            CUDA_TRY(launch_chained(k_sort_hist<true>, g.n_tiles, kSortThreads, 0, stream, x, nullptr, train_counts, N, cfg->F, cfg->R,
                1 << g.tile_shift, g.n_tiles, 0, g.bins[0], w.hist, gtot[0], plan->meta, w.partials, plan->z));
        else
            CUDA_TRY(launch_chained(k_sort_hist<false>, g.n_tiles, kSortThreads, 0, stream, nullptr, kin, nullptr, N, cfg->F, cfg->R,
                1 << g.tile_shift, g.n_tiles, g.shift[p], g.bins[p], w.hist, gtot[p], plan->meta, nullptr, nullptr));
        CUDA_TRY(launch_chained(k_sort_scan, g.bins[p], 128, 0, stream, w.hist, gtot[p], g.n_tiles));
        if (p == 0)
            CUDA_TRY(launch_chained(k_sort_scatter<true>, g.n_tiles, kSortThreads, 0, stream, x, cfg->R, kin, vin, kout, vout, N, g.tile_shift,
                                                                         g.n_tiles, g.shift[p], g.bins[p], w.hist));
        else
            CUDA_TRY(launch_chained(k_sort_scatter<false>, g.n_tiles, kSortThreads, 0, stream, x, cfg->R, kin, vin, kout, vout, N, g.tile_shift,
                                                                          g.n_tiles, g.shift[p], g.bins[p], w.hist));
        CUDA_TRY(cudaGetLastError());
        kin = kout; vin = vout;
        kout = (kout == w.kA) ? w.kB : w.kA;
        vout = (vout == w.vA) ? w.vB : w.vA;
    }
    const int32_t* keys_s = kin; const int32_t* vals_s = vin;
    CUDA_TRY(launch_chained(k_plan_heads, g.n_tiles, kSortThreads, 0, stream, keys_s, N, g.tile_shift, w.tile_heads));
    CUDA_TRY(launch_chained(k_plan_scatter, g.n_tiles, kSortThreads, 0, stream, keys_s, vals_s, w.tile_heads, N, g.tile_shift, plan->uniq,
                                                           plan->seg_off, plan->inverse, plan->occ, plan->pos_of,
                                                           plan->pos_rank, plan->meta));
    int grid2 = (N + 255) / 256;
    if (grid2 > 4 * kPlanGrid) grid2 = 4 * kPlanGrid;
    ClassBounds cbd{};
    cbd.n = cfg->n_classes;
    for (int i = 0; i < kMaxFields; ++i) cbd.bound[i] = cfg->class_bound[i];
    CUDA_TRY(launch_chained(k_plan_finish, grid2, 256, 0, stream, plan->uniq, plan->seg_off, plan->occ, plan->inverse, N, cfg->F,
                                             cbd, plan->partner, plan->urec, plan->class_off, plan->meta, plan->hot,
                                             (int)cut_list_capacity(cap.n_tiles)));
    return 0;
}
