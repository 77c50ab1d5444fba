// Batch plan: sorted unique rows, inverse map, segments, work items.
//
// Replaces the three torch.unique calls of vfm-torch.py:190-192 and the
// per-group ones of vfm-tomasrch.py:537-545.  Integer outputs are bit-identical
// to torch.unique(sorted=True): the (row id, occurrence) pairs are radix-sorted
// by row id with a stable sort, so inside a segment the occurrences stay in
// ascending batch order -- that fixed order is what makes the segmented
// gradient reduction deterministic.
//
// Round-1 implementation: the sort and the two prefix sums use CUB device
// primitives (library plumbing); keying, head flags, scatter of
// uniq/inverse/segments, the per-column normalisers and the work-item list are
// own kernels.
#include "common.cuh"
#include "internal.h"

#include <mutex>
#include <unordered_map>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

namespace vfmb {

// ---- P1: keys + per-column normaliser Z_f = sum_n 1 / cnt_train(x[n,f]) --------------------
// (equal to sum over uniq(x[:,f]) of cnt_f(u)/cnt_train(u), vfm-torch.py:305-306)
__global__ void __launch_bounds__(256)
k_plan_keys(const int64_t* __restrict__ x, const float* __restrict__ train_counts, int N, int F, int R,
            int32_t* __restrict__ keys, int32_t* __restrict__ vals, int32_t* __restrict__ meta,
            double* __restrict__ partials, int32_t* __restrict__ counter, float* __restrict__ z) {
    __shared__ double s_part[kMaxFields][8];
    __shared__ bool s_last;
    double acc[kMaxFields];
#pragma unroll
    for (int f = 0; f < kMaxFields; ++f) acc[f] = 0.0;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < N; o += gridDim.x * blockDim.x) {
        int64_t id = x[o];
        bool ok = id >= 0 && id < R;
        if (!ok) { atomicOr(&meta[2], 1); id = 0; }
        keys[o] = (int32_t)id;
        vals[o] = o;
        float t = 1.0f / __ldg(train_counts + id);
        int f = o % F;
#pragma unroll
        for (int g = 0; g < kMaxFields; ++g) if (g == f) acc[g] += (double)t;
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int f = 0; f < F; ++f) {
        double s = warp_sum(acc[f]);
        if (lane == 0) s_part[f][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < F) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += s_part[threadIdx.x][w];
        partials[(size_t)blockIdx.x * kMaxFields + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence();
        // warp f reduces column f: lane l adds blocks l, l+32, ... then a fixed shuffle tree
        if (warp < F) {
            double s = 0.0;
            for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(partials + (size_t)b * kMaxFields + warp);
            s = warp_sum(s);
            if (lane == 0) z[warp] = (float)s;
        }
        if (threadIdx.x == 0) *counter = 0;
    }
}

struct HeadFlag {
    const int32_t* keys;
    __host__ __device__ int32_t operator()(int i) const {
        return (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
    }
};

// ---- P3: scatter uniq / seg_off / inverse from the sorted pairs and their head-flag scan ---
__global__ void __launch_bounds__(256)
k_plan_scatter(const int32_t* __restrict__ keys_s, const int32_t* __restrict__ vals_s,
               const int32_t* __restrict__ rank_incl, int N, int32_t* __restrict__ uniq,
               int32_t* __restrict__ seg_off, int32_t* __restrict__ inverse, int32_t* __restrict__ occ,
               int32_t* __restrict__ pos_of, int32_t* __restrict__ pos_rank, int32_t* __restrict__ meta) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        int r = rank_incl[i] - 1;
        int k = keys_s[i];
        int o = vals_s[i];
        bool head = (i == 0) || (keys_s[i - 1] != k);
        if (head) { uniq[r] = k; seg_off[r] = i; }
        inverse[o] = r;
        occ[i] = o;
        pos_of[o] = i;
        pos_rank[i] = r;
        if (i == N - 1) { meta[0] = r + 1; seg_off[r + 1] = N; }
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
              int32_t* __restrict__ class_off, const int32_t* __restrict__ meta) {
    const int U = meta[0];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = tid; i < N; i += nth) {
        int o = occ[i];
        partner[i] = (F == 2) ? inverse[o ^ 1] : o / F;
    }
    for (int u = tid; u < U; u += nth) {
        const int seg0 = seg_off[u];
        const int rowid = uniq[u];
        reinterpret_cast<int4*>(urec)[u] = make_int4(rowid, seg_off[u + 1] - seg0, seg0, seg_off[u + 1] - seg0);
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

struct PlanWs {
    int32_t *keys, *vals, *keys_s, *vals_s, *rank;
    double* partials;
    int32_t* counter;
    void* cub;
    size_t cub_bytes;
    size_t total;
};

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static PlanWs carve(void* base, int N, int u_cap) {
    PlanWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return (char*)base + o; };
    w.keys = (int32_t*)take((size_t)N * 4);
    w.vals = (int32_t*)take((size_t)N * 4);
    w.keys_s = (int32_t*)take((size_t)N * 4);
    w.vals_s = (int32_t*)take((size_t)N * 4);
    w.rank = (int32_t*)take((size_t)N * 4);
    w.partials = (double*)take((size_t)kPlanGrid * kMaxFields * 8);
    w.counter = (int32_t*)take(256);
    // CUB temp-storage sizes: host-side queries, cached per N (they cost microseconds per call)
    static std::mutex mu;
    static std::unordered_map<int, size_t> cache;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it0 = cache.find(N);
        if (it0 == cache.end()) {
            size_t sort_b = 0, scan_b = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (int32_t*)nullptr, (int32_t*)nullptr,
                                            (int32_t*)nullptr, (int32_t*)nullptr, N, 0, 31);
            cub::TransformInputIterator<int32_t, HeadFlag, cub::CountingInputIterator<int>> it(
                cub::CountingInputIterator<int>(0), HeadFlag{nullptr});
            cub::DeviceScan::InclusiveSum(nullptr, scan_b, it, (int32_t*)nullptr, N);
            it0 = cache.emplace(N, sort_b > scan_b ? sort_b : scan_b).first;
        }
        w.cub_bytes = it0->second;
    }
    w.cub = take(w.cub_bytes);
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
    PlanWs w = carve(nullptr, (int)N, (int)out->u_cap);
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
    const int u_cap = (int)cap.u_cap;
    PlanWs w = carve(workspace, N, u_cap);

    CUDA_TRY(cudaMemsetAsync(plan->meta, 0, 8 * sizeof(int32_t), stream));
    CUDA_TRY(cudaMemsetAsync(w.counter, 0, 4, stream));
    int grid = (N + 255) / 256;
    if (grid > kPlanGrid) grid = kPlanGrid;
    k_plan_keys<<<grid, 256, 0, stream>>>(x, train_counts, N, cfg->F, cfg->R, w.keys, w.vals, plan->meta,
                                          w.partials, w.counter, plan->z);
    CUDA_TRY(cudaGetLastError());
    size_t cb = w.cub_bytes;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(w.cub, cb, w.keys, w.keys_s, w.vals, w.vals_s, N, 0,
                                             bits_for(cfg->R), stream));
    cub::TransformInputIterator<int32_t, HeadFlag, cub::CountingInputIterator<int>> heads(
        cub::CountingInputIterator<int>(0), HeadFlag{w.keys_s});
    cb = w.cub_bytes;
    CUDA_TRY(cub::DeviceScan::InclusiveSum(w.cub, cb, heads, w.rank, N, stream));
    int grid2 = (N + 255) / 256;
    if (grid2 > 4 * kPlanGrid) grid2 = 4 * kPlanGrid;
    k_plan_scatter<<<grid2, 256, 0, stream>>>(w.keys_s, w.vals_s, w.rank, N, plan->uniq, plan->seg_off,
                                              plan->inverse, plan->occ, plan->pos_of, plan->pos_rank, plan->meta);
    CUDA_TRY(cudaGetLastError());
    ClassBounds cbd{};
    cbd.n = cfg->n_classes;
    for (int i = 0; i < kMaxFields; ++i) cbd.bound[i] = cfg->class_bound[i];
    k_plan_finish<<<grid2, 256, 0, stream>>>(plan->uniq, plan->seg_off, plan->occ, plan->inverse, N, cfg->F,
                                             cbd, plan->partner, plan->urec, plan->class_off, plan->meta);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
