// Mode B (row-sharded tables) -- the glue between the plan / step kernels and
// the three all-to-alls of a step, as own kernels instead of framework tensor ops:
//
//   requester   vfmb_shard_bucket        unique ids -> fixed-capacity slots per owner (id mod P)
//   owner       vfmb_shard_owner_ids     received ids -> local row indices (plan input)
//               vfmb_shard_owner_pack    summed batch counts; sampled rows -> reply slots
//   requester   vfmb_shard_unpack_rows   reply slots -> sampled-row scratch in unique-rank order
//               vfmb_shard_pack_grads    row gradients -> slots; additive scalars of the rank
//   owner       vfmb_shard_unpack_grads  received gradient rows -> gather table + coefficients
//
// Slot layout of every exchange: [P, CAP, w] (w = 2 ints for ids, d+1 floats for rows), slot
// q*CAP + j holds the j-th id (ascending) this rank asks of owner q; -1 / zeros = empty.
#include "step_common.cuh"

namespace vfmb {

constexpr int kBkt = 256;          // threads per block
constexpr int kBktTile = 1024;     // unique ids per block

__device__ __forceinline__ int owner_of(int id, int P) { return id % P; }

// Destination of an exchange.  n == 0: a local send buffer in slot order (the caller runs an NCCL
// all-to-all on it).  n == P: the exchange buffers of all ranks, peer-mapped over NVLink -- slot
// (q, j) of this rank is written straight into chunk `rank` of rank q's buffer, so the pack kernel
// IS the all-to-all (stores over NVSwitch; a barrier separates it from the consumer).
template <typename T>
__device__ __forceinline__ T* slot_ptr(const Peers& pe, T* local, int slot, int CAP, int width) {
    if (pe.n == 0) return local + (size_t)slot * width;
    const int q = slot / CAP, j = slot - q * CAP;
    return reinterpret_cast<T*>(pe.base[q]) + ((size_t)pe.rank * CAP + j) * width;
}

// ---- bucket, pass 1: ids per owner in every tile of the sorted unique list
__global__ void __launch_bounds__(kBkt)
k_bucket_count(const int32_t* __restrict__ uniq, const int32_t* __restrict__ meta, int P,
               int32_t* __restrict__ blk_cnt /*[nblk][8]*/) {
    __shared__ int s_cnt[kMaxFields];
    const int U = meta[0];
    if (threadIdx.x < kMaxFields) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int lo = blockIdx.x * kBktTile;
    for (int c = 0; c < kBktTile; c += kBkt) {
        const int u = lo + c + threadIdx.x;
        const int q = u < U ? owner_of(uniq[u], P) : -1;
        for (int k = 0; k < P; ++k) {
            const unsigned b = __ballot_sync(0xffffffffu, q == k);
            if ((threadIdx.x & 31) == 0 && b) atomicAdd(&s_cnt[k], __popc(b));
        }
    }
    __syncthreads();
    if (threadIdx.x < kMaxFields) blk_cnt[blockIdx.x * kMaxFields + threadIdx.x] = s_cnt[threadIdx.x];
}

// ---- bucket, pass 2: slot of every unique rank (stable inside an owner), request records
__global__ void __launch_bounds__(kBkt)
k_bucket_place(const int32_t* __restrict__ uniq, const int32_t* __restrict__ urec, const int32_t* __restrict__ meta,
               int u_cap, int P, int CAP, const int32_t* __restrict__ blk_cnt, int32_t* __restrict__ send,
               int32_t* __restrict__ dest, int32_t* __restrict__ overflow, Peers pe) {
    __shared__ int s_base[kMaxFields];
    __shared__ int s_w[kBkt / 32][kMaxFields];
    const int U = meta[0], M = P * CAP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < kMaxFields) {
        int b = 0;
        for (int i = 0; i < (int)blockIdx.x; ++i) b += blk_cnt[i * kMaxFields + threadIdx.x];
        s_base[threadIdx.x] = b;
    }
    __syncthreads();
    const int lo = blockIdx.x * kBktTile;
    for (int c = 0; c < kBktTile; c += kBkt) {
        const int u = lo + c + threadIdx.x;
        const bool valid = u < U;
        const int id = valid ? uniq[u] : 0;
        const int q = valid ? owner_of(id, P) : -1;
        int rank = 0;
        for (int k = 0; k < P; ++k) {
            const unsigned b = __ballot_sync(0xffffffffu, q == k);
            if (q == k) rank = __popc(b & ((1u << lane) - 1u));
            if (lane == 0) s_w[warp][k] = __popc(b);
        }
        __syncthreads();
        int slot = M;
        if (valid) {
            int off = s_base[q];
            for (int w = 0; w < warp; ++w) off += s_w[w][q];
            const int ord = off + rank;
            if (ord < CAP) slot = q * CAP + ord; else atomicOr(overflow, 1);
            if (slot < M) {
                int32_t* rq = slot_ptr<int32_t>(pe, send, slot, CAP, 2);
                rq[0] = id;
                rq[1] = urec[4 * (size_t)u + 1];                  // occurrences in this rank's batch
            }
        }
        if (u < u_cap) dest[u] = slot;
        __syncthreads();
        if (threadIdx.x < P) {
            int tot = 0;
            for (int w = 0; w < kBkt / 32; ++w) tot += s_w[w][threadIdx.x];
            s_base[threadIdx.x] += tot;
        }
        __syncthreads();
    }
}

// ---- peer mode: this rank's chunk of every rank's request buffer := empty (-1)
__global__ void __launch_bounds__(256)
k_fill_peers(Peers pe, int words_per_chunk, int32_t value) {
    for (int q = 0; q < pe.n; ++q) {
        int32_t* dst = reinterpret_cast<int32_t*>(pe.base[q]) + (size_t)pe.rank * words_per_chunk;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < words_per_chunk; i += gridDim.x * blockDim.x) dst[i] = value;
    }
}

// ---- peer mode: small vectors (Z_f, scalar sums): put into slot `rank` of every rank, then every
// rank adds the P vectors in rank order -- an all-reduce whose result is bitwise equal everywhere
__global__ void k_put_small(Peers pe, const float* __restrict__ vec, int n, int pitch) {
    const int q = blockIdx.x;
    float* dst = reinterpret_cast<float*>(pe.base[q]) + (size_t)pe.rank * pitch;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = vec[i];
}
__global__ void k_sum_small(const float* __restrict__ slots, int P, int n, int pitch, float* __restrict__ out) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float acc = 0.f;
        for (int q = 0; q < P; ++q) acc += slots[(size_t)q * pitch + i];
        out[i] = acc;
    }
}

// ---- owner: received global ids -> local row index (padding -> the sentinel row R_loc); optional
// private copy of the requests (peer mode: the shared buffer is rewritten by the next step)
__global__ void __launch_bounds__(256)
k_owner_ids(const int32_t* __restrict__ recv, int M, int P, int R_loc, int64_t* __restrict__ loc,
            int32_t* __restrict__ copy) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < M; s += gridDim.x * blockDim.x) {
        const int2 rq = reinterpret_cast<const int2*>(recv)[s];
        loc[s] = rq.x >= 0 ? (int64_t)(rq.x / P) : (int64_t)R_loc;
        if (copy) reinterpret_cast<int2*>(copy)[s] = rq;
    }
}

// ---- owner: batch count of a row summed over the requesting ranks (integers, fixed order)
__global__ void __launch_bounds__(256)
k_owner_counts(const int32_t* __restrict__ recv, const int32_t* __restrict__ occ, const int32_t* __restrict__ meta,
               int32_t* __restrict__ urec) {
    const int U = meta[0];
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < U; u += gridDim.x * blockDim.x) {
        int len = urec[4 * (size_t)u + 1];
        const int seg = urec[4 * (size_t)u + 2];
        // a real row is requested at most once per rank; the padding slots all share the sentinel
        // row (tens of thousands of occurrences, count 0): only look at the first few
        if (len > kMaxFields) len = recv[2 * occ[seg]] >= 0 ? kMaxFields : 0;
        int tot = 0;
        for (int i = 0; i < len; ++i) {
            const int s = occ[seg + i];
            tot += recv[2 * s] >= 0 ? recv[2 * s + 1] : 0;
        }
        urec[4 * (size_t)u + 3] = tot;
    }
}

// ---- rows <-> slots: one warp per slot / unique rank, d+1 floats (sampled row | bias)
__global__ void __launch_bounds__(256)
k_pack_rows(const float* __restrict__ vs, const float* __restrict__ ws, const int32_t* __restrict__ inverse,
            const int32_t* __restrict__ ids, int M, int CAP, int d, float* __restrict__ out, Peers pe) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
    for (int s = gw; s < M; s += nw) {
        if (pe.n && __ldg(ids + 2 * s) < 0) continue;     // peer mode: nobody reads an empty slot's reply
        const int u = __ldg(inverse + s);
        const float* src = vs + (size_t)u * d;
        float* dst = slot_ptr<float>(pe, out, s, CAP, d + 1);
        for (int k = lane; k < d; k += 32) dst[k] = __ldg(src + k);
        if (lane == 0) dst[d] = __ldg(ws + u);
    }
}

__global__ void __launch_bounds__(256)
k_unpack_rows(const float* __restrict__ recv, const int32_t* __restrict__ dest, const int32_t* __restrict__ meta,
              int M, int d, float* __restrict__ vs, float* __restrict__ ws) {
    const int U = meta[0];
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
    for (int u = gw; u < U; u += nw) {
        const int s = min(__ldg(dest + u), M - 1);
        const float* src = recv + (size_t)s * (d + 1);
        float* dst = vs + (size_t)u * d;
        for (int k = lane; k < d; k += 32) dst[k] = __ldg(src + k);
        if (lane == 0) ws[u] = __ldg(src + d);
    }
}

// gradients of the requester's unique rows -> slots (empty slots were zeroed by the caller);
// thread 0 also assembles the additive scalars of this rank
__global__ void __launch_bounds__(256)
k_pack_grads(const float* __restrict__ grow, const float* __restrict__ gws, const int32_t* __restrict__ dest,
             const int32_t* __restrict__ meta, int M, int d, float* __restrict__ out,
             const float* __restrict__ stats_l, const float* __restrict__ stats_o, float n_local,
             float* __restrict__ tail, int t_nll, int t_resid, int t_sqerr, int t_klrows, int n_tail,
             int CAP, Peers pe) {
    const int U = meta[0];
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x < n_tail) {
        float v = 0.f;
        const int t = threadIdx.x;
        if (t == t_nll) v = stats_l[VFMB_ST_NLL_MEAN] * n_local;
        else if (t == t_resid) v = stats_l[VFMB_ST_SUM_RESID];
        else if (t == t_sqerr) v = stats_l[VFMB_ST_SUM_SQERR];
        else if (t == t_klrows) v = stats_o[VFMB_ST_KL_ROWS];
        tail[t] = v;
    }
    for (int u = gw; u < U; u += nw) {
        const int s = __ldg(dest + u);
        if (s >= M) continue;                              // overflowed (flagged at bucketing)
        const float* src = grow + (size_t)u * d;
        float* dst = slot_ptr<float>(pe, out, s, CAP, d + 1);
        for (int k = lane; k < d; k += 32) dst[k] = __ldg(src + k);
        if (lane == 0) dst[d] = __ldg(gws + u);
    }
}

// owner: received [M, d+1] -> gather table [M, d] (16-byte aligned rows) and the per-position
// coefficient (the bias gradient) in the plan's sorted-occurrence order
__global__ void __launch_bounds__(256)
k_unpack_grads(const float* __restrict__ recv, const int32_t* __restrict__ occ, const int32_t* __restrict__ ids,
               int M, int d, float* __restrict__ table, float* __restrict__ rsorted) {
    // ids != NULL (peer mode): empty slots were never written by their requester -> zeros
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
    for (int s = gw; s < M; s += nw) {
        const bool live = !ids || __ldg(ids + 2 * s) >= 0;
        const float* src = recv + (size_t)s * (d + 1);
        float* dst = table + (size_t)s * d;
        for (int k = lane; k < d; k += 32) dst[k] = live ? __ldg(src + k) : 0.f;
        if (lane == 0) {
            const int so = __ldg(occ + s);
            rsorted[s] = (!ids || __ldg(ids + 2 * so) >= 0) ? __ldg(recv + (size_t)so * (d + 1) + d) : 0.f;
        }
    }
}

// ---- fused mode B, requester side (id-only, off the critical path): where every row of the batch is
// read from / written to once the exchanges are in place --
//   inv_slot[o]      slot (in this rank's received-rows region) of the row of occurrence o = n*F + f
//   partner_slot[i]  F == 2: slot of the partner row of sorted position i  (F > 2: the sample index)
//   gptr[u]          where the gradient row of unique rank u goes: its owner's slot (over NVLink)
// Overflowed rows (dest = M) map to the spare slot M (zeros) / the dump row.
__global__ void __launch_bounds__(256)
k_route(const int32_t* __restrict__ dest, const int32_t* __restrict__ inverse, const int32_t* __restrict__ partner,
        const int32_t* __restrict__ meta, int N, int F, int u_cap, int M, int CAP, int SP, Peers pe, float* dump,
        int32_t* __restrict__ inv_slot, int32_t* __restrict__ partner_slot, float** __restrict__ gptr) {
    const int U = meta[0];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int o = tid; o < N; o += nth) {
        inv_slot[o] = min(__ldg(dest + __ldg(inverse + o)), M);
        const int pr = __ldg(partner + o);
        partner_slot[o] = (F == 2) ? min(__ldg(dest + pr), M) : pr;
    }
    for (int u = tid; u < u_cap; u += nth) {
        const int sl = u < U ? __ldg(dest + u) : M;
        float* p = dump;
        if (sl < M) {
            const int q = sl / CAP, j = sl - q * CAP;
            p = reinterpret_cast<float*>(pe.base[q]) + ((size_t)pe.rank * CAP + j) * SP;
        }
        gptr[u] = p;
    }
}

static inline int warp_grid(int64_t items) {
    int64_t g = (items + 7) / 8;
    if (g < 1) g = 1;
    if (g > 8 * kNumSMs) g = 8 * kNumSMs;
    return (int)g;
}

}  // namespace vfmb

using namespace vfmb;

extern "C" int64_t vfmb_shard_bucket_workspace(int32_t u_cap) {
    return (int64_t)((u_cap + kBktTile - 1) / kBktTile) * kMaxFields * 4;
}

static int make_peers(const void* const* peers, int32_t P, int32_t rank, Peers* out) {
    Peers pe{};
    if (peers) {
        if (P < 1 || P > kMaxFields || rank < 0 || rank >= P) return set_error(VFMB_EINVAL, "peer table: bad P / rank");
        for (int q = 0; q < P; ++q) {
            if (!peers[q]) return set_error(VFMB_EINVAL, "peer table: null buffer of rank %d", q);
            pe.base[q] = const_cast<void*>(peers[q]);
        }
        pe.n = P; pe.rank = rank;
    }
    *out = pe;
    return 0;
}

extern "C" int vfmb_shard_bucket(const vfmb_plan* plan, int32_t u_cap, int32_t P, int32_t CAP, int32_t* send,
                                 int32_t* dest, int32_t* overflow, void* workspace, const void* const* peers,
                                 int32_t rank, vfmb_stream stream_) {
    if (!plan || (!send && !peers) || !dest || !overflow || !workspace || P < 1 || P > kMaxFields || CAP < 1 || u_cap < 1)
        return set_error(VFMB_EINVAL, "vfmb_shard_bucket: bad argument (1 <= P <= %d)", kMaxFields);
    cudaStream_t stream = (cudaStream_t)stream_;
    Peers pe;
    int rc = make_peers(peers, P, rank, &pe);
    if (rc) return rc;
    const int nblk = (u_cap + kBktTile - 1) / kBktTile;
    if (pe.n) k_fill_peers<<<32, 256, 0, counted(stream)>>>(pe, CAP * 2, -1);                        // -1: empty slot
    else CUDA_TRY(cudaMemsetAsync(send, 0xFF, (size_t)P * CAP * 2 * sizeof(int32_t), stream));
    k_bucket_count<<<nblk, kBkt, 0, counted(stream)>>>(plan->uniq, plan->meta, P, (int32_t*)workspace);
    k_bucket_place<<<nblk, kBkt, 0, counted(stream)>>>(plan->uniq, plan->urec, plan->meta, u_cap, P, CAP,
                                              (const int32_t*)workspace, send, dest, overflow, pe);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_shard_route(const vfmb_plan* plan_l, const int32_t* dest, int32_t B, int32_t F, int32_t u_cap,
                                int32_t M, int32_t CAP, int32_t SP, const void* const* peers_grads, int32_t P,
                                int32_t rank, float* dump, int32_t* inv_slot, int32_t* partner_slot, float** gptr,
                                vfmb_stream stream_) {
    if (!plan_l || !dest || !dump || !inv_slot || !partner_slot || !gptr || B < 1 || F < 1 || CAP < 1 || M != P * CAP)
        return set_error(VFMB_EINVAL, "vfmb_shard_route: bad argument");
    Peers pe;
    int rc = make_peers(peers_grads, P, rank, &pe);
    if (rc) return rc;
    if (!pe.n) return set_error(VFMB_EINVAL, "vfmb_shard_route: peer table required");
    const int N = B * F;
    int g = (N + 255) / 256;
    if (g > 4 * kNumSMs) g = 4 * kNumSMs;
    k_route<<<g, 256, 0, counted((cudaStream_t)stream_)>>>(dest, plan_l->inverse, plan_l->partner, plan_l->meta, N, F, u_cap,
                                                          M, CAP, SP, pe, dump, inv_slot, partner_slot, gptr);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_shard_put_small(const float* vec, int32_t n, int32_t pitch, const void* const* peers, int32_t P,
                                    int32_t rank, vfmb_stream stream_) {
    Peers pe;
    int rc = make_peers(peers, P, rank, &pe);
    if (rc) return rc;
    if (!vec || !peers || n < 1 || n > pitch) return set_error(VFMB_EINVAL, "vfmb_shard_put_small: bad argument");
    k_put_small<<<P, 32, 0, counted((cudaStream_t)stream_)>>>(pe, vec, n, pitch);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_shard_sum_small(const float* slots, int32_t P, int32_t n, int32_t pitch, float* out,
                                    vfmb_stream stream_) {
    if (!slots || !out || P < 1 || n < 1 || n > pitch) return set_error(VFMB_EINVAL, "vfmb_shard_sum_small: bad argument");
    k_sum_small<<<1, 32, 0, counted((cudaStream_t)stream_)>>>(slots, P, n, pitch, out);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_shard_owner_ids(const int32_t* recv, int32_t M, int32_t P, int32_t R_loc, int64_t* loc,
                                    int32_t* recv_copy, vfmb_stream stream_) {
    if (!recv || !loc || M < 1 || P < 1) return set_error(VFMB_EINVAL, "vfmb_shard_owner_ids: bad argument");
    k_owner_ids<<<(M + 255) / 256 > 2 * kNumSMs ? 2 * kNumSMs : (M + 255) / 256, 256, 0, counted((cudaStream_t)stream_)>>>(
        recv, M, P, R_loc, loc, recv_copy);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_shard_owner_pack(const vfmb_plan* plan_o, const int32_t* recv, int32_t M, int32_t CAP, int32_t d,
                                     const float* vs, const float* ws, float* reply, int32_t counts_only,
                                     const void* const* peers, int32_t rank, vfmb_stream stream_) {
    if (!plan_o || !recv || M < 1 || CAP < 1) return set_error(VFMB_EINVAL, "vfmb_shard_owner_pack: bad argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (counts_only) {
        k_owner_counts<<<2 * kNumSMs, 256, 0, counted(stream)>>>(recv, plan_o->occ, plan_o->meta, plan_o->urec);
    } else {
        if (!vs || !ws || (!reply && !peers) || d < 1) return set_error(VFMB_EINVAL, "vfmb_shard_owner_pack: bad argument");
        Peers pe;
        int rc = make_peers(peers, M / CAP, rank, &pe);
        if (rc) return rc;
        k_pack_rows<<<warp_grid(M), 256, 0, counted(stream)>>>(vs, ws, plan_o->inverse, recv, M, CAP, d, reply, pe);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_shard_unpack_rows(const vfmb_plan* plan_l, const float* recv_rows, const int32_t* dest,
                                      int32_t u_cap, int32_t M, int32_t d, float* vs, float* ws, vfmb_stream stream_) {
    if (!plan_l || !recv_rows || !dest || !vs || !ws) return set_error(VFMB_EINVAL, "vfmb_shard_unpack_rows: bad argument");
    k_unpack_rows<<<warp_grid(u_cap), 256, 0, counted((cudaStream_t)stream_)>>>(recv_rows, dest, plan_l->meta, M, d, vs, ws);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_shard_pack_grads(const vfmb_plan* plan_l, const float* grow, const float* gws, const int32_t* dest,
                                     int32_t u_cap, int32_t M, int32_t CAP, int32_t d, float* out, const float* stats_local,
                                     const float* stats_owner, float n_local, float* tail, const int32_t* tail_idx,
                                     int32_t n_tail, const void* const* peers, int32_t rank, vfmb_stream stream_) {
    if (!plan_l || !grow || !gws || !dest || (!out && !peers) || !stats_local || !stats_owner || !tail || !tail_idx ||
        n_tail > 256 || CAP < 1)
        return set_error(VFMB_EINVAL, "vfmb_shard_pack_grads: bad argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    Peers pe;
    int rc = make_peers(peers, M / CAP, rank, &pe);
    if (rc) return rc;
    if (!pe.n) CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)M * (d + 1) * sizeof(float), stream));
    k_pack_grads<<<warp_grid(u_cap), 256, 0, counted(stream)>>>(grow, gws, dest, plan_l->meta, M, d, out, stats_local, stats_owner,
                                                       n_local, tail, tail_idx[0], tail_idx[1], tail_idx[2], tail_idx[3], n_tail,
                                                       CAP, pe);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_shard_unpack_grads(const vfmb_plan* plan_o, const float* recv_g, const int32_t* recv_ids, int32_t M,
                                       int32_t d, float* table, float* rsorted, vfmb_stream stream_) {
    if (!plan_o || !recv_g || !table || !rsorted) return set_error(VFMB_EINVAL, "vfmb_shard_unpack_grads: bad argument");
    k_unpack_grads<<<warp_grid(M), 256, 0, counted((cudaStream_t)stream_)>>>(recv_g, plan_o->occ, recv_ids, M, d, table, rsorted);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
