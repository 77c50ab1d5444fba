// Measurement tool (not part of the library): what HBM bandwidth a kernel of k_adam_rows' size and
// shape can reach at all on this GPU.  Three tables of `rows` rows x 512 B are read and written once
// (the traffic of the row update at ml20m: 59 K rows x 3 tables x 512 B x 2 = 182 MB) by
//   copy      plain float4 streaming copy of the same number of bytes (the STREAM-style figure
//             MEASURED_PEAKS.json quotes, but at THIS size instead of 4 GB),
//   rmw3      p, m, v read / modified / written in place, contiguous rows,
//   gather3   the same through a row-index list (random subset of a 2.8x larger table, ascending,
//             as the touched rows of a batch are).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_ceiling stream_ceiling.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * st < n; i += 4 * st) {
        float4 x0 = a[i], x1 = a[i + st], x2 = a[i + 2 * st], x3 = a[i + 3 * st];
        b[i] = x0; b[i + st] = x1; b[i + 2 * st] = x2; b[i + 3 * st] = x3;
    }
    for (; i < n; i += st) b[i] = a[i];
}

// one row = 32 lanes x float4 = 512 B; a warp handles a row of all three tables at a time
template <bool GATHER>
__global__ void __launch_bounds__(256, 4) k_rmw3(float4* __restrict__ p, float4* __restrict__ m, float4* __restrict__ v,
                                                  const int* __restrict__ rows, int n_rows) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
    for (int r0 = gw * 2; r0 < n_rows; r0 += nw * 2) {
        float4 P[2], M[2], V[2]; size_t off[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int r = min(r0 + q, n_rows - 1);
            off[q] = (size_t)(GATHER ? rows[r] : r) * 32 + lane;
            P[q] = p[off[q]]; M[q] = m[off[q]]; V[q] = v[off[q]];
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (r0 + q >= n_rows) break;
            P[q].x += 1e-3f * M[q].x; P[q].y += 1e-3f * M[q].y; P[q].z += 1e-3f * M[q].z; P[q].w += 1e-3f * M[q].w;
            M[q].x = 0.9f * M[q].x + V[q].x; M[q].y = 0.9f * M[q].y + V[q].y; M[q].z = 0.9f * M[q].z + V[q].z; M[q].w = 0.9f * M[q].w + V[q].w;
            V[q].x *= 0.999f; V[q].y *= 0.999f; V[q].z *= 0.999f; V[q].w *= 0.999f;
            p[off[q]] = P[q]; m[off[q]] = M[q]; v[off[q]] = V[q];
        }
    }
}

// the same gather, software-pipelined through shared memory with cp.async (LDGSTS): a warp keeps the
// NEXT chunk of R rows x 3 tables in flight (no registers tied up) while it updates the current one
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int R, int MINB>
__global__ void __launch_bounds__(256, MINB) k_rmw3_async(float4* __restrict__ p, float4* __restrict__ m, float4* __restrict__ v,
                                                           const int* __restrict__ rows, int n_rows) {
    extern __shared__ float4 smem[];                         // [8 warps][2 stages][3 tables][R rows][32 lanes]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int gw = blockIdx.x * 8 + w, nw = gridDim.x * 8;
    float4* my = smem + (size_t)w * (2 * 3 * R * 32);
    const int lo = (int)((long long)gw * n_rows / nw), hi = (int)((long long)(gw + 1) * n_rows / nw);
    if (lo >= hi) return;
    float4* tbl[3] = {p, m, v};
    auto issue = [&](int stage, int r0) {
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const size_t off = (size_t)rows[min(r0 + q, hi - 1)] * 32 + lane;
#pragma unroll
            for (int t = 0; t < 3; ++t) cp_async16(my + ((stage * 3 + t) * R + q) * 32 + lane, tbl[t] + off);
        }
        cp_commit();
    };
    issue(0, lo);
    int c = 0;
    for (int r0 = lo; r0 < hi; r0 += R, c ^= 1) {
        if (r0 + R < hi) { issue(c ^ 1, r0 + R); cp_wait<1>(); } else cp_wait<0>();
#pragma unroll
        for (int q = 0; q < R; ++q) {
            if (r0 + q >= hi) break;
            const size_t off = (size_t)rows[r0 + q] * 32 + lane;
            float4 P = my[((c * 3 + 0) * R + q) * 32 + lane], M = my[((c * 3 + 1) * R + q) * 32 + lane], V = my[((c * 3 + 2) * R + q) * 32 + lane];
            P.x += 1e-3f * M.x; P.y += 1e-3f * M.y; P.z += 1e-3f * M.z; P.w += 1e-3f * M.w;
            M.x = 0.9f * M.x + V.x; M.y = 0.9f * M.y + V.y; M.z = 0.9f * M.z + V.z; M.w = 0.9f * M.w + V.w;
            V.x *= 0.999f; V.y *= 0.999f; V.z *= 0.999f; V.w *= 0.999f;
            p[off] = P; m[off] = M; v[off] = V;
        }
    }
}

template <int R, int MINB>
static void run_async(float4* p, float4* m, float4* v, const int* rows, int n_rows) {
    const size_t smem = (size_t)8 * 2 * 3 * R * 512;
    static bool init = false;
    if (!init) { CK(cudaFuncSetAttribute(k_rmw3_async<R, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); init = true; }
    k_rmw3_async<R, MINB><<<148 * MINB, 256, smem>>>(p, m, v, rows, n_rows);
}

int main(int argc, char** argv) {
    const int n_rows = argc > 1 ? atoi(argv[1]) : 59336;
    const int table_rows = (int)(n_rows * 2.8);
    const size_t tb = (size_t)table_rows * 512;
    float4 *p, *m, *v, *flush, *src;
    CK(cudaMalloc(&p, tb)); CK(cudaMalloc(&m, tb)); CK(cudaMalloc(&v, tb));
    CK(cudaMalloc(&src, (size_t)n_rows * 512 * 3)); CK(cudaMemset(src, 0, (size_t)n_rows * 512 * 3));
    const size_t fb = 512u << 20;
    CK(cudaMalloc(&flush, fb));
    CK(cudaMemset(p, 0, tb)); CK(cudaMemset(m, 0, tb)); CK(cudaMemset(v, 0, tb));
    std::vector<int> idx(table_rows);
    for (int i = 0; i < table_rows; ++i) idx[i] = i;
    srand(1);
    for (int i = table_rows - 1; i > 0; --i) std::swap(idx[i], idx[rand() % (i + 1)]);
    idx.resize(n_rows);
    std::sort(idx.begin(), idx.end());
    int* rows;
    CK(cudaMalloc(&rows, n_rows * 4));
    CK(cudaMemcpy(rows, idx.data(), n_rows * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double bytes = (double)n_rows * 512 * 6;
    const int grid = 148 * 4;
    for (int variant = 0; variant < 7; ++variant) {
        for (int flush_l2 = 0; flush_l2 < 2; ++flush_l2) {
            float best = 1e9f, sum = 0.f;
            const int reps = 20;
            for (int it = 0; it < reps + 3; ++it) {
                if (flush_l2) CK(cudaMemsetAsync(flush, it, fb));
                CK(cudaEventRecord(e0));
                if (variant == 0) k_copy<<<grid * 2, 256>>>(src, (float4*)flush, (size_t)(bytes / 2 / 16));
                else if (variant == 1) k_rmw3<false><<<grid, 256>>>(p, m, v, rows, n_rows);
                else if (variant == 2) k_rmw3<true><<<grid, 256>>>(p, m, v, rows, n_rows);
                else if (variant == 3) run_async<1, 8>(p, m, v, rows, n_rows);      // 24 KB/block, 8 blocks/SM
                else if (variant == 4) run_async<2, 4>(p, m, v, rows, n_rows);      // 48 KB/block, 4 blocks/SM
                else if (variant == 5) run_async<4, 2>(p, m, v, rows, n_rows);      // 96 KB/block, 2 blocks/SM
                else run_async<8, 1>(p, m, v, rows, n_rows);                        // 192 KB/block, 1 block/SM
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (it >= 3) { best = std::min(best, ms); sum += ms; }
            }
            const char* nm[] = {"copy", "rmw3", "gather3", "gather3 cp.async R=1 x8", "gather3 cp.async R=2 x4",
                                "gather3 cp.async R=4 x2", "gather3 cp.async R=8 x1"};
            printf("%-24s rows %d  %s  mean %.1f us = %.0f GB/s   best %.1f us = %.0f GB/s\n", nm[variant], n_rows,
                   flush_l2 ? "L2 flushed" : "L2 warm   ", sum / reps * 1e3, bytes / (sum / reps * 1e-3) / 1e9,
                   best * 1e3, bytes / (best * 1e-3) / 1e9);
        }
    }
    CK(cudaGetLastError());
    return 0;
}
