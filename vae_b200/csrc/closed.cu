// Closed-form Gaussian VFM step (vfm-tomasrch.py:323-453, 569-594) and the plan-free
// posterior-mean prediction (vfm-torch.py:248-259, vfm-tomasrch.py:342-348).
#include "common.cuh"
#include "internal.h"

namespace vfmb {

// one lane group per sample; rows are read straight from the parameter table (means only)
template <int VEC, int LPR, int NV>
__global__ void __launch_bounds__(256)
k_predict_mean(int B, int F, int d, int R, const float* __restrict__ bias, const float* __restrict__ entity,
               float global_bias, const int64_t* __restrict__ x, float* __restrict__ out) {
    constexpr int GPW = kWarp / LPR;
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const int group = (threadIdx.x >> 5) * GPW + lane / LPR;
    const int groups_per_block = (blockDim.x >> 5) * GPW;
    for (int n = blockIdx.x * groups_per_block + group; n < B; n += gridDim.x * groups_per_block) {
        Vec<VEC> prod[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < VEC; ++j) prod[i].v[j] = 1.f;
        float bsum = 0.f;
        for (int f = 0; f < F; ++f) {
            int64_t id = x[(size_t)n * F + f];
            if (id < 0 || id >= R) id = 0;
            bsum += __ldg(bias + id * 2);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                int k = (gl + i * LPR) * VEC;
                if (k < d) {
                    Vec<VEC> a = ld_vec_nc<VEC>(entity + id * 2 * d + k);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) prod[i].v[j] *= a.v[j];
                }
            }
        }
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            int k = (gl + i * LPR) * VEC;
            if (k < d)
#pragma unroll
                for (int j = 0; j < VEC; ++j) part += prod[i].v[j];
        }
        float inter = group_sum<LPR>(part, group_mask<LPR>());
        if (gl == 0) out[n] = global_bias + bsum + inter;
    }
}

}  // namespace vfmb

using namespace vfmb;

#define VFMB_LAYOUT_SWITCH3(L, ...)                                                    \
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

extern "C" int vfmb_predict_mean(const vfmb_config* cfg, const float* bias, const float* entity,
                                 float global_bias, const int64_t* x, float* out, vfmb_stream stream_) {
    if (!cfg || !bias || !entity || !x || !out) return set_error(VFMB_EINVAL, "vfmb_predict_mean: null argument");
    if (cfg->B <= 0 || cfg->F < 1 || cfg->F > VFMB_MAX_FIELDS) return set_error(VFMB_EINVAL, "vfmb_predict_mean: bad B/F");
    Layout L;
    if (!pick_layout(cfg->d, &L)) return set_error(VFMB_ESHAPE, "unsupported embedding size %d", cfg->d);
    const int gpb = 8 * (32 / L.lpr);
    int64_t g = ((int64_t)cfg->B + gpb - 1) / gpb;
    if (g > kMaxGrid) g = kMaxGrid;
    VFMB_LAYOUT_SWITCH3(L, {
        k_predict_mean<VEC, LPR, NV><<<(int)g, 256, 0, (cudaStream_t)stream_>>>(
            cfg->B, cfg->F, cfg->d, cfg->R, bias, entity, global_bias, x, out);
    });
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_closed_forward(const vfmb_config*, const vfmb_tables*, const vfmb_plan*,
                                   const vfmb_step_io*, vfmb_stream) {
    return set_error(VFMB_ESHAPE, "vfmb_closed_forward: not built yet");
}
extern "C" int vfmb_closed_backward(const vfmb_config*, const vfmb_tables*, const vfmb_plan*,
                                    const vfmb_step_io*, const vfmb_adam*, int32_t, vfmb_stream) {
    return set_error(VFMB_ESHAPE, "vfmb_closed_backward: not built yet");
}
