// Shared device helpers for the VFM kernels (sm_100a).
//
// Lane mapping used by every row kernel: a table row (entity mean | raw scale,
// 2d floats) is handled by a *group* of LPR lanes of one warp; lane l owns the
// vectors j = l + i*LPR (i < NV) of VEC floats each, i.e. elements
// k = VEC*j .. VEC*j+VEC-1.  VEC = 4 (128-bit ld/st) whenever d % 4 == 0, which
// makes every row 16-byte aligned (row pitch 8d bytes); VEC = 1 is the fallback
// for the scripts' odd default widths (d = 5, 3: vfm-torch.py:18,35).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vfmb {

constexpr int kMaxFields = 8;
constexpr int kWarp = 32;

template <int VEC> struct Vec;
template <> struct Vec<4> { float v[4]; };
template <> struct Vec<1> { float v[1]; };

template <int VEC>
__device__ __forceinline__ Vec<VEC> ld_vec(const float* p) {
    Vec<VEC> r;
    if constexpr (VEC == 4) {
        float4 t = *reinterpret_cast<const float4*>(p);
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
        r.v[0] = *p;
    }
    return r;
}

// read-only path for data written by an *earlier* kernel (scratch rows, plan)
template <int VEC>
__device__ __forceinline__ Vec<VEC> ld_vec_nc(const float* p) {
    Vec<VEC> r;
    if constexpr (VEC == 4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
        r.v[0] = __ldg(p);
    }
    return r;
}

// L2-coherent load (bypasses L1): data written by OTHER blocks of the running kernel
template <int VEC>
__device__ __forceinline__ Vec<VEC> ld_vec_cg(const float* p) {
    Vec<VEC> r;
    if constexpr (VEC == 4) {
        float4 t = __ldcg(reinterpret_cast<const float4*>(p));
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
        r.v[0] = __ldcg(p);
    }
    return r;
}

// streaming variants for data touched once per step (Adam moments): evict-first in L2 so that
// the L2-resident scratch of the gather kernels is not displaced
template <int VEC>
__device__ __forceinline__ Vec<VEC> ld_vec_cs(const float* p) {
    Vec<VEC> r;
    if constexpr (VEC == 4) {
        float4 t = __ldcs(reinterpret_cast<const float4*>(p));
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
        r.v[0] = __ldcs(p);
    }
    return r;
}
template <int VEC>
__device__ __forceinline__ void st_vec_cs(float* p, const Vec<VEC>& r) {
    if constexpr (VEC == 4) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
    } else {
        __stcs(p, r.v[0]);
    }
}

template <int VEC>
__device__ __forceinline__ void st_vec(float* p, const Vec<VEC>& r) {
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    } else {
        *p = r.v[0];
    }
}

// Kernels launched with launch_chained() (internal.h): wait for the predecessor grid to complete (its writes are
// visible afterwards), then let the successor be scheduled behind this grid.  No-ops for a plain launch.
__device__ __forceinline__ void chain_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
}

// mask of the LPR-lane group the calling lane belongs to.  Groups of one warp run
// independent grid-stride loops (different trip counts), so every shuffle inside such a
// loop must name only its own group -- a full-warp mask would wait for lanes that left.
template <int LPR>
__device__ __forceinline__ unsigned group_mask() {
    if constexpr (LPR == 32) return 0xffffffffu;
    else return ((1u << LPR) - 1u) << (((threadIdx.x & 31) / LPR) * LPR);
}

// sum over the LPR lanes of a group (LPR is a power of two <= 32; groups are
// aligned inside the warp, so xor-shuffles below LPR never leave the group)
template <int LPR>
__device__ __forceinline__ float group_sum(float x, unsigned mask) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) x += __shfl_xor_sync(mask, x, o);
    return x;
}

__device__ __forceinline__ float warp_sum(float x) { return group_sum<32>(x, 0xffffffffu); }
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// ---------------------------------------------------------------- cheap special functions
// Single-MUFU approximations (<= 2 ulp; denormal inputs flush to zero).  The step kernels are
// instruction-issue bound, and the IEEE division / square-root sequences dominated the Adam
// epilogue; the results stay far inside the 1e-5 parity tolerance.
__device__ __forceinline__ float fast_rcp(float x) {
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}
__device__ __forceinline__ float fast_sqrt(float x) {
    float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}
// L2 prefetch of the 128-byte line holding p: no register is tied up while it is in flight
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

// L2 residency hints (tuning l2_keep): k_stage pulls the rows k_adam_rows will read ~50 us later into L2 with
// evict_last priority so that the scratch traffic in between does not displace them; the row update reads
// them with an evict_first policy, which hands the lines back.
__device__ __forceinline__ void prefetch_l2_keep(const void* p) {
    asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(p));
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol)); return pol;
}

// ---------------------------------------------------------------- link functions
// LINK 0: abs (vfm-torch.py:126, vfm-tomasrch.py:201), 1: softplus (vfm-torch.py:125)
// A raw scale of exactly 0 (it happens: N(0,1) initialisation of 10^8 values returns a few exact
// zeros, on the CPU too) is sigma = 0 under the abs link -- an infinite KL here and a ValueError in the
// reference (torch Normal rejects scale = 0).  sigma is floored at 1e-18 (sigma^2 stays a normal
// float): everything stays finite, and nothing changes for |raw| >= 1e-18.
constexpr float kSigmaFloor = 1e-18f;
template <int LINK> __device__ __forceinline__ float link_fn(float raw) {
    if constexpr (LINK == 0) return fmaxf(fabsf(raw), kSigmaFloor);
    else return raw > 20.f ? raw : log1pf(expf(raw));      // torch softplus threshold
}
template <int LINK> __device__ __forceinline__ float link_grad(float raw) {
    if constexpr (LINK == 0) return raw > 0.f ? 1.f : (raw < 0.f ? -1.f : 0.f);   // sign(0)=0
    else return raw > 20.f ? 1.f : 1.f / (1.f + expf(-raw));
}

// KL(N(m,s) || N(pm,ps)) as torch _kl_normal_normal
__device__ __forceinline__ float kl_normal(float m, float s, float pm, float ps) {
    float r = s / ps;
    float vr = r * r;
    float t = (m - pm) / ps;
    return 0.5f * (vr + t * t - 1.f - logf(vr));
}
__device__ __forceinline__ float kl_std_normal(float m, float s) {
    float vr = s * s;
    return 0.5f * (vr + m * m - 1.f - logf(vr));
}
// same with the MUFU logarithm (abs. error 2^-21.4 for vr in [0.5, 2]); used per element in
// k_stage where the precise logf sequence was a quarter of all instructions
__device__ __forceinline__ float kl_std_normal_fast(float m, float s) {
    float vr = s * s;
    return 0.5f * (vr + m * m - 1.f - __logf(vr));
}

// ---------------------------------------------------------------- Philox4x32-10
// Counter-based RNG (Salmon et al. 2011).  Keyed by (seed); the counter is
// (row id, vector index, step, stream tag) so that the draw for an entity does
// not depend on which batch position, kernel or GPU asks for it.
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        // one 32x32->64 multiply per lane pair (IMAD.WIDE.U32) gives both halves
        uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    }
    __host__ __device__ static inline void block(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
        for (int r = 0; r < 10; ++r) { round(c, k0, k1); k0 += W0; k1 += W1; }
    }
};

constexpr uint32_t kTagEntity = 0u, kTagBias = 1u, kTagGlobal = 2u;

// four N(0,1) draws from one Philox block (Box-Muller on two uniform pairs)
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2,
                                               uint32_t c3, float (&out)[4]) {
    uint32_t c[4] = {c0, c1, c2, c3};
    Philox::block(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float kInv32 = 2.3283064365386963e-10f;       // 2^-32
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        float u1 = ((float)c[2 * p] + 0.5f) * kInv32;   // (0, 1]
        float u2 = ((float)c[2 * p + 1] + 0.5f) * kInv32;
        u1 = fminf(u1, 1.0f);
        float r = sqrtf(-2.0f * __logf(u1));
        float s, co;
        __sincosf(6.283185307179586f * u2, &s, &co);
        out[2 * p] = r * co;
        out[2 * p + 1] = r * s;
    }
}

// tag word: kind in the low 2 bits, variational-sample index above
__device__ __forceinline__ uint32_t philox_tag(uint32_t kind, int s) { return kind | ((uint32_t)s << 2); }

}  // namespace vfmb
