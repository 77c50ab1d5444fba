// Sampled-ELBO VFM step, phase B: the row-update kernels (chain rule + KL gradient + Adam on the
// touched rows; vfm-torch.py:368-370) and the dense Adam sweep.  See sampled.cu for the step.
#include "sampled_common.cuh"

namespace vfmb {

// ------------------------------------------------------------------------------- k_adam_rows
// Backward, phase B (the HBM-bound kernel of the step): per unique row, chain rule from
// (g_v, g_w) to (mean, raw scale) + KL gradient, then Adam on the row -- parameters and both
// moments of every touched row are read and written exactly once.
//
// Row gradients are final in grow/gws (the gather kernel finished the rows cut by tile boundaries).
// FLAVOR 0  plain.
// FLAVOR 1  + the block that finishes last updates the scalar parameters and the step counter
//             (what k_final did as a separate launch).
// FLAVOR 2  + the count-rescaled KL of the rows (the row is in registers, the kernel is DRAM-bound
//             with idle issue slots) and, without injected noise, the Philox draws recomputed
//             instead of read back -- k_stage<LEAN> wrote neither.

// The scalar parameters (alpha, global bias) need only the sums k_score left in `stats`, not the row
// gradients, so their update is split from the end-of-kernel bookkeeping:
//   scalar_update   gradients of the scalar parameters + their Adam step (or the dense-gradient
//                   store).  One thread.  Leaves KL(global bias) of the PRE-update parameters in
//                   stats[VFMB_ST_KL] for finish_step.  Does not touch the step counter.
//   finish_step     by the block that finishes last: loss terms (KL of the rows), step counter.
// k_adam_rows_pipe runs scalar_update at kernel START on a block that takes no rows, which removes
// ~4 us of serial tail (a dozen dependent L2 round trips + two fp64 pow) from the critical path.
template <int LINK, int MODE>
__device__ __forceinline__ void scalar_update(const DevCfg& c, const FinalArgs& fa, const AdamDev& h,
                                              const int32_t* __restrict__ adam_step, float kl_scale, bool with_kl) {
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;            // Adam steps applied so far
    const uint32_t nstep = fa.noise_step ? (uint32_t)fa.noise_step[1] : 0u;   // noise index of the forward
    float* scalars = fa.scalars; float* stats = fa.stats;
    float alpha = scalars[VFMB_S_ALPHA], mu0 = scalars[VFMB_S_GB_MEAN], rho0 = scalars[VFMB_S_GB_SCALE];
    const float sig0 = link_fn<LINK>(rho0), ap = link_fn<LINK>(alpha);
    const double sr = (double)stats[VFMB_ST_SUM_RESID], sq = (double)stats[VFMB_ST_SUM_SQERR];
    // sum_s eps0_s * (sum_n dloss/dpred[s, n]); one sample: eps0 * sr
    double e0sr = 0.0;
    if (c.S > 1) {
        for (int q = 0; q < c.S; ++q) e0sr += (double)global_eps(fa.eps_global, c, nstep, q) * (double)stats[VFMB_ST_RESID_S + q];
    } else {
        e0sr = (double)global_eps(fa.eps_global, c, nstep) * sr;
    }
    if (with_kl) stats[VFMB_ST_KL] = kl_std_normal(mu0, sig0);   // pre-update; finish_step adds the rows' KL
                                                                 // (!with_kl: the forward already left the KL there)
    float g_mu0 = (float)(sr + (double)(kl_scale * mu0));
    float g_rho0 = link_grad<LINK>(rho0) * (float)(e0sr + (double)(kl_scale * (sig0 - 1.f / sig0)));
    float g_alpha = 0.f;
    if (fa.likelihood == VFMB_GAUSSIAN) {
        double sc = (double)c.n_train / ((double)c.S * (double)c.B);
        g_alpha = link_grad<LINK>(alpha) * (float)(sc * (0.5 * sq - 0.5 * (double)c.S * (double)c.B / (double)ap));
    }
    if (MODE == VFMB_ADAM_TOUCHED) {
        float ss, b2;
        adam_coeffs(h, (int)step + 1, &ss, &b2);
        adam_elem(mu0, fa.sm[VFMB_S_GB_MEAN], fa.sv[VFMB_S_GB_MEAN], g_mu0, h, ss, b2);
        adam_elem(rho0, fa.sm[VFMB_S_GB_SCALE], fa.sv[VFMB_S_GB_SCALE], g_rho0, h, ss, b2);
        scalars[VFMB_S_GB_MEAN] = mu0; scalars[VFMB_S_GB_SCALE] = rho0;
        if (fa.likelihood == VFMB_GAUSSIAN) {  // Bernoulli: alpha has no gradient, Adam skips it (N10)
            adam_elem(alpha, fa.sm[VFMB_S_ALPHA], fa.sv[VFMB_S_ALPHA], g_alpha, h, ss, b2);
            scalars[VFMB_S_ALPHA] = alpha;
        }
    } else if (fa.grad_scalars) {
        fa.grad_scalars[VFMB_S_ALPHA] = g_alpha;
        fa.grad_scalars[VFMB_S_GB_MEAN] = g_mu0;
        fa.grad_scalars[VFMB_S_GB_SCALE] = g_rho0;
    }
}

template <int MODE>
__device__ __forceinline__ void finish_step(const FinalArgs& fa, int32_t* __restrict__ adam_step, bool with_kl,
                                            double kl_rows, int U) {
    float* stats = fa.stats;
    if (with_kl) {                                      // loss terms of the pre-update parameters
        const float kl = stats[VFMB_ST_KL] + (float)kl_rows;
        stats[VFMB_ST_KL_ROWS] = (float)kl_rows;
        stats[VFMB_ST_KL] = kl;
        stats[VFMB_ST_LOSS] = (float)((double)stats[VFMB_ST_LOSS] + (double)kl);
        stats[VFMB_ST_U] = (float)U;
    }
    if (MODE == VFMB_ADAM_TOUCHED) adam_step[0] += 1;
}

template <int LINK, int MODE>
__device__ __forceinline__ void final_scalars(const DevCfg& c, const FinalArgs& fa, const AdamDev& h,
                                              int32_t* __restrict__ adam_step, float kl_scale,
                                              bool with_kl, double kl_rows, int U) {
    scalar_update<LINK, MODE>(c, fa, h, adam_step, kl_scale, with_kl);
    finish_step<MODE>(fa, adam_step, with_kl, kl_rows, U);
}

#ifndef VFMB_ADAM_MINB
#define VFMB_ADAM_MINB 4          // resident blocks/SM of the fused flavours (64 registers)
#endif
#ifndef VFMB_ADAM_SCHED
#define VFMB_ADAM_SCHED 0         // 0: grid-stride chunks of CH rows; 1: equal contiguous row ranges per warp
#endif
#ifndef VFMB_ADAM_PF
#define VFMB_ADAM_PF 1            // L2 prefetch of the chunk's parameter / moment rows up front
#endif
// bias row of one unique row: chain rule + KL gradient + Adam (or the dense-gradient store).
// klw: in c_u, out c_u * KL(N(a, tau) || N(0,1)) of the pre-update row (when KLF)
template <int LINK, int MODE, bool KLF>
__device__ __forceinline__ void bias_update(float* __restrict__ bias, float* __restrict__ bias_m,
                                            float* __restrict__ bias_v, float* __restrict__ grad_bias,
                                            int rowid, float gw, float eb, float cfac, const AdamDev& h,
                                            float step_size, float inv_bc2, float& klw) {
    const size_t boff = (size_t)rowid * 2;
    float2 ab = *reinterpret_cast<const float2*>(bias + boff);
    const float tau = link_fn<LINK>(ab.y);
    if (KLF) klw *= kl_std_normal(ab.x, tau);
    const float ga = fmaf(cfac, ab.x, gw);
    const float gb = link_grad<LINK>(ab.y) * fmaf(gw, eb, cfac * (tau - fast_rcp(tau)));
    if (MODE == VFMB_ADAM_TOUCHED) {
        float2 bm = *reinterpret_cast<const float2*>(bias_m + boff);
        float2 bv = *reinterpret_cast<const float2*>(bias_v + boff);
        adam_elem(ab.x, bm.x, bv.x, ga, h, step_size, inv_bc2);
        adam_elem(ab.y, bm.y, bv.y, gb, h, step_size, inv_bc2);
        *reinterpret_cast<float2*>(bias + boff) = ab;
        *reinterpret_cast<float2*>(bias_m + boff) = bm;
        *reinterpret_cast<float2*>(bias_v + boff) = bv;
    } else {
        *reinterpret_cast<float2*>(grad_bias + boff) = make_float2(ga, gb);
    }
}

template <int VEC, int LPR, int NV, int LINK, int MODE, int FLAVOR>
__global__ void __launch_bounds__(256, NV > 1 ? 2 : (FLAVOR == 0 ? 4 : VFMB_ADAM_MINB))
k_adam_rows(DevCfg c, float* __restrict__ bias, float* __restrict__ bias_m, float* __restrict__ bias_v,
            float* __restrict__ entity, float* __restrict__ entity_m, float* __restrict__ entity_v,
            const int32_t* __restrict__ urec, const int32_t* __restrict__ meta,
            const float* __restrict__ eps_bias, const float* __restrict__ eps_entity,
            const float* __restrict__ cq, const float* __restrict__ grow, const float* __restrict__ gws,
            AdamDev h, int32_t* __restrict__ adam_step, float kl_scale,
            float* __restrict__ grad_bias, float* __restrict__ grad_entity, FinalArgs fa) {
    constexpr int GPW = kWarp / LPR, CH = kRounds * GPW;
    constexpr bool KLF = FLAVOR == 2;
    const int U = meta[0];
    const int d = c.d;
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;
    const uint32_t nstep = fa.noise_step ? (uint32_t)fa.noise_step[1] : 0u;   // the forward's noise index
    const int lane = threadIdx.x & 31, gl = lane % LPR, gidx = lane / LPR;
    const unsigned gmask = group_mask<LPR>();
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    // bias corrections: fp64 pow per thread (~100 cycles per warp on the FP64 pipe) -- cheaper than
    // a block barrier in front of the first loads
    float step_size = 0.f, inv_bc2 = 1.f;
    if (MODE == VFMB_ADAM_TOUCHED) adam_coeffs(h, (int)step + 1, &step_size, &inv_bc2);
    float facc = 0.f;                                     // sum_u c_u * KL_u over this thread's rows

#if VFMB_ADAM_SCHED == 1
    // equal contiguous ranges: warp w handles rows [w U / nwarps, (w+1) U / nwarps), so all warps
    // finish together whatever U is
    const int lo = (int)((int64_t)gwarp * U / nwarps), hi = (int)((int64_t)(gwarp + 1) * U / nwarps);
    for (int base = lo; base < hi; base += CH) {
        const int ul = base + lane;
        const bool valid = lane < CH && ul < hi;
#else
    const int hi = U;
    for (int base = gwarp * CH; base < U; base += nwarps * CH) {
        // ---- lane-parallel: record, prefetch of the row's parameter / moment lines
        const int ul = base + lane;
        const bool valid = lane < CH && ul < U;
#endif
        int rowid_l = 0;
        float cfac_l = 0.f, klw_l = 0.f, gw_l = 0.f;      // KL weight c_u, c_u * KL(bias) of the lane's row, sum r_n
        if (valid) {
            const int4 rec = __ldg(reinterpret_cast<const int4*>(urec) + ul);
            rowid_l = rec.x;
            const size_t eoff = (size_t)rowid_l * 2 * d;
#if VFMB_ADAM_PF
            prefetch_row(entity + eoff, 8 * d);
            if (MODE == VFMB_ADAM_TOUCHED) {
                prefetch_row(entity_m + eoff, 8 * d);
                prefetch_row(entity_v + eoff, 8 * d);
            }
#endif
            const float cq_l = __ldg(cq + ul);
            cfac_l = kl_scale * cq_l;
            klw_l = cq_l;
            // bias row now: nothing of it stays live across the wide work
            gw_l = __ldg(gws + ul);
            bias_update<LINK, MODE, KLF>(bias, bias_m, bias_v, grad_bias, rowid_l, gw_l,
                                         __ldg(eps_bias + ul), cfac_l, h, step_size, inv_bc2, klw_l);
        }
        float klrow = 0.f;
        // ---- wide work: GPW rows per round
#pragma unroll 1
        for (int it = 0; it < kRounds; ++it) {
            const int sel = it * GPW + gidx;
            const int rowid = bcast(rowid_l, sel);
            const float cfac = bcast(cfac_l, sel);
            const float gwr = c.pairwise ? bcast(gw_l, sel) : 0.f;      // pairwise: own term (sum r_n) v_u
            const int u = base + sel;
            float kl = 0.f;
            if (u < hi) {
                const size_t eoff = (size_t)rowid * 2 * d;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    int k = (gl + i * LPR) * VEC;
                    if (k < d) {
                        Vec<VEC> e;                         // the noise k_stage used for this row
                        if (KLF) e = entity_eps<VEC>(eps_entity, c, u, rowid * c.row_stride + c.row_offset, k, nstep);
                        else e = ld_vec_nc<VEC>(eps_entity + (size_t)u * d + k);
                        Vec<VEC> mu = ld_vec<VEC>(entity + eoff + k), rho = ld_vec<VEC>(entity + eoff + d + k);
                        Vec<VEC> m1, m2, v1, v2;
                        if (MODE == VFMB_ADAM_TOUCHED) {
                            m1 = ld_vec_cs<VEC>(entity_m + eoff + k); m2 = ld_vec_cs<VEC>(entity_m + eoff + d + k);
                            v1 = ld_vec_cs<VEC>(entity_v + eoff + k); v2 = ld_vec_cs<VEC>(entity_v + eoff + d + k);
                        }
                        const Vec<VEC> g = ld_vec_nc<VEC>(grow + (size_t)u * d + k);
                        Vec<VEC> gmu, grho;
                        float quad = 0.f, prodv = 1.f;
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            const float sig = link_fn<LINK>(rho.v[j]);
                            const float gj = fmaf(-gwr, fmaf(e.v[j], sig, mu.v[j]), g.v[j]);   // gwr = 0 unless pairwise
                            gmu.v[j] = fmaf(cfac, mu.v[j], gj);
                            grho.v[j] = link_grad<LINK>(rho.v[j]) * fmaf(gj, e.v[j], cfac * (sig - fast_rcp(sig)));
                            if (KLF) {
                                const float vr = sig * sig;
                                quad += vr + mu.v[j] * mu.v[j] - 1.f;
                                prodv *= vr;
                            }
                        }
                        if (KLF) {       // sum_k KL(N(mu,sig)||N(0,1)), one logarithm per lane (as k_stage)
                            float lg = __logf(prodv);
                            if (!(prodv > 1e-30f && prodv < 1e30f)) {
                                lg = 0.f;
#pragma unroll
                                for (int j = 0; j < VEC; ++j) { const float sg = link_fn<LINK>(rho.v[j]); lg += logf(sg * sg); }
                            }
                            kl += 0.5f * (quad - lg);
                        }
                        if (MODE == VFMB_ADAM_TOUCHED) {
#pragma unroll
                            for (int j = 0; j < VEC; ++j) {
                                adam_elem(mu.v[j], m1.v[j], v1.v[j], gmu.v[j], h, step_size, inv_bc2);
                                adam_elem(rho.v[j], m2.v[j], v2.v[j], grho.v[j], h, step_size, inv_bc2);
                            }
                            st_vec<VEC>(entity + eoff + k, mu);        st_vec<VEC>(entity + eoff + d + k, rho);
                            st_vec_cs<VEC>(entity_m + eoff + k, m1);   st_vec_cs<VEC>(entity_m + eoff + d + k, m2);
                            st_vec_cs<VEC>(entity_v + eoff + k, v1);   st_vec_cs<VEC>(entity_v + eoff + d + k, v2);
                        } else {
                            st_vec<VEC>(grad_entity + eoff + k, gmu);  st_vec<VEC>(grad_entity + eoff + d + k, grho);
                        }
                    }
                }
                if (KLF) kl = group_sum<LPR>(kl, gmask);
            }
            if (KLF) hand_back<LPR>(klrow, kl, it, lane);
        }
        // ---- lane-parallel: KL of the rows -- klw_l is c_u * KL(bias row) (bias_update), klrow the
        // entity part of the row's KL
        if (KLF && valid) facc += fmaf(__ldg(cq + ul), klrow, klw_l);
    }
    if (FLAVOR >= 1) {
        // the block that finishes last owns the scalar parameters: every block has read the step
        // counter / Adam coefficients before it signalled, so updating them here is race-free
        double acc[1] = {(double)facc};
        if (block_partials<1>(acc, fa.partials, fa.counter)) {
            double tot[1] = {0.0};
            if (KLF) final_sums<1>(fa.partials, tot);
            if (threadIdx.x == 0) {
                final_scalars<LINK, MODE>(c, fa, h, adam_step, kl_scale, KLF, tot[0], U);
                *fa.counter = 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------- k_adam_rows_pipe
// The same row update (Adam on the touched rows, 128-bit rows: d % 4 == 0, one vector per lane),
// software-pipelined: k_adam_rows alternates "issue 6 loads -> wait -> ~250 instructions of chain rule,
// Philox and Adam -> 6 stores", so a warp has nothing in flight most of the time and the kernel sits at
// ~0.57 of the HBM peak while a bare gather of the same bytes reaches 0.76 (scripts/stream_ceiling.cu).
// Here the parameter / moment slices of the NEXT round travel with cp.async (LDGSTS: global -> shared
// memory, no registers held) while the current round is computed: every warp always has a round of
// rows (3 KB) in flight.  Each lane reads back exactly the 16-byte pieces it requested itself, so no
// barrier is involved -- only cp.async.wait_group.  Noise (Philox) and the gradient row do not depend on
// the staged data and are produced before the wait.
// Work split: equal contiguous ranges of unique rows per warp (all warps finish together); inside a
// range, chunks of 32 rows: the scalar work of a row (record, KL weight, bias row update) is done lane-
// parallel by all 32 lanes, then the warp walks the chunk in rounds of GPW rows.
// Shared memory: 8 warps x NSTAGE stages x 6 x 512 B = 48 KB per block at NSTAGE = 2 (4 blocks per SM).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16_hint(void* smem, const void* gmem, uint64_t pol) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(s), "l"(gmem), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

constexpr int kPipeStages = 2;

template <int LPR, int LINK, int FLAVOR>
__global__ void __launch_bounds__(256, 4)
k_adam_rows_pipe(DevCfg c, float* __restrict__ bias, float* __restrict__ bias_m, float* __restrict__ bias_v,
                 float* __restrict__ entity, float* __restrict__ entity_m, float* __restrict__ entity_v,
                 const int32_t* __restrict__ urec, const int32_t* __restrict__ meta,
                 const float* __restrict__ eps_bias, const float* __restrict__ eps_entity,
                 const float* __restrict__ cq, const float* __restrict__ grow, const float* __restrict__ gws,
                 AdamDev h, int32_t* __restrict__ adam_step, float kl_scale, FinalArgs fa) {
    chain_wait();                                           // launched with launch_chained()
    constexpr int VEC = 4, GPW = kWarp / LPR, NS = kPipeStages;
    constexpr bool KLF = FLAVOR == 2;
    constexpr bool DPF = FLAVOR == 3;    // mode B owner: scalar parameters + loss from the ranks' tail slots
    constexpr int MODE = VFMB_ADAM_TOUCHED;
    extern __shared__ float4 s_stage[];                   // [8 warps][NS][6][32 lanes]
    const int U = meta[0];
    const int d = c.d;
    const uint32_t step = (uint32_t)adam_step[0];
    const uint32_t nstep = fa.noise_step ? (uint32_t)fa.noise_step[1] : 0u;   // the forward's noise index
    const int lane = threadIdx.x & 31, gl = lane % LPR, gidx = lane / LPR, warp = threadIdx.x >> 5;
    const unsigned gmask = group_mask<LPR>();
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    float step_size, inv_bc2;
    adam_coeffs(h, (int)step + 1, &step_size, &inv_bc2);
    float4* my = s_stage + (size_t)warp * (NS * 6 * 32) + lane;
    const uint64_t pol = policy_evict_first();            // rows k_stage parked in L2 (fa.l2_demote) are handed back
    // FLAVOR >= 1: the last block takes no rows; its first thread updates the scalar parameters right
    // away (they depend on the sums k_score left, not on the row gradients) -- off the critical path
    const bool split = FLAVOR >= 1 && gridDim.x > 1;
    const int row_warps = split ? nwarps - (int)(blockDim.x >> 5) : nwarps;
    const bool scalar_block = split && blockIdx.x == gridDim.x - 1;
    auto scalars_now = [&]() {
        if (DPF) {                                        // the P ranks' additive scalars, added in rank order
            float tail[VFMB_DP_TAIL];
#pragma unroll
            for (int t = 0; t < VFMB_DP_TAIL; ++t) tail[t] = 0.f;
            for (int q = 0; q < fa.tail_P; ++q)
#pragma unroll
                for (int t = VFMB_DP_T_NLL; t <= VFMB_DP_T_OVERFLOW; ++t) tail[t] += __ldcg(fa.tail_slots + (size_t)q * fa.tail_pitch + t);
            fa.stats[VFMB_ST_KL_ROWS] = tail[VFMB_DP_T_KLROWS];
            dp_scalar_update<LINK>(fa.likelihood, c.S, fa.B_global, fa.n_train_global, c.seed, fa.scalars, fa.sm, fa.sv,
                                   tail, fa.eps_global, h, step, nstep, fa.stats);
            if (tail[VFMB_DP_T_OVERFLOW] > 0.f) fa.stats[VFMB_ST_LOSS] = __int_as_float(0x7fc00000);
        } else {
            scalar_update<LINK, MODE>(c, fa, h, adam_step, kl_scale, KLF);
        }
    };
    if (scalar_block && threadIdx.x == 0) scalars_now();
    const int lo = scalar_block ? 0 : (int)((int64_t)gwarp * U / row_warps);
    const int hi = scalar_block ? 0 : (int)((int64_t)(gwarp + 1) * U / row_warps);
    const int k = gl * VEC;                               // this lane's 4 elements of the mean / scale halves
    const bool kin = k < d;
    float facc = 0.f;                                     // sum_u c_u * KL_u over this thread's rows

    for (int cbase = lo; cbase < hi; cbase += 32) {
        // ---- lane-parallel: one unique row per lane
        const int ul = cbase + lane;
        const bool valid = ul < hi;
        int rowid_l = 0, seg0_l = 0, len_l = 0;
        float cfac_l = 0.f, klw_l = 0.f, cq_l = 0.f;
        if (valid) {
            const int4 rec = __ldg(reinterpret_cast<const int4*>(urec) + ul);
            rowid_l = rec.x; len_l = rec.y; seg0_l = rec.z;
            cq_l = __ldg(cq + ul);
            cfac_l = kl_scale * cq_l;
            klw_l = cq_l;
            if (DPF && rowid_l >= fa.n_real) len_l = 0;        // padding slots share a sentinel row: not a row
        }
        const int nrounds = (min(32, hi - cbase) + GPW - 1) / GPW;
        auto issue = [&](int r) {                          // warp-uniform call: the shuffle needs every lane
            const int sel = r * GPW + gidx;
            const int rowid = bcast(rowid_l, sel);
            if (cbase + sel < hi && kin) {
                const size_t eoff = (size_t)rowid * 2 * d + k;
                float4* dst = my + (r % NS) * (6 * 32);
                if (fa.l2_demote & 1) { cp_async16_hint(dst + 0 * 32, entity + eoff, pol); cp_async16_hint(dst + 1 * 32, entity + eoff + d, pol); }
                else { cp_async16(dst + 0 * 32, entity + eoff);     cp_async16(dst + 1 * 32, entity + eoff + d); }
                if (fa.l2_demote & 2) { cp_async16_hint(dst + 2 * 32, entity_m + eoff, pol); cp_async16_hint(dst + 3 * 32, entity_m + eoff + d, pol); }
                else { cp_async16(dst + 2 * 32, entity_m + eoff);   cp_async16(dst + 3 * 32, entity_m + eoff + d); }
                if (fa.l2_demote & 4) { cp_async16_hint(dst + 4 * 32, entity_v + eoff, pol); cp_async16_hint(dst + 5 * 32, entity_v + eoff + d, pol); }
                else { cp_async16(dst + 4 * 32, entity_v + eoff);   cp_async16(dst + 5 * 32, entity_v + eoff + d); }
            }
            cp_async_commit();
        };
        issue(0);
        // bias row of the lane's row while the first round is in flight
        float gw_l = 0.f;                                  // sum r_n of the lane's row
        if (DPF) {
            // mode B owner: the row's gradient = the slots its requesters stored (sorted occurrences of the
            // owner's plan, at most one per rank), added in source-rank order
            if (valid && len_l > 0) {
                for (int i = 0; i < len_l; ++i)
                    gw_l += __ldcg(fa.recv_grads + (size_t)__ldg(fa.occ + seg0_l + i) * fa.slot_pitch + d);
                bias_update<LINK, MODE, KLF>(bias, bias_m, bias_v, nullptr, rowid_l, gw_l, __ldg(eps_bias + ul), cfac_l, h,
                                             step_size, inv_bc2, klw_l);
            }
        } else if (valid) {
            gw_l = __ldg(gws + ul);
            bias_update<LINK, MODE, KLF>(bias, bias_m, bias_v, nullptr, rowid_l, gw_l, __ldg(eps_bias + ul),
                                         cfac_l, h, step_size, inv_bc2, klw_l);
        }
        float klrow = 0.f;
#pragma unroll 1
        for (int r = 0; r < nrounds; ++r) {
            if (r + 1 < nrounds) issue(r + 1); else cp_async_commit();   // one group per iteration, always
            const int sel = r * GPW + gidx;
            const int rowid = bcast(rowid_l, sel);
            const float cfac = bcast(cfac_l, sel);
            const int u = cbase + sel;
            const float gwr = c.pairwise ? bcast(gw_l, sel) : 0.f;      // pairwise: own term (sum r_n) v_u
            int n_req = 0, seg0 = 0;
            if (DPF) { n_req = bcast(len_l, sel); seg0 = bcast(seg0_l, sel); }
            const bool live = u < hi && kin && (!DPF || n_req > 0);
            // independent of the staged rows: the noise k_stage used for this row, the row gradient
            Vec<VEC> e, g;
            if (live) {
                if (KLF) e = entity_eps<VEC>(eps_entity, c, u, rowid * c.row_stride + c.row_offset, k, nstep);
                else e = ld_vec_nc<VEC>(eps_entity + (size_t)u * d + k);
                if (DPF) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) g.v[j] = 0.f;
                    for (int i = 0; i < n_req; ++i) {
                        const Vec<VEC> t = ld_vec_cg<VEC>(fa.recv_grads + (size_t)__ldg(fa.occ + seg0 + i) * fa.slot_pitch + k);
#pragma unroll
                        for (int j = 0; j < VEC; ++j) g.v[j] += t.v[j];
                    }
                } else {
                    g = ld_vec_nc<VEC>(grow + (size_t)u * d + k);
                }
            }
            cp_async_wait<1>();                            // round r has landed (round r+1 may still fly)
            float kl = 0.f;
            if (live) {
                const float4* src = my + (r % NS) * (6 * 32);
                const float4 a0 = src[0 * 32], a1 = src[1 * 32], a2 = src[2 * 32], a3 = src[3 * 32], a4 = src[4 * 32], a5 = src[5 * 32];
                Vec<VEC> mu = {a0.x, a0.y, a0.z, a0.w}, rho = {a1.x, a1.y, a1.z, a1.w};
                Vec<VEC> m1 = {a2.x, a2.y, a2.z, a2.w}, m2 = {a3.x, a3.y, a3.z, a3.w};
                Vec<VEC> v1 = {a4.x, a4.y, a4.z, a4.w}, v2 = {a5.x, a5.y, a5.z, a5.w};
                Vec<VEC> gmu, grho;
                float quad = 0.f, prodv = 1.f;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const float sig = link_fn<LINK>(rho.v[j]);
                    const float gj = fmaf(-gwr, fmaf(e.v[j], sig, mu.v[j]), g.v[j]);   // gwr = 0 unless pairwise
                    gmu.v[j] = fmaf(cfac, mu.v[j], gj);
                    grho.v[j] = link_grad<LINK>(rho.v[j]) * fmaf(gj, e.v[j], cfac * (sig - fast_rcp(sig)));
                    if (KLF) {
                        const float vr = sig * sig;
                        quad += vr + mu.v[j] * mu.v[j] - 1.f;
                        prodv *= vr;
                    }
                }
                if (KLF) {       // sum_k KL(N(mu,sig)||N(0,1)), one logarithm per lane (as k_stage)
                    float lg = __logf(prodv);
                    if (!(prodv > 1e-30f && prodv < 1e30f)) {
                        lg = 0.f;
#pragma unroll
                        for (int j = 0; j < VEC; ++j) { const float sg = link_fn<LINK>(rho.v[j]); lg += logf(sg * sg); }
                    }
                    kl = 0.5f * (quad - lg);
                }
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    adam_elem(mu.v[j], m1.v[j], v1.v[j], gmu.v[j], h, step_size, inv_bc2);
                    adam_elem(rho.v[j], m2.v[j], v2.v[j], grho.v[j], h, step_size, inv_bc2);
                }
                const size_t eoff = (size_t)rowid * 2 * d + k;
                st_vec<VEC>(entity + eoff, mu);        st_vec<VEC>(entity + eoff + d, rho);
                st_vec_cs<VEC>(entity_m + eoff, m1);   st_vec_cs<VEC>(entity_m + eoff + d, m2);
                st_vec_cs<VEC>(entity_v + eoff, v1);   st_vec_cs<VEC>(entity_v + eoff + d, v2);
            }
            if (KLF) {
                kl = group_sum<LPR>(kl, gmask);
                hand_back<LPR>(klrow, kl, r, lane);
            }
        }
        // KL of the rows: klw_l is c_u * KL(bias row) (bias_update), klrow the entity part
        if (KLF && valid) facc += fmaf(cq_l, klrow, klw_l);
    }
    cp_async_wait<0>();
    if (FLAVOR >= 1) {
        double acc[1] = {(double)facc};
        if (block_partials<1>(acc, fa.partials, fa.counter)) {
            double tot[1] = {0.0};
            if (KLF) final_sums<1>(fa.partials, tot);
            if (threadIdx.x == 0) {
                if (!split) scalars_now();
                finish_step<MODE>(fa, adam_step, KLF, tot[0], U);
                *fa.counter = 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------- k_adam_rows_multi
// S > 1: the row gradient is the sum over the S sampled copies of the row,
//   d/dmu = sum_s g_s + c_u mu,   d/drho = sign(rho) (sum_s g_s * eps_s + c_u (sigma - 1/sigma)),
//   d/da = S g_w + c_u a,         d/db = sign(b) (g_w sum_s eps^w_s + c_u (tau - 1/tau))
// (g_s = sum_n rho_n partner_s(n): k_gather on sample s; g_w = sum_n rho_n is the same for every s).
// One lane group per unique row, plain structure (S = 1 is the tuned path).  The block that finishes
// last updates the scalar parameters, as in k_adam_rows<FLAVOR 1>.
template <int VEC, int LPR, int NV, int LINK, int MODE>
__global__ void __launch_bounds__(256)
k_adam_rows_multi(DevCfg c, int u_stride, float* __restrict__ bias, float* __restrict__ bias_m,
                  float* __restrict__ bias_v, float* __restrict__ entity, float* __restrict__ entity_m,
                  float* __restrict__ entity_v, const int32_t* __restrict__ urec, const int32_t* __restrict__ meta,
                  const float* __restrict__ eps_bias, const float* __restrict__ eps_entity, int eps_stride,
                  const float* __restrict__ cq, const float* __restrict__ grow, const float* __restrict__ gws,
                  AdamDev h, int32_t* __restrict__ adam_step, float kl_scale,
                  float* __restrict__ grad_bias, float* __restrict__ grad_entity, FinalArgs fa) {
    constexpr int GPW = kWarp / LPR;
    const int U = meta[0], d = c.d, S = c.S;
    if (eps_stride < 0) eps_stride = U;                    // injected noise: [S, U, ...]; scratch: [S, u_stride, ...]
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const int group = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + lane / LPR;
    const int ngroups = gridDim.x * (blockDim.x >> 5) * GPW;
    float step_size = 0.f, inv_bc2 = 1.f;
    if (MODE == VFMB_ADAM_TOUCHED) adam_coeffs(h, (int)step + 1, &step_size, &inv_bc2);

    for (int u = group; u < U; u += ngroups) {
        const int rowid = __ldg(urec + 4 * (size_t)u);
        const float cfac = kl_scale * __ldg(cq + u);
        const size_t eoff = (size_t)rowid * 2 * d;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            int k = (gl + i * LPR) * VEC;
            if (k < d) {
                Vec<VEC> mu = ld_vec<VEC>(entity + eoff + k), rho = ld_vec<VEC>(entity + eoff + d + k);
                Vec<VEC> gs, ges;
#pragma unroll
                for (int j = 0; j < VEC; ++j) { gs.v[j] = 0.f; ges.v[j] = 0.f; }
                const float gwr = c.pairwise ? __ldg(gws + u) : 0.f;       // pairwise: own term (sum r_n) v_{s,u}
                for (int q = 0; q < S; ++q) {
                    const Vec<VEC> g = ld_vec_nc<VEC>(grow + ((size_t)q * u_stride + u) * d + k);
                    const Vec<VEC> e = ld_vec_nc<VEC>(eps_entity + ((size_t)q * eps_stride + u) * d + k);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        const float gj = fmaf(-gwr, fmaf(e.v[j], link_fn<LINK>(rho.v[j]), mu.v[j]), g.v[j]);
                        gs.v[j] += gj; ges.v[j] = fmaf(gj, e.v[j], ges.v[j]);
                    }
                }
                Vec<VEC> gmu, grho;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const float sig = link_fn<LINK>(rho.v[j]);
                    gmu.v[j] = fmaf(cfac, mu.v[j], gs.v[j]);
                    grho.v[j] = link_grad<LINK>(rho.v[j]) * (ges.v[j] + cfac * (sig - fast_rcp(sig)));
                }
                if (MODE == VFMB_ADAM_TOUCHED) {
                    Vec<VEC> m1 = ld_vec<VEC>(entity_m + eoff + k), m2 = ld_vec<VEC>(entity_m + eoff + d + k);
                    Vec<VEC> v1 = ld_vec<VEC>(entity_v + eoff + k), v2 = ld_vec<VEC>(entity_v + eoff + d + k);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        adam_elem(mu.v[j], m1.v[j], v1.v[j], gmu.v[j], h, step_size, inv_bc2);
                        adam_elem(rho.v[j], m2.v[j], v2.v[j], grho.v[j], h, step_size, inv_bc2);
                    }
                    st_vec<VEC>(entity + eoff + k, mu);     st_vec<VEC>(entity + eoff + d + k, rho);
                    st_vec<VEC>(entity_m + eoff + k, m1);   st_vec<VEC>(entity_m + eoff + d + k, m2);
                    st_vec<VEC>(entity_v + eoff + k, v1);   st_vec<VEC>(entity_v + eoff + d + k, v2);
                } else {
                    st_vec<VEC>(grad_entity + eoff + k, gmu);  st_vec<VEC>(grad_entity + eoff + d + k, grho);
                }
            }
        }
        if (gl == 0) {                                     // bias row
            const size_t boff = (size_t)rowid * 2;
            float2 ab = *reinterpret_cast<const float2*>(bias + boff);
            const float gw = __ldg(gws + u);
            float ebsum = 0.f;
            for (int q = 0; q < S; ++q) ebsum += __ldg(eps_bias + (size_t)q * eps_stride + u);
            const float tau = link_fn<LINK>(ab.y);
            const float ga = fmaf(cfac, ab.x, (float)S * gw);
            const float gb = link_grad<LINK>(ab.y) * fmaf(gw, ebsum, cfac * (tau - fast_rcp(tau)));
            if (MODE == VFMB_ADAM_TOUCHED) {
                float2 bm = *reinterpret_cast<const float2*>(bias_m + boff);
                float2 bv = *reinterpret_cast<const float2*>(bias_v + boff);
                adam_elem(ab.x, bm.x, bv.x, ga, h, step_size, inv_bc2);
                adam_elem(ab.y, bm.y, bv.y, gb, h, step_size, inv_bc2);
                *reinterpret_cast<float2*>(bias + boff) = ab;
                *reinterpret_cast<float2*>(bias_m + boff) = bm;
                *reinterpret_cast<float2*>(bias_v + boff) = bv;
            } else {
                *reinterpret_cast<float2*>(grad_bias + boff) = make_float2(ga, gb);
            }
        }
    }
    double acc[1] = {0.0};
    if (block_partials<1>(acc, fa.partials, fa.counter)) {
        if (threadIdx.x == 0) {
            final_scalars<LINK, MODE>(c, fa, h, adam_step, kl_scale, false, 0.0, U);
            *fa.counter = 0;
        }
    }
}

// ------------------------------------------------------------------------------- dense Adam
__global__ void __launch_bounds__(256)
k_adam_dense(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
             const float* __restrict__ g, int64_t n, AdamDev h, const int32_t* __restrict__ adam_step) {
    __shared__ float s_coef[2];
    if (threadIdx.x == 0) adam_coeffs(h, adam_step[0] + 1, &s_coef[0], &s_coef[1]);
    __syncthreads();
    const float ss = s_coef[0], b2 = s_coef[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float pi = p[i], mi = m[i], vi = v[i];
        adam_elem(pi, mi, vi, g[i], h, ss, b2);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}
__global__ void k_step_advance(int32_t* adam_step) { adam_step[0] += 1; }

}  // namespace vfmb

using namespace vfmb;

// flavor: see k_adam_rows
int launch_adam(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                       const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode, float kl_grad_scale,
                       int flavor, vfmb_stream stream_, const DpTail* dp) {
    Prep P;
    int rc = prep(cfg, "vfmb_sampled_adam_rows", stream_, 1, &P);
    if (rc) return rc;
    if (!tab || !plan || !io) return set_error(VFMB_EINVAL, "vfmb_sampled_adam_rows: null argument");
    if (mode == VFMB_ADAM_TOUCHED && (!adam || !tab->entity_m || !tab->entity_v || !tab->bias_m || !tab->bias_v || !tab->adam_step))
        return set_error(VFMB_EINVAL, "vfmb_sampled_adam_rows: Adam state required");
    if (mode == VFMB_GRAD_ONLY && (!io->grad_bias || !io->grad_entity))
        return set_error(VFMB_EINVAL, "vfmb_sampled_adam_rows: gradient outputs required");
    if (mode != VFMB_ADAM_TOUCHED && mode != VFMB_GRAD_ONLY) return set_error(VFMB_EINVAL, "vfmb_sampled_adam_rows: bad mode");
    if (((!io->grow || !io->gws) && flavor != 3) || !io->cq) return set_error(VFMB_EINVAL, "vfmb_sampled_adam_rows: scratch required");
    if (flavor >= 1 && mode == VFMB_ADAM_TOUCHED && (!tab->scalars_m || !tab->scalars_v))
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: Adam state required");
    if (flavor >= 1 && (!tab->scalars || !io->stats || !io->partials || !io->counters))
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: scalars / stats required");
    const Layout& L = P.L; cudaStream_t stream = P.stream; const int ch = P.ch; const auto& cap = P.cap;
    const DevCfg& dc = P.dc;
    AdamDev h = make_adam(adam);
    // the noise the forward used: injected arrays, or what k_stage wrote to scratch (Philox);
    // flavor 2 recomputes the Philox draws of the rows instead
    const float* eps_e = io->eps_entity ? io->eps_entity : (flavor == 2 ? nullptr : io->es);
    const float* eps_b = io->eps_bias ? io->eps_bias : io->ebs;
    if (!eps_b || (flavor != 2 && !eps_e)) return set_error(VFMB_EINVAL, "vfmb_sampled_adam_rows: noise of the forward required");
    FinalArgs fa{};
    fa.scalars = tab->scalars; fa.sm = tab->scalars_m; fa.sv = tab->scalars_v; fa.stats = io->stats;
    fa.eps_global = io->eps_global; fa.grad_scalars = io->grad_scalars; fa.partials = io->partials;
    fa.counter = io->counters ? io->counters + 2 : nullptr;
    fa.gslot = io->partials ? (const float*)io->partials + scratch_map(cfg->B, cfg->F, cfg->d, cap.u_cap).gslot_off : nullptr;
    fa.likelihood = cfg->likelihood; fa.noise_step = tab->noise_step;
    fa.l2_demote = (flavor == 2) ? tuning().l2_keep : 0;    // the fused step's lean k_stage parked the rows
    if (flavor == 3) {
        if (!dp || !dp->tail_slots || !dp->stats_out) return set_error(VFMB_EINVAL, "vfmb_shard_owner_update: tail slots required");
        if (!(mode == VFMB_ADAM_TOUCHED && P.L.vec == 4 && P.L.nv == 1 && tuning().adam_pipe != 0))
            return set_error(VFMB_ESHAPE, "vfmb_shard_owner_update: fused update needs d %% 4 == 0, d <= 128");
        fa.tail_slots = dp->tail_slots; fa.tail_P = dp->P; fa.tail_pitch = dp->pitch;
        fa.B_global = dp->B_global; fa.n_train_global = dp->n_train_global; fa.stats = dp->stats_out;
        fa.eps_global = dp->eps_global;
        fa.recv_grads = dp->recv_grads; fa.occ = plan->occ; fa.slot_pitch = dp->slot_pitch; fa.n_real = dp->n_real;
    }
    // the HBM-bound kernel keeps its full wave even next to the plan (measured: 113.8 vs 116.2 us/step)
    const bool adam_reserve = tuning().adam_reserve != 0;
    if (mode == VFMB_ADAM_TOUCHED && L.vec == 4 && L.nv == 1 && tuning().adam_pipe != 0) {
        // software-pipelined row update (cp.async staging), see k_adam_rows_pipe
        const size_t smem = (size_t)8 * kPipeStages * 6 * 32 * sizeof(float4);
        cudaEvent_t ev0, ev1;
        profile_events(&ev0, &ev1);
        if (ev0 && ev1) record_profile_event(ev0, stream);
#define LAUNCH_PIPE(LPR_, LINK, FLAVOR)                                                                   \
        do {                                                                                              \
            auto kern = k_adam_rows_pipe<LPR_, LINK, FLAVOR>;                                             \
            static int blocks = 0;                                                                        \
            if (blocks == 0) {                                                                            \
                CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                int per_sm = 0;                                                                           \
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));        \
                blocks = (per_sm < 1 ? 1 : per_sm) * kNumSMs;                                             \
            }                                                                                             \
            int grid = blocks;                                                                            \
            if (adam_reserve && grid_reserve() > 0 && grid / kNumSMs > grid_reserve() + 1) grid -= grid_reserve() * kNumSMs; \
            const int64_t need = (cap.u_cap + 63) / 64;        /* >= 8 rows per warp */                    \
            if (grid > need) grid = (int)(need < 1 ? 1 : need);                                           \
            CUDA_TRY(launch_chained(kern, grid, 256, smem, stream, dc, tab->bias, tab->bias_m, tab->bias_v, tab->entity, \
                tab->entity_m, tab->entity_v, plan->urec, plan->meta, eps_b, eps_e, io->cq, io->grow, io->gws, \
                h, tab->adam_step, kl_grad_scale, fa));                                                   \
        } while (0)
#define LAUNCH_PIPE_F(LPR_, LINK)                                                                         \
        do { if (flavor == 0) LAUNCH_PIPE(LPR_, LINK, 0); else if (flavor == 1) LAUNCH_PIPE(LPR_, LINK, 1); \
             else if (flavor == 2) LAUNCH_PIPE(LPR_, LINK, 2); else LAUNCH_PIPE(LPR_, LINK, 3); } while (0)
#define LAUNCH_PIPE_L(LPR_)                                                                               \
        do { if (cfg->link == VFMB_LINK_ABS) LAUNCH_PIPE_F(LPR_, 0); else LAUNCH_PIPE_F(LPR_, 1); } while (0)
        if (L.lpr == 4) LAUNCH_PIPE_L(4);
        else if (L.lpr == 8) LAUNCH_PIPE_L(8);
        else if (L.lpr == 16) LAUNCH_PIPE_L(16);
        else LAUNCH_PIPE_L(32);
#undef LAUNCH_PIPE_L
#undef LAUNCH_PIPE_F
#undef LAUNCH_PIPE
        if (ev0 && ev1) record_profile_event(ev1, stream);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
#define LAUNCH_ADAM(LINK, MODE, FLAVOR)                                                                  \
    k_adam_rows<VEC, LPR, NV, LINK, MODE, FLAVOR><<<grid_resident(k_adam_rows<VEC, LPR, NV, LINK, MODE, FLAVOR>, cap.u_cap, ch, 0, adam_reserve), 256, 0, counted(stream)>>>( \
        dc, tab->bias, tab->bias_m, tab->bias_v, tab->entity, tab->entity_m, tab->entity_v,              \
        plan->urec, plan->meta, eps_b, eps_e, io->cq, io->grow, io->gws, h, tab->adam_step,              \
        kl_grad_scale, io->grad_bias, io->grad_entity, fa)
#define LAUNCH_ADAM_F(LINK, MODE)                                                                        \
    do { if (flavor == 0) LAUNCH_ADAM(LINK, MODE, 0); else if (flavor == 1) LAUNCH_ADAM(LINK, MODE, 1);  \
         else LAUNCH_ADAM(LINK, MODE, 2); } while (0)
    VFMB_LAYOUT_SWITCH(L, {
        cudaEvent_t ev0, ev1;
        profile_events(&ev0, &ev1);
        if (ev0 && ev1) record_profile_event(ev0, stream);
        if (cfg->link == VFMB_LINK_ABS) {
            if (mode == VFMB_ADAM_TOUCHED) LAUNCH_ADAM_F(0, VFMB_ADAM_TOUCHED); else LAUNCH_ADAM_F(0, VFMB_GRAD_ONLY);
        } else {
            if (mode == VFMB_ADAM_TOUCHED) LAUNCH_ADAM_F(1, VFMB_ADAM_TOUCHED); else LAUNCH_ADAM_F(1, VFMB_GRAD_ONLY);
        }
        if (ev0 && ev1) record_profile_event(ev1, stream);
    });
#undef LAUNCH_ADAM_F
#undef LAUNCH_ADAM
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_sampled_adam_rows(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                                      const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode,
                                      float kl_grad_scale, vfmb_stream stream) {
    VFMB_PHASE_S1("vfmb_sampled_adam_rows");
    return launch_adam(cfg, tab, plan, io, adam, mode, kl_grad_scale, 0, stream);
}

// S > 1: k_adam_rows_multi on the [S][u_cap] gradients / noise
int launch_adam_multi(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                             const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode, float kl_grad_scale,
                             vfmb_stream stream_) {
    Prep P;
    int rc = prep(cfg, "vfmb_sampled_backward", stream_, 2, &P);
    if (rc) return rc;
    if (mode == VFMB_ADAM_TOUCHED && (!adam || !tab->entity_m || !tab->entity_v || !tab->bias_m || !tab->bias_v ||
                                      !tab->adam_step || !tab->scalars_m || !tab->scalars_v))
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: Adam state required");
    if (mode == VFMB_GRAD_ONLY && (!io->grad_bias || !io->grad_entity))
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: gradient outputs required");
    if (mode != VFMB_ADAM_TOUCHED && mode != VFMB_GRAD_ONLY) return set_error(VFMB_EINVAL, "vfmb_sampled_backward: bad mode");
    if (!io->grow || !io->gws || !io->cq || !tab->scalars || !io->stats || !io->partials || !io->counters)
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: scratch required");
    // the noise of the forward: injected arrays [S, U, ...] or what k_stage kept [S, u_cap, ...]
    const float* eps_e = io->eps_entity ? io->eps_entity : io->es;
    const float* eps_b = io->eps_bias ? io->eps_bias : io->ebs;
    if (!eps_e || !eps_b || (io->eps_entity != nullptr) != (io->eps_bias != nullptr))
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: noise of the forward required (both arrays injected, or none)");
    const int eps_stride = io->eps_entity ? -1 : (int)P.cap.u_cap;
    const Layout& L = P.L; const DevCfg& dc = P.dc; cudaStream_t stream = P.stream; const auto& cap = P.cap;
    AdamDev h = make_adam(adam);
    FinalArgs fa{};
    fa.scalars = tab->scalars; fa.sm = tab->scalars_m; fa.sv = tab->scalars_v; fa.stats = io->stats;
    fa.eps_global = io->eps_global; fa.grad_scalars = io->grad_scalars; fa.partials = io->partials;
    fa.counter = io->counters + 2; fa.gslot = nullptr; fa.likelihood = cfg->likelihood; fa.noise_step = tab->noise_step;
    int grid = (int)((cap.u_cap + 8 * (32 / L.lpr) - 1) / (8 * (32 / L.lpr)));
    if (grid > kGridCap) grid = kGridCap;
#define LAUNCH_AM(LINK, MODE)                                                                            \
    k_adam_rows_multi<VEC, LPR, NV, LINK, MODE><<<grid, 256, 0, counted(stream)>>>(                               \
        dc, (int)cap.u_cap, tab->bias, tab->bias_m, tab->bias_v, tab->entity, tab->entity_m, tab->entity_v, \
        plan->urec, plan->meta, eps_b, eps_e, eps_stride, io->cq, io->grow, io->gws, h, tab->adam_step,  \
        kl_grad_scale, io->grad_bias, io->grad_entity, fa)
    VFMB_LAYOUT_SWITCH(L, {
        if (cfg->link == VFMB_LINK_ABS) {
            if (mode == VFMB_ADAM_TOUCHED) LAUNCH_AM(0, VFMB_ADAM_TOUCHED); else LAUNCH_AM(0, VFMB_GRAD_ONLY);
        } else {
            if (mode == VFMB_ADAM_TOUCHED) LAUNCH_AM(1, VFMB_ADAM_TOUCHED); else LAUNCH_AM(1, VFMB_GRAD_ONLY);
        }
    });
#undef LAUNCH_AM
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_adam_dense(float* p, float* m, float* v, const float* g, int64_t n, const vfmb_adam* adam,
                               const int32_t* adam_step, vfmb_stream stream_) {
    if (!p || !m || !v || !g || !adam || !adam_step || n < 0) return set_error(VFMB_EINVAL, "vfmb_adam_dense: bad argument");
    if (n == 0) return 0;
    AdamDev h = make_adam(adam);
    int64_t grid = (n + 255) / 256;
    if (grid > 16 * kNumSMs) grid = 16 * kNumSMs;
    k_adam_dense<<<(int)grid, 256, 0, counted((cudaStream_t)stream_)>>>(p, m, v, g, n, h, adam_step);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_adam_step_advance(int32_t* adam_step, vfmb_stream stream_) {
    if (!adam_step) return set_error(VFMB_EINVAL, "vfmb_adam_step_advance: null");
    k_step_advance<<<1, 1, 0, counted((cudaStream_t)stream_)>>>(adam_step);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

