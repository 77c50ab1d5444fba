// Host-side internals shared by the translation units of libvfm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vfm_b200.h"

namespace vfmb {

constexpr int kTile = 32;           // sorted positions per backward tile (k_gather)
constexpr int kHotPartials = 32;    // a row spanning more tiles than this is listed as "hot" in the plan
constexpr int kNumSMs = 148;        // B200
constexpr int kPlanGrid = 2 * kNumSMs;
constexpr int kMaxGrid = 4 * kNumSMs;   // grid-stride kernels: one resident wave at 4 blocks/SM
constexpr int kPriorGrid = 2 * kNumSMs; // blocks that carry prior-gradient partials (closed form)

// entries of vfmb_plan.hot: every backward-tile boundary cuts at most one row (front of the list),
// a hot row spans > kHotPartials tiles (back of the list)
static inline int64_t cut_list_capacity(int64_t n_tiles) { return n_tiles + n_tiles / (kHotPartials - 2) + 4; }

int set_error(int code, const char* fmt, ...);
// Process-wide launch tuning (vfmb_set_tuning / vfmb_set_grid_reserve); read at launch time on the
// host, never by device code.  Defaults are the measured best for the BASELINE shapes.
struct Tuning {
    int grid_reserve = 0;    // block slots per SM the persistent step kernels leave free
    int adam_reserve = 1;    // 1: k_adam_rows also leaves the reserved slot free (round 2, pipelined kernel: 104.2 vs
                             //   108.4 us per ml20m step next to the plan; 0 was better for the register-staged kernel)
    int adam_pipe = 1;       // Adam on the touched rows: cp.async-pipelined kernel (k_adam_rows_pipe) when the row
                             //   layout allows (d % 4 == 0, d <= 128); 0 = the register-staged k_adam_rows
    int gather_dyn = 1;      // backward gather: tiles after a group's first are handed out by an atomic counter
                             //   (0 = static stride); results do not depend on it
    int stage_chunk = 16;    // k_stage / k_score: units (unique rows / samples) per warp and pass, 16 or 32 (measured:
    int score_chunk = 32;    //   k_stage 42 -> 36 us on sideinfo with 16, neutral on ml20m; k_score loses with 16)
    int stage_wide = 0;      // k_stage: half the lanes per row, two vectors per lane
    int score_wide = -1;     // k_score: half the lanes per row, two vectors per lane; -1 = when F > 2 (another summation order of the
                             //   dot product: last-bit differences, so one setting per process)
    int gather_wide = 1;     // backward gather: half the lanes per row, two vectors per lane.  This also halves the group
                             //   tile, i.e. moves the cuts of a long row's sum: last-bit differences, one setting per process
    int gather_keep = 32;    // backward gather: rows of up to this many occurrences are never cut by a tile boundary
    int gather_fence = 1;    // finisher of cut rows: 1 = fence.acq_rel.gpu, 0 = __threadfence() (fence.sc)
    int pdl = 0;             // programmatic dependent launch along the step's and the plan's kernel chains: the next
                             //   kernel is scheduled behind its predecessor and waits (griddepcontrol.wait) for it.
                             //   Measured: 115.5 vs 104.6 us per ml20m step -- the grids are sized to the resident
                             //   capacity, so the early blocks only take slots; off by default
    int l2_keep = 1;         // fused step: k_stage parks the rows of entity (bit 0) / m (bit 1) / v (bit 2) in L2 with
                             //   evict_last priority for the row update, which reads them evict_first.  Measured per ml20m
                             //   step: 0: 104.9 us (update 49.3), 1: 102.7 (45.6), 3: 107.9 (42.8), 7: 112.2 (41.8) -- the
                             //   moments do not fit next to the scratch; having k_score prefetch them instead lost too
    int prefetch_mv = 0;     // fused step: earlier phases pull the Adam moments of the touched rows into L2
                             //   (bit 0: k_stage fetches m, bit 1: k_stage fetches v, bit 2: k_gather fetches m,
                             //    bit 3: k_gather fetches v)
};
const Tuning& tuning();
static inline int grid_reserve() { return tuning().grid_reserve; }
// every kernel launch of the library passes its stream through counted(): vfmb_launch_count() is the
// number of own kernels enqueued (or captured into a CUDA graph) by this process so far
extern unsigned long long g_launch_count;
static inline cudaStream_t counted(cudaStream_t s) { ++g_launch_count; return s; }
// Launch of a kernel in a dependency chain.  With tuning().pdl the launch carries the programmatic-serialization
// attribute: the grid may be scheduled once every block of its predecessor in the stream has passed its own
// chain_wait() (or exited); it must therefore call chain_wait() (device side, common.cuh) before touching anything
// an earlier kernel wrote.  Captured into a CUDA graph this becomes a programmatic edge.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chained(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream,
                                         Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem; cfg.stream = counted(stream);
    cudaLaunchAttribute at[1];
    if (tuning().pdl) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// optional events around the dominant kernel (see vfmb_profile_events)
void profile_events(cudaEvent_t* start, cudaEvent_t* stop);
// record a measurement event; on a capturing stream as an external event-record NODE (a plain record would
// only mark a dependency inside the capture and never fire at replay)
static inline void record_profile_event(cudaEvent_t ev, cudaStream_t stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) == cudaSuccess && st == cudaStreamCaptureStatusActive)
        cudaEventRecordWithFlags(ev, stream, cudaEventRecordExternal);
    else
        cudaEventRecord(ev, stream);
}

#define CUDA_TRY(expr)                                                                  \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess)                                                          \
            return ::vfmb::set_error((int)_e, "%s failed: %s (%s:%d)", #expr,           \
                                     cudaGetErrorString(_e), __FILE__, __LINE__);       \
    } while (0)

// lane layout for a given embedding size (see common.cuh)
struct Layout {
    int vec;   // 4 or 1
    int lpr;   // lanes per row: 4, 8, 16, 32
    int nv;    // vectors per lane: 1 or 2
};
bool pick_layout(int d, Layout* out);

}  // namespace vfmb
