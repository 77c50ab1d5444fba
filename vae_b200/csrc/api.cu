// C-ABI glue: error reporting, version, closed-form scalar-block layout.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "internal.h"

namespace vfmb {

static thread_local char g_error[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

static thread_local cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;
void profile_events(cudaEvent_t* start, cudaEvent_t* stop) { *start = g_prof_start; *stop = g_prof_stop; }

unsigned long long g_launch_count = 0;
static Tuning g_tuning;
const Tuning& tuning() { return g_tuning; }

}  // namespace vfmb

extern "C" int vfmb_set_grid_reserve(int blocks_per_sm) {
    if (blocks_per_sm < 0 || blocks_per_sm > 8) return vfmb::set_error(VFMB_EINVAL, "vfmb_set_grid_reserve: 0..8");
    vfmb::g_tuning.grid_reserve = blocks_per_sm;
    return 0;
}

extern "C" int vfmb_set_tuning(const char* key, int value) {
    if (!key) return vfmb::set_error(VFMB_EINVAL, "vfmb_set_tuning: null key");
    vfmb::Tuning& t = vfmb::g_tuning;
    if (!strcmp(key, "grid_reserve")) return vfmb_set_grid_reserve(value);
    if (!strcmp(key, "fuse_score")) return 0;               // (the fused score + gather kernel is gone; accepted, ignored)
    if (!strcmp(key, "adam_reserve")) { t.adam_reserve = value != 0; return 0; }
    if (!strcmp(key, "adam_pipe")) { t.adam_pipe = value != 0; return 0; }
    if (!strcmp(key, "l2_keep")) { if (value < 0 || value > 7) return vfmb::set_error(VFMB_EINVAL, "vfmb_set_tuning: l2_keep 0..7"); t.l2_keep = value; return 0; }
    if (!strcmp(key, "pdl")) { t.pdl = value != 0; return 0; }
    if (!strcmp(key, "stage_chunk") || !strcmp(key, "score_chunk")) {
        if (value != 16 && value != 32) return vfmb::set_error(VFMB_EINVAL, "vfmb_set_tuning: %s must be 16 or 32", key);
        (key[1] == 't' ? t.stage_chunk : t.score_chunk) = value; return 0;
    }
    if (!strcmp(key, "stage_wide")) { t.stage_wide = value != 0; return 0; }
    if (!strcmp(key, "score_wide")) { t.score_wide = value < 0 ? -1 : (value != 0); return 0; }
    if (!strcmp(key, "gather_wide")) { t.gather_wide = value != 0; return 0; }
    if (!strcmp(key, "gather_dyn")) { t.gather_dyn = value != 0; return 0; }
    if (!strcmp(key, "gather_fence")) { t.gather_fence = value != 0; return 0; }
    if (!strcmp(key, "gather_keep")) { if (value < 1 || value > 4096) return vfmb::set_error(VFMB_EINVAL, "vfmb_set_tuning: gather_keep 1..4096"); t.gather_keep = value; return 0; }
    if (!strcmp(key, "prefetch_mv")) { if (value < 0 || value > 15) return vfmb::set_error(VFMB_EINVAL, "vfmb_set_tuning: prefetch_mv 0..15"); t.prefetch_mv = value; return 0; }
    return vfmb::set_error(VFMB_EINVAL, "vfmb_set_tuning: unknown key '%s'", key);
}

extern "C" int vfmb_profile_events(void* start_event, void* stop_event) {
    vfmb::g_prof_start = (cudaEvent_t)start_event;
    vfmb::g_prof_stop = (cudaEvent_t)stop_event;
    return 0;
}

extern "C" int64_t vfmb_launch_count(void) { return (int64_t)vfmb::g_launch_count; }
extern "C" const char* vfmb_last_error(void) { return vfmb::g_error; }
extern "C" int vfmb_version(void) { return 100; }

// closed-form scalar block: [0..8) globals, then bias prior mean[G], bias prior scale[G],
// (padded to a multiple of 4) entity prior mean[G*d], entity prior scale[G*d]
static int32_t entity_base(int32_t G) { return VFMB_C_GROUP_BASE + ((2 * G + 3) / 4) * 4; }
extern "C" int32_t vfmb_closed_off_bias_prior_mean(int32_t G, int32_t d, int32_t g) { (void)d; (void)G; return VFMB_C_GROUP_BASE + g; }
extern "C" int32_t vfmb_closed_off_bias_prior_scale(int32_t G, int32_t d, int32_t g) { (void)d; return VFMB_C_GROUP_BASE + G + g; }
extern "C" int32_t vfmb_closed_off_entity_prior_mean(int32_t G, int32_t d, int32_t g) { return entity_base(G) + g * d; }
extern "C" int32_t vfmb_closed_off_entity_prior_scale(int32_t G, int32_t d, int32_t g) { return entity_base(G) + G * d + g * d; }
extern "C" int32_t vfmb_closed_scalar_count(int32_t G, int32_t d) { return entity_base(G) + 2 * G * d; }
