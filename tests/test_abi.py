"""CPU-side checks of the C-ABI shared library: it builds for sm_100a, loads,
and exports every symbol ``include/vfm_b200.h`` declares (no compute calls)."""
import ctypes as C
import os
import re

from vae_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vfm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vfmb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    L.build()
    lib = L.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in vfm_b200.h but not exported"
    assert sorted(L.SYMBOLS) == declared, "ctypes table and header disagree"
    assert lib.vfmb_version() >= 100


def test_host_only_queries_and_error_reporting():
    lib = L.lib()
    cap = L.PlanCapacity()
    assert lib.vfmb_plan_capacity(65536, 2, 165237, C.byref(cap)) == 0
    assert cap.u_cap == 131072 and cap.n_tiles == 131072 // cap.tile and cap.workspace_bytes > 0
    assert lib.vfmb_plan_capacity(8, 2, 5, C.byref(cap)) == 0 and cap.u_cap == 5
    rc = lib.vfmb_plan_capacity(0, 2, 10, C.byref(cap))
    assert rc == 10001 and b"vfmb_plan_capacity" in lib.vfmb_last_error()
    assert lib.vfmb_plan_capacity(4, 9, 10, C.byref(cap)) != 0          # F > VFMB_MAX_FIELDS
    G, d = 3, 8
    offs = [lib.vfmb_closed_off_bias_prior_mean(G, d, 0), lib.vfmb_closed_off_bias_prior_scale(G, d, 0),
            lib.vfmb_closed_off_entity_prior_mean(G, d, 0), lib.vfmb_closed_off_entity_prior_scale(G, d, 0)]
    assert offs == sorted(offs) and offs[0] == 8 and offs[2] % 4 == 0
    assert lib.vfmb_closed_scalar_count(G, d) == offs[3] + G * d


def test_struct_sizes_match_header():
    # the ctypes mirrors must have the C layout: 8 int32 + 8 int32 + 8 float + float (+pad) + u64
    assert C.sizeof(L.Config) == 8 * 4 + 8 * 4 + 8 * 4 + 4 + 4 + 8 + 4 + 4
    assert C.sizeof(L.Tables) == 12 * 8
    assert C.sizeof(L.Plan) == 12 * 8
    assert C.sizeof(L.StepIO) == 24 * 8
    assert C.sizeof(L.Adam) == 32


def test_capacity_invariants_and_host_side_argument_checks():
    """Host-only entry points: capacities grow with the batch, the cut-row list can hold one row per
    backward-tile boundary plus the hot rows, scratch covers the block partials; bad arguments are
    reported through the return code + vfmb_last_error (nothing is launched)."""
    lib = L.lib()
    prev = 0
    for B, F, R in ((1, 1, 1), (512, 2, 100), (65536, 2, 165237), (65536, 8, 1_000_000), (100_000, 8, 100_000_000)):
        cap = L.PlanCapacity()
        assert lib.vfmb_plan_capacity(B, F, R, C.byref(cap)) == 0
        n = B * F
        assert cap.u_cap == min(n, R) and cap.tile == 32 and cap.n_tiles == (n + 31) // 32
        assert cap.cut_rows_cap >= cap.n_tiles + cap.n_tiles // 30
        assert cap.workspace_bytes >= 4 * 4 * n and cap.workspace_bytes >= prev          # 4 key/value arrays
        prev = cap.workspace_bytes
        cfg = L.Config()
        cfg.B, cfg.F, cfg.d, cfg.R, cfg.S = B, F, 64, R, 1
        assert lib.vfmb_partials_doubles(C.byref(cfg)) >= 2048 * 16                      # block partials of 2048 blocks
    assert lib.vfmb_plan_capacity(1 << 28, 8, 10, C.byref(L.PlanCapacity())) != 0         # B*F >= 2^30
    assert lib.vfmb_set_grid_reserve(1) == 0 and lib.vfmb_set_grid_reserve(0) == 0
    assert lib.vfmb_set_grid_reserve(9) != 0 and b"vfmb_set_grid_reserve" in lib.vfmb_last_error()
    assert lib.vfmb_shard_bucket_workspace(131072) == 128 * 8 * 4
    # null arguments never reach a launch
    assert lib.vfmb_plan_build(None, None, None, None, None, 0, None) != 0
    assert lib.vfmb_sampled_step(None, None, None, None, None, None) != 0
    assert lib.vfmb_shard_bucket(None, 1, 2, 4, None, None, None, None, None, 0, None) != 0
    cfg = L.Config()
    cfg.B, cfg.F, cfg.d, cfg.R, cfg.S, cfg.n_classes = 8, 2, 4, 10, 9, 2                  # S out of range
    assert lib.vfmb_sampled_forward(C.byref(cfg), C.byref(L.Tables()), C.byref(L.Plan()), C.byref(L.StepIO()), None) != 0
    assert b"variational samples" in lib.vfmb_last_error()


def test_launch_knobs_accept_known_keys_and_reject_the_rest():
    """vfmb_set_tuning is host-side state only: every documented key round-trips, bad keys / values are
    refused with an error text (no GPU needed)."""
    lib = L.lib()
    good = {"grid_reserve": (1, 0), "adam_reserve": (0, 1), "adam_pipe": (0, 1), "l2_keep": (0, 7, 1),
            "gather_dyn": (0, 1), "gather_fence": (0, 1), "gather_wide": (0, 1), "gather_keep": (1, 32),
            "stage_wide": (1, 0), "score_wide": (1, 0, -1), "stage_chunk": (32, 16), "score_chunk": (16, 32),
            "pdl": (1, 0), "prefetch_mv": (3, 0), "fuse_score": (1, 0)}
    for key, values in good.items():
        for v in values:                                       # the last value of every tuple is the default
            assert lib.vfmb_set_tuning(key.encode(), v) == 0, (key, v, lib.vfmb_last_error())
    for key, v in (("no_such_knob", 1), ("l2_keep", 8), ("stage_chunk", 24), ("gather_keep", 0), ("prefetch_mv", 99)):
        assert lib.vfmb_set_tuning(key.encode(), v) != 0
        assert key.split("_")[0].encode() in lib.vfmb_last_error() or b"unknown" in lib.vfmb_last_error()
    assert lib.vfmb_set_tuning(None, 0) != 0
