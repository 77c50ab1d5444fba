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
    assert C.sizeof(L.Tables) == 11 * 8
    assert C.sizeof(L.Plan) == 12 * 8
    assert C.sizeof(L.StepIO) == 24 * 8
    assert C.sizeof(L.Adam) == 32
