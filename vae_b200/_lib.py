"""ctypes binding of ``libvfm_b200.so`` (the C ABI declared in ``include/vfm_b200.h``).

There is no fallback: if the shared library is missing or a call fails, a
``RuntimeError`` is raised -- the product never computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
# experiment hook: VFMB_VARIANT=name + VFMB_NVCC_EXTRA="-DX=1 ..." builds / loads libvfm_b200_name.so
_VARIANT = os.environ.get("VFMB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, f"libvfm_b200{'_' + _VARIANT if _VARIANT else ''}.so")
SOURCES = ["api.cu", "plan.cu", "sampled.cu", "sampled_adam.cu", "closed.cu", "predict.cu", "dp.cu", "shard.cu"]
HEADERS = ["common.cuh", "internal.h", "step_common.cuh", "sampled_common.cuh", os.path.join(_ROOT, "include", "vfm_b200.h")]
# -prec-div/-prec-sqrt=false: MUFU-based division and square root (<= 2 ulp) instead of the IEEE
# slow paths, which made the Adam epilogue instruction-bound; denormals and expf/logf stay precise
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-prec-div=false", "-prec-sqrt=false", "-Xcompiler", "-fPIC", "-shared", "-w"] + \
             os.environ.get("VFMB_NVCC_EXTRA", "").split()

MAX_FIELDS = 8
GAUSSIAN, BERNOULLI = 0, 1
LINK_ABS, LINK_SOFTPLUS = 0, 1
INTER_PROD, INTER_PAIRWISE = 0, 1
ADAM_TOUCHED, GRAD_ONLY = 0, 1
STATS = 32
ST_RESID_S, ST_W0_S, MAX_SAMPLES = 8, 16, 8
DP_TAIL, DP_T_NLL, DP_T_RESID, DP_T_SQERR, DP_T_KLROWS, DP_T_OVERFLOW = 16, 8, 9, 10, 11, 12
ST_LOSS, ST_NLL_MEAN, ST_KL, ST_SUM_RESID, ST_SUM_SQERR, ST_KL_ROWS, ST_W0, ST_U = range(8)
S_ALPHA, S_GB_MEAN, S_GB_SCALE, S_COUNT = 0, 1, 2, 4
C_ALPHA, C_GB_MEAN, C_GB_SCALE, C_GB_PRIOR_MEAN, C_GB_PRIOR_SCALE = 0, 1, 2, 3, 4

_f32p, _i32p, _i64p, _f64p = (C.c_void_p,) * 4     # raw device addresses (tensor.data_ptr())


class Config(C.Structure):
    _fields_ = [("B", C.c_int32), ("F", C.c_int32), ("d", C.c_int32), ("R", C.c_int32),
                ("S", C.c_int32), ("likelihood", C.c_int32), ("link", C.c_int32),
                ("n_classes", C.c_int32), ("class_bound", C.c_int32 * MAX_FIELDS),
                ("class_size", C.c_float * MAX_FIELDS), ("n_train", C.c_float),
                ("row_stride", C.c_int32), ("seed", C.c_uint64), ("row_offset", C.c_int32),
                ("interaction", C.c_int32)]


class Tables(C.Structure):
    _fields_ = [("bias", _f32p), ("bias_m", _f32p), ("bias_v", _f32p),
                ("entity", _f32p), ("entity_m", _f32p), ("entity_v", _f32p),
                ("train_counts", _f32p),
                ("scalars", _f32p), ("scalars_m", _f32p), ("scalars_v", _f32p),
                ("adam_step", _i32p), ("noise_step", _i32p)]


class Adam(C.Structure):
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double)]


class Plan(C.Structure):
    _fields_ = [("uniq", _i32p), ("inverse", _i32p), ("seg_off", _i32p), ("occ", _i32p),
                ("pos_of", _i32p), ("pos_rank", _i32p), ("partner", _i32p), ("urec", _i32p),
                ("class_off", _i32p), ("z", _f32p), ("meta", _i32p), ("hot", _i32p)]


class PlanCapacity(C.Structure):
    _fields_ = [("u_cap", C.c_int64), ("n_tiles", C.c_int64), ("workspace_bytes", C.c_int64),
                ("tile", C.c_int32), ("cut_rows_cap", C.c_int32)]


class StepIO(C.Structure):
    _fields_ = [("y", _f32p), ("eps_global", _f32p), ("eps_bias", _f32p), ("eps_entity", _f32p),
                ("vs", _f32p), ("ws", _f32p), ("es", _f32p), ("ebs", _f32p), ("cq", _f32p),
                ("grow", _f32p), ("gws", _f32p), ("msg", _f32p), ("pred", _f32p), ("mean", _f32p),
                ("resid", _f32p), ("rsorted", _f32p), ("partials", _f64p), ("counters", _i32p), ("stats", _f32p),
                ("kl_bias_out", _f32p), ("kl_entity_out", _f32p),
                ("grad_bias", _f32p), ("grad_entity", _f32p), ("grad_scalars", _f32p)]


# every symbol include/vfm_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "vfmb_version": (C.c_int, []),
    "vfmb_profile_events": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vfmb_set_grid_reserve": (C.c_int, [C.c_int]),
    "vfmb_set_tuning": (C.c_int, [C.c_char_p, C.c_int]),
    "vfmb_launch_count": (C.c_int64, []),
    "vfmb_last_error": (C.c_char_p, []),
    "vfmb_closed_off_bias_prior_mean": (C.c_int32, [C.c_int32] * 3),
    "vfmb_closed_off_bias_prior_scale": (C.c_int32, [C.c_int32] * 3),
    "vfmb_closed_off_entity_prior_mean": (C.c_int32, [C.c_int32] * 3),
    "vfmb_closed_off_entity_prior_scale": (C.c_int32, [C.c_int32] * 3),
    "vfmb_closed_scalar_count": (C.c_int32, [C.c_int32] * 2),
    "vfmb_plan_capacity": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P(PlanCapacity)]),
    "vfmb_plan_build": (C.c_int, [_P(Config), C.c_void_p, C.c_void_p, _P(Plan), C.c_void_p,
                                  C.c_size_t, C.c_void_p]),
    "vfmb_partials_doubles": (C.c_int64, [_P(Config)]),
    "vfmb_sampled_forward": (C.c_int, [_P(Config), _P(Tables), _P(Plan), _P(StepIO), C.c_void_p]),
    "vfmb_sampled_backward": (C.c_int, [_P(Config), _P(Tables), _P(Plan), _P(StepIO), _P(Adam),
                                        C.c_int32, C.c_float, C.c_void_p]),
    "vfmb_sampled_stage": (C.c_int, [_P(Config), _P(Tables), _P(Plan), _P(StepIO), C.c_void_p]),
    "vfmb_sampled_score": (C.c_int, [_P(Config), _P(Tables), _P(Plan), _P(StepIO), C.c_void_p]),
    "vfmb_sampled_gather": (C.c_int, [_P(Config), _P(Plan), _P(StepIO), C.c_void_p, C.c_int32, C.c_void_p]),
    "vfmb_sampled_adam_rows": (C.c_int, [_P(Config), _P(Tables), _P(Plan), _P(StepIO), _P(Adam), C.c_int32,
                                         C.c_float, C.c_void_p]),
    "vfmb_dp_final": (C.c_int, [_P(Config), _P(Tables), C.c_void_p, C.c_void_p, _P(Adam), C.c_void_p, C.c_void_p]),
    "vfmb_sampled_step": (C.c_int, [_P(Config), _P(Tables), _P(Plan), _P(StepIO), _P(Adam), C.c_void_p]),
    "vfmb_dp_scatter_counts": (C.c_int, [_P(Config), _P(Plan), _P(StepIO), C.c_void_p, C.c_void_p, C.c_void_p]),
    "vfmb_dp_apply_sampled": (C.c_int, [_P(Config), _P(Tables), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, _P(Adam), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "vfmb_shard_bucket_workspace": (C.c_int64, [C.c_int32]),
    "vfmb_shard_bucket": (C.c_int, [_P(Plan), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "vfmb_shard_put_small": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "vfmb_shard_sum_small": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "vfmb_shard_owner_ids": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vfmb_shard_owner_pack": (C.c_int, [_P(Plan), C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "vfmb_shard_unpack_rows": (C.c_int, [_P(Plan), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "vfmb_shard_pack_grads": (C.c_int, [_P(Plan), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "vfmb_shard_unpack_grads": (C.c_int, [_P(Plan), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "vfmb_shard_route": (C.c_int, [_P(Plan), C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vfmb_shard_stage_put": (C.c_int, [_P(Config), _P(Tables), _P(Plan), _P(StepIO), C.c_int32, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "vfmb_shard_score": (C.c_int, [_P(Config), _P(Tables), _P(Plan), _P(StepIO), C.c_void_p, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_void_p]),
    "vfmb_shard_gather_put": (C.c_int, [_P(Config), _P(Plan), _P(StepIO), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "vfmb_shard_owner_update": (C.c_int, [_P(Config), _P(Tables), _P(Plan), _P(StepIO), _P(Adam), C.c_void_p, C.c_int32,
                                          C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "vfmb_adam_dense": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                  _P(Adam), C.c_void_p, C.c_void_p]),
    "vfmb_adam_step_advance": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vfmb_closed_forward": (C.c_int, [_P(Config), _P(Tables), _P(Plan), _P(StepIO), C.c_void_p]),
    "vfmb_closed_backward": (C.c_int, [_P(Config), _P(Tables), _P(Plan), _P(StepIO), _P(Adam),
                                       C.c_int32, C.c_void_p]),
    "vfmb_closed_backward_weighted": (C.c_int, [_P(Config), _P(Tables), _P(Plan), _P(StepIO), _P(Adam),
                                                C.c_int32, C.c_void_p, C.c_float, C.c_float, C.c_void_p]),
    "vfmb_predict_mean": (C.c_int, [_P(Config), C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "vfmb_predict_sampled": (C.c_int, [_P(Config), _P(Tables), C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "vfmb_philox_normals": (C.c_int, [_P(Config), C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def _source_digest() -> str:
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = [os.path.join(CSRC, s) for s in SOURCES] + \
           [h_ if os.path.isabs(h_) else os.path.join(CSRC, h_) for h_ in HEADERS]
    for p in deps:
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _stale() -> bool:
    """True when the .so is missing or was built from different sources (content
    hash, not mtimes: the snapshot sent to the GPU box does not keep mtimes)."""
    if not os.path.isfile(LIB_PATH) or not os.path.isfile(LIB_PATH + ".sha256"):
        return True
    with open(LIB_PATH + ".sha256") as fh:
        return fh.read().strip() != _source_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into ``vae_b200/libvfm_b200.so``
    (in-tree, so that it travels to the GPU box).  No-op when up to date."""
    if not force and not _stale():
        return LIB_PATH
    # one builder at a time: the ranks of a torchrun job all get here when the library is stale (it happened:
    # eight concurrent builds wrote the same objects and the first ranks loaded a half-linked library)
    import fcntl
    os.makedirs(os.path.join(_ROOT, "build"), exist_ok=True)
    with open(os.path.join(_ROOT, "build", ".lock" + ("_" + _VARIANT if _VARIANT else "")), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():                   # another process built it while this one waited
                return LIB_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    # one object per translation unit, compiled concurrently (the template instantiations make
    # each .cu ~20-60 s of ptxas time), then one link step
    obj_dir = os.path.join(_ROOT, "build", "obj" + ("_" + _VARIANT if _VARIANT else ""))
    os.makedirs(obj_dir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"]
    jobs = []
    for s in SOURCES:
        obj = os.path.join(obj_dir, s.replace(".cu", ".o"))
        cmd = [nvcc, *cflags, "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        jobs.append((obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    errs = []
    for obj, p in jobs:
        out, _ = p.communicate()
        if p.returncode != 0:
            errs.append(out)
    if errs:
        raise RuntimeError("nvcc failed:\n" + "\n".join(errs))
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, *NVCC_FLAGS, "-o", tmp] + [obj for obj, _ in jobs]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)                                # readers never see a partly written library
    with open(LIB_PATH + ".sha256", "w") as fh:
        fh.write(_source_digest())
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library.  Builds it when sources are newer and nvcc exists;
    raises if it cannot be loaded (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if _stale():
        try:
            build()
        except FileNotFoundError as exc:          # no nvcc on this machine
            if not os.path.isfile(LIB_PATH):
                raise RuntimeError(f"{LIB_PATH} is missing and nvcc is not available") from exc
    try:
        handle = C.CDLL(LIB_PATH)
    except OSError as exc:
        raise RuntimeError(f"cannot load {LIB_PATH}: {exc} -- the CUDA extension is required, "
                           "there is no CPU fallback") from exc
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(handle, name)
        fn.restype, fn.argtypes = res, args
    _lib = handle
    return handle


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().vfmb_last_error().decode(errors="replace")
        raise RuntimeError(f"{what or 'vfm_b200'} failed (code {rc}): {msg}")


def ptr(t) -> Optional[int]:
    """Device address of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
