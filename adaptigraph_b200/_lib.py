"""ctypes binding of libadaptigraph_b200.so (include/adaptigraph_b200.h).

There is no fallback: if the shared object is missing or a symbol does not
resolve, importing this module raises.  Build it with `python -m adaptigraph_b200.build`
(or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGX_LIB", os.path.join(HERE, "libadaptigraph_b200.so"))

AGX_FP = 160
AGX_MAX_TOPK = 32
AGX_SEM_BATCH, AGX_SEM_SINGLE = 0, 1
AGX_Y_MIN, AGX_Y_MASKED_MEAN = 0, 1
AGX_PREC_FP32, AGX_PREC_TC_F16X3, AGX_PREC_TC_MIXED = 0, 1, 2
AGX_ERROR_CHAMFER, AGX_ERROR_BOX = 0, 1
AGX_PENALTY_ROPE, AGX_PENALTY_CLOTH, AGX_PENALTY_GRANULAR = 0, 1, 2
AGX_NUM_LAYERS = 11
AGX_ERR_ARG, AGX_ERR_CAPACITY, AGX_ERR_CUDA = -1, -2, -3

# reference state_dict prefixes in AgxWeights index order (model.py:103-122)
LAYER_NAMES = [
    "particle_encoder.model.0", "particle_encoder.model.2", "particle_encoder.model.4",
    "relation_encoder.model.0", "relation_encoder.model.2", "relation_encoder.model.4",
    "particle_propagator.linear", "relation_propagator.linear",
    "non_rigid_predictor.linear_0", "non_rigid_predictor.linear_1", "non_rigid_predictor.linear_2",
]

# every extern "C" symbol the header declares; tests check the library exports all of them
EXPORTS = [
    "agx_version", "agx_last_error", "agx_launch_count",
    "agx_packed_weights_bytes", "agx_pack_weights",
    "agx_graph_workspace_bytes", "agx_graph_build", "agx_onehot_to_ids", "agx_edges_to_onehot", "agx_fps", "agx_fps_radii", "agx_chamfer",
    "agx_running_cost_workspace_bytes", "agx_running_cost",
    "agx_forward_workspace_bytes", "agx_forward",
    "agx_rollout_workspace_bytes", "agx_rollout",
    "agx_profile_enable", "agx_profile_read", "agx_kind_name",
    "agx_train_saved_bytes", "agx_train_scratch_bytes", "agx_train_saved_offsets", "agx_forward_train", "agx_backward", "agx_adam_step",
]
AGX_NUM_KINDS = 12


class AgxModelDims(C.Structure):
    _fields_ = [("F", C.c_int32), ("n_his", C.c_int32), ("d_attr", C.c_int32), ("d_phys", C.c_int32),
                ("d_act", C.c_int32), ("pstep", C.c_int32)]


class AgxWeights(C.Structure):
    _fields_ = [("weight", C.c_void_p * AGX_NUM_LAYERS), ("bias", C.c_void_p * AGX_NUM_LAYERS)]


class AgxWeightGrads(C.Structure):
    _fields_ = [("weight", C.c_void_p * AGX_NUM_LAYERS), ("bias", C.c_void_p * AGX_NUM_LAYERS)]


class AgxGraphIn(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("n_p", C.c_int32),
                ("state", C.c_void_p), ("attrs", C.c_void_p), ("action", C.c_void_p),
                ("p_instance", C.c_void_p), ("physics", C.c_void_p),
                ("row_ptr", C.c_void_p), ("send", C.c_void_p), ("recv", C.c_void_p), ("E_cap", C.c_int64)]


class AgxRolloutIn(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("n_p", C.c_int32),
                ("state", C.c_void_p), ("attrs", C.c_void_p), ("action", C.c_void_p),
                ("p_instance", C.c_void_p), ("physics", C.c_void_p),
                ("mask", C.c_void_p), ("tool_mask", C.c_void_p), ("thr2", C.c_void_p),
                ("topk", C.c_int32), ("connect_tools_all", C.c_int32), ("n_steps", C.c_int32),
                ("y_mode", C.c_int32), ("gripper_raise", C.c_float), ("E_cap", C.c_int64)]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library is the product and there is no fallback. "
            "Build it with `python -m adaptigraph_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    P = C.POINTER
    sigs = {
        "agx_version": (C.c_int, []),
        "agx_last_error": (C.c_char_p, []),
        "agx_launch_count": (i64, []),
        "agx_packed_weights_bytes": (sz, [P(AgxModelDims)]),
        "agx_pack_weights": (C.c_int, [P(AgxModelDims), P(AgxWeights), vp, vp]),
        "agx_graph_workspace_bytes": (sz, [i32, i32, i32]),
        "agx_graph_build": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, i64, vp, vp, vp, sz, vp]),
        "agx_onehot_to_ids": (C.c_int, [vp, i32, i32, i32, vp, vp]),
        "agx_edges_to_onehot": (C.c_int, [vp, vp, i32, i32, i32, vp, vp, vp]),
        "agx_fps": (C.c_int, [vp, vp, i32, i32, i32, vp, C.c_double, vp, vp, vp]),
        "agx_fps_radii": (C.c_int, [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]),
        "agx_chamfer": (C.c_int, [vp, vp, i32, i32, i32, i32, vp, vp]),
        "agx_running_cost_workspace_bytes": (sz, [i32, i32]),
        "agx_running_cost": (C.c_int, [vp, vp, i32, vp, vp, i32, vp, i32, i32, C.c_float, i32, i32, i32, vp, sz, vp, vp]),
        "agx_forward_workspace_bytes": (sz, [P(AgxModelDims), i32, i32, i64]),
        "agx_forward": (C.c_int, [P(AgxModelDims), vp, P(AgxGraphIn), vp, i64, vp, i32, vp, sz, vp]),
        "agx_rollout_workspace_bytes": (sz, [P(AgxModelDims), i32, i32, i64, i32]),
        "agx_rollout": (C.c_int, [P(AgxModelDims), vp, P(AgxRolloutIn), vp, vp, vp, i32, vp, sz, vp]),
        "agx_train_saved_bytes": (sz, [P(AgxModelDims), i32, i32, i64]),
        "agx_train_scratch_bytes": (sz, [P(AgxModelDims), i32, i32, i64]),
        "agx_train_saved_offsets": (C.c_int, [P(AgxModelDims), i32, i32, i64, P(i64), i32]),
        "agx_forward_train": (C.c_int, [P(AgxModelDims), vp, P(AgxGraphIn), vp, i64, vp, vp, sz, vp]),
        "agx_backward": (C.c_int, [P(AgxModelDims), vp, P(AgxGraphIn), vp, vp, vp, vp, vp, vp, P(AgxWeightGrads), vp, vp, sz, vp]),
        "agx_adam_step": (C.c_int, [vp, vp, vp, vp, i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_float, vp, vp]),
        "agx_profile_enable": (C.c_int, [i32]),
        "agx_profile_read": (C.c_int, [P(C.c_double), P(i64)]),
        "agx_kind_name": (C.c_char_p, [i32]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()


def check(rc: int, what: str) -> None:
    """Maps the ABI's error codes onto the exceptions the reference's Python raises."""
    if rc == 0:
        return
    msg = lib.agx_last_error().decode(errors="replace")
    if rc == AGX_ERR_ARG:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg} (code {rc})")
