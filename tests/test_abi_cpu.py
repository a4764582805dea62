"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol the header
declares, and rejects bad arguments before touching the device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as ge
    ge.build()
    import adaptigraph_b200._lib as lib
    return lib


def test_header_symbols_are_exported(L):
    hdr = open(os.path.join(ROOT, "include", "adaptigraph_b200.h")).read()
    declared = re.findall(r"AGX_API\s+[\w\s\*]+?\b(agx_\w+)\s*\(", hdr)
    assert len(declared) >= 13
    assert sorted(declared) == sorted(L.EXPORTS)
    for name in declared:
        assert getattr(L.lib, name) is not None
    assert L.lib.agx_version() == 100


def test_struct_layouts_match_header(L):
    assert C.sizeof(L.AgxModelDims) == 24
    assert C.sizeof(L.AgxWeights) == 2 * 11 * 8
    assert C.sizeof(L.AgxGraphIn) == 16 + 8 * 8 + 8
    assert L.AgxRolloutIn.E_cap.offset % 8 == 0


def test_workspace_sizes_are_host_only_and_monotone(L):
    dims = L.AgxModelDims(150, 4, 2, 1, 3, 3)
    a = L.lib.agx_forward_workspace_bytes(C.byref(dims), 4, 100, 2000)
    b = L.lib.agx_forward_workspace_bytes(C.byref(dims), 8, 100, 4000)
    assert 0 < a < b
    # 5 node buffers + nfeat + C_e at the padded stride
    assert a >= (5 * 400 * 160 + 400 * 16 + 2000 * 160) * 4
    assert L.lib.agx_graph_workspace_bytes(4, 100, 10) > 4 * 100 * 10 * 4
    assert L.lib.agx_graph_workspace_bytes(0, 100, 10) == 0
    assert L.lib.agx_rollout_workspace_bytes(C.byref(dims), 4, 100, 2000, 10) > a
    assert L.lib.agx_packed_weights_bytes(C.byref(dims)) >= 252903 * 4


def test_bad_arguments_return_error_codes_without_a_device(L):
    dims = L.AgxModelDims(150, 4, 2, 1, 3, 3)
    g = L.AgxGraphIn()
    rc = L.lib.agx_forward(C.byref(dims), None, C.byref(g), None, 0, None, 0, None, 0, None)
    assert rc == L.AGX_ERR_ARG and b"graph" in L.lib.agx_last_error()
    bad = L.AgxModelDims(150, 3, 2, 1, 3, 3)   # n_his the kernels are not specialised for
    rc = L.lib.agx_forward(C.byref(bad), None, C.byref(g), None, 0, None, 0, None, 0, None)
    assert rc == L.AGX_ERR_ARG and b"n_his" in L.lib.agx_last_error()
    rc = L.lib.agx_graph_build(None, None, None, None, 1, 10, 5, 0, 0, None, None, None, 0, None, None, None, 0, None)
    assert rc == L.AGX_ERR_ARG
    with pytest.raises(ValueError):
        L.check(rc, "agx_graph_build")


def test_ops_reject_cpu_tensors(L):
    import torch
    from adaptigraph_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.graph_build(torch.zeros(1, 4, 3), torch.ones(1, 4, dtype=torch.bool), torch.zeros(1, 4, dtype=torch.bool),
                        torch.ones(1), 2, False, 0, 16)


def test_module_keeps_reference_state_dict_and_init(L):
    import torch
    import agx_helpers as H
    from adaptigraph_b200 import DynamicsPredictor, synthetic as syn
    torch.manual_seed(0)
    m = DynamicsPredictor(*syn.configs("cloth", 3), "cpu")
    gw = H.golden_weights()
    sd = m.state_dict()
    assert list(sd.keys()) == list(gw.keys())
    assert all(torch.equal(sd[k], gw[k]) for k in gw)
    m.load_state_dict(gw)
    assert sum(p.numel() for p in m.parameters()) == 252903
    mc, matc, dc = syn.configs("rope", 3)
    mc["state_dim"] = 3
    with pytest.raises(NotImplementedError):
        DynamicsPredictor(mc, matc, dc, "cpu")


def test_widened_entry_points_validate_arguments_without_a_device(L):
    """agx_fps / agx_chamfer / agx_adam_step / the training entries return AGX_ERR_ARG before any launch."""
    lib = L.lib
    assert lib.agx_fps(None, None, 1, 10, 4, None, -1.0, None, None, None) == L.AGX_ERR_ARG
    buf = (C.c_float * 64)()
    ibuf = (C.c_int32 * 16)()
    p, ip = C.cast(buf, C.c_void_p), C.cast(ibuf, C.c_void_p)
    assert lib.agx_fps(p, None, 0, 10, 4, ip, -1.0, ip, ip, None) == L.AGX_ERR_ARG           # B = 0
    assert lib.agx_fps(p, None, 1, 300000, 4, ip, -1.0, ip, ip, None) == L.AGX_ERR_ARG      # beyond the cluster staging limit
    assert b"staging limit" in lib.agx_last_error()
    assert lib.agx_chamfer(None, p, 1, 4, 4, 0, p, None) == L.AGX_ERR_ARG
    assert lib.agx_chamfer(p, p, 1, 20000, 4, 0, p, None) == L.AGX_ERR_ARG                   # N + M beyond shared memory
    assert lib.agx_adam_step(None, p, p, p, 16, 1e-3, 0.9, 0.999, 1e-8, 1.0, ip, None) == L.AGX_ERR_ARG
    assert lib.agx_adam_step(p, p, p, p, 0, 1e-3, 0.9, 0.999, 1e-8, 1.0, ip, None) == L.AGX_ERR_ARG
    off = C.c_void_p(p.value + 4)                                                            # misaligned bucket
    assert lib.agx_adam_step(off, p, p, p, 8, 1e-3, 0.9, 0.999, 1e-8, 1.0, ip, None) == L.AGX_ERR_ARG
    dims = L.AgxModelDims(150, 4, 2, 1, 3, 3)
    assert lib.agx_train_saved_bytes(C.byref(dims), 2, 50, 400) > 0
    assert lib.agx_train_scratch_bytes(C.byref(dims), 2, 50, 400) > 0
