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
