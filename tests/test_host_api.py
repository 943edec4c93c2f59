"""CPU tests of the host layer: the reference-facing API mirror, topology folding, state_dict schema,
the epoch state machine and the C-ABI library's symbols (no compute: this box has no GPU)."""
import ctypes
import os
import math
import re

import numpy as np
import pytest
import torch

from conftest import FLOW_CASES, ROOT
from oracle import flow as oflow
from oracle import nis as onis

from nf_b200 import _cabi
from nf_b200.flowspec import FlowSequential
from nf_b200.normalizing_flows.manager import AffineManager, EpochState, PWLinManager, PWQuadManager, get_bin
from nf_b200.PhaseSpace.flat_phase_space_generator import FlatInvertiblePhasespace, PhaseSpaceGeneratorError


def make_manager(meta):
    torch.manual_seed(0)
    if meta["kind"] == "quad":
        NF = PWQuadManager(n_flow=meta["n_flow"])
        NF.create_model(meta["n_cells"], meta["n_bins"], meta["NN"])
    elif meta["kind"] == "affine":
        NF = AffineManager(n_flow=meta["n_flow"])
        NF.create_model(meta["n_pass_through"], meta["n_cells"], meta["NN"], meta["roll_step"])
    else:
        NF = PWLinManager(n_flow=meta["n_flow"])
        NF.create_model(meta["n_pass_through"], meta["n_cells"], meta["n_bins"], meta["NN"], meta["roll_step"])
    return NF


@pytest.mark.parametrize("case", FLOW_CASES)
def test_create_model_reproduces_reference_topology_and_schema(golden, case):
    g = golden("flow_" + case)
    NF = make_manager(g.meta)
    model = NF._model
    assert isinstance(model, torch.nn.Sequential) and isinstance(model, FlowSequential)
    assert [n for n, _ in model.named_children()] == g.meta["children"]
    ref_sd = g.state_dict()
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref_sd[k].shape), k
    model.load_state_dict(ref_sd)            # reference checkpoints load unchanged (float64 -> float32)
    assert NF.best_model is NF.model
    # (the affine cell's hidden-layer biases stay outside the kernels' parameter block: flowspec._hidden_biases)
    assert sum(p.numel() for p in model.parameters()) == \
        model.spec().n_params + sum(b.numel() for b in model.spec().hidden_biases)


@pytest.mark.parametrize("case", FLOW_CASES)
def test_flowspec_tables_equal_oracle_compile(golden, case):
    g = golden("flow_" + case)
    meta = g.meta
    spec = make_manager(meta)._model.spec()
    if meta["kind"] == "quad":
        layers = oflow.pwquad_layers(meta["n_flow"], meta["n_cells"])
    else:
        layers = oflow.pwlin_layers(meta["n_flow"], meta["n_pass_through"], meta["n_cells"], meta["roll_step"])
    cells, out_perm = oflow.compile_layers(layers, meta["n_flow"])
    assert spec.out_perm == out_perm
    assert [(n, f, t) for n, _, f, t in spec.cells] == [(c["name"], c["feed_idx"], c["trafo_idx"]) for c in cells]
    d = spec.desc
    assert d.n_cells == len(cells) and d.n_bins == meta["n_bins"] and d.depth == len(meta["NN"])
    lib = _cabi.load()
    off = 0
    for i in range(d.n_cells):
        assert d.cells[i].param_off == off
        off += lib.nis_flow_cell_param_count(ctypes.byref(d), i)
    assert off == spec.n_params
    assert lib.nis_flow_workspace_bytes(ctypes.byref(d), 1 << 16) > 0
    assert lib.nis_flow_bn_saved_count(ctypes.byref(d)) > 0


def test_parameters_are_views_of_one_arena_and_survive_optimizer_and_deepcopy():
    import copy
    NF = PWQuadManager(n_flow=3)
    NF.create_model(3, 4, [8, 8])
    spec = NF._model.spec()
    flat = spec.param_arena.get(torch.device("cpu"))
    assert flat.numel() == spec.n_params
    p0 = spec.params[3]
    assert p0.data_ptr() == flat.data_ptr() + 4 * spec.param_arena.offsets[3]
    opt = torch.optim.Adamax(NF._model.parameters(), lr=1e-2)
    for p in NF._model.parameters():
        p.grad = torch.ones_like(p)
    before = flat.clone()
    opt.step()
    assert spec.param_arena.get(torch.device("cpu")) is flat and not torch.equal(flat, before)
    twin = copy.deepcopy(NF._model)
    f2 = twin.spec().param_arena.get(torch.device("cpu"))
    assert torch.equal(f2, flat) and f2.data_ptr() != flat.data_ptr()
    NF._model.double()                                   # user-side dtype change: slow path, still consistent
    f3 = spec.param_arena.get(torch.device("cpu"))
    assert f3.dtype == torch.float32 and torch.allclose(f3, flat)


def test_model_property_and_errors():
    NF = PWQuadManager(n_flow=2)
    with pytest.raises(AttributeError, match="No model was instantiated"):
        NF.model
    assert NF.integrate(lambda x: x[:, 0], 2, 10) == (0, 0)          # "No model has been trained"
    assert get_bin(5, 4) == [0, 1, 0, 1] and get_bin(7) == [1, 1, 1]
    NF.create_model(2, 4, [3] * 3)
    if not torch.cuda.is_available():
        with pytest.raises(_cabi.NisBackendError):                    # no silent CPU fallback
            NF._model(NF.format_input(torch.rand(4, 2)))
    out = NF.format_input(torch.rand(4, 2))
    assert out.shape == (4, 3) and out.dtype == torch.float64 and torch.all(out[:, -1] == 1)


def test_phase_space_host_surface():
    with pytest.raises(PhaseSpaceGeneratorError):
        FlatInvertiblePhasespace([100.0], [100.0] * 3)
    with pytest.raises(PhaseSpaceGeneratorError):
        FlatInvertiblePhasespace([1.0, 1.0, 1.0], [100.0] * 3)
    ps = FlatInvertiblePhasespace([100.0] * 2, [100.0] * 4, pdf=None, pdf_active=False)
    assert ps.nDimPhaseSpace() == 8
    assert math.isclose(FlatInvertiblePhasespace.get_flatWeights(1000.0, 4) / 2e6, 0.06648282151394422, rel_tol=1e-13)
    assert FlatInvertiblePhasespace.get_flatWeights(1000.0, 1) == 1.0
    with pytest.raises(PhaseSpaceGeneratorError):
        ps.generateKinematics_batch(1000.0, torch.full((4, 8), float("nan"), dtype=torch.double))
    if not torch.cuda.is_available():
        with pytest.raises(_cabi.NisBackendError):
            ps.generateKinematics_batch(1000.0, torch.rand(4, 8, dtype=torch.double))


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_epoch_state_equals_oracle_state_machine_on_reference_traces(golden, case):
    g = golden("train_" + case)
    m = g.meta
    losses = [float(x) for x in g["losses"]]
    # extend the trace with synthetic tails to drive every branch
    rng = np.random.default_rng(3)
    tails = [losses, losses + list(losses[-1] * (1 + 0.2 * rng.random(40))), [1.0] * 30, list(np.linspace(1, 0.1, 400))]
    for seq in tails:
        ours = EpochState(float(g["int_loss"]), m["preburn_time"], m["kill_counter"], 1e-2)
        ref = onis.EpochStateMachine(float(g["int_loss"]), preburn_time=m["preburn_time"], kill_counter=m["kill_counter"])
        best_epoch = 0
        for i, loss in enumerate(seq):
            if ours.improved(loss, True):
                best_epoch = i
            stop_o = ours.advance(i, loss)
            stop_r = ref.step(i, loss)
            assert stop_o == stop_r and ours.preburner == ref.preburner and ours.counter == ref.counter, (case, i)
            if stop_o:
                break
        assert best_epoch == ref.best_epoch and ours.best_loss == ref.best_loss
    assert best_epoch >= 0


def test_cabi_library_exports_every_declared_symbol():
    lib = _cabi.load()
    header = open(ROOT + "/include/nis_b200.h").read()
    declared = set(re.findall(r"\b(nis_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_cabi.EXPORTS)
    assert lib.nis_sizeof_flow_desc() == ctypes.sizeof(_cabi.NisFlowDesc)
    assert lib.nis_sizeof_rambo_desc() == ctypes.sizeof(_cabi.NisRamboDesc)
    assert lib.nis_strerror(-2).decode().startswith("workspace")
    bad = _cabi.NisFlowDesc()
    assert lib.nis_flow_workspace_bytes(ctypes.byref(bad), 10) == 0       # invalid descriptor rejected on the host
    assert lib.nis_flow_cell_param_count(ctypes.byref(bad), 0) == -1


def test_entry_points_validate_their_arguments_on_the_host():
    """Argument checks of the forward / backward entry points (plain and _cached) run before any CUDA call: a missing
    parameter pack is NIS_EINVAL (-1), a short workspace NIS_EWORKSPACE (-2), an empty batch NIS_OK -- no GPU needed."""
    from nf_b200.normalizing_flows.manager import PWQuadManager
    lib = _cabi.load()
    NF = PWQuadManager(n_flow=4)
    NF.create_model(4, 8, [16, 16])
    d = NF._model.spec().desc
    dp = ctypes.byref(d)
    buf = (ctypes.c_float * 64)()
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    need = lib.nis_flow_workspace_bytes(dp, 0)
    assert need > 0
    F32, TRAIN = 0, _cabi.BN_TRAIN
    # forward: params missing -> EINVAL; workspace too small -> EWORKSPACE; B == 0 -> OK
    for fn, extra in ((lib.nis_flow_forward, ()), (lib.nis_flow_forward_cached, (None,))):
        args = lambda params, wsb, B: (dp, params, None, ptr, F32, 4, ptr, F32, None, None, None) + extra + (TRAIN, ptr, wsb, B, None)
        assert fn(*args(None, need, 0)) == -1
        assert fn(*args(ptr, need - 1, 0)) == -2
        assert fn(*args(ptr, need, 0)) == 0
    # backward: saved states missing -> EINVAL; B == 0 -> OK
    for fn, extra in ((lib.nis_flow_backward, ()), (lib.nis_flow_backward_cached, (None,))):
        args = lambda saved, wsb, B: (dp, ptr, None, saved, ptr) + extra + (ptr, F32, ptr, None, TRAIN, ptr, wsb, B, None)
        assert fn(*args(None, need, 0)) == -1
        assert fn(*args(ptr, need - 1, 0)) == -2
        assert fn(*args(ptr, need, 0)) == 0


def test_activation_cache_count_is_host_code():
    """nis_flow_act_saved_count: cells * depth * width floats per point (tiles of 128) for the shape whose forward AND backward
    run the streamed-weights kernels (cfg5), 0 for the resident-weights shapes (cfg2, cfg4), for a batch below the
    tensor-core threshold and for a single hidden layer; FlowSpec.act_saved_count applies NIS_ACT_CACHE_MAX_BYTES."""
    from nf_b200.normalizing_flows.manager import PWLinManager, PWQuadManager
    lib = _cabi.load()
    NF = PWQuadManager(n_flow=16)
    NF.create_model(8, 64, [256] * 4)
    spec = NF._model.spec()
    B = (1 << 14) + 5
    tiles = (B + 127) // 128
    assert lib.nis_flow_act_saved_count(ctypes.byref(spec.desc), B) == 8 * 4 * tiles * 256 * 128
    assert lib.nis_flow_act_saved_count(ctypes.byref(spec.desc), 100) == 0
    assert spec.act_saved_count(lib, B) == 8 * 4 * tiles * 256 * 128
    os.environ["NIS_ACT_CACHE_MAX_BYTES"] = "1000"
    try:
        assert spec.act_saved_count(lib, B) == 0
    finally:
        del os.environ["NIS_ACT_CACHE_MAX_BYTES"]
    NF = PWLinManager(n_flow=8)
    NF.create_model(4, 6, 32, [64] * 3, 4)
    assert lib.nis_flow_act_saved_count(ctypes.byref(NF._model.spec().desc), 1 << 16) == 0
    NF = PWQuadManager(n_flow=8)
    NF.create_model(6, 32, [64] * 3)
    assert lib.nis_flow_act_saved_count(ctypes.byref(NF._model.spec().desc), 1 << 16) == 0
    NF = PWQuadManager(n_flow=6)
    NF.create_model(6, 12, [128])
    assert lib.nis_flow_act_saved_count(ctypes.byref(NF._model.spec().desc), 1 << 14) == 0


@pytest.mark.parametrize("kind,n_flow,args,scratch_per_point", [
    ("lin", 8, (4, 6, 32, [64] * 3, 4), 7 * 64 * 4),          # cfg2: z_1..z_3, dL/dlogits (2 tiles), two dL/dh buffers
    ("quad", 8, (6, 32, [64] * 3), 6 * 64 * 4),               # cfg4: streamed-weights backward, width 64
    ("quad", 16, (8, 64, [256] * 4), 7 * 256 * 4),            # cfg5: z_1..z_4, two dL/dh, dz + the logits-gradient tile
])
def test_workspace_covers_the_tensor_core_scratch(kind, n_flow, args, scratch_per_point):
    """nis_flow_workspace_bytes is host code: without a GPU it must still size the workspace for every kernel path
    the shape can take (forward activation buffers, tensor-core backward scratch), growing linearly with the batch."""
    from nf_b200.normalizing_flows.manager import PWLinManager, PWQuadManager
    NF = (PWLinManager if kind == "lin" else PWQuadManager)(n_flow=n_flow)
    NF.create_model(*args)
    d = NF._model.spec().desc
    lib = _cabi.load()
    b1, b2 = 1 << 14, 1 << 16
    w1 = lib.nis_flow_workspace_bytes(ctypes.byref(d), b1)
    w2 = lib.nis_flow_workspace_bytes(ctypes.byref(d), b2)
    assert w2 > w1 > 0
    per_point = (w2 - w1) / (b2 - b1)
    assert per_point >= scratch_per_point, per_point
    assert per_point <= 3 * scratch_per_point + 4096, per_point
    small = lib.nis_flow_workspace_bytes(ctypes.byref(d), 100)      # below the tensor-core threshold: generic kernels only
    assert 0 < small < w1
