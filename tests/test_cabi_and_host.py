"""CPU-side checks: the C-ABI library loads and exports every symbol the
header declares; the drop-in classes construct like the reference (same seed
-> same weights, same state_dict keys) and refuse to run without a GPU."""
import ctypes
import os

import pytest
import torch

from golden_io import unflatten
from helpers import load_golden
import recommendsystemproject_b200 as tt
from recommendsystemproject_b200 import _lib, ops, synth
from recommendsystemproject_b200._lib import TTError


def test_library_exports_every_declared_symbol():
    names = _lib.declared_symbols()
    assert len(names) >= 17
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tt_b200.h but not exported"
    assert set(names) == set(_lib._SIGNATURES), "ctypes signature table out of sync with the header"
    assert _lib.load().tt_abi_version() == 2


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    n = ctypes.c_size_t(0)
    assert lib.tt_emb_segment_grad_workspace(0, 8, ctypes.byref(n)) == -1
    assert b"bad" in lib.tt_last_error()
    assert lib.tt_emb_segment_grad_workspace(1000, 64, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.tt_ce_workspace(512, 0, 10, 128, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.tt_score_topk_workspace(64, 1000, 128, 10, ctypes.byref(n)) == 0 and n.value > 0


@pytest.mark.parametrize("name", ["pool_small", "seq_small"])
def test_same_seed_same_weights_and_keys(name):
    npz, cfg = load_golden(name)
    init = unflatten(npz, "state_init")
    torch.manual_seed(int(npz["seed"]))
    model = tt.TwoTowerModel(tt.GenericTower(cfg, "user_tower"), tt.GenericTower(cfg, "item_tower"))
    sd = model.state_dict()
    assert list(sd.keys()) == list(init.keys())
    for k, v in init.items():
        assert torch.equal(sd[k], v), k
    model.load_state_dict(unflatten(npz, "state0"))  # reference checkpoints load unchanged


def test_shipped_config_parameter_count():
    cfg = synth.config_c2()
    model = tt.TwoTowerModel(tt.GenericTower(cfg, "user_tower"), tt.GenericTower(cfg, "item_tower"))
    assert sum(p.numel() for p in model.parameters()) == 865760


def test_config_errors_match_reference():
    with pytest.raises(ValueError):
        tt.GenericTower({"two_tower": {}}, "user_tower")
    cfg = synth.config_c1()
    del cfg["two_tower"]["user_tower"]["sparse_features"][0]["vocab_size"]
    with pytest.raises(ValueError):
        tt.GenericTower(cfg, "user_tower")
    cfg = synth.config_c2()
    cfg["two_tower"]["user_tower"]["transformer_parameters"]["n_head"] = 5
    with pytest.raises(ValueError):
        tt.GenericTower(cfg, "user_tower")


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    cfg = synth.config_c1()
    model = tt.TwoTowerModel(tt.GenericTower(cfg, "user_tower"), tt.GenericTower(cfg, "item_tower"), *synth.MAPS_C1)
    batch = synth.make_batch_c1(B=8, L=5)
    with pytest.raises(TTError):
        model(batch)
    with pytest.raises(TTError):
        model.compute_loss(torch.eye(2), torch.eye(2))
    with pytest.raises(TTError):
        ops.score_topk(torch.eye(2), torch.eye(2), 1)
    with pytest.raises(TTError):
        _lib.require_device()


def test_synthetic_batches_follow_collate_contract():
    b = synth.make_batch_c2(B=16, n_neg=2)
    assert b["user_tower"]["sparse"].dtype == torch.int64 and b["user_tower"]["sparse"].shape == (16, 1)
    assert b["user_tower"]["dense"].dtype == torch.float32
    assert b["user_tower"]["sequence"]["hist_genre_ids"].shape == (16, 20, 3)
    assert len(b["hard_negatives"]) == 2 and b["hard_negatives"][0]["sparse"].shape == (16, 2)
    h = b["user_tower"]["sequence"]["hist_movie_ids"]
    # right padded: once a 0 appears everything after it is 0
    assert ((h == 0).long().diff(dim=1) >= 0).all()


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the reference's own CPU step on a bounded sample of the headline workload; the only
    arm that runs without a GPU) prints exactly one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    # "reference" = the unmodified reference from baseline/_ref (installed by __graft_entry__.build()), else the oracle port
    ref_installed = os.path.isdir(os.path.join(root, "baseline", "_ref", "project", "models", "TwoTower"))
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_installed else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["steps"] == 1 and d["warmup"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_torch_custom_ops_are_registered_and_cuda_only():
    """torch.ops.tt_b200.* (torch.library custom ops over the C ABI): schemas exist, fake (meta) implementations infer
    shapes, and a CPU tensor is refused by the dispatcher -- there is no CPU kernel to fall back to."""
    ns = torch.ops.tt_b200
    for name in ("gather_pool", "segment_grad", "inbatch_ce", "inbatch_ce_backward", "score_topk"):
        assert hasattr(ns, name), name
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        out = ns.gather_pool(torch.empty(100, 16), torch.empty(7, 5, dtype=torch.int64), ops.POOL_MEAN, 0)
        assert out.shape == (7, 16) and out.dtype == torch.float32
        loss, lse = ns.inbatch_ce(torch.empty(8, 64), torch.empty(8, 64), torch.empty(8, dtype=torch.int64), 0.1, 0)
        assert loss.shape == () and lse.shape == (8,)
        s, i = ns.score_topk(torch.empty(4, 64), torch.empty(50, 64), 10, 0)
        assert s.shape == (4, 10) and s.dtype == torch.float64 and i.dtype == torch.int64
    with pytest.raises((NotImplementedError, RuntimeError)):
        ns.gather_pool(torch.zeros(10, 8), torch.zeros(3, 2, dtype=torch.int64), ops.POOL_SUM, -1)


def test_loss_kernel_host_rules():
    """Pure host logic of the tensor-core loss: the declared id range and the single-pass range rule (include/tt_b200.h:
    |u||w| / T * log2(e) <= 96 for unit-norm embeddings <=> T >= 0.0155 with the rounding margin)."""
    from recommendsystemproject_b200 import ops
    assert [ops.id_bits_for(v) for v in (1, 2, 3, 256, 257, 10_000_001, 100_000_001)] == [1, 1, 2, 8, 9, 24, 27]
    assert ops.CE_FLAG_ID_RANGE == 8 and ops.CE_FLAG_LOGIT_RANGE == 16
    if ops.SINGLE_PASS_DEFAULT:
        assert ops.single_pass_ok(0.05) and ops.single_pass_ok(0.0155) and not ops.single_pass_ok(0.015)
