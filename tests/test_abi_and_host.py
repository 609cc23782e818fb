"""CPU suite: the C-ABI library loads and exports every symbol include/ldt_b200.h declares; the host mirrors keep
the reference's state_dict layout; host-side logic (config validation, torch-RNG launch geometry, sharding)."""
import ctypes as C
import json
import os
import re

import pytest
import torch

from tests.helpers import GOLDEN, airplane_config, ns

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ldt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ldt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ldt_b200 import _lib
    from ldt_b200.build import build
    build()
    lib = C.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ldt_b200.h but not exported"
    # and the ctypes prototypes cover exactly the declared entry points
    assert sorted(_lib.PROTOTYPES) == syms
    L = _lib.load()
    header = open(os.path.join(ROOT, "include", "ldt_b200.h")).read()
    assert L.ldt_abi_version() == int(re.search(r"#define LDT_ABI_VERSION (\d+)", header).group(1)) == 2
    assert isinstance(L.ldt_last_error_string(), bytes)


def test_argument_validation_without_gpu():
    """Entry points validate shapes before touching the device: error codes + messages, no CUDA call needed."""
    from ldt_b200 import _lib
    L = _lib.load()
    assert L.ldt_nn_distance(-1, 4, None, 4, None, None, None, None, None, None) == -1
    assert b"negative" in L.ldt_last_error_string()
    assert L.ldt_nn_distance(0, 4, None, 4, None, None, None, None, None, None) == 0  # empty batch is a no-op
    assert L.ldt_nn_distance(2, 4, None, 0, None, None, None, None, None, None) == -1  # one empty set: refuse
    assert L.ldt_pairwise_cd(4, 4, 8, 8, None, None, 3, 2, None, None) == -1
    assert L.ldt_pairwise_cd(4, 4, 8, 8, None, None, 2, 2, None, None) == 0  # empty row range
    a = _lib.GemmArgs(M=128, N=128, K=100)
    assert L.ldt_gemm_bf16(C.byref(a), None) == -1 and b"multiple of 64" in L.ldt_last_error_string()
    assert L.ldt_attention_nk32(1, 4, 32, 48, None, 0, None, None, 0, None, None) == -3
    assert L.ldt_layernorm_mod_bf16(4, 100, None, None, None, 0, 1, None, None, 1e-6, None, None) == -1
    assert L.ldt_sde_step(9, 16, 1, 1, None, 1, None, 0, 0, 0, None, 0, 1, None, None) == -1
    assert L.ldt_pairwise_cd_upper(8, 16, None, 2, 2, None, None) == -1 and b"row_first" in L.ldt_last_error_string()
    assert L.ldt_pairwise_cd_upper(0, 16, None, 0, 1, None, None) == 0
    plan = _lib.ScorePlan(batch=2, tokens=32, z_dim=120, z_pad=128, hidden=128, heads=4, mlp_hidden=512, num_blocks=0)
    assert L.ldt_score_forward(C.byref(plan), None, None, 0, None, None) == -3 and b"head dim 64" in L.ldt_last_error_string()
    assert L.ldt_score_forward(None, None, None, 0, None, None) == -1
    assert L.ldt_sample_loop(None, None) == -1
    assert L.ldt_decoder_forward(None, None, None, None, None) == -1
    assert L.ldt_group_features(2, 64, 4, 8, 16, None, None, None, None, 3, None, None, None, None, 64, None) == -1
    assert b"normalize" in L.ldt_last_error_string()
    assert L.ldt_group_features(2, 64, 4, 8, 16, None, None, None, None, 2, None, None, None, None, 32, None) == -1
    assert b"ld_out" in L.ldt_last_error_string()
    assert L.ldt_group_features(0, 64, 4, 8, 16, None, None, None, None, 0, None, None, None, None, 64, None) == 0
    assert L.ldt_group_max(0, 8, 16, None, 16, None, 16, None) == 0 and L.ldt_group_max(4, 0, 16, None, 16, None, 16, None) == -1
    assert L.ldt_split_tf32(4, 40, None, 40, None, 32, 0, 0, None) == -1 and b"ld_part" in L.ldt_last_error_string()
    a = _lib.GemmArgs(M=128, N=128, K=96, operand_type=1, epilogue=1)
    a.lda = a.ldw = 96
    a.ldo = 128
    assert L.ldt_gemm_bf16(C.byref(a), None) < 0   # bf16 epilogues are not available to f32 operands (and operands are null)


def test_state_dict_layout_matches_reference():
    """Keys, shapes AND order equal the reference modules' (checkpoint contract)."""
    from ldt_b200 import Compressor, Score
    lay = json.load(open(os.path.join(GOLDEN, "state_dict_layout.json")))
    cfg = ns(airplane_config())
    cfg.score.num_blocks = 2
    mine = [[k, list(v.shape)] for k, v in Score(cfg.score).state_dict().items()]
    assert mine == lay["score_2blocks"]
    mine = [[k, list(v.shape)] for k, v in Compressor(cfg.compressor).state_dict().items()]
    assert mine == lay["compressor"]
    c = Compressor(cfg.compressor)
    assert float(c.conv_in.initialized) == 0.0
    c.init()
    assert float(c.conv_in.initialized) == 1.0


def test_config_validation_mirrors_reference_errors():
    from ldt_b200 import Compressor, DiffusionVPSDE, Score
    cfg = ns(airplane_config())
    del cfg.score.condition  # the reference raises AttributeError on missing keys (score.py:56)
    with pytest.raises(AttributeError):
        Score(cfg.score)
    cfg = ns(airplane_config())
    cfg.score.AdaLN = False   # only the AdaLN blocks of the shipped configs are built
    with pytest.raises(NotImplementedError):
        Score(cfg.score)
    cfg = ns(airplane_config())
    cfg.score.unet, cfg.score.num_blocks, cfg.score.hidden_size, cfg.score.num_heads = True, 2, 128, 2
    assert any(k.startswith("Transformer_Down.0.adaLN2.1.") for k in Score(cfg.score).state_dict())
    sde = DiffusionVPSDE(cfg.sde, device="cpu")
    assert sde.betas.dtype == torch.float32 and sde.betas.shape == (1000,)
    with pytest.raises(NotImplementedError, match="preditor not Implemented"):  # message of diffusion_continuous.py:327
        sde.sample_discrete(None, 1, 10, "nope", None, 1, (32, 120), 1e-6, False, True, 0.01, "cpu")
    with pytest.raises(NotImplementedError, match="corrector not Implemented"):
        sde.sample_discrete(None, 1, 10, "ancestral", "nope", 1, (32, 120), 1e-6, False, True, 0.01, "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):   # no CPU fallback for the encoder path either
        Compressor(cfg.compressor).forward(torch.zeros(1, 64, 3))


def test_batchnorm_folding_equals_conv_then_batchnorm_on_cpu():
    """grouping.fold_conv_bn: Conv1d(k=1) followed by an eval-mode BatchNorm1d is one affine map (W*s, (b-mean)*s+beta); this
    is what lets the prologue layers run as single contractions with a ReLU epilogue (Compressor/layers.py:115-160)."""
    from ldt_b200.grouping import fold_conv_bn
    torch.manual_seed(0)
    conv, bn = torch.nn.Conv1d(37, 16, 1), torch.nn.BatchNorm1d(16)
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.3)
        bn.running_var.uniform_(0.5, 1.5)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
    bn.eval()
    x = torch.randn(5, 37, 11)
    W, b = fold_conv_bn(conv, bn)
    with torch.no_grad():
        want = bn(conv(x))
    got = torch.einsum("oc,bcn->bon", W, x) + b[None, :, None]
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    W0, b0 = fold_conv_bn(torch.nn.Linear(8, 4))
    assert W0.shape == (4, 8) and b0.shape == (4,)


def test_resnet_trunk_im2col_wiring_on_cpu(monkeypatch):
    """grouping.resnet_trunk_maxpool (ConditionNet's image branch, scorenet/score.py:24-26,33-35): im2col ordering, stride /
    padding arithmetic, the max-pool by unfold and the BasicBlock residual wiring, with the two device kernels stood in by
    their torch definitions (the kernels themselves are checked on the GPU against the reference golden)."""
    from torchvision import models
    from ldt_b200 import grouping, ops
    from ldt_b200._lib import EPI_BIAS_F32, EPI_BIAS_RELU_F32, EPI_RESID_RELU_F32
    torch.manual_seed(1)
    trunk = torch.nn.Sequential(*list(models.resnet18(weights=None).children())[:-4])
    with torch.no_grad():
        for m in trunk.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.2)
    trunk.eval()

    def conv_rows(x, packed, epilogue=EPI_BIAS_F32, resid=None):
        W, b = packed
        y = x.double() @ W.double().t() + b.double()
        if epilogue == EPI_RESID_RELU_F32:
            y = y + resid.double()
        if epilogue in (EPI_BIAS_RELU_F32, EPI_RESID_RELU_F32):
            y = y.clamp_min(0)
        return y.float()
    monkeypatch.setattr(grouping, "pack_tf32", grouping.fold_conv_bn)
    monkeypatch.setattr(grouping, "conv_rows", conv_rows)
    monkeypatch.setattr(ops, "group_max", lambda x, k, c=None: x.reshape(x.shape[0] // k, k, x.shape[1]).amax(dim=1))
    img = torch.rand(2, 3, 64, 96)
    with torch.no_grad():
        want = torch.nn.functional.adaptive_max_pool2d(trunk(img), 1).reshape(2, 128)
        got = grouping.resnet_trunk_maxpool(trunk, img)
    assert got.shape == (2, 128) and torch.allclose(got, want, rtol=1e-4, atol=1e-5), (got - want).abs().max()


def test_product_path_never_imports_oracle():
    """The oracle is test infrastructure: nothing under ldt_b200/ may import or reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ldt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/_ref", "").lower() or f == "nn_distance.cu", (dirpath, f)


def test_cpu_tensors_fail_loudly():
    from ldt_b200 import Score, ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.nn_distance_idx(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))
    cfg = ns(airplane_config())
    cfg.score.num_blocks = 1
    cfg.score.hidden_size, cfg.score.num_heads, cfg.score.t_dim = 128, 2, 128
    with pytest.raises(RuntimeError, match="CUDA"):
        Score(cfg.score).eval()(torch.zeros(1, 32, 120), torch.zeros(1))


def test_sde_coefficient_table_matches_oracle_on_cpu():
    """step_coefficients reproduces the per-step scalars of every predictor (bit-exact vs the oracle's op sequence)."""
    from ldt_b200 import DiffusionVPSDE
    from oracle import ldt_oracle as O
    c = airplane_config()["sde"]
    sde = DiffusionVPSDE(ns(airplane_config()).sde, device="cpu")
    osde = O.VPSDE(c["beta_start"], c["beta_end"], c["sigma2_0"], c["sample_N"])
    coef, ts = sde.step_coefficients("ancestral", 1000, 1e-6, False, torch.device("cpu"))
    idx = (ts * 999).long()
    assert torch.equal(coef[:, 0], torch.sqrt(osde.var(ts)))
    assert torch.equal(coef[:, 1], osde.betas[idx])
    assert torch.equal(coef[:, 2], torch.sqrt(1.0 - osde.betas[idx]))
    assert torch.equal(coef[:, 3], torch.sqrt(osde.betas[idx]))
    coef, ts = sde.step_coefficients("ddim", 1000, 1e-6, False, torch.device("cpu"))
    assert float(coef[-1, 1]) == 1.0 and float(coef[-1, 4]) == 0.0  # last step: at_next = 1 (:169-170)
    assert torch.equal(coef[:-1, 1], osde.alphas_cump[idx[:-1] - 1].sqrt())
    raw, _ = sde.step_coefficients("eulermaruyama", 10, 1e-6, True, torch.device("cpu"), raw_score=True)
    assert torch.all(raw[:, 0] == -1.0) and torch.all(raw[:, 4] == 0.0)


def test_shard_ranges_cover_exactly():
    from ldt_b200.distributed import shard_range
    for total in (0, 1, 7, 2048, 2049):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("tiles_m,tn1,tn2,kb1,kb2,pairs", [(32, 16, 4, 16, 64, 74), (8, 16, 4, 16, 64, 74), (1, 16, 4, 16, 64, 16),
                                                           (5, 4, 1, 4, 16, 20), (16, 2, 2, 8, 8, 32), (33, 16, 4, 16, 64, 74),
                                                           (32, 16, 4, 16, 64, 66), (7, 3, 5, 4, 12, 74)])
def test_fused_mlp_static_schedule_covers_every_tile_once(tiles_m, tn1, tn2, kb1, kb2, pairs):
    """Host view of mlp.cu's static work lists (ldt_mlp_schedule_item): every fc1 and fc2 tile appears exactly once, a pair
    never starts an fc2 tile before its last fc1 tile (the deadlock-freedom argument), and the k-block load is balanced."""
    import ctypes as C
    from ldt_b200 import _lib
    lib = _lib.load()
    seen1, seen2, loads = set(), set(), []
    for p in range(pairs):
        n, ph, mb, nt = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        assert lib.ldt_mlp_schedule_item(tiles_m, tn1, tn2, kb1, kb2, pairs, p, 0, C.byref(n), C.byref(ph), C.byref(mb), C.byref(nt)) == 0
        in_phase2, load = False, 0
        for it in range(n.value):
            lib.ldt_mlp_schedule_item(tiles_m, tn1, tn2, kb1, kb2, pairs, p, it, C.byref(n), C.byref(ph), C.byref(mb), C.byref(nt))
            if ph.value == 0:
                assert not in_phase2
                continue
            assert 0 <= mb.value < tiles_m
            if ph.value == 1:
                assert not in_phase2, "fc1 tile after an fc2 tile"
                assert 0 <= nt.value < tn1 and (mb.value, nt.value) not in seen1
                seen1.add((mb.value, nt.value))
                load += kb1
            else:
                in_phase2 = True
                assert 0 <= nt.value < tn2 and (mb.value, nt.value) not in seen2
                seen2.add((mb.value, nt.value))
                load += kb2
        loads.append(load)
    assert len(seen1) == tiles_m * tn1 and len(seen2) == tiles_m * tn2
    ideal = (tiles_m * tn1 * kb1 + tiles_m * tn2 * kb2) / pairs
    if (tiles_m * tn2) % pairs != 0 and kb2 % kb1 == 0 and tiles_m * tn1 >= 4 * pairs:   # the balanced regime
        assert max(loads) <= ideal + kb2 / 2 + kb1, (max(loads), ideal)
