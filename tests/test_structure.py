"""Structural parity with the golden data the reference ships:
Keras summary dumps (models/*/X3D_*.txt) and checkpoint indices (models/*/model.index)."""
import math

import pytest

from x3d_tf_b200 import arch as A
from x3d_tf_b200.config import get_config, three_crop_size, variants
from oracle import x3d_oracle as O

CANON = {"X3D_XS": (4, 160), "X3D_S": (13, 160), "X3D_M": (16, 224), "X3D_L": (16, 312),
         "X3D_XL": (16, 312)}


@pytest.mark.parametrize("variant", list(CANON))
def test_param_counts_match_keras_summary(variant, summaries):
    g = summaries[variant]
    arch = A.build_arch(get_config(variant))
    pc = A.param_counts(arch)
    by_name = {r["name"]: r for r in g["layers"]}
    for k in ("conv_1", "res_stage_2", "res_stage_3", "res_stage_4", "res_stage_5", "conv_5",
              "fc_1", "fc_2"):
        assert pc[k] == by_name[k]["params"], k
    assert pc["total"] == g["total"]
    assert pc["trainable"] == g["trainable"]
    assert pc["non_trainable"] == g["non_trainable"]


@pytest.mark.parametrize("variant", list(CANON))
def test_output_shapes_match_keras_summary(variant, summaries):
    g = {r["name"]: r["shape"] for r in summaries[variant]["layers"]}
    T, S = CANON[variant]
    assert g["input_1"] == [None, T, S, S, 3]
    arch = A.build_arch(get_config(variant))
    plan = A.plan_shapes(arch, T, S, S)
    assert g["conv_1"] == [None, T, plan.stem.H, plan.stem.W, arch.stem_channels]
    for s in range(4):
        lv = plan.stages[s]
        assert g[f"res_stage_{s + 2}"] == [None, T, lv.H, lv.W, arch.stage_dims[s][3]]
    assert g["conv_5"][-1] == arch.conv5_channels
    assert g["fc_1"][-1] == 2048 and g["fc_2"][-1] == 400


@pytest.mark.parametrize("variant", ["X3D_XS", "X3D_S", "X3D_M"])
def test_variables_match_checkpoint_index(variant, checkpoint_index):
    """Every model variable name and shape equals the shipped bundle's (476 tensors)."""
    suffix = "/.ATTRIBUTES/VARIABLE_VALUE"
    model_keys = {k[:-len(suffix)]: tuple(shape) for k, dt, shape, off, size in checkpoint_index["keys"]
                  if k.endswith(suffix) and "/.OPTIMIZER_SLOT/" not in k
                  and not k.startswith("optimizer/")}
    assert len(model_keys) == 476
    ours = dict(A.variable_shapes(A.build_arch(get_config(variant))))
    assert ours == model_keys
    # the oracle derives the same set independently
    assert O.OracleSpec.from_cfg(get_config(variant)).variable_shapes() == model_keys
    # byte total: 15 183 320 = 4 * 3 795 830
    assert sum(4 * math.prod(s) for s in ours.values()) == 15183320


def test_se_placement_xsm():
    arch = A.build_arch(get_config("X3D_M"))
    got = {s: [b.index for b in arch.stage_blocks(s) if b.has_se] for s in range(4)}
    assert got == {0: [0, 2], 1: [1, 3], 2: [0, 2, 4, 6, 8, 10], 3: [1, 3, 5]}
    assert [b.se_width for b in arch.blocks if b.has_se and b.index < 2] == [8, 8, 16, 32]


def test_se_placement_l_xl():
    arch = A.build_arch(get_config("X3D_L"))
    got = {s: [b.index for b in arch.stage_blocks(s) if b.has_se] for s in range(4)}
    assert got[0] == [0, 2, 4] and got[1] == [1, 3, 5, 7, 9]
    assert got[2] == list(range(1, 25, 2)) and got[3] == list(range(0, 15, 2))
    xl = A.build_arch(get_config("X3D_XL"))
    assert xl.stem_channels == 32
    assert xl.stage_dims == [(5, 32, 72, 32), (10, 32, 162, 72), (25, 72, 306, 136),
                             (15, 136, 630, 280)]
    assert sorted({b.se_width for b in xl.blocks if b.has_se}) == [8, 16, 24, 40]


def test_global_counter_quirk():
    """model.py:326: the block counter is process-global; a second X3D-L starts at 56."""
    second = A.build_arch(get_config("X3D_L"), first_block_index=56)
    assert [b.index for b in second.stage_blocks(0) if b.has_se] == [1, 3]


def test_round_width_repeats():
    assert A.round_width(12, 2) == 24 and A.round_width(12, 2.9) == 32
    assert A.round_width(54, 0.0625) == 8 and A.round_width(432, 0.0625) == 32
    assert A.round_width(630, 0.0625) == 40 and A.round_width(24, 0) == 24
    assert A.round_repeats(5, 2.2) == 11 and A.round_repeats(3, 5.0) == 15
    for w in range(1, 700, 7):
        for m in (0.0625, 1.0, 2.0, 2.9):
            assert A.round_width(w, m) == O.round_width(w, m)


@pytest.mark.parametrize("size,k,s,exp", [
    (91, 3, 2, (46, 1, 1)), (46, 3, 2, (23, 0, 1)), (45, 3, 2, (23, 1, 1)),
    (23, 3, 2, (12, 1, 1)), (39, 3, 2, (20, 1, 1)), (112, 3, 2, (56, 0, 1)),
    (16, 3, 1, (16, 1, 1)), (7, 3, 1, (7, 1, 1))])
def test_same_pad(size, k, s, exp):
    assert A.same_pad(size, k, s) == exp
    assert O.tf_same_pads(size, k, s) == exp[1:]


def test_spatial_plan_table():
    """SURVEY.md Appendix B."""
    arch = A.build_arch(get_config("X3D_S"))
    p = A.plan_shapes(arch, 13, 182, 182)
    assert (p.stem.H, [l.H for l in p.stages], p.pads) == \
        (91, [46, 23, 12, 6], [(1, 1), (0, 0), (1, 1), (0, 0)])
    arch = A.build_arch(get_config("X3D_L"))
    p = A.plan_shapes(arch, 16, 356, 356)
    assert (p.stem.H, [l.H for l in p.stages], p.pads) == \
        (178, [89, 45, 23, 12], [(0, 0), (1, 1), (1, 1), (1, 1)])


def test_config_tree():
    cfg = get_config("X3D_M")
    assert cfg.NETWORK.BN.EPS == 1e-5 and cfg.TEST.NUM_TEMPORAL_VIEWS == 10
    assert cfg.DATA.TEMP_DURATION == 16 and three_crop_size("X3D_M") == 256
    with pytest.raises(AttributeError):
        cfg.NETWORK.NUM_CLASSES = 3          # frozen
    c2 = cfg.clone()
    c2.NETWORK.NUM_CLASSES = 3
    assert cfg.NETWORK.NUM_CLASSES == 400
    with pytest.raises(KeyError):
        c2.merge_from_dict({"NETWORK": {"NOPE": 1}})
    assert set(variants()) == {"X3D_XS", "X3D_S", "X3D_M", "X3D_L", "X3D_XL"}


def test_config_merges_reference_style_yaml(tmp_path):
    p = tmp_path / "x.yaml"
    p.write_text("NETWORK:\n  WIDTH_FACTOR: 2.9\n  DEPTH_FACTOR: 5.0\n  SCALE_RES2: True\n"
                 "  WEIGHT_DECAY: 5e-5\n  BN:\n    EPS: 1e-5\nTEST:\n  NUM_TEMPORAL_VIEWS: 3\n")
    from x3d_tf_b200.config import get_default_config
    cfg = get_default_config()
    cfg.merge_from_file(str(p))
    cfg.freeze()
    assert cfg.NETWORK.SCALE_RES2 is True and cfg.NETWORK.WEIGHT_DECAY == 5e-5
    assert cfg.NETWORK.BN.EPS == 1e-5


def test_training_lr_schedule_matches_train_py():
    """train.py:114-125 -- linear warm-up then half-cosine, per epoch."""
    import math
    from x3d_tf_b200.config import get_config
    from x3d_tf_b200.training import lr_schedule
    cfg = get_config("X3D_M")
    t = cfg.TRAIN
    assert lr_schedule(cfg, 0) == pytest.approx(t.WARMUP_LR)
    assert lr_schedule(cfg, t.WARMUP_EPOCHS) == pytest.approx(t.BASE_LR)
    e = t.WARMUP_EPOCHS + 10
    assert lr_schedule(cfg, e) == pytest.approx(t.BASE_LR * 0.5 * (math.cos(math.pi * e / t.EPOCHS) + 1))


def test_pixel_pairing_host_logic():
    """PointwiseConv's pixel pairing (host side, no kernel call): which layers are paired, that pairs
    never straddle clips, and that the paired weight is blockdiag(W, ..., W) of the BN-folded kernel in
    the packed [Npad, Kpad] layout with the bias repeated per pixel slot."""
    import numpy as np
    import torch
    from x3d_tf_b200 import model as M
    rng = np.random.default_rng(3)
    saved = (M.Options.pair_pixels, M.Options.pair_aligned, M.Options.pair_max_k, M.Options.pair_max_n)
    try:
        M.Options.pair_pixels, M.Options.pair_aligned, M.Options.pair_max_k, M.Options.pair_max_n = 2, False, 256, 256
        cpu = torch.device("cpu")
        mk = lambda K, N: M.PointwiseConv(rng.normal(size=(1, 1, 1, K, N)).astype(np.float32),
                                          rng.uniform(0.5, 1.5, N), rng.normal(size=N), cpu)
        a2, c2, a3, a4 = mk(24, 54), mk(54, 24), mk(48, 108), mk(96, 216)
        assert a2._pair_factor(1000) == 2 and c2._pair_factor(1000) == 2     # 48- / 112-byte rows
        assert a3._pair_factor(1000) == 1                                    # 96 / 224 bytes: sector-aligned
        assert a4._pair_factor(1000) == 1                                    # 2 * 216 > one MMA tile
        assert a2._pair_factor(1001) == 1                                    # odd row count
        assert a2._pair_factor(2 * 13 * 91 * 91, 13 * 91 * 91) == 1          # odd clip: a pair would straddle clips
        assert a2._pair_factor(2 * 16 * 64 * 64, 16 * 64 * 64) == 2
        M.Options.pair_pixels = 4
        assert a2._pair_factor(1000) == 4 and a2._pair_factor(1002) == 2
        M.Options.pair_pixels = 1
        assert a2._pair_factor(1000) == 1
        M.Options.pair_aligned, M.Options.pair_pixels = True, 2
        assert a3._pair_factor(1000) == 2
        wp, bias = a2._paired_weights(2, cpu)
        Ks, Ns = a2.Ks, a2.Ns
        assert wp.shape == ((2 * Ns + 15) // 16 * 16, (2 * Ks + 63) // 64 * 64) and wp.dtype == torch.bfloat16
        w = wp.float().numpy()
        single = a2.wp.float().numpy()[:Ns, :Ks]
        assert np.array_equal(w[:Ns, :Ks], single) and np.array_equal(w[Ns:2 * Ns, Ks:2 * Ks], single)
        assert not w[:Ns, Ks:].any() and not w[Ns:2 * Ns, :Ks].any() and not w[2 * Ns:].any()
        assert torch.equal(bias, torch.cat([a2.bias, a2.bias]))
    finally:
        M.Options.pair_pixels, M.Options.pair_aligned, M.Options.pair_max_k, M.Options.pair_max_n = saved


def test_precision_policy_and_strategy_host_logic(monkeypatch):
    """utils.get_precision / get_strategy (utils.py:144-192): policy names and the no-CPU rule."""
    import torch
    from x3d_tf_b200 import runtime as R
    assert R.get_precision(False) == "float32"
    assert R.get_precision(True) == ("mixed_bfloat16" if torch.cuda.is_available() else "float32")
    assert R.policy_dtype("mixed_bfloat16") == "bfloat16" and R.policy_dtype("mixed_float16") == "bfloat16"
    assert R.policy_dtype("float32") == "float32"
    with pytest.raises(ValueError):
        R.policy_dtype("float64")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            R.get_strategy(1)                       # the reference would fall back to the CPU; this build must not
    st = R.Strategy(rank=1, world=2, device=torch.device("cpu"))
    assert st.num_replicas_in_sync == 2 and st.shard(7) == (4, 7)
