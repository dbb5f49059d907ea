"""CPU: the C-ABI library loads and exports every symbol include/accx.h declares; the Python
mirror of the reference interface keeps its constructors, state_dict layout and error behaviour.
No compute calls here (no GPU in the build container)."""
import ctypes
import inspect
import os

import pytest
import torch

from helpers import load_case, module_cases


def test_library_exports_every_declared_symbol():
    from accx import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in include/accx.h but not exported by libaccx.so"
    _lib.load()
    assert _lib.load().accx_version() >= 100
    assert ctypes.sizeof(_lib.Operand) == 72          # accx_operand_t layout (8-byte aligned fields)


def test_invalid_arguments_return_error_codes_not_crashes():
    from accx import _lib
    lib = _lib.load()
    rc = lib.accx_bn_finalize(0, 1.0, None, None, None, None, 1e-5, 0.1, 1, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"bn_finalize" in lib.accx_last_error()
    with pytest.raises(_lib.AccxError):
        _lib.call("accx_pool_sum", 0, 0, 1, 3, 3, 8, 1, 1.0, 1, 1, 8, None)     # 3x3 not divisible by 2


def test_train_step_entry_points_validate_arguments():
    """rows f1 / f2 of the scope table: bad arguments come back as error codes before anything touches the GPU"""
    from accx import _lib
    lib = _lib.load()
    err = lambda: lib.accx_last_error().decode()
    assert lib.accx_adam_step(3, None, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, None) == -1
    assert "adam_step" in err()
    assert lib.accx_maxpool2_bwd(1, 1, 7, 8, 8, 1, 1, 1, None) == -1 and "even" in err()          # odd height
    assert lib.accx_maxpool2_fwd(1, 1, 8, 8, 8, None, None, None) == -1 and "maxpool2_fwd" in err()
    assert lib.accx_upshuffle(1, 1, 1, 4, 4, 7, 16, None, 16, 14, None, None) == -1 and "even" in err()   # odd Co
    assert lib.accx_copy_cols(1, 4, 8, 16, 4, 16, 8, None) == -1 and "copy_cols" in err()            # ld_src < C
    assert lib.accx_dice_bce_fwd(1, 2, 16, None, None, 0.5, 0.5, None, None, None) == -1 and "dice_bce_fwd" in err()
    assert lib.accx_dice_bce_bwd(1, 7, 2, 16, 16, 16, 16, 0.5, 0.5, None, None, None) == -1 or "dice_bce_bwd" in err()
    assert lib.accx_set_knob(999, 1) == -1 and "out of range" in err()
    assert lib.accx_set_knob(0, 0) == 0


def test_every_entry_point_is_mapped_to_the_reference_in_integration_md():
    from accx import _lib
    doc = open(os.path.join(_lib.ROOT, "INTEGRATION.md")).read()
    missing = [n for n in _lib.parse_header() if f"`{n}`" not in doc]
    assert not missing, f"INTEGRATION.md does not say which reference call these replace: {missing}"


def test_reference_signatures_are_kept():
    import accx
    sig = lambda c: list(inspect.signature(c.__init__).parameters)[1:]
    assert sig(accx.ChannelSELayer) == ["num_channels"]
    assert sig(accx.HANCLayer) == ["in_chnl", "out_chnl", "k"]
    assert sig(accx.HANCBlock) == ["n_filts", "out_channels", "k", "inv_fctr"]
    assert sig(accx.ResPath) == ["in_chnls", "n_lvl"]
    assert sig(accx.MLFC)[:5] == ["in_filters1", "in_filters2", "in_filters3", "in_filters4", "lenn"]
    assert sig(accx.ACC_UNet)[:3] == ["n_channels", "n_classes", "n_filts"]
    d = inspect.signature(accx.HANCBlock.__init__).parameters
    assert d["k"].default == 3 and d["inv_fctr"].default == 3


@pytest.mark.parametrize("name", module_cases())
def test_state_dict_layout_matches_reference(name):
    from test_modules_gpu import build
    case = load_case(name)
    mod = build(name)
    sd = mod.state_dict()
    assert list(sd.keys()) == list(case["sd"].keys())
    for k, v in case["sd"].items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    mod.load_state_dict(case["sd"])


def test_default_init_reproduces_reference_weights():
    """same submodule construction order => torch.manual_seed(2) gives the reference's weights"""
    import accx
    z = load_case("init_seed2_f8")["raw"]
    torch.manual_seed(2)
    sd = accx.ACC_UNet(3, 1, 8).state_dict()
    assert [str(n) for n in z["names"]] == list(sd.keys())
    for n, s in zip(z["names"], z["sums"]):
        assert abs(float(sd[str(n)].double().sum()) - s) <= 1e-9 + 1e-9 * abs(s), n


def test_harness_flavour_reproduces_the_training_harness_model():
    """`from nets.ACC_UNet import ACC_UNet` (Experiments/train_model.py:24) is NOT the canonical model: cnv72 is built
    with inv_fctr=3 and the forward returns logits (Experiments/nets/ACC_UNet.py:584,596-597,655).  The shim under
    acc-unet-unext_b200/nets/ resolves to accx.ACC_UNet_Harness, whose state_dict (names, shapes, seed-2 weights)
    is the harness model's -- so train_model.py checkpoints load -- while accx.ACC_UNet keeps the canonical layout."""
    import accx
    from nets.ACC_UNet import ACC_UNet as HarnessNet
    assert HarnessNet is accx.ACC_UNet_Harness
    z = load_case("init_seed2_f8_harness")["raw"]
    torch.manual_seed(2)
    m = HarnessNet(3, 1, 8)
    sd = m.state_dict()
    assert [str(n) for n in z["names"]] == list(sd.keys())
    for n, shp, s in zip(z["names"], z["shapes"], z["sums"]):
        assert str(tuple(sd[str(n)].shape)) == str(shp), n
        assert abs(float(sd[str(n)].double().sum()) - s) <= 1e-9 + 1e-9 * abs(s), n
    assert m.last_activation is None
    assert tuple(sd["cnv72.conv1.weight"].shape) == (96, 32, 1, 1)
    torch.manual_seed(2)
    canon = accx.ACC_UNet(3, 1, 8)
    assert tuple(canon.state_dict()["cnv72.conv1.weight"].shape) == (1088, 32, 1, 1)
    assert isinstance(canon.last_activation, torch.nn.Sigmoid)
    with pytest.raises(RuntimeError):
        canon.load_state_dict(sd)                 # the two flavours are not interchangeable


def test_dropin_module_names_importable():
    import ACC_UNet, ACC_UNet_lite, ACC_UNet_w   # noqa: E401
    assert ACC_UNet.MLFC(8, 8, 8, 8).variant == "base"
    assert ACC_UNet_w.MLFC(8, 8, 8, 8).variant == "w" and hasattr(ACC_UNet_w.MLFC(8, 8, 8, 8), "W")
    assert ACC_UNet_lite.MLFC(8, 8, 8, 8).variant == "lite"
    for cls in ("ChannelSELayer", "HANCLayer", "HANCBlock", "ResPath", "MLFC", "ACC_UNet"):
        assert hasattr(ACC_UNet, cls)


def test_no_cpu_path():
    import accx
    with pytest.raises(accx.AccxError):
        accx.HANCBlock(8, 8)(torch.randn(1, 8, 4, 4))


def test_product_never_imports_the_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "acc-unet-unext_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dp, f)).read().lower(), f


# ---- host-side launch planning of the tensor-core contraction (no device needed) -----------------------------------
PLAN = ("fold", "conv3", "bn", "n_tiles", "m_tiles", "stages", "smem", "resident", "addst_w", "tmem_cols", "n_kb", "grid")


def _plan(B, H, W, N, ops, adds=(), res_ld=None, ldy=None, stats=True, dtype=1, out_dtype=1):
    """ops: [(K, ld, dy, dx, act)]; -> dict of accx_pw_fwd_tc_plan (dtype 1 = bf16, 0 = fp32, include/accx.h)"""
    from accx import _lib
    lib = _lib.load()
    arr = (_lib.Operand * len(ops))()
    for o, (K, ld, dy, dx, act) in zip(arr, ops):
        o.K, o.ld, o.dy, o.dx, o.act = K, ld, dy, dx, act
    al = (ctypes.c_int * max(1, len(adds)))(*adds)
    out = (ctypes.c_int * 12)()
    rc = lib.accx_pw_fwd_tc_plan(dtype, out_dtype, B, H, W, N, arr, len(ops), len(adds), al, int(res_ld is not None),
                                 res_ld or 0, N if ldy is None else ldy, int(stats), out, 12)
    assert rc == 0, lib.accx_last_error().decode()
    return dict(zip(PLAN, out))


def test_plan_pixel_folding_rules():
    """two pixels per row only for narrow, contiguous, unshifted bf16 contractions"""
    from accx import _lib
    lib = _lib.load()
    F32 = 0                                   # ACCX_F32 (include/accx.h); the default of _plan is ACCX_BF16 = 1
    base = dict(B=16, H=224, W=224, N=32, ops=[(32, 32, 0, 0, 2)])
    p = _plan(**base)
    assert p["fold"] == 1 and p["bn"] == 64 and p["n_kb"] == 1 and p["m_tiles"] == 16 * 224 * 224 // 256
    assert p["resident"] == 1 and p["smem"] <= 227 * 1024 and p["stages"] >= 3 and p["grid"] == 148
    assert _plan(**{**base, "N": 96})["fold"] == 1 and _plan(**{**base, "N": 96})["bn"] == 192
    assert _plan(**{**base, "ops": [(96, 96, 0, 0, 2)]})["fold"] == 1                      # narrow output
    assert _plan(**{**base, "ops": [(32, 32, 0, 0, 2), (32, 32, 0, 0, 0)]})["fold"] == 1      # several operands
    assert _plan(**base, res_ld=32)["fold"] == 1
    assert _plan(**base, adds=(1, 2))["fold"] == 1
    # ... and not otherwise
    assert _plan(**{**base, "N": 64, "ops": [(64, 64, 0, 0, 2)]})["fold"] == 0               # no side <= 32 channels
    assert _plan(**{**base, "ops": [(32, 64, 0, 0, 2)]})["fold"] == 0                        # column slice of a wider tensor
    assert _plan(**{**base, "ops": [(32, 32, 0, 1, 2)]})["fold"] == 0                        # shifted tap
    assert _plan(**base, ldy=64)["fold"] == 0                                                # sliced output
    assert _plan(**base, res_ld=64)["fold"] == 0
    assert _plan(**base, adds=(0,))["fold"] == 0                                             # addend at full resolution
    assert _plan(B=1, H=3, W=5, N=32, ops=[(32, 32, 0, 0, 2)])["fold"] == 0                  # odd pixel count
    assert _plan(B=2, H=3, W=5, N=32, ops=[(32, 32, 0, 0, 2)], adds=())["fold"] == 1
    assert _plan(**{**base, "N": 40})["fold"] == 0                                           # N % 16 != 0
    assert _plan(**{**base, "ops": [(160, 160, 0, 0, 2)]})["fold"] == 0                      # sum K > 128
    assert _plan(**base, dtype=F32, out_dtype=F32)["fold"] == 0                              # fp32 storage
    assert _plan(**base, out_dtype=F32, stats=False)["fold"] == 0                            # fp32 output
    assert lib.accx_set_knob(23, 1) == 0
    try:
        assert _plan(**base)["fold"] == 0                                                    # knob 23 = 1: off
    finally:
        lib.accx_set_knob(23, 0)


def test_plan_slab_mode_and_addend_staging():
    taps = [(32, 32, dy, dx, 2) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    p = _plan(16, 224, 224, 32, taps)
    assert p["conv3"] == 1 and p["fold"] == 0 and p["resident"] == 1 and p["tmem_cols"] >= 2 * 3 * p["bn"]
    assert _plan(16, 56, 56, 128, [(128, 128, dy, dx, 2) for dy in (-1, 0, 1) for dx in (-1, 0, 1)])["conv3"] == 0   # > 64 channels
    # first addend through shared memory: narrow single column tile with all chunks full, four stages left under the cap
    p = _plan(16, 224, 224, 32, [(32, 32, 0, 0, 2)], adds=(1,))
    assert p["fold"] == 1 and p["addst_w"] == 32 and p["stages"] >= 4
    p = _plan(16, 112, 112, 48, [(64, 64, 0, 0, 2)], adds=(1, 2))
    assert p["fold"] == 0 and p["addst_w"] == 48 and p["stages"] >= 4 and p["smem"] <= 161 * 1024
    p = _plan(16, 112, 112, 64, [(64, 64, 0, 0, 2)], adds=(1, 2))
    assert p["fold"] == 0 and p["addst_w"] == 0 and p["stages"] >= 4    # 70 KB of staging rows would leave three stages
    p = _plan(16, 224, 224, 64, [(192, 192, 0, 0, 2)], adds=(1,))
    assert p["addst_w"] == 0 and p["stages"] >= 4                       # staging would leave two stages: direct loads
    assert _plan(16, 56, 56, 128, [(384, 384, 0, 0, 2)], adds=(1,))["addst_w"] == 0          # 128-channel tile: not staged
    assert _plan(16, 224, 224, 40, [(64, 64, 0, 0, 2)], adds=(1,))["addst_w"] == 0           # ragged last chunk


def test_plan_always_fits_the_sm():
    """every geometry the planner returns can be launched: shared memory <= 227 KB, TMEM <= 512 columns, >= 1 stage, and the
    persistent grid never exceeds the tiles; swept over the model's shapes and some hostile ones"""
    import itertools
    Ks = (8, 16, 32, 64, 96, 128, 192, 256, 384, 512, 768, 1536, 4352)
    Ns = (8, 16, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1536, 4352)
    for (K, N), (B, H, W), n_add, res, f32out in itertools.product(
            itertools.product(Ks, Ns), ((16, 224, 224), (16, 14, 14), (1, 7, 5), (2, 56, 56)), (0, 1, 3), (False, True), (False, True)):
        if n_add and (H % 8 or W % 8):
            continue
        if res and N % 8:
            continue
        p = _plan(B, H, W, N, [(K, K, 0, 0, 2)], adds=tuple(range(1, n_add + 1)), res_ld=N if res else None,
                  stats=not f32out, out_dtype=0 if f32out else 1)
        what = (K, N, B, H, W, n_add, res, f32out, p)
        assert 1 <= p["stages"] <= 8 and 0 < p["smem"] <= 227 * 1024, what
        assert p["tmem_cols"] <= 512 and p["tmem_cols"] & (p["tmem_cols"] - 1) == 0, what
        assert p["bn"] % 8 == 0 and p["bn"] * p["n_tiles"] >= N * (2 if p["fold"] else 1), what
        assert 1 <= p["grid"] <= p["m_tiles"] * p["n_tiles"], what
        assert not p["fold"] or p["resident"], what
    # fp32 storage (3 x TF32): 32-channel k-blocks, two weight tiles each
    for K, N in ((32, 32), (64, 192), (4352, 128), (128, 4352)):
        p = _plan(16, 56, 56, N, [(K, K, 0, 0, 2)], dtype=0, out_dtype=0)
        assert p["fold"] == 0 and p["n_kb"] == (K + 31) // 32 and p["smem"] <= 227 * 1024 and p["stages"] >= 1, (K, N, p)


WPLAN = ("fold", "nb", "n_tiles", "k_tiles", "dy_blocks", "stage_px", "splits", "stages", "smem", "tmem_cols", "grid")


def _wplan(B, H, W, N, K, ld=None, ldy=None, taps=((0, 0),), act=2, expect_rc=0):
    from accx import _lib
    lib = _lib.load()
    o = _lib.Operand()
    o.K, o.ld, o.act = K, K if ld is None else ld, act
    tdy = (ctypes.c_int * len(taps))(*[t[0] for t in taps])
    tdx = (ctypes.c_int * len(taps))(*[t[1] for t in taps])
    out = (ctypes.c_int * 11)()
    rc = lib.accx_pw_wgrad_tc_plan(B, H, W, N, ctypes.byref(o), len(taps), tdy, tdx, N if ldy is None else ldy, out, 11)
    assert rc == expect_rc, lib.accx_last_error().decode()
    return dict(zip(WPLAN, out))


def test_wgrad_plan_folding_stages_and_launchability():
    """host-side plan of the tensor-core weight gradient: two pixels per row for narrow contiguous single-tap operands,
    256-pixel stages over >= 1e5 (folded) rows, and every plan fits an SM (227 KB, 512 TMEM columns, >= 1 stage)"""
    p = _wplan(16, 224, 224, 32, 32)
    assert p["fold"] == 1 and p["nb"] == 64 and p["stage_px"] == 256 and p["n_tiles"] == p["k_tiles"] == 1 and p["dy_blocks"] == 1
    assert p["grid"] == p["splits"] and 1 <= p["grid"] <= 2 * 148
    assert _wplan(16, 224, 224, 64, 32)["fold"] == 1 and _wplan(16, 224, 224, 32, 96)["fold"] == 1
    assert _wplan(16, 224, 224, 64, 64)["fold"] == 0                      # no narrow side
    assert _wplan(16, 224, 224, 96, 32)["fold"] == 0                      # N > 64
    assert _wplan(16, 224, 224, 32, 32, ld=64)["fold"] == 0               # column slice
    assert _wplan(16, 224, 224, 32, 32, ldy=96)["fold"] == 0
    assert _wplan(16, 224, 224, 32, 32, taps=((0, 1),))["fold"] == 0      # shifted tap
    assert _wplan(1, 15, 15, 32, 32)["fold"] == 0                         # odd pixel count
    assert _wplan(16, 224, 224, 32, 24)["fold"] == 0                      # K % 16 != 0
    assert _wplan(3, 181, 187, 32, 32)["stage_px"] == 256 and _wplan(2, 56, 56, 32, 32)["stage_px"] == 128
    nine = tuple((dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1))
    p = _wplan(16, 224, 224, 32, 32, taps=nine[:5])
    assert p["fold"] == 0 and p["nb"] == 32 and p["tmem_cols"] == 256 and p["stage_px"] == 128
    _wplan(16, 56, 56, 128, 128, taps=nine, expect_rc=-1)                 # 9 x 64 columns exceed TMEM: the engine sends <= 5 taps
    for K in (8, 16, 32, 64, 96, 128, 192, 256, 384, 768, 1536, 4352):
        for N in (8, 32, 64, 96, 128, 256, 768, 4352):
            for B, H, W in ((16, 224, 224), (16, 14, 14), (1, 7, 5), (16, 56, 56)):
                for taps in (((0, 0),), nine[:4], nine[4:]):
                    if len(taps) > 1 and K > 256:
                        continue
                    p = _wplan(B, H, W, N, K, taps=taps)
                    what = (K, N, B, H, W, len(taps), p)
                    assert 1 <= p["stages"] <= 4 and 0 < p["smem"] <= 227 * 1024, what
                    assert p["tmem_cols"] <= 512 and p["tmem_cols"] >= len(taps) * p["nb"], what
                    assert p["grid"] == p["n_tiles"] * p["k_tiles"] * p["splits"] and p["grid"] >= 1, what
                    assert p["nb"] % 16 == 0 and p["nb"] * p["k_tiles"] >= K * (2 if p["fold"] else 1), what
