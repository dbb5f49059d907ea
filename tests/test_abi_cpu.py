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
