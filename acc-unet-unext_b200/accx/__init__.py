"""accx: B200-native (sm_100a) kernels behind ACC-UNet's HANC/MLFC blocks.

Public surface = the reference's own module classes (same names and signatures):
    from accx import ChannelSELayer, HANCLayer, HANCBlock, ResPath, MLFC, Conv2d_batchnorm
    from accx import ACC_UNet, ACC_UNet_W, ACC_UNet_Lite
or, as a drop-in for `from ACC_UNet import ...`, put acc-unet-unext_b200/ on sys.path (it holds
ACC_UNet.py / ACC_UNet_w.py / ACC_UNet_lite.py shims with the reference's module names, and nets/ACC_UNet.py for
the training harness's `from nets.ACC_UNet import ACC_UNet`, which is a DIFFERENT architecture: ACC_UNet_Harness).
"""
from .modules import ChannelSELayer, Conv2d_batchnorm, HANCBlock, HANCLayer, MLFC, ResPath  # noqa: F401
from .model import ACC_UNet, ACC_UNet_Harness, ACC_UNet_Lite, ACC_UNet_W  # noqa: F401
from ._lib import AccxError, load as load_library  # noqa: F401

__all__ = ["ChannelSELayer", "Conv2d_batchnorm", "HANCBlock", "HANCLayer", "MLFC", "ResPath",
           "ACC_UNet", "ACC_UNet_W", "ACC_UNet_Lite", "ACC_UNet_Harness", "AccxError", "load_library"]
