"""Drop-in for the reference module `ACC_UNet_lite` (ACC_UNet/ACC_UNet_lite.py)."""
from accx.modules import ChannelSELayer, Conv2d_batchnorm, HANCBlock, HANCLayer, ResPath  # noqa: F401
from accx.modules import MLFC as _MLFC
from accx.model import ACC_UNet_Lite  # noqa: F401


class MLFC(_MLFC):
    def __init__(self, in_filters1, in_filters2, in_filters3, in_filters4, lenn=1):
        super().__init__(in_filters1, in_filters2, in_filters3, in_filters4, lenn, variant="lite")
