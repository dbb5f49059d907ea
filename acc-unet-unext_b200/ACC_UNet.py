"""Drop-in for the reference module `ACC_UNet` (ACC_UNet/ACC_UNet.py): `from ACC_UNet import MLFC` etc.
resolve to the accx (sm_100a CUDA) implementations with unchanged constructors and signatures."""
from accx.modules import ChannelSELayer, Conv2d_batchnorm, HANCBlock, HANCLayer, MLFC, ResPath  # noqa: F401
from accx.model import ACC_UNet  # noqa: F401
