"""Drop-in for the training harness's model module (Experiments/nets/ACC_UNet.py, imported as
`from nets.ACC_UNet import ACC_UNet` at Experiments/train_model.py:24): same blocks, but cnv72 with inv_fctr=3 (:584)
and logits out (:596-597,655).  NOT interchangeable with ACC_UNet/ACC_UNet.py: the state_dict shapes differ."""
from accx.modules import ChannelSELayer, Conv2d_batchnorm, HANCBlock, HANCLayer, MLFC, ResPath  # noqa: F401
from accx.model import ACC_UNet_Harness as ACC_UNet  # noqa: F401
