"""ACC-UNet assembled from the accx drop-in blocks (caller glue of ACC_UNet.py:530-659).

The five hot-path module types run the accx kernels, and so does the glue between them (SURVEY.md
section 8, rows a8 / f1): MaxPool2d(2), ConvTranspose2d(2, 2, stride 2) + the skip concat (one
contraction + an interleave), the final 1x1 conv.  Only the optional final sigmoid is a torch op.

`compute_dtype=torch.bfloat16` stores activations in bf16 (fp32 accumulation, statistics and
parameters); default None follows the input's dtype (fp32 = the reference's arithmetic).
"""
from __future__ import annotations

import torch
from torch import nn

from . import engine as E
from .modules import HANCBlock, MLFC, ResPath, maxpool2, out_conv, run_parallel, up_cat


class _ACCUNetBase(nn.Module):
    variant = "base"
    cnv72_inv_fctr = 34        # ACC_UNet/ACC_UNet.py:584; the training harness's copy builds cnv72 with 3 (ACC_UNet_Harness)
    sigmoid_out = True         # ACC_UNet/ACC_UNet.py:594-599: sigmoid on the one-class output

    def __init__(self, n_channels, n_classes, n_filts=32, compute_dtype=None):
        super().__init__()
        self.n_channels = n_channels
        self.n_classes = n_classes
        self.compute_dtype = compute_dtype
        f = n_filts
        self.pool = nn.MaxPool2d(2)
        self.cnv11 = HANCBlock(n_channels, f, k=3, inv_fctr=3)
        self.cnv12 = HANCBlock(f, f, k=3, inv_fctr=3)
        self.cnv21 = HANCBlock(f, f * 2, k=3, inv_fctr=3)
        self.cnv22 = HANCBlock(f * 2, f * 2, k=3, inv_fctr=3)
        self.cnv31 = HANCBlock(f * 2, f * 4, k=3, inv_fctr=3)
        self.cnv32 = HANCBlock(f * 4, f * 4, k=3, inv_fctr=3)
        self.cnv41 = HANCBlock(f * 4, f * 8, k=2, inv_fctr=3)
        self.cnv42 = HANCBlock(f * 8, f * 8, k=2, inv_fctr=3)
        self.cnv51 = HANCBlock(f * 8, f * 16, k=1, inv_fctr=3)
        self.cnv52 = HANCBlock(f * 16, f * 16, k=1, inv_fctr=3)
        self.rspth1 = ResPath(f, 4)
        self.rspth2 = ResPath(f * 2, 3)
        self.rspth3 = ResPath(f * 4, 2)
        self.rspth4 = ResPath(f * 8, 1)
        self.mlfc1 = MLFC(f, f * 2, f * 4, f * 8, lenn=1, variant=self.variant)
        self.mlfc2 = MLFC(f, f * 2, f * 4, f * 8, lenn=1, variant=self.variant)
        self.mlfc3 = MLFC(f, f * 2, f * 4, f * 8, lenn=1, variant=self.variant)
        self.up6 = nn.ConvTranspose2d(f * 16, f * 8, kernel_size=(2, 2), stride=2)
        self.cnv61 = HANCBlock(f * 8 + f * 8, f * 8, k=2, inv_fctr=3)
        self.cnv62 = HANCBlock(f * 8, f * 8, k=2, inv_fctr=3)
        self.up7 = nn.ConvTranspose2d(f * 8, f * 4, kernel_size=(2, 2), stride=2)
        self.cnv71 = HANCBlock(f * 4 + f * 4, f * 4, k=3, inv_fctr=3)
        self.cnv72 = HANCBlock(f * 4, f * 4, k=3, inv_fctr=self.cnv72_inv_fctr)
        self.up8 = nn.ConvTranspose2d(f * 4, f * 2, kernel_size=(2, 2), stride=2)
        self.cnv81 = HANCBlock(f * 2 + f * 2, f * 2, k=3, inv_fctr=3)
        self.cnv82 = HANCBlock(f * 2, f * 2, k=3, inv_fctr=3)
        self.up9 = nn.ConvTranspose2d(f * 2, f, kernel_size=(2, 2), stride=2)
        self.cnv91 = HANCBlock(f + f, f, k=3, inv_fctr=3)
        self.cnv92 = HANCBlock(f, f, k=3, inv_fctr=3)
        if n_classes == 1:
            self.out = nn.Conv2d(f, n_classes, kernel_size=(1, 1))
            self.last_activation = nn.Sigmoid() if self.sigmoid_out else None
        else:
            self.out = nn.Conv2d(f, n_classes + 1, kernel_size=(1, 1))
            self.last_activation = None

    def allreduce_phases(self):
        """Data-parallel training: [(trigger, modules)] -- when the backward of `trigger` (a module, or "group" = the first
        group of parallel chains) starts, every parameter gradient of `modules` is final.  Backward runs decoder ->
        MLFC -> (ResPaths || encoder), and the parameters of each part are contiguous in parameter order, so TrainStep
        all-reduces the decoder's 60 % of the gradient buffer under the MLFC + encoder backward and MLFC's under the
        encoder's (ACC_UNet.py:620-631 read backwards)."""
        dec = [self.up6, self.cnv61, self.cnv62, self.up7, self.cnv71, self.cnv72, self.up8, self.cnv81, self.cnv82,
               self.up9, self.cnv91, self.cnv92, self.out]
        return [(self.mlfc3, dec), ("group", [self.mlfc1, self.mlfc2, self.mlfc3])]

    def forward(self, x):
        E.require_cuda(x)
        cd = self.compute_dtype or x.dtype
        if x.dtype != cd or not x.permute(0, 2, 3, 1).is_contiguous():
            x = E.to_nchw_view(E.input_to_nhwc(x, cd)) if not x.requires_grad else \
                x.to(cd).contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(cd == torch.bfloat16)):
            # Same dataflow as ACC_UNet.forward (ACC_UNet.py:605-631), issued so that independent chains overlap: the
            # ResPath of level l only needs that level's encoder output, so it runs (forward AND backward, one
            # autograd node per pair) on a parallel stream lane next to the encoder blocks of level l+1.
            e2 = self.cnv12(self.cnv11(x))
            x2, e3 = run_parallel([self.rspth1, [self.cnv21, self.cnv22]], [e2, maxpool2(e2)])
            x3, e4 = run_parallel([self.rspth2, [self.cnv31, self.cnv32]], [e3, maxpool2(e3)])
            x4, e5 = run_parallel([self.rspth3, [self.cnv41, self.cnv42]], [e4, maxpool2(e4)])
            x5, x6 = run_parallel([self.rspth4, [self.cnv51, self.cnv52]], [e5, maxpool2(e5)])
            x2, x3, x4, x5 = self.mlfc1(x2, x3, x4, x5)
            x2, x3, x4, x5 = self.mlfc2(x2, x3, x4, x5)
            x2, x3, x4, x5 = self.mlfc3(x2, x3, x4, x5)
            x7 = self.cnv62(self.cnv61(up_cat(x6, x5, self.up6)))
            x8 = self.cnv72(self.cnv71(up_cat(x7, x4, self.up7)))
            x9 = self.cnv82(self.cnv81(up_cat(x8, x3, self.up8)))
            x10 = self.cnv92(self.cnv91(up_cat(x9, x2, self.up9)))
            logits = out_conv(x10, self.out)                       # fp32 logits
        if self.last_activation is not None:
            logits = self.last_activation(logits)
        return logits


class ACC_UNet(_ACCUNetBase):
    """ACC_UNet/ACC_UNet.py:530-659"""
    variant = "base"


class ACC_UNet_W(_ACCUNetBase):
    """ACC_UNet/ACC_UNet_w.py: MLFC merge blended by a learned scalar W"""
    variant = "w"


class ACC_UNet_Lite(_ACCUNetBase):
    """ACC_UNet/ACC_UNet_lite.py: MLFC reduced to its four SE layers (its conv parameters stay unused)"""
    variant = "lite"


class ACC_UNet_Harness(_ACCUNetBase):
    """The flavour the reference's TRAINING HARNESS builds (`from nets.ACC_UNet import ACC_UNet`,
    Experiments/train_model.py:24): Experiments/nets/ACC_UNet.py differs from the canonical model file in two
    places -- cnv72 is built with inv_fctr=3 (:584; conv1.weight [384, 128, 1, 1] instead of [4352, 128, 1, 1]) and
    forward returns LOGITS (:596-597,655; the harness loss is logit based).  Checkpoints written by train_model.py
    load into this class, not into ACC_UNet."""
    variant = "base"
    cnv72_inv_fctr = 3
    sigmoid_out = False
