"""ELKUNet: the LinK U-Net segmentation model (67.72 mIoU configuration of the reference README).

Same constructor kwargs, sub-module names and parameter shapes as the reference
(segmentation/core/models/semantic_kitti/linkunet.py:188-385).  The encoder half is shared with
ELKEncoder; here the decoder branches run: transposed 2^3/stride-2 convs (reusing the down convs'
kernel maps, conv.py:132-142), skip concatenation (`torchsparse.cat`), residual blocks with a 1x1
shortcut, and a Linear classifier.  The LinK blocks use the UNet phase rule (cos_x is NOT divided
by the tensor stride, linkunet.py:165)."""
import torch
import torch.nn as nn

import link_b200
from link_b200.linkencoder import _ELKBackbone
from link_b200.tensor import SparseTensor

__all__ = ['ELKUNet', 'LinKUNet']


class ELKUNet(_ELKBackbone):

    block_variant = 'unet'

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.classifier = nn.Sequential(nn.Linear(self.cs[8], kwargs['num_classes']))
        self.weight_initialization()

    def forward(self, x: SparseTensor) -> torch.Tensor:
        x0, x1, x2, x3, x4 = self.forward_levels(x)
        y = x4
        for up, skip in ((self.up1, x3), (self.up2, x2), (self.up3, x1), (self.up4, x0)):
            y = up[0](y)
            y = link_b200.cat([y, skip])
            y = up[1](y)
        return self.classifier(y.F)


LinKUNet = ELKUNet
