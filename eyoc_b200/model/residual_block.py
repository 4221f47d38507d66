"""Residual block with the reference's model/residual_block.py:13-77 interface (BN and IN variants)."""
import torch.nn as nn

from ..nn import MinkowskiConvolution, conv_bn_act
from .common import get_norm


class BasicBlockBase(nn.Module):
    expansion = 1
    NORM_TYPE = 'BN'

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, D=3):
        super(BasicBlockBase, self).__init__()
        if stride != 1 or dilation != 1 or downsample is not None:
            raise NotImplementedError('only the stride-1 residual block of ResUNet2 is on the hot path')
        self.conv1 = MinkowskiConvolution(inplanes, planes, kernel_size=3, stride=stride, dimension=D)
        self.norm1 = get_norm(self.NORM_TYPE, planes, bn_momentum=bn_momentum, D=D)
        self.conv2 = MinkowskiConvolution(planes, planes, kernel_size=3, stride=1, dilation=dilation, bias=False, dimension=D)
        self.norm2 = get_norm(self.NORM_TYPE, planes, bn_momentum=bn_momentum, D=D)
        self.downsample = downsample

    def forward(self, x):
        """residual_block.py:37-53 as two fused launches: conv+BN+ReLU, conv+BN+residual+ReLU."""
        out = conv_bn_act(x, self.conv1, self.norm1, relu=True)
        return conv_bn_act(out, self.conv2, self.norm2, residual=x, relu=True)


class BasicBlockBN(BasicBlockBase):
    NORM_TYPE = 'BN'


class BasicBlockIN(BasicBlockBase):
    """model/residual_block.py:60-61: conv -> instance norm -> ReLU -> conv -> instance norm -> + x -> ReLU (conv_bn_act runs
    the convolution alone and fuses residual + ReLU into the normalisation pass)."""
    NORM_TYPE = 'IN'


def get_block(norm_type, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, D=3):
    if norm_type == 'BN':
        return BasicBlockBN(inplanes, planes, stride, dilation, downsample, bn_momentum, D)
    elif norm_type == 'IN':
        return BasicBlockIN(inplanes, planes, stride, dilation, downsample, bn_momentum, D)
    else:
        raise ValueError(f'Type {norm_type}, not defined')
