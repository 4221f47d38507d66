"""FCGF ResUNet feature extractor with the reference's model/resunet.py interface, on sm_100a kernels.

``ResUNet2`` mirrors model/resunet.py:10-193 (same constructor, same sub-module names, MinkowskiEngine
state-dict keys), and ``ResUNetBN2C`` (:206-209) is the default model of the reference (config.py:82).
``forward`` issues 23 fused launches (conv + folded BN + residual + ReLU, skip concatenation read in
place, final L2 normalisation in the last epilogue) instead of ~90 separate MinkowskiEngine / torch ops.
"""
import torch
import torch.nn as nn

from .. import nn as enn
from ..nn import MinkowskiConvolution, MinkowskiConvolutionTranspose, conv_bn_act
from ..sparse import SparseTensor
from .common import get_norm
from .residual_block import get_block  # noqa: F401  (ResUNetExpanded builds its second blocks with it)


class ResUNet2(nn.Module):
    NORM_TYPE = None
    BLOCK_NORM_TYPE = 'BN'
    CHANNELS = [None, 32, 64, 128, 256]
    TR_CHANNELS = [None, 32, 64, 64, 128]

    def __init__(self, in_channels=3, out_channels=32, bn_momentum=0.1, normalize_feature=None, conv1_kernel_size=None,
                 D=3):
        super().__init__()
        self.D = D
        NORM_TYPE, BLOCK_NORM_TYPE = self.NORM_TYPE, self.BLOCK_NORM_TYPE
        CHANNELS, TR_CHANNELS = self.CHANNELS, self.TR_CHANNELS
        self.normalize_feature = normalize_feature

        def conv(cin, cout, k, stride=1, bias=False):
            return MinkowskiConvolution(cin, cout, kernel_size=k, stride=stride, dilation=1, bias=bias, dimension=D)

        def conv_tr(cin, cout):
            return MinkowskiConvolutionTranspose(cin, cout, kernel_size=3, stride=2, dilation=1, bias=False, dimension=D)

        def norm(c):
            return get_norm(NORM_TYPE, c, bn_momentum=bn_momentum, D=D)

        def block(c):
            return get_block(BLOCK_NORM_TYPE, c, c, bn_momentum=bn_momentum, D=D)

        self.conv1 = conv(in_channels, CHANNELS[1], conv1_kernel_size)
        self.norm1 = norm(CHANNELS[1])
        self.block1 = block(CHANNELS[1])
        self.conv2 = conv(CHANNELS[1], CHANNELS[2], 3, stride=2)
        self.norm2 = norm(CHANNELS[2])
        self.block2 = block(CHANNELS[2])
        self.conv3 = conv(CHANNELS[2], CHANNELS[3], 3, stride=2)
        self.norm3 = norm(CHANNELS[3])
        self.block3 = block(CHANNELS[3])
        self.conv4 = conv(CHANNELS[3], CHANNELS[4], 3, stride=2)
        self.norm4 = norm(CHANNELS[4])
        self.block4 = block(CHANNELS[4])
        self.conv4_tr = conv_tr(CHANNELS[4], TR_CHANNELS[4])
        self.norm4_tr = norm(TR_CHANNELS[4])
        self.block4_tr = block(TR_CHANNELS[4])
        self.conv3_tr = conv_tr(CHANNELS[3] + TR_CHANNELS[4], TR_CHANNELS[3])
        self.norm3_tr = norm(TR_CHANNELS[3])
        self.block3_tr = block(TR_CHANNELS[3])
        self.conv2_tr = conv_tr(CHANNELS[2] + TR_CHANNELS[3], TR_CHANNELS[2])
        self.norm2_tr = norm(TR_CHANNELS[2])
        self.block2_tr = block(TR_CHANNELS[2])
        self.conv1_tr = conv(CHANNELS[1] + TR_CHANNELS[2], TR_CHANNELS[1], 1)
        self.final = conv(TR_CHANNELS[1], out_channels, 1, bias=True)

    # The default data path keeps activations as fp16 hi/lo pairs (|x| < 65504).  The kernels raise a device flag when a value
    # leaves that range (the fp32 reference would carry on); forward() reads it once per call (one 4-byte D2H) and, if set,
    # runs the network again through the fp32-activation tensor-core path ('tf32x3').  Set False to skip the read.
    RANGE_CHECK = True

    def forward(self, x):
        """model/resunet.py:142-193 (see _forward) + the split-half range guard."""
        if self.training:
            raise NotImplementedError('eyoc_b200 implements the inference path only: call model.eval()')
        out = self._forward(x)
        if self.RANGE_CHECK and enn.CONV_MODE == 'f16x3':
            mgr = x.coordinate_manager
            if int(mgr.range_status.item()) & 1:
                out = self.forward_fp32_activations(x)
        return out

    def forward_fp32_activations(self, x):
        """The fallback of the range guard: the same forward pass through the fp32-activation tensor-core path."""
        import warnings
        warnings.warn('activations beyond the fp16 hi/lo range (|x| >= 65504 or non-finite): re-running this forward '
                      "pass with fp32 activations (CONV_MODE 'tf32x3')")
        x.coordinate_manager.range_status.zero_()
        enn.CONV_MODE = 'tf32x3'
        try:
            return self._forward(x)
        finally:
            enn.CONV_MODE = 'f16x3'

    # ---- the forward pass in two stages, for callers that overlap stage 1 of one block with other work (pipeline.py):
    #      stage 1 = everything bound by dependent-access latency (first convolution with its neighbour search, kernel maps, tile
    #      orders), stage 2 = the tensor-core convolutions.  forward(x) == trunk(x, prepare(x)).
    def prepare(self, x):
        if self.training:
            raise NotImplementedError('eyoc_b200 implements the inference path only: call model.eval()')
        with torch.no_grad():
            y1 = conv_bn_act(x, self.conv1, self.norm1)
        self.build_maps(x.coordinate_manager)
        return y1

    def build_maps(self, mgr):
        """Queue every kernel map (tile order + tile masks for the tensor-core convolutions) the trunk will ask for: 3^3 maps on
        strides 1, 2, 4, 8, the strided ones between them and their transposes (model/resunet.py:31-140)."""
        tiled = enn.TILE_ORDER and enn.CONV_MODE in ('f16x3', 'tf32x3')
        want = [(1, 1, False), (1, 2, False), (2, 2, False), (2, 4, False), (4, 4, False), (4, 8, False), (8, 8, False),
                (8, 4, True), (4, 2, True), (2, 1, True)]
        mgr.ensure_levels(8)
        for ts_in, ts_out, tr in want:
            if mgr.num_rows(ts_in) == 0 or mgr.num_rows(ts_out) == 0:
                continue
            if tiled:
                mgr.tiled_map(ts_in, ts_out, 3, transposed=tr)
                if enn.CONV_MODE == 'f16x3':
                    mgr.tile_masks(ts_in, ts_out, 3, transposed=tr)
            else:
                mgr.kernel_map(ts_in, ts_out, 3, transposed=tr)

    def trunk(self, x, y1):
        """Stage 2.  Returns (output, pending read of the fp16 range flag or None): the caller resolves the read when
        convenient and calls forward_fp32_activations(x) if bit 0 is set."""
        out = self._forward(x, y1)
        read = None
        if self.RANGE_CHECK and enn.CONV_MODE == 'f16x3':
            from ..sparse import _AsyncRead
            read = _AsyncRead(x.coordinate_manager.range_status)
        return out, read

    def _forward(self, x, y1=None):
        """model/resunet.py:142-193.  The MEF.relu calls after each block (:146,151,156,161,166,173,180) are
        idempotent (the block already ends in ReLU, residual_block.py:51) and therefore cost nothing here."""
        with torch.no_grad():
            out_s1 = self.block1(y1 if y1 is not None else conv_bn_act(x, self.conv1, self.norm1))
            out_s2 = self.block2(conv_bn_act(out_s1, self.conv2, self.norm2))
            out_s4 = self.block3(conv_bn_act(out_s2, self.conv3, self.norm3))
            out_s8 = self.block4(conv_bn_act(out_s4, self.conv4, self.norm4))

            out_s4_tr = self.block4_tr(conv_bn_act(out_s8, self.conv4_tr, self.norm4_tr))
            out_s2_tr = self.block3_tr(conv_bn_act(out_s4_tr, self.conv3_tr, self.norm3_tr, skip=out_s4))
            out_s1_tr = self.block2_tr(conv_bn_act(out_s2_tr, self.conv2_tr, self.norm2_tr, skip=out_s2))

            out = conv_bn_act(out_s1_tr, self.conv1_tr, relu=True, skip=out_s1)
            # final 1x1 conv + bias; L2 normalisation (no epsilon, :189) fused into its epilogue
            return conv_bn_act(out, self.final, l2norm=bool(self.normalize_feature))


class ResUNetBN2(ResUNet2):
    NORM_TYPE = 'BN'


class ResUNetBN2B(ResUNet2):
    NORM_TYPE = 'BN'
    CHANNELS = [None, 32, 64, 128, 256]
    TR_CHANNELS = [None, 64, 64, 64, 64]


class ResUNetBN2C(ResUNet2):
    NORM_TYPE = 'BN'
    CHANNELS = [None, 32, 64, 128, 256]
    TR_CHANNELS = [None, 64, 64, 64, 128]


class ResUNetBN2D(ResUNet2):
    NORM_TYPE = 'BN'
    CHANNELS = [None, 32, 64, 128, 256]
    TR_CHANNELS = [None, 64, 64, 128, 128]


class ResUNetBN2E(ResUNet2):
    NORM_TYPE = 'BN'
    CHANNELS = [None, 128, 128, 128, 256]
    TR_CHANNELS = [None, 64, 128, 128, 128]


class ResUNetFatBN(ResUNet2):
    """model/resunet.py:224-227 (used by scripts/test_waymo.sh:13)."""
    NORM_TYPE = 'BN'
    CHANNELS = [None, 32, 64, 128, 256]
    TR_CHANNELS = [None, 128, 128, 128, 256]


class ResUNetIN2(ResUNet2):
    """model/resunet.py:229-232: batch norm after the strided / transposed convolutions, instance norm inside the blocks."""
    NORM_TYPE = 'BN'
    BLOCK_NORM_TYPE = 'IN'


class ResUNetIN2B(ResUNetBN2B):
    NORM_TYPE = 'BN'
    BLOCK_NORM_TYPE = 'IN'


class ResUNetIN2C(ResUNetBN2C):
    NORM_TYPE = 'BN'
    BLOCK_NORM_TYPE = 'IN'


class ResUNetIN2D(ResUNetBN2D):
    NORM_TYPE = 'BN'
    BLOCK_NORM_TYPE = 'IN'


class ResUNetIN2E(ResUNetBN2E):
    NORM_TYPE = 'BN'
    BLOCK_NORM_TYPE = 'IN'


class ResUNetExpanded(ResUNet2):
    """model/resunet.py:254-486: every level runs two residual blocks with a stand-alone norm in between
    (norm -> block -> ReLU -> norm_2 -> block_2 -> ReLU).  Same sub-module names / state-dict keys as the reference."""
    NORM_TYPE = None
    BLOCK_NORM_TYPE = 'BN'
    CHANNELS = [None, 32, 64, 128, 256]
    TR_CHANNELS = [None, 32, 64, 64, 128]

    def __init__(self, in_channels=3, out_channels=32, bn_momentum=0.1, normalize_feature=None, conv1_kernel_size=None, D=3):
        super().__init__(in_channels, out_channels, bn_momentum, normalize_feature, conv1_kernel_size, D)
        C, T = self.CHANNELS, self.TR_CHANNELS
        for name, c in (('1', C[1]), ('2', C[2]), ('3', C[3]), ('4', C[4]), ('4_tr', T[4]), ('3_tr', T[3]), ('2_tr', T[2])):
            setattr(self, f'norm{name}_2', get_norm(self.NORM_TYPE, c, bn_momentum=bn_momentum, D=D))
            setattr(self, f'block{name}_2', get_block(self.BLOCK_NORM_TYPE, c, c, bn_momentum=bn_momentum, D=D))

    def _level(self, x, name):
        """block -> (ReLU: idempotent) -> norm_2 -> block_2 (-> ReLU: idempotent), resunet.py:398-408."""
        x = getattr(self, f'block{name}')(x)
        return getattr(self, f'block{name}_2')(getattr(self, f'norm{name}_2')(x))

    def _forward(self, x, y1=None):
        """model/resunet.py:396-486."""
        with torch.no_grad():
            out_s1 = self._level(y1 if y1 is not None else conv_bn_act(x, self.conv1, self.norm1), '1')
            out_s2 = self._level(conv_bn_act(out_s1, self.conv2, self.norm2), '2')
            out_s4 = self._level(conv_bn_act(out_s2, self.conv3, self.norm3), '3')
            out_s8 = self._level(conv_bn_act(out_s4, self.conv4, self.norm4), '4')
            out_s4_tr = self._level(conv_bn_act(out_s8, self.conv4_tr, self.norm4_tr), '4_tr')
            out_s2_tr = self._level(conv_bn_act(out_s4_tr, self.conv3_tr, self.norm3_tr, skip=out_s4), '3_tr')
            out_s1_tr = self._level(conv_bn_act(out_s2_tr, self.conv2_tr, self.norm2_tr, skip=out_s2), '2_tr')
            out = conv_bn_act(out_s1_tr, self.conv1_tr, relu=True, skip=out_s1)
            return conv_bn_act(out, self.final, l2norm=bool(self.normalize_feature))


class ResUNetExpBN2C(ResUNetExpanded):
    """model/resunet.py:489-492."""
    NORM_TYPE = 'BN'
    CHANNELS = [None, 32, 64, 128, 256]
    TR_CHANNELS = [None, 64, 64, 64, 128]
