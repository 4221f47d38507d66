"""Norm factory with the reference's model/common.py:4-10 interface (BN branch; IN is out of scope)."""
from ..nn import MinkowskiBatchNorm


def get_norm(norm_type, num_feats, bn_momentum=0.05, D=-1):
    if norm_type == 'BN':
        return MinkowskiBatchNorm(num_feats, momentum=bn_momentum)
    elif norm_type == 'IN':
        raise NotImplementedError('InstanceNorm variants are not on the inference hot path (no shipped config uses them)')
    else:
        raise ValueError(f'Type {norm_type}, not defined')
