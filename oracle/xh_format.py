"""TEST INFRASTRUCTURE (see oracle/__init__.py): numpy restatement of the split-half activation format of the fp16
tensor-core convolution (include/eyoc_b200.h, csrc/sparse_conv_h.cu xh_split / xh_join).

A row of C channels (C % 32 == 0) is C / 32 chunks of 64 fp16 values: 32 "hi" values followed by 32 "lo'" values with
    hi = fp16(x)            (round to nearest even)
    lo' = fp16((x - hi) * 2^11)
    x  ~= hi + lo' * 2^-11  (evaluated as one fp32 FMA)
This is not part of the reference (MinkowskiEngine keeps fp32 features); it is the internal layout the kernels exchange
between layers, restated here so that the tests can pin the kernel's pack / unpack bit for bit and bound its error.
"""
import numpy as np

LO_SCALE = np.float32(2048.0)
LO_INV = np.float32(1.0 / 2048.0)


def pack(x):
    """fp32 [n, c] -> fp16 [n, 2 c] (per 32-channel chunk: 32 hi | 32 lo')."""
    x = np.asarray(x, np.float32)
    n, c = x.shape
    assert c % 32 == 0
    with np.errstate(over='ignore', invalid='ignore'):
        hi = x.astype(np.float16)
        lo = ((x - hi.astype(np.float32)) * LO_SCALE).astype(np.float16)
    out = np.empty((n, c // 32, 2, 32), np.float16)
    out[:, :, 0] = hi.reshape(n, c // 32, 32)
    out[:, :, 1] = lo.reshape(n, c // 32, 32)
    return out.reshape(n, 2 * c)


def unpack(xh):
    """fp16 [n, 2 c] -> fp32 [n, c]:  fma(lo', 2^-11, hi) in fp32 (exact: hi and lo' 2^-11 do not overlap)."""
    xh = np.asarray(xh, np.float16)
    n, c2 = xh.shape
    v = xh.reshape(n, c2 // 64, 2, 32).astype(np.float32)
    # lo' * 2^-11 is exact in fp32 and the sum has at most 22 significant bits: plain fp32 arithmetic equals the kernel's FMA
    return (v[:, :, 1] * LO_INV + v[:, :, 0]).reshape(n, c2 // 2)
