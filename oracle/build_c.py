"""Build the oracle's C helpers (TEST INFRASTRUCTURE) with gcc into oracle/_build/ (git-ignored; travels to the GPU box).

    python -m oracle.build_c
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'knn_oracle.c')
OUT = os.path.join(HERE, '_build', 'libknn_oracle.so')


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # -ffp-contract=off: no fusing beyond the explicit fmaf; no -ffast-math: IEEE semantics are the point
    subprocess.run(['gcc', '-O2', '-ffp-contract=off', '-fopenmp', '-shared', '-fPIC', SRC, '-o', OUT, '-lm'], check=True)
    return OUT


def load():
    """ctypes handle of the helper library, or None when gcc / the built file is unavailable (callers fall back to numpy)."""
    import ctypes
    try:
        lib = ctypes.CDLL(build())
    except Exception:      # noqa: BLE001
        return None
    lib.knn_sq_seq.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                               ctypes.c_void_p, ctypes.c_void_p]
    lib.knn_sq_seq.restype = None
    return lib


if __name__ == '__main__':
    print(build(force=True))
