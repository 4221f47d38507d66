"""Wall-clock helpers with the reference's lib/timer.py:5-73 interface (Timer, AverageMeter).

Unlike the reference, ``Timer.toc`` can synchronise the CUDA device first (``sync=True``) so GPU
work is actually included; the default keeps the reference's behaviour.
"""
import time

import numpy as np
import torch


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val, self.avg, self.sum, self.sq_sum, self.count = 0, 0, 0.0, 0.0, 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count
        self.sq_sum += val ** 2 * n
        self.var = self.sq_sum / self.count - self.avg ** 2


class Timer(object):
    def __init__(self, binary_fn=None, init_val=0, sync=False):
        self.total_time, self.calls, self.start_time, self.diff, self.avg = 0., 0, 0., 0., 0.
        self.binary_fn, self.tmp, self.sync = binary_fn, init_val, sync

    def reset(self):
        self.total_time, self.calls, self.start_time, self.diff, self.avg = 0, 0, 0, 0, 0

    def tic(self):
        if self.sync and torch.cuda.is_available():
            torch.cuda.synchronize()
        self.start_time = time.time()

    def toc(self, average=True):
        if self.sync and torch.cuda.is_available():
            torch.cuda.synchronize()
        self.diff = time.time() - self.start_time
        self.total_time += self.diff
        self.calls += 1
        self.avg = self.total_time / self.calls
        if self.binary_fn:
            self.tmp = self.binary_fn(self.tmp, self.diff)
        return self.avg if average else self.diff
