"""Build libeyoc_b200.so in-tree with nvcc for sm_100a (no torch headers: plain C ABI).

    python -m eyoc_b200.csrc.build [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ['capi.cu', 'knn.cu', 'sc2pcr.cu', 'coordmap.cu', 'sparse_conv.cu', 'sparse_conv_tc.cu', 'sparse_conv_h.cu', 'irls.cu', 'host_plan.cu', 'gather4_probe.cu', 'instance_norm.cu']
LIB = os.path.join(os.path.dirname(HERE), 'libeyoc_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--fmad=true']


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.cuh', '.h'))]
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for s in srcs:
        src = os.path.join(HERE, s)
        obj = os.path.join(objdir, s.replace('.cu', '.o'))
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        r = subprocess.run([NVCC, *FLAGS, '-c', src, '-o', obj], capture_output=True, text=True, cwd=HERE)
        log = obj.replace('.o', '.ptxas.log')
        with open(log, 'w') as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(compile_one, jobs))
    objs = [os.path.join(objdir, s.replace('.cu', '.o')) for s in srcs]
    if force or jobs or _stale(LIB, objs):
        r = subprocess.run([NVCC, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB, *objs, '-lcudart', '-Xcompiler', '-pthread'],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
