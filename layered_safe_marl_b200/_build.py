"""Build the sm_100a shared library in-tree (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'liblsm_b200.so')
SOURCES = [os.path.join(CSRC, 'lsm_kernels.cu'), os.path.join(CSRC, 'lsm_capi.cu')]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cuh', '.h'))] + \
    [os.path.join(os.path.dirname(HERE), 'include', f) for f in ('lsm_b200.h', 'lsm_math.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              # no FMA contraction: float64 intermediates must round like the reference's numpy arithmetic
              '-fmad=false',
              # host code too (lsm_math.h on the host must round like the device and the oracle build)
              '-Xcompiler', '-ffp-contract=off',
              '-Xcompiler', '-fPIC', '-shared']


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in DEPS)


EXP_LIB_PATH = os.path.join(HERE, 'liblsm_b200_exp.so')


def build_experiments(extra_flags=(), verbose: bool = False) -> str:
    """The lab build of the same sources: -DLSM_EXPERIMENTS compiles the ablation switches (LSM_DEBUG, LSM_WPE, LSM_PAIR,
    LSM_AGENT_MINB, ... read from the environment) and the alternative kernel instantiations that the product library
    does not contain. Tools select it with LSM_LIB=<path>."""
    build(force=True, verbose=verbose, extra_flags=['-DLSM_EXPERIMENTS'] + list(extra_flags), out_path=EXP_LIB_PATH)
    return EXP_LIB_PATH


def build(force: bool = False, verbose: bool = False, extra_flags=(), out_path: str = LIB_PATH) -> str:
    """`extra_flags` / `out_path` build an A/B variant (extra -D switches) next to the product library; a variant
    is selected at run time with LSM_LIB=<path> (experiments only)."""
    if not force and not needs_build() and out_path == LIB_PATH:
        return LIB_PATH
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (['-Xptxas', '-v'] if verbose else []) + ['-o', out_path] + SOURCES
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout)
    if verbose:
        print(proc.stdout)
    return out_path
