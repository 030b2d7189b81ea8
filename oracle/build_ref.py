"""TEST INFRASTRUCTURE -- builds oracle/_ref/ from the reference's own sources where they lie under /root/reference.

The only compiled code of the reference near the path is its CPU marching-cubes extractor
(third_parties/coslam/external/NumpyMarchingCubes/marching_cubes/src/marching_cubes.cpp, used by save_mesh through
src/slam/coslam/coslam_utils.py:145).  Its own build is a Cython extension; here the one algorithm file is compiled
directly with g++ (the extension's flags: -std=c++11, setup.py:47) behind oracle/mc_ref_shim.cpp.  Outputs go to
oracle/_ref/ only (git-ignored, travels to the GPU box with the snapshot).  Build container only."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/third_parties/coslam/external/NumpyMarchingCubes/marching_cubes/src'
OUT = os.path.join(HERE, '_ref', 'libmc_ref.so')


def reference_present():
    return os.path.exists(os.path.join(REF_SRC, 'marching_cubes.cpp'))


def build(force=False):
    if not reference_present():
        return OUT if os.path.exists(OUT) else None
    src = os.path.join(HERE, 'mc_ref_shim.cpp')
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ['g++', '-O2', '-std=c++11', '-w', '-shared', '-fPIC', '-I', REF_SRC, src, '-o', OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
