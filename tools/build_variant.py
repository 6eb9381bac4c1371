"""Build a variant of the library with extra compile flags into deepgemm_b200/lib/<name>.so (development: A/B experiments on the
GPU box select it with DGB200_LIB). usage: build_variant.py <name> <flag>..."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepgemm_b200 import _lib  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
lib_dir = os.path.dirname(_lib.LIB_PATH)
obj_dir = os.path.join(lib_dir, 'obj_' + name)
os.makedirs(obj_dir, exist_ok=True)
nvcc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'bin', 'nvcc')


def compile_one(src):
    obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + '.o')
    subprocess.run([nvcc] + _lib.COMPILE_FLAGS + flags + ['-c', src, '-o', obj], check=True)
    return obj


with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as pool:
    objs = list(pool.map(compile_one, _lib.UNITS))
out = os.path.join(lib_dir, name + '.so')
subprocess.run([nvcc, '-shared', '-o', out] + objs + ['-lcudart'], check=True)
print(out)
