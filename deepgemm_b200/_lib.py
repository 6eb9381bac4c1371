"""ctypes binding of the dgb200 C ABI (include/dgb200.h).

The CUDA library is the product: there is no CPU or eager-PyTorch fallback. If ``libdgb200.so`` is missing
or a call fails, the error is raised to the caller (the reference raises RuntimeError from DG_HOST_ASSERT,
csrc/utils/exception.hpp:12-40).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
LIB_PATH = os.environ.get('DGB200_LIB') or os.path.join(_HERE, 'lib', 'libdgb200.so')   # env: development only
_CSRC = os.path.join(_HERE, 'csrc')
# translation units (host API first, then the kernel-instance units) and every header they include: a change to any
# header rebuilds everything, a change to one unit rebuilds that unit
UNITS = [os.path.join(_CSRC, 'dgb200_api.cu')] + sorted(
    os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.startswith('gemm_') and f.endswith('.cu'))
SOURCES = UNITS + sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(('.cuh', '.h'))) + \
    [os.path.join(_REPO, 'include', 'dgb200.h')]

COMPILE_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
                 '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC']
NVCC_FLAGS = COMPILE_FLAGS + ['-shared']   # (kept for tools that compile a single file)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the AOT library for sm_100a with nvcc (cross-compiles without a GPU). The kernel instances live in
    several translation units (csrc/gemm_*.cu) that are compiled in parallel, then linked into one shared object."""
    from concurrent.futures import ThreadPoolExecutor
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(s) for s in SOURCES if os.path.exists(s))
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    nvcc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'bin', 'nvcc')
    obj_dir = os.path.join(os.path.dirname(LIB_PATH), 'obj')
    os.makedirs(obj_dir, exist_ok=True)
    headers = [s for s in SOURCES if not s.endswith('.cu')]
    newest_header = max(os.path.getmtime(h) for h in headers)

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + '.o')
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), newest_header):
            return obj
        cmd = [nvcc] + COMPILE_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'nvcc failed:\n{" ".join(cmd)}\n{res.stdout}\n{res.stderr}')
        if verbose:
            print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, UNITS))
    cmd = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-lcudart']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'link failed:\n{" ".join(cmd)}\n{res.stdout}\n{res.stderr}')
    return LIB_PATH


class _Config(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ('block_m', 'cluster', 'num_stages', 'num_sms', 'smem_bytes', 'num_tiles',
                                            'num_splits', 'cluster_split', 'tma_store', 'swap_ab')]


_P, _I, _L = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64

# name -> (restype, argtypes); must list every symbol include/dgb200.h declares (tests check this)
SIGNATURES = {
    'dgb200_last_error': (ctypes.c_char_p, []),
    'dgb200_version': (_I, []),
    'dgb200_set_num_sms': (_I, [_I]),
    'dgb200_get_num_sms': (_I, []),
    'dgb200_set_tc_util': (_I, [_I]),
    'dgb200_get_tc_util': (_I, []),
    'dgb200_set_pdl': (_I, [_I]),
    'dgb200_get_pdl': (_I, []),
    'dgb200_set_split_k': (_I, [_I]),
    'dgb200_get_split_k': (_I, []),
    'dgb200_set_mk_alignment_for_contiguous_layout': (_I, [_I]),
    'dgb200_get_mk_alignment_for_contiguous_layout': (_I, []),
    'dgb200_get_theoretical_mk_alignment_for_contiguous_layout': (_I, [_I]),
    'dgb200_get_tma_aligned_size': (_I, [_I, _I]),
    'dgb200_pack_sf_ue8m0': (_I, [_P, _P, _I, _I, _I, _I, _L, _L, _L, _P, _I, _I, _P]),
    'dgb200_transpose_sf_fp32': (_I, [_P, _P, _I, _I, _I, _L, _L, _L, _P]),
    'dgb200_pack_sf_ue8m0_k_grouped': (_I, [_P, _P, _I, _P, _I, _I, _I, _I, _P]),
    'dgb200_fp8_gemm_nt': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _L, _L, _L, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _P]),
    'dgb200_workspace_bytes': (_L, [_I, _I]),
    'dgb200_fp8_gemm_nt_skip_head_mid': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _L, _L, _L, _I, _I, _I, _I, _I, _I, _P]),
    'dgb200_fp8_bmm': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _L, _L, _L, _L, _L, _L, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'dgb200_per_token_cast_to_fp8': (_I, [_P, _L, _P, _L, _P, _I, _I, _I, _I, _P]),
    'dgb200_bf16_gemm_nt': (_I, [_P, _P, _P, _I, _I, _I, _L, _L, _L, _I, _I, _I, _I, _P]),
    'dgb200_m_grouped_bf16_gemm_nt_contiguous': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _L, _L, _L, _I, _I, _I, _I, _P]),
    'dgb200_bf16_bmk_bnk_mn': (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    'dgb200_bf16_bmm': (_I, [_P, _P, _P, _I, _I, _I, _I, _L, _L, _L, _L, _L, _L, _I, _P]),
    'dgb200_k_grouped_bf16_gemm_tn_contiguous': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'dgb200_m_grouped_bf16_gemm_nt_masked': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'dgb200_debug_fp8_peak': (_I, [_I, _I, _I, _P]),
    'dgb200_m_grouped_fp8_gemm_nt_contiguous': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _L, _L, _L, _I, _I, _I,
                                                     _I, _I, _I, _I, _I, _P]),
    'dgb200_m_grouped_fp8_gemm_nt_masked': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'dgb200_k_grouped_fp8_gemm_tn_contiguous': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    'dgb200_plan': (_I, [_I, _I, _I, _I, _I, _I, _I, _I, ctypes.POINTER(_Config)]),
    'dgb200_last_config': (_I, [ctypes.POINTER(_Config)]),
    'dgb200_debug_set_timestamps': (_I, [_P]),
    'dgb200_launch_count': (_L, []),
    'dgb200_ep_buffer_bytes': (_L, [_I, _I, _I, _I]),
    'dgb200_ep_buffer_offsets': (_I, [_I, _I, _I, _I, ctypes.POINTER(_L)]),
    'dgb200_ep_alloc': (_I, [_L, ctypes.POINTER(_P)]),
    'dgb200_ep_free': (_I, [_P]),
    'dgb200_ep_export': (_I, [_P, _P]),
    'dgb200_ep_import': (_I, [_P, ctypes.POINTER(_P)]),
    'dgb200_ep_unimport': (_I, [_P]),
    'dgb200_ep_dispatch': (_I, [_P, _L, _P, _L, _L, _P, _I, _I, _I, _I, _I, _I, _I, ctypes.POINTER(_P), _I, _I, _P, _P, _I, _P]),
    'dgb200_ep_combine': (_I, [_P, _L, _P, _P, _I, _I, _I, _P, _I, _I, _I, _I, _I, ctypes.POINTER(_P), ctypes.POINTER(_P), _L, _P]),
    'dgb200_ep_grouped_gemm': (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _I, _L, _L, _I, _I, _I, _I, _I, _P]),
}

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                               f'(there is no CPU fallback for the FP8 GEMM path)')
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise RuntimeError(lib().dgb200_last_error().decode())


def last_config() -> dict:
    cfg = _Config()
    check(lib().dgb200_last_config(ctypes.byref(cfg)))
    return {n: getattr(cfg, n) for n, _ in _Config._fields_}


def plan(gemm_type: int, m: int, n: int, k: int, num_groups: int = 1, expected_m: int = 0, alignment: int = 1,
         num_sms: int = 148) -> dict:
    cfg = _Config()
    check(lib().dgb200_plan(gemm_type, m, n, k, num_groups, expected_m, alignment, num_sms, ctypes.byref(cfg)))
    return {n_: getattr(cfg, n_) for n_, _ in _Config._fields_}


def launch_count() -> int:
    return int(lib().dgb200_launch_count())
