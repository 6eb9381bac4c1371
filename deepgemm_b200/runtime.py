"""Process-wide knobs (reference: csrc/apis/runtime.hpp:11-49 and csrc/apis/layout.hpp:142-150).

All state lives in the C library; these are thin forwards so that C callers and Python callers share it.
"""
from typing import Optional

from ._lib import check, lib


def set_num_sms(num_sms: int) -> None:
    check(lib().dgb200_set_num_sms(int(num_sms)))


def get_num_sms() -> int:
    n = lib().dgb200_get_num_sms()
    if n == 0:
        import torch
        n = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return n


def set_tc_util(tc_util: int) -> None:
    check(lib().dgb200_set_tc_util(int(tc_util)))


def get_tc_util() -> int:
    return lib().dgb200_get_tc_util()


def set_pdl(enabled: bool) -> None:
    check(lib().dgb200_set_pdl(int(bool(enabled))))


def get_pdl() -> bool:
    return bool(lib().dgb200_get_pdl())


def set_split_k(enabled: bool) -> None:
    """No reference equivalent: allow (default) or forbid cutting K of small dense problems into slices. Off, every
    output is bit-identical to the reference's SM100 kernel; on, small-M shapes agree to FP32 rounding and run faster."""
    check(lib().dgb200_set_split_k(int(bool(enabled))))


def get_split_k() -> bool:
    return bool(lib().dgb200_get_split_k())


def set_ignore_compile_dims(value: bool) -> None:
    """JIT hint in the reference (heuristics/runtime.hpp:18-24); shapes are always run-time values here."""


def set_block_size_multiple_of(block_m_multiple_of: int, block_n_multiple_of: int) -> None:
    """JIT search-space hint in the reference (heuristics/runtime.hpp:26-37); accepted and ignored."""


def set_mk_alignment_for_contiguous_layout(alignment: int) -> None:
    check(lib().dgb200_set_mk_alignment_for_contiguous_layout(int(alignment)))


def get_mk_alignment_for_contiguous_layout() -> int:
    return lib().dgb200_get_mk_alignment_for_contiguous_layout()


def get_theoretical_mk_alignment_for_contiguous_layout(expected_m: Optional[int] = None) -> int:
    return lib().dgb200_get_theoretical_mk_alignment_for_contiguous_layout(-1 if expected_m is None else int(expected_m))


def get_tma_aligned_size(x: int, element_size: int) -> int:
    r = lib().dgb200_get_tma_aligned_size(int(x), int(element_size))
    if r < 0:
        raise RuntimeError(f'Assertion error: 16 % element_size == 0 (element_size={element_size})')
    return r
