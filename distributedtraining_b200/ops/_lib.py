"""ctypes binding of ``build/libdtb200.so`` (the sm_100a kernel library).

Policy (the driver checks which ``.so`` files the GPU tests load):
  * CUDA available and library present  -> kernels are used (``have_kernels() == True``)
  * CUDA available and library missing  -> **loud failure** (no silent eager fallback on a GPU box)
  * no CUDA (CPU dev box / gloo tests)  -> PyTorch reference implementations in :mod:`ops.reference`
"""
from __future__ import annotations

import ctypes
import os
from functools import lru_cache

import torch

from .build import LIB_PATH


class KernelLibraryMissing(RuntimeError):
    pass


@lru_cache(maxsize=1)
def lib() -> ctypes.CDLL:
    alt = os.environ.get("DTB200_LIB")  # A/B harness: load another build of the same C ABI
    if alt:
        return ctypes.CDLL(alt, mode=ctypes.RTLD_GLOBAL)
    if not LIB_PATH.exists():
        raise KernelLibraryMissing(
            f"{LIB_PATH} not found: run `python -m distributedtraining_b200.ops.build` (or __graft_entry__.build()) first")
    return ctypes.CDLL(str(LIB_PATH), mode=ctypes.RTLD_GLOBAL)


def force_reference() -> bool:
    return os.environ.get("DTB200_FORCE_REFERENCE", "0") == "1"


def have_kernels() -> bool:
    """True iff the hand-written kernels must be used for CUDA tensors."""
    if force_reference() or not torch.cuda.is_available():
        return False
    lib()  # raises loudly on a GPU box without the library
    return True


def use_kernels(t: torch.Tensor) -> bool:
    return t.is_cuda and have_kernels()


def stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: torch.Tensor | None) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


@lru_cache(maxsize=8)
def num_sms(device_index: int = -1) -> int:
    if device_index < 0:
        device_index = torch.cuda.current_device()
    return torch.cuda.get_device_properties(device_index).multi_processor_count


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed with code {rc}")
