"""ctypes binding of libtaseg_b200.so — the C ABI declared in include/taseg_b200.h.

The prototypes are parsed from the header itself, so the Python side cannot drift from the contract.
There is NO fallback: if the shared library is missing the first call raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "taseg_b200.h")
LIB_PATH = os.environ.get("TSG_LIB") or os.path.join(HERE, "libtaseg_b200.so")   # TSG_LIB: profiling build

TSG_OK, TSG_ERR_INVALID, TSG_ERR_CUDA, TSG_ERR_WORKSPACE, TSG_ERR_RANGE, TSG_ERR_UNSUPPORTED = range(6)
TSG_F32, TSG_BF16, TSG_F16 = 0, 1, 2
DTYPES = {torch.float32: TSG_F32, torch.bfloat16: TSG_BF16, torch.float16: TSG_F16}

_SCALARS = {"uint64_t": ctypes.c_uint64, "int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "size_t": ctypes.c_size_t,
            "float": ctypes.c_float, "uint32_t": ctypes.c_uint32, "tsg_stream_t": ctypes.c_void_p}


class Frame(ctypes.Structure):
    """tsg_frame (include/taseg_b200.h)."""
    _fields_ = [("offset", ctypes.c_int64), ("count", ctypes.c_int64), ("sample", ctypes.c_int32),
                ("is_cur", ctypes.c_int32), ("pose0", ctypes.c_float * 16), ("pose", ctypes.c_float * 16)]


class Sweep(ctypes.Structure):
    """tsg_sweep (include/taseg_b200.h)."""
    _fields_ = [("offset", ctypes.c_int64), ("count", ctypes.c_int64), ("sample", ctypes.c_int32), ("is_key", ctypes.c_int32),
                ("R", ctypes.c_double * 9), ("T", ctypes.c_double * 3), ("dt", ctypes.c_float), ("pad_", ctypes.c_float)]


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """name -> (restype, argtypes) for every function declared in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"^\s*(const char \*|int64_t|size_t|int)\s*(tsg_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.M | re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3)
        restype = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "size_t": ctypes.c_size_t,
                   "const char *": ctypes.c_char_p}[ret]
        argtypes = []
        for a in [x.strip() for x in args.split(",")]:
            if a in ("void", ""):
                continue
            if "*" in a:
                argtypes.append(ctypes.c_void_p)
            else:
                ty = a.replace("const ", "").split()[0]
                argtypes.append(_SCALARS[ty])
        protos[name] = (restype, argtypes)
    return protos


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m taseg_b200.build` (there is no CPU or eager fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in parse_header().items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = restype, argtypes
    return _lib


def ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return ctypes.c_void_p(t.data_ptr())
    return t


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream() -> ctypes.c_void_p:
    """torch's current CUDA stream of the current device as a raw cudaStream_t (every kernel is launched on it)."""
    if _raw_stream is not None:      # one C call instead of building a torch.cuda.Stream object per launch
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(status: int, what: str = "") -> None:
    if status == TSG_OK:
        return
    msg = lib().tsg_last_error().decode()
    if status == TSG_ERR_INVALID:
        raise ValueError(msg or what)       # the reference throws std::invalid_argument -> ValueError
    raise RuntimeError(f"{what}: status {status}: {msg}")


# kernels launched per C-ABI call (for the benchmark's gpu_launches claim; conservative: sorts launch 1 histogram kernel
# + one kernel per 9-bit pass, counted here at their minimum for the key widths of the benchmark)
KERNELS_PER_CALL = {"tsg_table_build": 2, "tsg_coord_table_build": 2, "tsg_kmap_pairs": 2, "tsg_kmap_transpose": 2,
                    "tsg_sort_pairs": 4, "tsg_unique_coords": 7, "tsg_unique_hash": 11, "tsg_aggregate_quantize": 4,
                    "tsg_compact_rows": 3, "tsg_kmap_sort_rows": 5, "tsg_aggregate_quantize_nus": 4, "tsg_aggregate_quantize_dev": 4,
                    "tsg_unique_coords_dev": 7, "tsg_coord_table_build_dev": 2, "tsg_kmap_sort_rows_dev": 4, "tsg_kmap_sort_rows_dev2": 3,
                    "tsg_kmap_transpose_dev": 2, "tsg_voxelize_plan": 5, "tsg_kmap_pair_list": 4, "tsg_bn_stats": 2, "tsg_bn_stats2": 2, "tsg_bn_backward": 3}
launch_count = 0


def call(name: str, *args) -> None:
    global launch_count
    launch_count += KERNELS_PER_CALL.get(name, 1)
    check(getattr(lib(), name)(*args), name)


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("taseg_b200 runs on CUDA tensors only (no CPU fallback); got a tensor on " + str(t.device))
