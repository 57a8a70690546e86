"""ctypes binding of libd3b.so — the C ABI declared in include/d3rlpy_b200.h.

Prototypes are parsed from the header so the Python side can never drift from the
declared ABI.  There is NO fallback: if the library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "d3rlpy_b200.h")
LIB_PATH = os.path.join(HERE, "libd3b.so")

_SCALARS = {
    "int": ctypes.c_int, "unsigned": ctypes.c_uint, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float, "double": ctypes.c_double,
}


class D3BError(RuntimeError):
    pass


_SYNC_EACH = os.environ.get("D3B_SYNC_EACH", "0") == "1"


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """Returns {name: (return_type, [(ctype, argname), ...])} for every d3b_* prototype."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(d3b_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.+?)\s*(\w+)$", a)
                parsed.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, parsed)
    return protos


def _ctype(t: str):
    if "*" in t:
        return ctypes.c_void_p
    t = t.replace("const", "").strip()
    return _SCALARS[t]


class Lib:
    dry = False
    _prof = None
    _FLOP_ARGS = {"linear_forward": (11, 12, 13, 14), "linear_backward_data": (12, 13, 14, 15),
                  "linear_backward_weight": (11, 12, 13, 14), "head_forward": (11, 12, 13, 14),
                  "head_backward_data": (12, 13, 14, 15), "head_backward_weight": (11, 12, 13, 14),
                  "umma_gemm": (6, 7, 8, 9), "umma_gemm_tn": (6, 7, 8, 9)}

    def start_profile(self, stream_obj) -> None:
        """Bracket every launch with CUDA events recorded on `stream_obj` (the launching stream)."""
        self._prof = []
        self._prof_stream = stream_obj

    def stop_profile(self):
        import torch

        torch.cuda.synchronize()
        recs = [(n, f, a.elapsed_time(b)) for n, f, a, b in self._prof]
        self._prof = None
        return recs

    def _profiled(self, name, fn, args):
        import torch

        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(self._prof_stream)
        rc = fn(*args)
        b.record(self._prof_stream)
        if rc != 0:
            raise D3BError(f"d3b_{name} failed ({rc}): {self.last_error()}")
        flops = 0.0
        if name in self._FLOP_ARGS:
            i = self._FLOP_ARGS[name]
            flops = 2.0 * args[i[0]] * args[i[1]] * args[i[2]] * args[i[3]]
        elif name == "umma_gemm_tn_batched":  # m[], n[] at 7, 8; k, members at 9, 10
            flops = 2.0 * args[9] * args[10] * sum(a * b for a, b in zip(list(args[7]), list(args[8])))
        elif name == "mlp_backward_bf16":  # rows, members, n_layers, dims[]; dgrad chain (+ head) (+ dx columns)
            dims = list(args[3])
            mac = sum(a * b for a, b in zip(dims[1:-1], dims[2:])) + dims[-1] * args[16]
            if args[22]:
                mac += dims[1] * args[26]
            flops = 2.0 * args[0] * args[1] * mac
        elif name == "mlp_forward_bf16":  # rows, members, n_layers, dims[] (+ head: n_head at 18)
            dims = list(args[6])
            mac = sum(a * b for a, b in zip(dims[:-1], dims[1:])) + dims[-1] * args[18]
            flops = 2.0 * args[3] * args[4] * mac
        self._prof.append((name, flops, a, b))
        return rc

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise D3BError(
                f"{path} not found: build it with `python -m d3rlpy_b200.build` (nvcc, sm_100a). "
                "d3rlpy_b200 has no CPU or PyTorch fallback.")
        self._dll = ctypes.CDLL(path)
        self.protos = parse_header()
        self._dll.d3b_last_error.restype = ctypes.c_char_p
        for name, (ret, args) in self.protos.items():
            fn = getattr(self._dll, name)  # AttributeError if the .so does not export a declared symbol
            if name == "d3b_last_error":
                continue
            fn.argtypes = [_ctype(t) for t, _ in args]
            fn.restype = ctypes.c_int64 if ret == "int64_t" else ctypes.c_int
        if self._dll.d3b_abi_version() != 1:
            raise D3BError("libd3b.so ABI version mismatch; rebuild")

    def last_error(self) -> str:
        return self._dll.d3b_last_error().decode()

    def raw(self, name: str):
        return getattr(self._dll, name)

    def __getattr__(self, name: str):
        fn = getattr(self._dll, "d3b_" + name)

        def call(*args):
            if self.dry:  # allocation-only pass before CUDA-graph capture (no launches)
                return 0
            if self._prof is not None:
                return self._profiled(name, fn, args)
            rc = fn(*args)
            if rc != 0:
                raise D3BError(f"d3b_{name} failed ({rc}): {self.last_error()}")
            if _SYNC_EACH:  # debugging aid (D3B_SYNC_EACH=1, eager mode only): attribute an asynchronous fault to its launch
                import torch

                try:
                    torch.cuda.synchronize()
                except Exception as e:  # noqa: BLE001
                    raise D3BError(f"device fault after d3b_{name}{tuple(a if isinstance(a, (int, float)) else '.' for a in args)}: {e}")
            return rc

        call.__name__ = name
        setattr(self, name, call)
        return call

    def launch_count(self) -> int:
        return int(self._dll.d3b_launch_count())


_LIB = None


def lib() -> Lib:
    global _LIB
    if _LIB is None:
        _LIB = Lib()
    return _LIB
