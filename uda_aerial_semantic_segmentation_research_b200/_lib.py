"""ctypes binding of the C-ABI library ``libuda_b200.so`` (declared in ``include/uda_b200.h``).

The library is built in-tree by ``make`` / ``__graft_entry__.build()``.  There is no fallback: if the
shared object is missing, or a call returns a negative code, a ``RuntimeError`` is raised.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libuda_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "uda_b200.h")

F32, BF16, I64, U8 = 0, 1, 2, 3

_lib = None


class UdaError(RuntimeError):
    pass


def header_symbols(header_path=HEADER_PATH):
    """Names of every function declared in include/uda_b200.h (used by the symbol-export test)."""
    src = open(header_path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uda_[a-z0-9_]+)\s*\(", src)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UdaError(
                f"{LIB_PATH} not found: build the sm_100a extension first (`make` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.uda_last_error.restype = ctypes.c_char_p
        _lib.uda_seg_loss_workspace_bytes.restype = ctypes.c_size_t
    return _lib


def _conv(a):
    # pointers travel as c_void_p, python floats as c_float, ints as c_longlong-safe ints
    if isinstance(a, float):
        return ctypes.c_float(a)
    if a is None:
        return ctypes.c_void_p(0)
    return a


#: optional per-entry-point CUDA-event profiler (bench.py): name -> list of (start_event, stop_event)
PROFILE = None


def call(name, *args, unsupported_ok=False):
    """Call ``uda_<name>``; raise with the library's thread-local message on a negative return code.  With
    ``unsupported_ok`` a return code of UDA_ERR_UNSUPPORTED (-2: nothing was launched) returns False instead."""
    fn = getattr(lib(), "uda_" + name)
    if PROFILE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*[_conv(a) for a in args])
        e1.record()
        PROFILE.setdefault(name, []).append((e0, e1))
    else:
        rc = fn(*[_conv(a) for a in args])
    if rc == -2 and unsupported_ok:
        if PROFILE is not None:
            PROFILE[name].pop()
        return False
    if rc != 0:
        msg = lib().uda_last_error()
        raise UdaError(f"uda_{name} failed ({rc}): {msg.decode() if msg else '?'}")
    return True


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def ll(v):
    return ctypes.c_longlong(int(v))


def ci(v):
    return ctypes.c_int(int(v))
