"""ctypes binding of libpackppi_b200.so (C ABI: include/packppi_b200.h).

There is NO fallback: if the shared library is missing or the device is not a compute-capability-10 GPU, every
kernel call raises RuntimeError.  Build with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C packppi_b200/csrc`.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PACKPPI_B200_LIB") or os.path.join(_HERE, "csrc", "libpackppi_b200.so")  # override: A/B builds

_P, _I, _F = ctypes.c_void_p, ctypes.c_int64, ctypes.c_float

# name -> argument kinds in header order: p = device pointer, i = int64, f = float, s = stream
SIGNATURES = {
    "pp_knn_build": "pp" "iii" "ppppp" "s",
    "pp_knn_build_cells": "pp" "iii" "ppppp" "pp" "s",
    "pp_geometry_build": "p" "i" "p" "s",
    "pp_edge_embed": "ppppp" "ii" "p" "s",
    "pp_node_embed": "ppppppp" "iii" "p" "s",
    "pp_ipmp_layer": "p" "i" "ppppp" "iii" "pp" "i" "p" "i" "pppp" "s",
    "pp_ipmp_node_pre": "p" "ii" "pppp" "iii" "pppp" "s",
    "pp_ipmp_edge_node": "p" "i" "pppp" "iii" "p" "i" "pppp" "s",
    "pp_ipmp_node_post": "p" "i" "ppppp" "iii" "pp" "s",
    "pp_ipmp_edge_edge": "p" "i" "pppp" "iii" "p" "i" "pppp" "s",
    "pp_ipmp_edge_tc": "p" "ii" "ppppp" "iii" "p" "i" "pppp" "ii" "p" "pp" "s",
    "pp_ipmp_node_pre_tc": "p" "ii" "pp" "ii" "pppp" "p" "pp" "s",
    "pp_ipmp_node_post_tc32": "p" "i" "ppp" "iii" "pp" "p" "pp" "s",
    "pp_decode_step": "pp" "ii" "p" "i" "ff" "ppp" "ppp" "f" "ii" "s",
    "pp_atom14_fwd": "pppp" "ii" "p" "s",
    "pp_clash_neighbours": "ppppp" "ii" "f" "i" "pppp" "s",
    "pp_clash_reach": "pppp" "i" "p" "s",
    "pp_clash_neighbours_cells": "ppp" "ii" "ff" "i" "ppp" "pp" "s",
    "pp_clash_fwd_bwd": "ppppppppp" "ii" "ff" "i" "p" "pp" "ppp" "s",
    "pp_prox_init": "ppppppppp" "iii" "p" "ff" "p" "pppp" "pp" "ppp" "p" "s",
    "pp_prox_step": "ppppppppp" "ppppp" "iii" "p" "ffffffff" "ppp" "ppp" "p" "p" "i" "s",
    "pp_prox_loss": "p" "iii" "p" "f" "i" "p" "s",
    "pp_prox_init_from_mean": "pppp" "i" "ppppp" "s",
    "pp_featurize": "pppppp" "ii" "ppp" "pppppppppppppp" "s",
    "pp_selftest_umma_f16": "ppp" "iii" "s",
    "pp_selftest_gather4": "p" "iiiiiii" "p" "s",
}

_KIND = {"p": _P, "i": _I, "f": _F, "s": _P}
_lib = None

# kernels launched per entry point (pp_ipmp_layer: 3, or 5 with the edge update - the caller passes `kernels=`)
KERNELS = {"pp_knn_build": 1, "pp_knn_build_cells": 5, "pp_geometry_build": 1, "pp_edge_embed": 1, "pp_node_embed": 1, "pp_ipmp_layer": 5,
           "pp_ipmp_node_pre": 1, "pp_ipmp_edge_node": 1, "pp_ipmp_node_post": 1, "pp_ipmp_edge_edge": 1, "pp_ipmp_edge_tc": 1, "pp_ipmp_node_pre_tc": 1, "pp_ipmp_node_post_tc32": 1,
           "pp_decode_step": 1, "pp_atom14_fwd": 1, "pp_clash_neighbours": 1, "pp_clash_reach": 1, "pp_clash_neighbours_cells": 5, "pp_clash_fwd_bwd": 2,
           "pp_prox_init": 4, "pp_prox_step": 2, "pp_prox_loss": 1, "pp_prox_init_from_mean": 1, "pp_featurize": 1, "pp_selftest_umma_f16": 1, "pp_selftest_gather4": 1}
LAUNCHES = 0      # running count of kernels launched through call()
PROFILE = None    # {entry name: []} -> call() appends (start event, end event, rows) around those entries


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: the CUDA extension has not been built "
                           "(run __graft_entry__.build() or `make -C packppi_b200/csrc`); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    lib.pp_last_error.restype = ctypes.c_char_p
    lib.pp_abi_version.restype = ctypes.c_int
    for fn in ("pp_layout_count", "pp_layout_total_floats", "pp_geo_stride", "pp_table_stride", "pp_tc_stream_floats",
               "pp_tc_pre_stream_floats",
               "pp_knn_cells_max"):
        getattr(lib, fn).restype = _I
    lib.pp_layout_entry.argtypes = [_I, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(_I), ctypes.POINTER(_I)]
    for name, sig in SIGNATURES.items():
        f = getattr(lib, name)
        f.restype = ctypes.c_int
        f.argtypes = [_KIND[c] for c in sig]
    if lib.pp_abi_version() != 2:
        raise RuntimeError("libpackppi_b200.so: ABI version mismatch, rebuild the extension")
    _lib = lib
    return lib


def layout():
    """name -> (offset, size) of the packed weight blob, as compiled into the library."""
    lib = load()
    out = {}
    name, off, size = ctypes.c_char_p(), _I(), _I()
    for i in range(lib.pp_layout_count()):
        lib.pp_layout_entry(i, ctypes.byref(name), ctypes.byref(off), ctypes.byref(size))
        out[name.value.decode()] = (off.value, size.value)
    return out, lib.pp_layout_total_floats()


_device_ok = set()


def _check_device(lib, dev):
    if dev in _device_ok:
        return
    with torch.cuda.device(dev):
        if lib.pp_check_device() != 0:
            raise RuntimeError(lib.pp_last_error().decode())
    _device_ok.add(dev)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("packppi_b200 kernels need CUDA tensors; there is no CPU fallback "
                           "(use the reference implementation for --device cpu)")
    if not t.is_contiguous():
        raise RuntimeError("packppi_b200 kernels need contiguous tensors")
    return t.data_ptr()


def call(name, *args, device=None, kernels=None, rows=0, tag=None):
    """Invoke an entry point on torch's current stream; tensors are passed as device pointers."""
    global LAUNCHES
    lib = load()
    dev = device
    conv = []
    for a in args:
        if torch.is_tensor(a):
            if dev is None:
                dev = a.device
            conv.append(ptr(a))
        else:
            conv.append(a)
    if dev is None:
        raise RuntimeError(f"{name}: no tensor argument to take the device from")
    _check_device(lib, dev.index if dev.index is not None else torch.cuda.current_device())
    with torch.cuda.device(dev):
        cur = torch.cuda.current_stream()
        prof = PROFILE.get(name if tag is None else f"{name}:{tag}") if PROFILE is not None else None
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
        rc = getattr(lib, name)(*conv, cur.cuda_stream)
        if prof is not None:
            e1.record(cur)
            prof.append((e0, e1, rows))
    LAUNCHES += KERNELS[name] if kernels is None else kernels
    if rc != 0:
        raise RuntimeError(f"{name} failed: {lib.pp_last_error().decode()}")
