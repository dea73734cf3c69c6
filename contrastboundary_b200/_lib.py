"""ctypes loader for libcbops.so — the ONLY compute backend of this package.

There is deliberately no CPU or PyTorch fallback: if the CUDA library is missing or a call fails,
the operators raise (`CbopsError`).  `python -m contrastboundary_b200.build` builds the library.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcbops.so")


class CbopsError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CbopsError(
                "contrastboundary_b200/libcbops.so is missing — build it with "
                "`python -m contrastboundary_b200.build` (there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.cb_last_error_string.restype = C.c_char_p
        _lib.cb_knn_workspace_bytes.restype = C.c_size_t
        _lib.cb_knn_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        for name in ("cb_cbl_workspace_bytes", "cb_pt_bnbuf_floats", "cb_pt_stats_doubles", "cb_pt_bwd_scratch_floats"):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = C.c_size_t
        if os.environ.get("CB_UMMA"):            # developer knob: CB_UMMA=0 -> mma.sync kernels instead of tcgen05 for the linear layers
            _lib.cb_linear_set_umma(C.c_int(int(os.environ["CB_UMMA"])))
        if os.environ.get("CB_GRID_FUSED"):      # developer knob: 0 -> the multi-kernel grid build
            _lib.cb_grid_set_fused(C.c_int(int(os.environ["CB_GRID_FUSED"])))
        if os.environ.get("CB_PDL"):             # developer knob: 0 -> no programmatic dependent launches
            _lib.cb_set_pdl(C.c_int(int(os.environ["CB_PDL"])))
        if os.environ.get("CB_FPS_MODE"):        # developer knob (profilers that cannot launch cluster kernels): see cbops.h
            _lib.cb_fps_set_mode(C.c_int(int(os.environ["CB_FPS_MODE"])), C.c_int(8192))
    return _lib


def launch_count():
    f = lib().cb_launch_count
    f.restype = C.c_ulonglong
    return int(f())


def ptr(t):
    """device pointer of a tensor (None -> NULL)"""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream(device=None):
    """torch's current stream (of `device`, default: the current device) as a C pointer"""
    if _raw_stream is not None:          # ~10x cheaper than building a torch.cuda.Stream object; called once per kernel launch
        if device is None:
            idx = torch.cuda.current_device()
        else:
            idx = device if isinstance(device, int) else torch.device(device).index
            if idx is None:
                idx = torch.cuda.current_device()
        return C.c_void_p(_raw_stream(idx))
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def check(rc, what):
    if rc != 0:
        msg = lib().cb_last_error_string()
        raise CbopsError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise CbopsError("contrastboundary_b200 operators run on CUDA tensors only (no CPU fallback)")


_fn_cache = {}
_Tensor = torch.Tensor


def call(name, *args):
    f = _fn_cache.get(name)
    if f is None:
        f = _fn_cache[name] = getattr(lib(), name)
    conv = []
    checked = False
    for a in args:
        if isinstance(a, _Tensor):
            if not checked and a.is_cuda:
                # the stream argument is the CURRENT device's stream: the tensors must live there (checked on the first CUDA
                # tensor of a call — the operands of one call share a device by construction)
                checked = True
                if a.get_device() != torch.cuda.current_device():
                    raise CbopsError("%s: tensor on cuda:%d but the current device is cuda:%d (call torch.cuda.set_device)"
                                     % (name, a.get_device(), torch.cuda.current_device()))
            conv.append(C.c_void_p(a.data_ptr()))
        elif isinstance(a, float):
            conv.append(C.c_float(a))
        elif isinstance(a, int):
            conv.append(C.c_int(a)) if -2 ** 31 < a < 2 ** 31 else conv.append(C.c_size_t(a))
        else:
            conv.append(a)
    rc = f(*conv)
    if rc != 0:
        check(rc, name)


class SizeT:
    """wrap an int that must be passed as size_t"""
    def __init__(self, v):
        self.v = int(v)


_ws_cache = {}
WS_NO_CACHE = False      # set while a CUDA graph is being captured: every capture owns its workspaces (graph pool memory)


def workspace(nbytes, device, tag="default"):
    """Cached, 256-byte aligned byte workspace per (device, stream, tag); grows monotonically."""
    if WS_NO_CACHE:
        return torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
    key = (device.index, torch.cuda.current_stream().cuda_stream, tag)
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _ws_cache[key] = t
    return t
