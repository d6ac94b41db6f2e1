"""ctypes binding of libnfdpf.so (the C-ABI declared in include/nfdpf.h).

No CPU fallback: if the library is missing or a call fails, a RuntimeError / ValueError is raised
(SURVEY.md 8b error conventions)."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnfdpf.so")

_P, _I, _F, _D, _I64 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_int64
SIGNATURES = {
    "nfdpf_version": (_I, []),
    "nfdpf_last_error": (C.c_char_p, []),
    "nfdpf_launch_count": (_I64, []),
    "nfdpf_soft_resample_fwd": (_I, [_P, _P, _P, _P, _D, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "nfdpf_soft_resample_bwd": (_I, [_P, _P, _P, _P, _P, _D, _I, _I, _I, _P, _P, _P, _P, _P]),
    "nfdpf_motion_moments": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _I, _I, _P]),
    "nfdpf_proposal_terms_fwd": (_I, [_P, _P, _P, _P, _P, _P, _F, _I64, _P, _P, _P]),
    "nfdpf_proposal_terms_bwd": (_I, [_P, _P, _P, _P, _F, _I64, _P, _P, _P, _P]),
    "nfdpf_weight_update_fwd": (_I, [_P, _P, _P, _P, _F, _I, _I, _P, _P, _P, _P]),
    "nfdpf_weight_update_bwd": (_I, [_P, _P, _P, _P, _F, _I, _I, _P, _P, _P, _P, _P]),
    "nfdpf_sum4": (_I, [_P, _P, _P, _P, _I64, _P, _P]),
    "nfdpf_row_moments": (_I, [_P, _I, _I, _I, _P, _I, _I, _P]),
    "nfdpf_row_moments_head": (_I, [_P, _I, _I, _I, _P, _I, _P, _I, _P]),
    "nfdpf_ot_workspace": (_I64, [_I, _I]),
    "nfdpf_ot_resample_fwd": (_I, [_P, _P, _F, _F, _F, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "nfdpf_ot_resample_bwd": (_I, [_P, _P, _F, _I, _I, _I, _P, _P, _P]),
    "nfdpf_gate_weights_fwd": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "nfdpf_gate_weights_bwd": (_I, [_P, _P, _P, _P, _I, _I, _P, _P]),
    "nfdpf_ess_gate": (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _P, _P]),
    "nfdpf_motion_moments_rng": (_I, [_P, _P, _P, _F, _I, _I, _I, _P, _P, _P, _I, _I, _P]),
    "nfdpf_init_particles_rng": (_I, [_P, _I, _P, _F, _I, _I, _I, _I, _P, _P]),
    "nfdpf_weighted_mean_fwd": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "nfdpf_weighted_mean_bwd": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "nfdpf_block_density_fwd": (_I, [_P, _I64, _I64, _P, _I64, _I64, _P, _I64, _I64, _P, _I64, _I64, _I, _I, _I, _I, _P, _P, _P, _P]),
    "nfdpf_block_density_bwd": (_I, [_P, _P, _I64, _I64, _P, _I64, _I64, _P, _I, _I, _I, _I, _P, _P, _P, _I64, _I64, _P]),
    "nfdpf_peak_probe": (_I64, [_I, _I, _P, _P]),
    "nfdpf_coupling_fwd": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "nfdpf_coupling_bwd_workspace": (_I64, [_I, _I, _I, _I, _I, _I]),
    "nfdpf_measure_fwd": (_I, [_I, _P, _P, _I, _F, _F, _P, _P, _I, _I, _I, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nfdpf_measure_bwd_workspace": (_I64, [_I, _I, _I, _I]),
    "nfdpf_measure_bwd": (_I, [_I, _P, _P, _I, _F, _F, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nfdpf_coupling_bwd": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nfdpf_coupling_bwd_block_floats": (_I64, [_I, _I, _I, _I, _I]),
    "nfdpf_coupling_bwd_deferred": (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "nfdpf_coupling_bwd_reduce": (_I, [_I, _I, _I, _I, _I, _P, _I, _P, _P]),
}
_lib = None


def load():
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libnfdpf.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "libnfdpf needs contiguous CUDA tensors"
    if t.device.index != torch.cuda.current_device():   # launches go to the current device: a foreign pointer would fault there
        raise RuntimeError("libnfdpf: tensor lives on cuda:%d but the current device is cuda:%d (use torch.cuda.device(...))"
                           % (t.device.index, torch.cuda.current_device()))
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.nfdpf_last_error().decode()
        if rc == -1:
            raise ValueError("%s: %s" % (name, msg))
        raise RuntimeError("%s failed (%d): %s" % (name, rc, msg))


def launch_count():
    return int(load().nfdpf_launch_count())


def f32(t):
    """fp32 contiguous CUDA view/copy of t (the reference casts FCNN inputs with .float(), nf/flows.py:114)."""
    if not t.is_cuda:
        raise RuntimeError("libnfdpf has no CPU path: tensor is on %s" % t.device)
    return t.detach().to(torch.float32).contiguous()
