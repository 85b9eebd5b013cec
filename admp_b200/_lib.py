"""ctypes binding of libadmp_b200.so (the C ABI declared in include/admp_b200.h).

The product path has NO CPU fallback: if the library is missing or no CUDA device is
present, every compute entry point raises (loudly) instead of computing elsewhere.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ADMP_LIB: alternative build of the same library (kernel A/B measurements); never a different implementation
LIB_PATH = os.environ.get('ADMP_LIB') or os.path.join(_HERE, 'lib', 'libadmp_b200.so')

F64, F32 = 0, 1
CK_COULOMB, CK_DISP6, CK_DISP8, CK_DISP10 = 1, 6, 8, 10

# scalar block slots (include/admp_b200.h)
S_E_REAL, S_E_RECIP, S_E_SELF, S_E_PEN = 0, 1, 2, 3
S_DBOX, S_DNSTAR, S_TK, S_DMSCALE, S_DPSCALE, S_MAXFIELD, S_COUNT = 4, 13, 22, 28, 33, 38, 48

WANT_GRAD, WANT_VIRIAL, WANT_PGRAD, SCF, SCF_HOSTSYNC = 1, 2, 4, 8, 16
SCF_CG = 32                 # beyond the reference: conjugate-gradient SCF (include/admp_b200.h)
REUSE_PAIR_TILES = 0x20000000

EXPORTS = [
    'admp_last_error', 'admp_version', 'admp_ctx_create', 'admp_ctx_destroy', 'admp_ctx_set_pme',
    'admp_ctx_set_topology', 'admp_ctx_set_kvec_order', 'admp_ctx_set_pair_cluster', 'admp_ctx_set_in_flight', 'admp_ctx_pair_cluster_active', 'admp_ctx_set_spread', 'admp_ctx_spread_bricks', 'admp_ctx_workspace_bytes', 'admp_ctx_scf_graph_active',
    'admp_frames_fwd', 'admp_frames_bwd', 'admp_rotate', 'admp_pme_real', 'admp_pme_recip', 'admp_pme_self',
    'admp_pme_spread', 'admp_pme_spread_only', 'admp_pme_fft', 'admp_pme_convolve', 'admp_pme_gather', 'admp_ctx_buffer', 'admp_ctx_buffer_io', 'admp_pme_fft_convolve', 'admp_pme_fft_pass', 'admp_set_box', 'admp_mesh_zero', 'admp_pme_spread_range',
    'admp_pme_gather_range', 'admp_pme_self_range', 'admp_frames_bwd_range', 'admp_scf_step', 'admp_virial_finalize', 'admp_ctx_fft_backend', 'admp_ctx_set_fft_backend',
    'admp_pme_eval', 'admp_disp_eval', 'admp_tt_pair', 'admp_tt_pair_c10', 'admp_pair_geometry', 'admp_pair_geometry_bwd', 'admp_nblist_build', 'admp_nblist_build_hostbox', 'admp_fp_peak',
    'admp_ipc_export', 'admp_ipc_open', 'admp_ipc_close', 'admp_ctx_set_peers', 'admp_slab_zero', 'admp_slab_spread', 'admp_slab_fft',
    'admp_slab_gather',
]


class AdmpLibraryError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library once; raise AdmpLibraryError with build instructions if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AdmpLibraryError(
            'admp_b200: CUDA library %s is missing. Build it with `python -c "import __graft_entry__ as g; g.build()"` '
            'or `make -C admp_b200/csrc`. There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u32, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32, ctypes.c_double
    lib.admp_last_error.restype = ctypes.c_char_p
    lib.admp_last_error.argtypes = []
    lib.admp_version.restype = i32
    lib.admp_ctx_create.argtypes = [ctypes.POINTER(vp), i32, i32]
    lib.admp_ctx_destroy.argtypes = [vp]
    lib.admp_ctx_set_pme.argtypes = [vp, dbl, i32, i32, i32, i32]
    lib.admp_ctx_set_topology.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.admp_ctx_set_kvec_order.argtypes = [vp, i32]
    lib.admp_ctx_set_pair_cluster.argtypes = [vp, i32, i32]
    lib.admp_ctx_set_in_flight.argtypes = [vp, i32]
    lib.admp_ctx_pair_cluster_active.argtypes = [vp]
    lib.admp_ctx_set_spread.argtypes = [vp, i32]
    lib.admp_ctx_spread_bricks.argtypes = [vp]
    lib.admp_ctx_workspace_bytes.restype = i64
    lib.admp_ctx_workspace_bytes.argtypes = [vp]
    lib.admp_ctx_scf_graph_active.argtypes = [vp]
    lib.admp_frames_fwd.argtypes = [vp] * 8
    lib.admp_frames_bwd.argtypes = [vp] * 9
    lib.admp_rotate.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp]
    lib.admp_pme_real.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, i32, u32, vp, vp, vp, vp, vp, vp]
    lib.admp_pme_recip.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, i32, i32, u32, vp, vp, i32, vp, vp]
    lib.admp_pme_spread.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp]
    lib.admp_pme_spread_only.argtypes = [vp, vp, vp, vp, i32, i32, vp]
    lib.admp_pme_fft.argtypes = [vp, vp, i32]
    lib.admp_pme_convolve.argtypes = [vp, vp, i32, u32, vp]
    lib.admp_pme_gather.argtypes = [vp, vp, vp, vp, i32, i32, vp, i32, u32, vp, vp, i32, vp, vp]
    lib.admp_ctx_buffer.argtypes = [vp, i32]
    lib.admp_ctx_buffer_io.argtypes = [vp, vp, i32, vp, i64, i32]
    lib.admp_pme_fft_convolve.argtypes = [vp, vp, i32, u32, vp]
    lib.admp_pme_fft_pass.argtypes = [vp, vp, i32, i32, vp]
    lib.admp_set_box.argtypes = [vp, vp, vp]
    lib.admp_mesh_zero.argtypes = [vp, vp]
    lib.admp_pme_spread_range.argtypes = [vp, vp, vp, vp, i32, i32, vp, i32, i32]
    lib.admp_pme_gather_range.argtypes = [vp, vp, vp, vp, i32, i32, vp, i32, u32, vp, vp, i32, vp, vp, i32, i32]
    lib.admp_pme_self_range.argtypes = [vp, vp, vp, vp, vp, u32, vp, vp, vp, vp, i32, i32]
    lib.admp_frames_bwd_range.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32]
    lib.admp_scf_step.argtypes = [vp, vp, vp, vp, vp, vp, i32, dbl, u32, vp, vp]
    lib.admp_virial_finalize.argtypes = [vp, vp, vp]
    lib.admp_ctx_fft_backend.argtypes = [vp]
    lib.admp_ctx_set_fft_backend.argtypes = [vp, i32]
    lib.admp_ctx_buffer.restype = vp
    lib.admp_pme_self.argtypes = [vp, vp, vp, vp, vp, u32, vp, vp, vp, vp]
    lib.admp_pme_eval.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, u32, i32, dbl, vp, vp, vp, vp, vp, vp, vp]
    lib.admp_disp_eval.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, i32, u32, vp, vp, vp]
    lib.admp_tt_pair.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, u32, vp, vp, vp]
    lib.admp_tt_pair_c10.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, u32, vp, vp, vp]
    lib.admp_pair_geometry.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp]
    lib.admp_pair_geometry_bwd.argtypes = [vp, vp, vp, vp, vp, i64, vp, u32, vp, vp]
    lib.admp_nblist_build.argtypes = [vp, vp, vp, vp, i32, dbl, vp, i64, vp]
    lib.admp_nblist_build_hostbox.argtypes = [vp, vp, vp, vp, vp, i32, dbl, vp, i64, vp]
    lib.admp_fp_peak.argtypes = [vp, i32, ctypes.POINTER(dbl)]
    lib.admp_ipc_export.argtypes = [vp, vp]
    lib.admp_ipc_open.argtypes = [vp, ctypes.POINTER(vp)]
    lib.admp_ipc_close.argtypes = [vp]
    lib.admp_ctx_set_peers.argtypes = [vp, i32, i32, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    lib.admp_slab_zero.argtypes = [vp, vp]
    lib.admp_slab_spread.argtypes = [vp, vp, vp, vp, i32, i32, vp, i32]
    lib.admp_slab_fft.argtypes = [vp, vp, i32, i32, u32, vp]
    lib.admp_slab_gather.argtypes = [vp, vp, vp, vp, i32, i32, vp, i32, u32, vp, vp, i32, vp, vp, i32]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ('admp_last_error', 'admp_ctx_workspace_bytes', 'admp_ctx_buffer'):
            fn.restype = i32
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise AdmpLibraryError(load().admp_last_error().decode('utf-8', 'replace'))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise AdmpLibraryError('admp_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback.')


def ptr(t):
    """Device (or host) address of a tensor / None -> NULL."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
