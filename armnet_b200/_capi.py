"""ctypes binding of libarmnet_b200.so (include/armnet_b200.h).

This is the binding a maintainer of the reference would add. There is no fallback: if the shared
library is missing the import fails, and every op raises on non-CUDA tensors.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('ARMNET_B200_LIB') or os.path.join(_HERE, 'libarmnet_b200.so')   # env: tuning builds only

OK = 0
ERR_NAMES = {-1: 'ARMNET_ERR_NULL', -2: 'ARMNET_ERR_SHAPE', -3: 'ARMNET_ERR_UNSUPPORTED', -4: 'ARMNET_ERR_CUDA',
             -5: 'ARMNET_ERR_ALIGN'}
SOLVER_AUTO = 0
SOLVER_BISECT = 1

# name -> (restype, argtypes); mirrors include/armnet_b200.h one to one
_P, _I, _L, _F, _Z = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
SIGNATURES = {
    'armnet_version': (_I, []),
    'armnet_last_error_string': (C.c_char_p, []),
    'armnet_last_launch_count': (_I, []),
    'armnet_set_tuning': (_I, [C.c_char_p, _I]),
    'armnet_device_info': (_I, [C.POINTER(_I), C.POINTER(_I)]),
    'armnet_embed_gather_f32': (_I, [_P, _I, _P, _P, _L, _L, _L, _I, _I, _P, _I, _F, _F, _I, _P, _P]),
    'armnet_linear_gather_f32': (_I, [_P, _I, _P, _P, _L, _L, _I, _P, _P, _P, _P]),
    'armnet_entmax_f32': (_I, [_P, _L, _I, _F, _I, _I, _P, _P]),
    'armnet_entmax_bwd_f32': (_I, [_P, _P, _L, _I, _F, _P, _P]),
    'armnet_fused_workspace_bytes': (_Z, [_I, _I, _I, _I]),
    'armnet_fused_bwd_supported': (_I, [_I, _I]),
    'armnet_fused_fwd_kernel_kind': (_I, [_I, _I, _I, _I, _F, _I]),
    'armnet_fused_bwd_finish_workspace_bytes': (_Z, [_I, _I, _I]),
    'armnet_fused_bwd_finish_f32': (_I, [_P, _I, _P, _L, _L, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'armnet_fused_bwd_f32': (_I, [_P, _I, _P, _P, _L, _L, _P, _P, _P, _I, _F, _L, _I, _I, _I, _I, _I,
                                  _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'armnet_mlp_split_weight_f32': (_I, [_P, _L, _P, _P, _P]),
    'armnet_mlp_linear_splits': (_I, [_L, _I, _I]),
    'armnet_mlp_linear_tf32x3': (_I, [_P, _L, _I, _P, _P, _I, _I, _P, _P]),
    'armnet_linear_tf32x3_dense': (_I, [_P, _L, _I, _P, _P, _I, _P, _P, _P]),
    'armnet_transpose_f32': (_I, [_P, _L, _L, _P, _P]),
    'armnet_mlp_hidden_tc_f32': (_I, [_P, _I, _L, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _I, _P, _P, _P]),
    'armnet_mlp_tail_packed_floats': (_Z, [_I, _I, _I]),
    'armnet_mlp_tail_f32': (_I, [_P, _I, _L, _I, _I, _I, _P, _P, _P]),
    'armnet_bn_workspace_floats': (_Z, [_L, _I, _I]),
    'armnet_bn_train_fwd_f32': (_I, [_P, _L, _I, _I, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P]),
    'armnet_bn_train_bwd_f32': (_I, [_P, _P, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    'armnet_libsvm_count_lines': (_I, [C.c_char_p, C.POINTER(_L)]),
    'armnet_libsvm_parse': (_I, [C.c_char_p, _I, _L, _P, _P, _P, C.POINTER(_L), C.POINTER(_L)]),
    'armnet_clamp_adam_f32': (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _F, _L, _P]),
    'armnet_fused_prepare_f32': (_I, [_P, _P, _P, _I, _F, _I, _I, _I, _I, _I, _P, _P]),
    'armnet_fused_fwd_prepared_f32': (_I, [_P, _I, _P, _P, _L, _L, _F, _I, _I, _L, _I, _I, _I, _I, _I,
                                           _I, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'armnet_fused_fwd_f32': (_I, [_P, _I, _P, _P, _L, _L, _P, _P, _P, _I, _F, _I, _I, _L, _I, _I, _I, _I, _I,
                                  _I, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} not found: the CUDA extension is not built. Run `make -C armnet_b200/csrc` '
            f'(or `python -c "import __graft_entry__ as g; g.build()"`). armnet_b200 has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class ArmnetError(RuntimeError):
    pass


def check(rc, what):
    if rc != OK:
        msg = lib.armnet_last_error_string().decode('utf-8', 'replace')
        raise ArmnetError(f'{what} failed: {ERR_NAMES.get(rc, rc)}: {msg}')


def version():
    return lib.armnet_version()
