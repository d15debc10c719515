"""ctypes loader of the plain-C oracle (oracle/armnet_oracle.c). TEST INFRASTRUCTURE: tests/ and bench.py's
cpu_baseline leg only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, '_build', 'liboracle.so')


def load():
    if not os.path.exists(_LIB):
        subprocess.run(['make', '-C', _HERE], check=True, capture_output=True)
    lib = C.CDLL(_LIB)
    lib.armnet_oracle_hot_path.restype = C.c_int
    lib.armnet_oracle_hot_path.argtypes = [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 3 + [
        C.c_int, C.c_float, C.c_int, C.c_int64] + [C.c_int] * 5 + [C.c_void_p] * 5
    lib.armnet_oracle_num_threads.restype = C.c_int
    return lib


def hot_path(state, alpha, ids, values, n_iter=50, want=('e', 'g', 'p', 's', 'z')):
    """numpy in / numpy out. `values` (float32 array) is clamped in place like the reference does."""
    lib = load()
    one_head = 'attn_layer.bilinear_w.weight' in state
    f32 = lambda t: np.ascontiguousarray(np.asarray(t, dtype=np.float32))
    table = f32(state['embedding.embedding.weight'])
    W = f32(state['attn_layer.bilinear_w.weight' if one_head else 'attn_layer.bilinear_w'])
    Q = f32(state['attn_layer.query'])
    Vv = f32(state['attn_layer.values'])
    ids = np.ascontiguousarray(np.asarray(ids, dtype=np.int64))
    assert values.dtype == np.float32 and values.flags['C_CONTIGUOUS']
    B, F = ids.shape
    V, E = table.shape
    if one_head:
        D, K, O = W.shape[0], 1, Q.shape[0]
    else:
        K, O, D = Q.shape
    R = K * O
    shapes = {'e': (B, F, E), 'g': (B, R, F), 'p': (B, R, F), 's': (B, R, E), 'z': (B, R, E)}
    out = {k: np.empty(shapes[k], dtype=np.float32) for k in want}
    ptr = lambda k: out[k].ctypes.data if k in out else None
    rc = lib.armnet_oracle_hot_path(ids.ctypes.data, values.ctypes.data, table.ctypes.data, V, W.ctypes.data,
                                    Q.ctypes.data, Vv.ctypes.data, int(one_head), float(alpha), n_iter, B, F, E, D,
                                    K, O, ptr('e'), ptr('g'), ptr('p'), ptr('s'), ptr('z'))
    if rc != 0:
        raise IndexError('index out of range in self')
    return out


def num_threads():
    return load().armnet_oracle_num_threads()
