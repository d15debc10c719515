"""CPU-side checks: the C-ABI library loads and exports every symbol include/armnet_b200.h declares (no compute
calls without a GPU), argument validation returns error codes instead of crashing, and the drop-in modules keep
the reference's constructor / state_dict / init-order surface."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol():
    from armnet_b200 import _capi
    hdr = open(os.path.join(ROOT, 'include', 'armnet_b200.h')).read()
    body = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(armnet_[a-z0-9_]+)\s*\(', body))
    assert declared, 'no declarations parsed'
    raw = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f'{name} declared in the header but not exported'
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    assert _capi.version() == 100


def test_argument_validation_without_gpu():
    from armnet_b200 import _capi
    lib = _capi.lib
    # row-pair tables with odd strides, then the tensor-memory kernel's operands: 512 packed M' rows of 128 bytes and
    # 512 field-packed value rows of 22 float2
    assert lib.armnet_fused_workspace_bytes(39, 10, 4, 128) == 256 * (11 + 39) * 2 * 4 + 512 * 128 + 512 * 22 * 8
    assert lib.armnet_fused_workspace_bytes(39, 10, 4, 100) == 200 * (11 + 39) * 2 * 4     # no tensor-memory instance
    assert lib.armnet_fused_workspace_bytes(39, 100, 1, 32) > 0
    assert lib.armnet_fused_workspace_bytes(65, 10, 4, 128) == 0        # > 64 fields: no instance
    assert lib.armnet_fused_workspace_bytes(39, 129, 4, 128) == 0
    rc = lib.armnet_entmax_f32(None, 4, 8, 1.5, 0, 50, None, None)
    assert rc == -1 and b'null' in lib.armnet_last_error_string()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.armnet_entmax_f32(p, 4, 0, 1.5, 0, 50, p, None) == -2   # F == 0
    assert lib.armnet_entmax_f32(p, 4, 8, 0.5, 0, 50, p, None) == -2   # alpha < 1
    assert lib.armnet_entmax_f32(p, 4, 65, 1.5, 0, 50, p, None) == -3  # too many fields
    assert lib.armnet_embed_gather_f32(p, 0, p, p, 10, 4, 2, 3, 8, p, 0, 0.0, 0.0, 0, None, None) == -2  # ld < E


def test_ops_refuse_cpu_tensors():
    from armnet_b200 import ops
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.entmax(torch.zeros(2, 4), 1.5)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.fused_forward(torch.zeros(2, 3, dtype=torch.int64), torch.ones(2, 3), torch.zeros(5, 4),
                          torch.zeros(1, 4, 4), torch.zeros(1, 2, 4), torch.zeros(1, 2, 3), 1.5)


@pytest.mark.parametrize('name', ['c1_1h_frappe', 'c2a_init'])
def test_same_seed_gives_reference_parameters(name):
    """Same constructor order and init calls as armnet.py:62-75 / armnet_1h.py:60-74: seeding like train.py:144
    reproduces the reference's state_dict bit for bit."""
    import armnet_b200 as ab
    g = load_golden(name)
    c = g.cfg
    torch.manual_seed(2025)
    if c['model'] == 'armnet':
        m = ab.ARMNetModel(c['nfield'], c['nfeat'], c['nemb'], c['nhead'], c['alpha'], c['nhid'], c['mlp_nlayer'],
                           c['mlp_nhid'], 0.0, bool(c['ensemble']), 2, 16)
    else:
        m = ab.ARMNet1H(c['nfield'], c['nfeat'], c['nemb'], c['alpha'], c['nhid'], c['d_k'], c['mlp_nlayer'],
                        c['mlp_nhid'], 0.0, bool(c['ensemble']), 2, 16)
    sd = m.state_dict()
    assert set(sd) == set(g.state)
    for k in sd:
        assert torch.equal(sd[k], g.state[k]), k


def test_state_dict_round_trip_all_fixtures(golden):
    import armnet_b200 as ab
    c = golden.cfg
    if c['model'] == 'armnet':
        m = ab.ARMNetModel(c['nfield'], c['nfeat'], c['nemb'], c['nhead'], c['alpha'], c['nhid'], c['mlp_nlayer'],
                           c['mlp_nhid'], 0.0, bool(c['ensemble']), 2, 16)
    else:
        m = ab.ARMNet1H(c['nfield'], c['nfeat'], c['nemb'], c['alpha'], c['nhid'], c['d_k'], c['mlp_nlayer'],
                        c['mlp_nhid'], 0.0, bool(c['ensemble']), 2, 16)
    m.load_state_dict(golden.state)             # strict: names and shapes are the reference's
    with pytest.raises(RuntimeError, match='CUDA only'):
        m({'id': golden.ids, 'value': golden.values.clone()})


def test_create_model_surface():
    import argparse
    import logging
    import armnet_b200 as ab
    args = argparse.Namespace(model='armnet', nfield=10, nfeat=100, nemb=10, nattn_head=4, alpha=1.7, h=8,
                              mlp_nlayer=2, mlp_nhid=16, dropout=0.0, ensemble=False, dnn_nlayer=2, dnn_nhid=16)
    m = ab.create_model(args, logging.getLogger('t'))
    assert isinstance(m, ab.ARMNetModel) and m.attn_layer.values.shape == (4, 8, 10)
    args.model = 'armnet_1h'
    assert isinstance(ab.create_model(args, logging.getLogger('t')), ab.ARMNet1H)
    args.model = 'nope'
    with pytest.raises(ValueError, match='unknown model'):
        ab.create_model(args, logging.getLogger('t'))


def test_train_cli_surface_and_helpers(tmp_path):
    """train.py keeps the reference's flags/defaults (train.py:15-50) and accepts run.sh's stale aliases."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('armnet_train_cli', os.path.join(ROOT, 'train.py'))
    t = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(t)          # guarded by __main__: importing must not parse argv or load data
    a = t.get_args([])
    assert (a.model, a.nfeat, a.nfield, a.nemb, a.h, a.nattn_head, a.alpha) == ('armnet', 5500, 10, 10, 128, 4, 1.7)
    assert (a.mlp_nlayer, a.mlp_nhid, a.batch_size, a.lr, a.seed, a.patience) == (2, 256, 4096, 0.003, 2025, 1)
    b = t.get_args(['--nlayer', '3', '--mlp_hid', '200', '--dnn_hid', '200', '--model', 'armnet_1h'])
    assert (b.mlp_nlayer, b.mlp_nhid, b.dnn_nhid) == (3, 200, 200)
    # libsvm parsing incl. a malformed line (skipped like data_loader.py:37-44)
    f = tmp_path / 'x.libsvm'
    f.write_text('1 3:1 7:0.5\n0 2:1 9:1\nbroken line\n1 4:x 5:1\n')
    ids, vals, y = t.load_libsvm(str(f), 2)
    assert ids.tolist() == [[3, 7], [2, 9]] and vals[0, 1].item() == 0.5 and y.tolist() == [1.0, 0.0]
    # AUC rank statistic
    logits = torch.tensor([0.1, 0.4, 0.35, 0.8])
    target = torch.tensor([0., 0., 1., 1.])
    assert abs(t.auc_on_device(logits, target) - 0.75) < 1e-12
    assert t.auc_on_device(logits, torch.ones(4)) == 0.0


def test_argument_validation_of_mlp_training_and_data_entries_without_gpu():
    """The newer entries also fail with error codes (never crash) before touching the device."""
    from armnet_b200 import _capi
    lib = _capi.lib
    buf = ctypes.create_string_buffer(256)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.armnet_mlp_tail_packed_floats(256, 1, 1) == 2 * 256 + 256 * 256 + 2 * 256 + 256 + 1
    assert lib.armnet_mlp_tail_packed_floats(0, 1, 1) == 0
    assert lib.armnet_mlp_split_weight_f32(None, 8, p, p, None) == -1
    assert lib.armnet_mlp_linear_tf32x3(None, 4, 8, p, p, 8, 1, p, None) == -1
    assert lib.armnet_mlp_linear_tf32x3(p, 4, 8, p, p, 8, 0, p, None) == -2          # splits < 1
    assert lib.armnet_mlp_linear_tf32x3(p, 4, 6, p, p, 8, 1, p, None) == -5          # K % 4 != 0: TMA cannot address rows
    assert lib.armnet_mlp_tail_f32(p, 1, 4, 6, 0, 1, p, p, None) == -3               # H % 4 != 0
    assert lib.armnet_mlp_tail_f32(p, 1, 4, 1024, 0, 1, p, p, None) == -3            # H > 512
    # armnet_mlp_hidden_tc_f32(in, splits, B, H_in, a_in, c_in, w_hi, w_lo, H_out, a_out, c_out, wf, bf, NO, y, out, stream)
    assert lib.armnet_mlp_hidden_tc_f32(None, 1, 4, 8, p, p, p, p, 8, p, p, p, p, 1, p, None, None) == -1
    assert lib.armnet_mlp_hidden_tc_f32(p, 1, 4, 8, p, p, p, p, 8, None, None, None, None, 0, None, None, None) == -1  # no output
    assert lib.armnet_mlp_hidden_tc_f32(p, 1, 4, 8, p, p, p, p, 8, None, p, p, p, 1, p, None, None) == -1    # y without a_out
    assert lib.armnet_mlp_hidden_tc_f32(p, 0, 4, 8, p, p, p, p, 8, p, p, p, p, 1, p, None, None) == -2       # splits < 1
    assert lib.armnet_mlp_hidden_tc_f32(p, 1, 4, 8, p, p, p, p, 512, p, p, p, p, 1, p, None, None) == -3     # H_out > 256
    assert lib.armnet_mlp_hidden_tc_f32(p, 5, 4, 8, p, p, p, p, 8, p, p, p, p, 1, p, None, None) == -3       # > 4 split-K slices
    assert lib.armnet_mlp_hidden_tc_f32(p, 1, 4, 8, p, p, p, p, 8, p, p, p, p, 5, p, None, None) == -3       # > 4 outputs
    assert lib.armnet_mlp_hidden_tc_f32(p, 1, 4, 6, p, p, p, p, 8, p, p, p, p, 1, p, None, None) == -5       # H_in % 4 != 0
    assert lib.armnet_clamp_adam_f32(p, p, p, p, 16, 1.0, 1.0, 1e-3, 0.9, 0.999, 1e-8, 0, None) == -2   # step < 1
    assert lib.armnet_clamp_adam_f32(None, p, p, p, 16, 1.0, 1.0, 1e-3, 0.9, 0.999, 1e-8, 1, None) == -1
    assert lib.armnet_bn_train_fwd_f32(p, 1, 4, 1, None, None, None, None, 0.1, 1e-5, p, p, p, p, None) == -2
    assert b'more than 1 value per channel' in lib.armnet_last_error_string()
    assert lib.armnet_bn_workspace_floats(0, 4, 4) == 0
    assert lib.armnet_linear_gather_f32(None, 0, p, p, 10, 4, 3, None, p, None, None) == -1
    n = ctypes.c_int64(0)
    assert lib.armnet_libsvm_count_lines(b'/nonexistent/file.libsvm', ctypes.byref(n)) == -2
    assert b'cannot open' in lib.armnet_last_error_string()


def test_training_helpers_refuse_cpu():
    from armnet_b200.parallel import FlatAdam
    from armnet_b200 import BatchScorer
    import armnet_b200 as ab
    lin = torch.nn.Linear(4, 2)
    with pytest.raises(RuntimeError, match='CUDA'):
        FlatAdam(lin.parameters())
    m = ab.ARMNetModel(10, 100, 10, 2, 1.7, 4, 1, 8, 0.0, False, 1, 8).eval()
    with pytest.raises(RuntimeError, match='CUDA'):
        BatchScorer(m, 8, 10)


def test_fused_kernel_selection_is_host_logic():
    """armnet_fused_fwd_kernel_kind: pure host-side dispatch (no GPU).  Default: the tensor-memory kernel (3) where its
    shape constraints hold, else the FP32-pipe kernel (1) or the warp-MMA kernel (2, default at nemb 16); 0 when no
    instance is compiled.  armnet_set_tuning switches kinds (the environment is read once per process)."""
    from armnet_b200 import ops
    try:
        ops.set_tuning('tmem', -1)
        ops.set_tuning('mma', -1)
        assert ops.fused_fwd_kernel_kind(39, 10, 4, 128, 1.7) == 3       # C2a
        assert ops.fused_fwd_kernel_kind(40, 8, 3, 128, 1.3) == 3        # K*O = 384, general alpha
        assert ops.fused_fwd_kernel_kind(39, 10, 4, 64, 2.0) == 1        # C2b: alpha = 2 -> armnet_fwd_kernel by default
        assert ops.fused_fwd_kernel_kind(40, 8, 2, 128, 1.5) == 1
        ops.set_tuning('tmem', 1)
        assert ops.fused_fwd_kernel_kind(39, 10, 4, 64, 2.0) == 3        # forced on
        assert ops.fused_fwd_kernel_kind(40, 8, 2, 128, 1.0) == 3
        assert ops.fused_fwd_kernel_kind(39, 10, 4, 128, 2.5) == 1       # never with the literal bisection
        assert ops.fused_fwd_kernel_kind(39, 16, 4, 128, 1.7) == 3       # nemb 16: wide packed layout, opt-in
        assert ops.fused_fwd_kernel_kind(40, 13, 2, 128, 1.7) == 3
        assert ops.fused_fwd_kernel_kind(39, 16, 5, 128, 1.7) != 3       # 5 x 48 + 320 tensor-memory columns > 512
        assert ops.fused_fwd_kernel_kind(39, 17, 4, 128, 1.7) != 3       # nemb > 16
        ops.set_tuning('tmem', -1)
        assert ops.fused_fwd_kernel_kind(39, 16, 4, 128, 1.7) == 2       # default at nemb 16: armnet_fwd_mma_kernel
        assert ops.fused_fwd_kernel_kind(39, 10, 4, 128, 2.5) == 1       # alpha > 2 -> literal bisection
        assert ops.fused_fwd_kernel_kind(39, 10, 4, 128, 1.7, ops.SOLVER_BISECT) == 1
        assert ops.fused_fwd_kernel_kind(39, 10, 4, 100, 1.7) == 1       # K*O % 256 != 0
        assert ops.fused_fwd_kernel_kind(39, 10, 8, 128, 1.7) == 1       # K*O = 1024: does not fit tensor memory
        assert ops.fused_fwd_kernel_kind(38, 10, 4, 128, 1.7) == 1       # field count not compiled
        ops.set_tuning('tmem', 0)
        assert ops.fused_fwd_kernel_kind(39, 10, 4, 128, 1.7) == 1       # nemb 10 warp-MMA: measured slower, off by default
        assert ops.fused_fwd_kernel_kind(39, 16, 4, 128, 1.7) in (1, 2)  # per-instance default
        assert ops.fused_fwd_kernel_kind(10, 10, 1, 10, 1.7) == 1
        assert ops.fused_fwd_kernel_kind(22, 100, 1, 32, 1.5) == 1
        assert ops.fused_fwd_kernel_kind(100, 10, 4, 128, 1.7) == 0      # more than 64 fields: no instance
        assert ops.fused_fwd_kernel_kind(39, 10, 0, 128, 1.7) == 0
        ops.set_tuning('mma', 1)
        for alpha in (1.0, 1.3, 1.5, 1.7, 2.0):
            assert ops.fused_fwd_kernel_kind(39, 10, 4, 128, alpha) == 2
        assert ops.fused_fwd_kernel_kind(33, 10, 4, 64, 2.0) == 2
        assert ops.fused_fwd_kernel_kind(40, 10, 1, 64, 1.7) == 2
        assert ops.fused_fwd_kernel_kind(32, 10, 4, 128, 1.7) == 1       # field bucket not compiled
        assert ops.fused_fwd_kernel_kind(39, 16, 4, 128, 1.7) == 2       # nemb 16 instance (config 4)
        assert ops.fused_fwd_kernel_kind(39, 12, 4, 128, 1.7) == 1       # nemb bucket not compiled
        assert ops.fused_fwd_kernel_kind(39, 10, 4, 100, 1.7) == 1       # K*O % 64 != 0
        assert ops.fused_fwd_kernel_kind(39, 10, 4, 128, 2.5) == 1       # alpha > 2 -> literal bisection
        assert ops.fused_fwd_kernel_kind(39, 10, 4, 128, 1.7, ops.SOLVER_BISECT) == 1
        ops.set_tuning('mma', 0)
        assert ops.fused_fwd_kernel_kind(39, 10, 4, 128, 1.7) == 1
        assert ops.fused_fwd_kernel_kind(39, 16, 4, 128, 1.7) == 1
        with pytest.raises(Exception):
            ops.set_tuning('no_such_switch', 1)
    finally:
        ops.set_tuning('tmem', -1)
        ops.set_tuning('mma', -1)
