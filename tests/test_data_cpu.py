"""Input pipeline host logic (armnet_b200/data.py + the native libsvm parser of the C ABI): no GPU needed."""
import os

import numpy as np
import pytest
import torch


def _write(path, rows, extra_lines=()):
    with open(path, 'w') as f:
        for i, (y, ids, vals) in enumerate(rows):
            f.write(f'{y:g} ' + ' '.join(f'{a}:{b:g}' for a, b in zip(ids, vals)) + '\n')
            if i < len(extra_lines):
                f.write(extra_lines[i] + '\n')


def _python_reference_parse(path, nfield):
    """data_loader.py:27-47 restated: per line `label id:val ...`; lines that fail are skipped."""
    ids, vals, ys = [], [], []
    for line in open(path):
        try:
            cols = line.split(' ')
            pairs = [c.split(':') for c in cols[1:]]
            i, v = [int(a) for a, _ in pairs], [float(b) for _, b in pairs]
            if len(i) != nfield:
                raise ValueError
            y = float(cols[0])
        except Exception:
            continue
        ids.append(i), vals.append(v), ys.append(y)
    return np.array(ids, np.int64).reshape(-1, nfield), np.array(vals, np.float32).reshape(-1, nfield), np.array(ys, np.float32)


def test_native_parser_matches_reference_semantics(tmp_path):
    from armnet_b200.data import parse_libsvm
    rng = np.random.default_rng(0)
    F = 10
    rows = [(int(rng.integers(0, 2)), rng.integers(0, 5382, F).tolist(), np.round(rng.random(F), 4).tolist())
            for _ in range(500)]
    bad = ['', '1 3:1 4:1', 'x 1:1 2:1 3:1 4:1 5:1 6:1 7:1 8:1 9:1 10:1', '0 1:1 2:1 3:1 4:1 5:1 6:1 7:1 8:1 9:1 10',
           '1 1:1 2:1 3:1 4:1 5:1 6:1 7:1 8:1 9:1 10:1 11:1']
    p = str(tmp_path / 'train.libsvm')
    _write(p, rows, bad)
    ids, vals, y, skipped = parse_libsvm(p, F)
    rid, rval, ry = _python_reference_parse(p, F)
    assert skipped == len(bad)
    assert ids.dtype == np.int32 and np.array_equal(ids.astype(np.int64), rid)
    assert np.array_equal(vals, rval) and np.array_equal(y, ry)
    with pytest.raises(Exception):
        parse_libsvm(str(tmp_path / 'missing.libsvm'), F)


def test_frappe_style_line_and_no_trailing_newline(tmp_path):
    from armnet_b200.data import parse_libsvm
    p = str(tmp_path / 'a.libsvm')
    with open(p, 'w') as f:                                    # data/frappe/train.libsvm:1 format, values all 1
        f.write('-1 451:1 4149:1 5041:1 5046:1 5053:1 5055:1 5058:1 5060:1 5073:1 5183:1\n')
        f.write('1 1:0.5 2:1e-3 3:1 4:1 5:1 6:1 7:1 8:1 9:1 10:2.5')
    ids, vals, y, skipped = parse_libsvm(p, 10)
    assert skipped == 0 and ids.shape == (2, 10) and y.tolist() == [-1.0, 1.0]
    assert ids[0].tolist() == [451, 4149, 5041, 5046, 5053, 5055, 5058, 5060, 5073, 5183]
    assert vals[1, 1] == np.float32(1e-3) and vals[1, 9] == np.float32(2.5)


def test_binary_cache_round_trip_and_invalidation(tmp_path):
    from armnet_b200.data import load_split
    F = 4
    p = str(tmp_path / 'valid.libsvm')
    _write(p, [(1, [1, 2, 3, 4], [1, 1, 1, 1]), (0, [5, 6, 7, 8], [0.5, 0.25, 1, 1])])
    a = load_split(p, F)
    assert os.path.exists(p + '.armnet_bin/meta.json')
    b = load_split(p, F)                                        # served from the memmapped cache
    for x, z in zip(a, b):
        assert np.array_equal(np.asarray(x), np.asarray(z))
    assert isinstance(b[0], np.memmap)
    _write(p, [(1, [9, 9, 9, 9], [1, 1, 1, 1])])                # the text file changed -> cache rebuilt
    os.utime(p, ns=(1, 1))
    c = load_split(p, F)
    assert np.asarray(c[0]).tolist() == [[9, 9, 9, 9]]


def test_device_split_batches_cover_every_row_once_and_shard_like_parallel():
    from armnet_b200.data import DeviceSplit
    from armnet_b200.parallel import shard_bounds
    n, F = 103, 5
    ids = np.arange(n * F, dtype=np.int32).reshape(n, F)
    split = DeviceSplit(ids, np.ones((n, F), np.float32), np.arange(n, dtype=np.float32), 'cpu')
    seen = torch.cat([b['y'] for b in split.batches(16, shuffle=True, generator=torch.Generator().manual_seed(3))])
    assert sorted(seen.tolist()) == list(range(n))
    # two ranks walking the same seeded order partition every global batch exactly like parallel.shard_bounds
    parts = [list(split.batches(16, True, torch.Generator().manual_seed(3), rank=r, world=2)) for r in range(2)]
    full = list(split.batches(16, True, torch.Generator().manual_seed(3)))
    for g, a, b in zip(full, *parts):
        assert torch.equal(torch.cat([a['y'], b['y']]), g['y']) and a['global_rows'] == g['y'].numel()
        lo, hi = shard_bounds(g['y'].numel(), 0, 2)
        assert a['y'].numel() == hi - lo
        assert torch.equal(a['id'], g['id'][lo:hi])


@pytest.mark.skipif(not os.path.exists('/root/reference/data/frappe/test.libsvm'),
                    reason='bundled Frappe data only exists in the dev container')
def test_native_parser_equals_reference_loader_on_frappe(tmp_path):
    """The unmodified reference's LibsvmDataset (data_loader.py:12-55) vs the native parser on the head of the bundled
    Frappe test split."""
    import sys
    from armnet_b200.data import parse_libsvm
    head = str(tmp_path / 'head.libsvm')
    with open('/root/reference/data/frappe/test.libsvm') as f, open(head, 'w') as g:
        for _, line in zip(range(3000), f):
            g.write(line)
    sys.path.insert(0, '/root/reference')
    try:
        from data_loader import LibsvmDataset
        ref = LibsvmDataset(head, 10)
    finally:
        sys.path.remove('/root/reference')
    ids, vals, y, skipped = parse_libsvm(head, 10)
    n = ref.nsamples
    assert skipped == 0 and ids.shape[0] == n == 3000
    assert np.array_equal(ids.astype(np.int64), ref.feat_id[:n].numpy())
    assert np.array_equal(vals, ref.feat_value[:n].numpy()) and np.array_equal(y, ref.y[:n].numpy())


def test_parser_edge_cases(tmp_path):
    """Empty and all-bad files, CRLF line ends, the largest id the int32 arrays hold, a single field -- same rows as
    data_loader.py:27-47 yields.  One documented leniency: the reference's `line.split(' ')` drops a line with a doubled
    or trailing blank (an empty token fails its `id:val` unpacking); the native parser tolerates the extra blanks."""
    from armnet_b200.data import parse_libsvm

    def write(name, text):
        p = str(tmp_path / name)
        with open(p, 'w', newline='') as f:
            f.write(text)
        return p

    ids, vals, y, skipped = parse_libsvm(write('empty.libsvm', ''), 3)
    assert ids.shape == (0, 3) and vals.shape == (0, 3) and y.shape == (0,) and skipped == 0
    ids, vals, y, skipped = parse_libsvm(write('bad.libsvm', '\n\nfoo\n1 2:3\n'), 3)
    assert ids.shape == (0, 3) and skipped == 4
    p = write('crlf.libsvm', '1 1:1 2:0.5 3:1\r\n0 4:1 5:1 6:1\r\n')
    ids, vals, y, skipped = parse_libsvm(p, 3)
    rid, rval, ry = _python_reference_parse(p, 3)
    assert skipped == 0 and np.array_equal(ids, rid) and np.array_equal(vals, rval) and np.array_equal(y, ry)
    ids, vals, y, skipped = parse_libsvm(write('big.libsvm', '1 2147483647:1 0:0 7:1e-30\n'), 3)
    assert ids.tolist() == [[2147483647, 0, 7]] and vals[0, 2] == np.float32(1e-30) and skipped == 0
    ids, vals, y, skipped = parse_libsvm(write('one.libsvm', '0 5:2\n1 6:3\n'), 1)
    assert ids.tolist() == [[5], [6]] and vals.tolist() == [[2.0], [3.0]] and y.tolist() == [0.0, 1.0]
    p = write('blanks.libsvm', '1  1:1 2:1 3:1\n1 1:1 2:1 3:1 \n')
    ids, vals, y, skipped = parse_libsvm(p, 3)
    assert ids.shape == (2, 3) and skipped == 0                  # reference: both lines skipped
    assert _python_reference_parse(p, 3)[0].shape == (0, 3)
