"""Input pipeline for the hot path (replaces data_loader.py:12-73 + the per-batch H2D of
train.py:101-106 when a split fits in HBM).

  libsvm text --(armnet_libsvm_parse, native)--> dense arrays --(cache)--> <file>.armnet_bin/{ids.i32,values.f32,y.f32}
  --> DeviceSplit: the whole split resident on the GPU (Criteo: 45 M rows x 39 fields x 8 B = 14 GB of a B200's
  180 GB), batches cut by an index permutation ON the device -- no DataLoader workers, no pinned staging, no per-batch
  host work.  ids are kept as int32 (the kernels take int32 or int64).

The reference re-parses the text files at every start (data_loader.py:25-47) with 4 worker processes feeding pinned
batches; its batch dict {'id', 'value', 'y'} (data_loader.py:52-55) is what DeviceSplit.batches() yields.
"""
import ctypes as C
import json
import os

import numpy as np
import torch

from ._capi import check, lib

__all__ = ['parse_libsvm', 'load_split', 'save_binary', 'load_binary', 'DeviceSplit']

_META = 'meta.json'


def parse_libsvm(path, nfield):
    """`label id:val ...` per line -> (ids [N,F] int32, values [N,F] f32, y [N] f32, n_skipped). Lines without exactly
    nfield well-formed pairs are skipped (data_loader.py:37-44)."""
    n = C.c_int64(0)
    check(lib.armnet_libsvm_count_lines(os.fsencode(path), C.byref(n)), 'armnet_libsvm_count_lines')
    cap = int(n.value)
    ids = np.empty((cap, nfield), dtype=np.int32)
    vals = np.empty((cap, nfield), dtype=np.float32)
    y = np.empty((cap,), dtype=np.float32)
    rows, bad = C.c_int64(0), C.c_int64(0)
    check(lib.armnet_libsvm_parse(os.fsencode(path), nfield, cap, ids.ctypes.data, vals.ctypes.data, y.ctypes.data,
                                  C.byref(rows), C.byref(bad)), 'armnet_libsvm_parse')
    r = int(rows.value)
    return ids[:r], vals[:r], y[:r], int(bad.value)


def save_binary(dirname, ids, vals, y, source=None):
    os.makedirs(dirname, exist_ok=True)
    np.ascontiguousarray(ids, dtype=np.int32).tofile(os.path.join(dirname, 'ids.i32'))
    np.ascontiguousarray(vals, dtype=np.float32).tofile(os.path.join(dirname, 'values.f32'))
    np.ascontiguousarray(y, dtype=np.float32).tofile(os.path.join(dirname, 'y.f32'))
    meta = {'rows': int(ids.shape[0]), 'nfield': int(ids.shape[1])}
    if source is not None:
        st = os.stat(source)
        meta.update(source_size=st.st_size, source_mtime_ns=st.st_mtime_ns)
    with open(os.path.join(dirname, _META), 'w') as f:
        json.dump(meta, f)


def load_binary(dirname, mmap=True):
    with open(os.path.join(dirname, _META)) as f:
        meta = json.load(f)
    n, nf = meta['rows'], meta['nfield']

    def arr(name, dtype, shape):
        p = os.path.join(dirname, name)
        if n == 0:
            return np.empty(shape, dtype=dtype)
        return np.memmap(p, dtype=dtype, mode='r', shape=shape) if mmap else np.fromfile(p, dtype=dtype).reshape(shape)

    return arr('ids.i32', np.int32, (n, nf)), arr('values.f32', np.float32, (n, nf)), arr('y.f32', np.float32, (n,)), meta


def load_split(path, nfield, cache=True, mmap=True):
    """One libsvm file as dense arrays, through the binary cache next to it (rebuilt when the text file changed)."""
    cdir = path + '.armnet_bin'
    if cache and os.path.exists(os.path.join(cdir, _META)):
        ids, vals, y, meta = load_binary(cdir, mmap=mmap)
        st = os.stat(path) if os.path.exists(path) else None
        fresh = st is None or (meta.get('source_size') == st.st_size and meta.get('source_mtime_ns') == st.st_mtime_ns)
        if fresh and meta['nfield'] == nfield:
            return ids, vals, y
    ids, vals, y, _ = parse_libsvm(path, nfield)
    if cache:
        try:
            save_binary(cdir, ids, vals, y, source=path)
        except OSError:
            pass                                   # read-only data directory: parse every time
    return ids, vals, y


class DeviceSplit:
    """A whole split resident on one device; batches are gathered there from an index permutation."""

    def __init__(self, ids, vals, y, device):
        as_t = lambda a: torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
        self.ids = as_t(ids).to(device)
        self.values = as_t(vals).to(device=device, dtype=torch.float32)
        self.y = as_t(y).to(device=device, dtype=torch.float32)
        self.device = torch.device(device)

    def __len__(self):
        return self.y.shape[0]

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.ids, self.values, self.y))

    def batches(self, bsz, shuffle=False, generator=None, rank=0, world=1):
        """Yields {'id','value','y'} device batches (data_loader.py:52-55). With world > 1 every rank walks the SAME
        global order (same seeded generator) and keeps its contiguous share of each global batch, like
        parallel.shard_batch."""
        n = len(self)
        if shuffle:
            order = torch.randperm(n, generator=generator).to(self.device)     # 8 B per row, once per epoch
        else:
            order = torch.arange(n, device=self.device)
        for i in range(0, n, bsz):
            sel = order[i:i + bsz]
            if world > 1:
                base, rem = divmod(sel.numel(), world)
                lo = rank * base + min(rank, rem)
                sel = sel[lo:lo + base + (1 if rank < rem else 0)]
            yield {'id': self.ids[sel], 'value': self.values[sel], 'y': self.y[sel], 'global_rows': min(bsz, n - i)}
