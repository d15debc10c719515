"""Build the binary cache of the reference's bundled Frappe splits (data/frappe/*.libsvm, 288 609 rows x 10 fields) under
data/frappe/ of THIS repo so that BASELINE config 1 can train on the GPU box, where /root/reference does not exist.
The cache (ids int32, values f32, labels f32; armnet_b200/data.py format) is git-ignored data, not source: it travels
with gpurun only.   python tools/make_frappe_cache.py [/root/reference/data/frappe]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from armnet_b200.data import parse_libsvm, save_binary

src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/data/frappe'
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'data', 'frappe')
os.makedirs(dst, exist_ok=True)
for name in ('train', 'valid', 'test'):
    ids, vals, y, meta = parse_libsvm(os.path.join(src, name + '.libsvm'), 10)
    save_binary(os.path.join(dst, name + '.libsvm.armnet_bin'), ids, vals, y, source=None)
    print(name, ids.shape, 'max id', int(ids.max()), 'positives', float(y.mean()))
