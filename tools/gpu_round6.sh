#!/bin/bash
mkdir -p gpurun_out data/tiny
python - <<'PY'
import numpy as np
rng = np.random.default_rng(0)
w = rng.normal(size=600)
for name, n in (('train', 20000), ('valid', 3000), ('test', 3000)):
    ids = rng.integers(0, 600, (n, 10)); s = w[ids].sum(1) / 3
    y = (rng.random(n) < 1 / (1 + np.exp(-s))).astype(int)
    with open(f'data/tiny/{name}.libsvm', 'w') as f:
        for i in range(n):
            f.write(f'{y[i]} ' + ' '.join(f'{a}:1' for a in ids[i]) + '\n')
PY
echo "--- train.py on a libsvm dataset (native parser + device-resident splits), armnet_1h"
timeout 600 python train.py --model armnet_1h --nfield 10 --nfeat 600 --nemb 10 --h 10 --alpha 1.7 --lr 0.003 --dataset tiny --data_dir data/ --batch_size 1024 --epoch 3 --patience 3 --report_freq 100 --exp_name tiny1h --log_dir gpurun_out/log/ 2>&1 | grep -E "resident|train\s|val\s|test\s|best|Total" | tail -12
echo "--- same with --host_data"
timeout 600 python train.py --model armnet_1h --nfield 10 --nfeat 600 --nemb 10 --h 10 --alpha 1.7 --lr 0.003 --dataset tiny --data_dir data/ --batch_size 1024 --epoch 1 --host_data --report_freq 100 --exp_name tiny1h_host --log_dir gpurun_out/log/ 2>&1 | grep -E "train\s|val\s|Total" | tail -4
echo "--- zoo model through train.py"
timeout 600 python train.py --model xdfm --nfield 10 --nfeat 600 --nemb 10 --h 8 --k 2 --dataset tiny --data_dir data/ --batch_size 1024 --epoch 1 --report_freq 100 --exp_name tinyxdfm --log_dir gpurun_out/log/ 2>&1 | grep -E "train\s|val\s|Total" | tail -4
rm -rf data/tiny
bash tools/gpu_sanitize.sh 2>&1 | grep -E "rc=|passed|failed|SUMMARY"
