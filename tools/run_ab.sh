bash tools/ab_hot.sh default a0 a1
ARMNET_B200_LIB=$PWD/armnet_b200/tuning/liba1.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
