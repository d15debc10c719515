for v in t0000 t11; do echo "== trace $v"; ARMNET_B200_LIB=$PWD/armnet_b200/tuning/lib$v.so python tools/trace_tmem.py --regime init 2>&1 | tail -24; done
bash tools/ab_hot.sh base n0000 n1000 n0100 n1100 n1111
