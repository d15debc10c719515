#!/bin/bash
# A/B of tuning builds on one box: tools/ab_hot.sh base new ...   (names of armnet_b200/tuning/libNAME.so; "default" = shipped lib)
for rep in 1 2; do
for n in "$@"; do
  if [ "$n" == default ]; then python tools/time_hot.py --tag default; else ARMNET_B200_LIB=$PWD/armnet_b200/tuning/lib$n.so python tools/time_hot.py --tag $n; fi
done
done
