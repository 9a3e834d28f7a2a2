#!/bin/bash
# round 2, GPU visit V: one warp per CTA as the default (L1), packed re-match walk per flavour,
# one warp per CTA for the LevelBalanced kernels; checksums against the four-warp build of the morning
set -u
O=gpurun_out
mkdir -p $O
L=minlz_b200/libminlz_cuda
AB_L2_REPS=2 timeout 400 python profiles/ab_variants.py ${L}_w4.so $L.so ${L}_rm0.so ${L}_rm1.so ${L}_l2w1.so > $O/ab_variants_v.log 2>&1
cat $O/ab_variants_v.log | cut -c1-500
