#!/bin/bash
# round 2, GPU visit U: encode_l1 with one warp per CTA (MZ_ENC_L1_WARPS=1: constant shared-memory
# addresses, no spills), the packed re-match walk (MZ_ENC_RM_PACK), combinations; checksums vs the product library
set -u
O=gpurun_out
mkdir -p $O
L=minlz_b200/libminlz_cuda
timeout 400 python profiles/ab_variants.py $L.so ${L}_w1.so ${L}_w2.so ${L}_rw.so ${L}_w1rw.so ${L}_w1pf3.so ${L}_w1rwext.so $L.so > $O/ab_variants_u.log 2>&1
cat $O/ab_variants_u.log | cut -c1-400
