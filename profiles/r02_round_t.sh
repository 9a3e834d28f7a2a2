#!/bin/bash
# round 2, GPU visit T: encode_l1 source-ring variants (MZ_ENC_SRC_PF 1/2/3) and the extension prefetch
# (MZ_ENC_EXT_PF), one process, checksums against the product library
set -u
O=gpurun_out
mkdir -p $O
L=minlz_b200/libminlz_cuda
timeout 400 python profiles/ab_variants.py $L.so ${L}_pf1.so ${L}_pf2.so ${L}_pf3.so ${L}_ext.so ${L}_pf3ext.so ${L}_pf1ext.so $L.so > $O/ab_variants_t.log 2>&1
cat $O/ab_variants_t.log | cut -c1-400
