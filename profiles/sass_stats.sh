#!/bin/bash
# Static SASS statistics of the built library (no GPU needed): resource usage per kernel, spill and
# bulk-copy mnemonics.  usage: bash profiles/sass_stats.sh > profiles/r02_sass_stats.txt
L=${1:-minlz_b200/libminlz_cuda.so}
echo "# SASS statistics of $L (cuobjdump -sass / -res-usage; $(nvcc --version | tail -2 | head -1))"
echo
cuobjdump -lelf $L | head -3
echo
echo "## resource usage"
cuobjdump -res-usage $L 2>&1 | grep -A1 "Function" | grep -v "^--" | paste - - | sed 's/ Function //' | cut -c1-200
echo
echo "## whole library: instruction mnemonics of interest (static counts)"
T=$(mktemp)
cuobjdump -sass $L > $T
for m in "LDL" "STL" "UBLKCP" "LDGSTS" "LDG.E.*256" "STG.E.*256" "SYNCS" "ATOMS" "MATCH.ANY" "REDUX" "BSSY" "S2R"; do
  printf "%-14s %s\n" "$m" "$(grep -cE "\b$m" $T)"
done
echo
echo "## per kernel: static instructions / LDL / STL / BSSY"
for k in encode_l1_asm_kernelILb0 encode_l1_asm_kernelILb1 encode_l1_kernelILb0 encode_l1_kernelILb1 encode_l2_asm_kernel encode_l2_kernel decode_pc_kernel crc32c_blocks_kernel pack_blocks_kernel scan_lengths_kernel validate_compare_kernel validate_ranges_kernel; do
  K=$(awk -v k="$k" '$0 ~ "Function : .*"k{on=1;next} /Function :/{on=0} on' $T)
  printf "%-28s %6s instr  LDL %3s  STL %3s  BSSY %3s  UBLKCP %2s\n" $k "$(echo "$K" | grep -c '^\s*/\*[0-9a-f]*\*/')" "$(echo "$K" | grep -c LDL)" "$(echo "$K" | grep -c STL)" "$(echo "$K" | grep -c BSSY)" "$(echo "$K" | grep -c UBLKCP)"
done
rm -f $T
