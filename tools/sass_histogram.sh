#!/bin/bash
# SASS evidence: opcode histogram of the shipped library (sm_100a), the mnemonics that show TMA / mbarrier / cluster /
# packed-fp32 code is really there.   tools/sass_histogram.sh > profiles/rN_sass_opcodes.txt
LIB=${1:-tf_seq2seq_losses_b200/libctc_b200.so}
echo "# cuobjdump -sass $LIB | opcode histogram ($(date -u +%Y-%m-%d))"
cuobjdump -sass "$LIB" > /tmp/sass_all.txt
echo "# architectures: $(grep -o 'arch = sm_[0-9a-z]*' /tmp/sass_all.txt | sort | uniq -c | tr '\n' ' ')"
echo "# kernels: $(grep -c 'Function :' /tmp/sass_all.txt)"
for op in UBLKCP SYNCS.PHASECHK SYNCS.ARRIVE SYNCS.EXCH ARRIVES.LDGSTSBAR NANOSLEEP.SYNCS UCGABAR_ARV UCGABAR_WAIT MAPA LDS.*CLUSTER FADD2 FMUL2 FFMA2 MUFU.EX2 MUFU.LG2 ATOMS.CAST.SPIN REDUX CREDUX SHFL LDGSTS LDG.E.*128 STG.E.NA.128 LDS.128 STS.128 BAR.SYNC WARPSYNC UTCHMMA UTCQMMA HMMA LDL STL; do
  printf "%-40s %8d\n" "$op" "$(grep -c -E "^\s+/\*[0-9a-f]+\*/\s+.*\b($op)" /tmp/sass_all.txt)"
done
echo "# (UTC*MMA / HMMA = 0 is expected: the path has no contraction; LDL/STL = register spills)"
