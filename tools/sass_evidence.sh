#!/bin/sh
# SASS evidence for the design claims (no GPU needed): instruction counts per kernel from cuobjdump -sass of the built library.
LIB=openmm_agbnp_plugin_b200/lib/libagbnp_b200.so
out=profiles/r2_sass_evidence.txt
echo "# cuobjdump -sass $LIB (sm_100a), instruction counts per kernel; tools/sass_evidence.sh" > $out
for k in k_gb k_tree k_deriv k_born k_tree_gamma k_peer_allreduce_final; do
  for f in $(cuobjdump -sass $LIB 2>/dev/null | grep "Function :" | grep "$k" | sed 's/.*Function : //'); do
    cuobjdump -sass -fun "$f" $LIB > /tmp/sass_one.txt 2>/dev/null
    tot=$(grep -c "^\s*/\*[0-9a-f]\{4\}\*/" /tmp/sass_one.txt)
    echo "" >> $out
    echo "== $(echo $f | c++filt)   ($tot SASS instructions)" >> $out
    for pat in "FFMA2" "FMUL2" "FADD2" "FFMA " "MUFU.EX2" "MUFU.RSQ" "MUFU.RCP" "MUFU.SQRT" "MUFU.LG2" "DFMA" "DMUL" "DADD" "LDGSTS" "LDS" "STS" "LDG" "STG" "RED.E" "REDG" "ATOMG" "ATOMS" "REDUX" "SHFL" "VOTE" "F2I" "I2F" "BAR.SYNC" "WARPSYNC" "ACQBULK\|griddepcontrol\|PREEXIT"; do
      c=$(grep -c "$pat" /tmp/sass_one.txt)
      [ "$c" -gt 0 ] && echo "   $pat: $c" >> $out
    done
    grep "RED\.\|REDG\|ATOM" /tmp/sass_one.txt | sed 's/^\s*//' | cut -c1-110 | sort | uniq -c | sort -rn | head -4 | sed 's/^/      /' >> $out
  done
done
echo "" >> $out
echo "# excerpt: inner loop of k_gb<false> (packed FFMA2/FMUL2/FADD2 on register pairs, MUFU.EX2/RSQ, no MOV between loads and math)" >> $out
cuobjdump -sass -fun '_ZN15agbnp_b200_impl4k_gbILb0EEEvNS_6GBArgsE' $LIB 2>/dev/null | grep -n "FFMA2" | head -1 | cut -d: -f1 > /tmp/l0
l0=$(cat /tmp/l0); cuobjdump -sass -fun '_ZN15agbnp_b200_impl4k_gbILb0EEEvNS_6GBArgsE' $LIB 2>/dev/null | sed -n "$((l0-6)),$((l0+40))p" | sed 's/^\s*//' | cut -c1-120 >> $out
echo "" >> $out
echo "# excerpt: LDGSTS (cp.async) staging of the column tiles in k_gb<false>" >> $out
cuobjdump -sass -fun '_ZN15agbnp_b200_impl4k_gbILb0EEEvNS_6GBArgsE' $LIB 2>/dev/null | grep "LDGSTS\|LDGDEPBAR\|DEPBAR" | sed 's/^\s*//' | cut -c1-120 | head -12 >> $out
echo "" >> $out
echo "# excerpt: k_tree<true>: redux.or head mask of the candidate enumeration, FP64 chain of the exact phase" >> $out
cuobjdump -sass -fun '_ZN15agbnp_b200_impl6k_treeILb1EEEvNS_8TreeArgsE' $LIB 2>/dev/null | grep "REDUX\|MUFU.RCP64H\|MUFU.RSQ64H\|DFMA" | sed 's/^\s*//' | cut -c1-120 | head -14 >> $out
wc -l $out
