#!/bin/bash
# Per-kernel SASS evidence of the shipped library: tcgen05 / TMA / TMEM mnemonics (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG =
# TMA load / store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = bulk copy, SYNCS = mbarrier) and legacy HMMA.
#   tools/sass_summary.sh > profiles/<name>_sass_summary.txt
LIB=${1:-viscy_b200/libviscy_b200.so}
echo "# cuobjdump -sass $LIB ($(date -u +%Y-%m-%d)), sm_100a; counts of instructions per kernel"
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3 }
/^[ \t]+\/\*[0-9a-f]+\*\// {
  op=$2; sub(/\..*/, "", op); gsub(/;/, "", op);
  if (op ~ /^@/) { op=$3; sub(/\..*/, "", op); gsub(/;/, "", op) }
  tot[fn]++
  if (op=="UTCHMMA"||op=="UTCQMMA"||op=="UTCOMMA") mma[fn]++
  if (op=="UTMALDG") tl[fn]++
  if (op=="UTMASTG") ts[fn]++
  if (op=="LDTM") lt[fn]++
  if (op=="UTCBAR") cb[fn]++
  if (op=="UBLKCP") bc[fn]++
  if (op=="SYNCS") sy[fn]++
  if (op=="HMMA") hm[fn]++
  if (op=="FFMA2"||op=="FMUL2"||op=="FADD2") f2[fn]++
}
END {
  printf "%-110s %7s %7s %7s %7s %6s %6s %6s %6s %5s %6s\n", "kernel (mangled)", "instrs", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "HMMA", "F*2"
  for (f in tot) { T+=tot[f]; M+=mma[f]; L+=tl[f]; S+=ts[f]; D+=lt[f]; B+=cb[f]; K+=bc[f]; Y+=sy[f]; Hh+=hm[f]; F+=f2[f]; if (mma[f]>0) nm++ }
  printf "%-110s %7d %7d %7d %7d %6d %6d %6d %6d %5d %6d\n", "TOTAL (" length(tot) " kernels, " nm " with UTCHMMA)", T, M, L, S, D, B, K, Y, Hh, F
  for (f in tot) printf "%-110s %7d %7d %7d %7d %6d %6d %6d %6d %5d %6d\n", substr(f,1,110), tot[f], mma[f], tl[f], ts[f], lt[f], cb[f], bc[f], sy[f], hm[f], f2[f] | "sort"
}'
