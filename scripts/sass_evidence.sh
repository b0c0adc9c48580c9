#!/bin/bash
# SASS mnemonics that prove which hardware paths the built library uses -> profiles/r2_sass_evidence.txt
so=poismf_b200/libpoismf_b200.so
sass=$(mktemp)
cuobjdump -sass $so > $sass
{
echo "# SASS evidence in $so (cuobjdump -sass | grep -c)"
for m in UTCHMMA LDTM UTCBAR UTMALDG UBLKCP SYNCS CREDUX LDGSTS ATOMS FMNMX3; do echo "$m $(grep -c "$m" $sass)"; done
for pair in "UTMALDG:TMA tensor copy, cp.async.bulk.tensor.2d" "UTCHMMA:tcgen05.mma" "LDTM:tcgen05.ld" "UBLKCP:TMA bulk copy, cp.async.bulk" "CREDUX:redux.sync.min.f32"; do
  m=${pair%%:*}; what=${pair#*:}
  echo; echo "# kernels containing $m ($what):"
  awk -v m="$m" '/Function : /{f=$3} index($0, m){c[f]++} END{for (k in c) printf "%7d %s\n", c[k], k}' $sass | sort -k2 | head -40
done
} > profiles/r2_sass_evidence.txt
rm -f $sass
