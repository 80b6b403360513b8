#!/bin/bash
# Opcode histogram of the built library (cuobjdump -sass): the Blackwell-only instructions the kernels are made of.
# usage: tools/sass_histogram.sh > profiles/r02_sass_opcodes.txt
cd "$(dirname "$0")/.."
SO=xeofs_b200/libxeofs_b200.so
echo "# cuobjdump -sass $SO  ($(date -u +%Y-%m-%dT%H:%MZ), git $(git rev-parse --short HEAD))"
echo "# arch: $(cuobjdump -lelf $SO | head -3 | tr '\n' ' ')"
cuobjdump -sass $SO > /tmp/xeofs_sass.txt
echo "# tensor-core / TMEM / TMA / bulk-copy opcodes (count over all kernels)"
for op in UTCHMMA UTCQMMA UTCBAR LDTM STTM UTMALDG UTMASTG UBLKCP UBLKPF UTCATOMSWS SYNCS ELECT DMMA; do
  printf "%-12s %6d\n" $op $(grep -c -E "[[:space:]]$op(\.|[[:space:]])" /tmp/xeofs_sass.txt)
done
echo "# kernels holding UTCHMMA (tcgen05.mma):"
awk '/Function :/ {fn=$3} /UTCHMMA/ {c[fn]++} END {for (f in c) printf "%6d  %s\n", c[f], f}' /tmp/xeofs_sass.txt | sort -k2 | c++filt | cut -c1-160
echo "# top 25 opcodes overall"
grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" /tmp/xeofs_sass.txt | awk '{print $2}' | cut -d. -f1 | sort | uniq -c | sort -rn | head -25
