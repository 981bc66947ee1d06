#!/bin/bash
# Blackwell-specific SASS opcodes in the built library (proof of tcgen05 / TMEM / TMA use): profiles/r02_sass_opcodes.txt
cd "$(dirname "$0")/.."
{
  echo "# cuobjdump -sass alphazero_openspiel_b200/libaz_b200.so | opcode histogram (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,"
  echo "# UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = bulk copy, UTCBAR = tcgen05.commit, SYNCS = mbarrier)"
  cuobjdump -sass alphazero_openspiel_b200/libaz_b200.so | grep -oE '\b(UTCHMMA|UTCQMMA|UTCOMMA|UTCMMA|LDTM|STTM|UTMALDG[.A-Z0-9_]*|UTMASTG[.A-Z0-9_]*|UBLKCP[.A-Z0-9_]*|UTCBAR[.A-Z0-9_]*|UTCCP[.A-Z0-9_]*|UTMAREDG[.A-Z0-9_]*|SYNCS[.A-Z0-9_]*|ELECT|FFMA2|FADD2|HMNMX2[.A-Z0-9_]*|HMUL2[.A-Z0-9_]*|HMMA[.A-Z0-9_]*)' | sort | uniq -c | sort -rn
  echo "# per kernel: tcgen05.mma count"
  cuobjdump -sass alphazero_openspiel_b200/libaz_b200.so | awk '/Function :/ {fn=$3} /UTCHMMA/ {c[fn]++} END {for (f in c) print c[f], f}' | sort -rn
} > profiles/r02_sass_opcodes.txt
cat profiles/r02_sass_opcodes.txt
