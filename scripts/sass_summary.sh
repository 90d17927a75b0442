#!/bin/bash
# Per-kernel counts of the SASS mnemonics that tell the tensor-core / TMA generation apart (B200_PROFILING.md):
#   UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit,
#   HMMA = mma.sync (legacy warp-level tensor path), LDGSTS = cp.async.
cd "$(dirname "$0")/.."
echo "# cuobjdump -sass pipeedge_b200/libpipeedge_b200.so (sm_100a), instruction counts per kernel"
printf "%-70s %8s %8s %8s %6s %6s %7s %6s %7s\n" kernel UTCHMMA UTMALDG UTMASTG LDTM STTM UTCBAR HMMA LDGSTS
cuobjdump -sass pipeedge_b200/libpipeedge_b200.so | awk '
  /Function :/ { if (name != "") print_row(); name=$3; delete c }
  { for (k in keys) if (index($0, k)) c[k]++ }
  function print_row() { printf "%-70s %8d %8d %8d %6d %6d %7d %6d %7d\n", substr(name,1,70), c["UTCHMMA"], c["UTMALDG"], c["UTMASTG"], c["LDTM"], c["STTM"], c["UTCBAR"], c["HMMA."], c["LDGSTS"] }
  BEGIN { keys["UTCHMMA"]; keys["UTMALDG"]; keys["UTMASTG"]; keys["LDTM"]; keys["STTM"]; keys["UTCBAR"]; keys["HMMA."]; keys["LDGSTS"] }
  END { print_row() }' | c++filt | sort
