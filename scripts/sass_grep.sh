#!/usr/bin/env bash
# Census of the Blackwell-native SASS in the built library (no GPU needed): writes profiles/r02_sass_grep.txt
set -eu
cd "$(dirname "$0")/.."
LIB=pyprob_b200/lib/libpyprob_b200.so
TMP=$(mktemp)
cuobjdump -sass "$LIB" > "$TMP"
{
  echo "# cuobjdump -sass $LIB (sm_100a): instructions per mnemonic (word match), round 2"
  for m in UTCHMMA UTCQMMA LDTM STTM UTCBAR UTCATOMSWS UBLKCP UTMALDG UTMASTG SYNCS UCGABAR_ARV UCGABAR_WAIT PREEXIT ACQBULK LDGSTS HMMA HGMMA; do
    printf "%-14s %s\n" "$m" "$(grep -cE "[^A-Z]$m[. ]" "$TMP" || true)"
  done
  echo "# PREEXIT = griddepcontrol.launch_dependents, ACQBULK = griddepcontrol.wait (programmatic dependent launch)"
  echo
  echo "# tcgen05.mma (UTCHMMA) per kernel:"
  awk '/Function :/{f=$3} /[^A-Z]UTCHMMA/{c[f]++} END{for(k in c) print c[k], k}' "$TMP" | sort -k2 | c++filt | cut -c1-140
  echo
  echo "# cluster barriers (UCGABAR_ARV: barrier.cluster.arrive) per kernel — distributed-shared-memory split-K kernels:"
  awk '/Function :/{f=$3} /UCGABAR_ARV/{c[f]++} END{for(k in c) print c[k], k}' "$TMP" | sort -k2 | c++filt | cut -c1-140
} > profiles/r02_sass_grep.txt
rm -f "$TMP"
cat profiles/r02_sass_grep.txt
