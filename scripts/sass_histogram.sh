#!/bin/bash
# SASS opcode evidence of the built library (no GPU needed): tcgen05 = UTC*MMA, TMEM ld/st = LDTM/STTM,
# TMA = UTMALDG/UTMASTG/UBLKCP, legacy tensor path = HMMA.  Usage: scripts/sass_histogram.sh > profiles/rN_sass_opcodes.txt
so=hnd_ghnd_object_detectors_b200/libghnd_b200.so
echo "# cuobjdump -sass $so  ($(date -u +%Y-%m-%dT%H:%MZ), $(git rev-parse --short HEAD))"
echo "# --- Blackwell-native opcodes over the whole library ---"
cuobjdump -sass $so | grep -oE "\b(UTC[A-Z]*MMA[A-Z0-9_.]*|UTCBAR[A-Z.]*|UTCATOMSWS[A-Z_.]*|LDTM[A-Z0-9_.x]*|STTM[A-Z0-9_.x]*|UTMALDG[A-Z0-9_.]*|UTMASTG[A-Z0-9_.]*|UTMAPF[A-Z0-9_.]*|UBLKCP[A-Z0-9_.]*|HMMA[A-Z0-9_.]*|SYNCS[A-Z0-9_.]*|LDGSTS[A-Z0-9_.]*|ATOMS[A-Z0-9_.]*|REDG[A-Z0-9_.]*)" | sort | uniq -c | sort -rn
echo "# --- per kernel: instructions, UTCHMMA, UTMALDG, UTMASTG, LDTM, HMMA ---"
cuobjdump -sass $so | awk '
/Function :/ { if (name != "") printf "%-110s %6d %5d %5d %5d %5d %5d\n", name, n, mma, ldg, stg, ldtm, hmma; name=$3; n=0; mma=0; ldg=0; stg=0; ldtm=0; hmma=0 }
/\/\*[0-9a-f]+\*\/ +[A-Z@]/ { n++ }
/UTC[A-Z]*MMA/ { mma++ } /UTMALDG/ { ldg++ } /UTMASTG/ { stg++ } /LDTM/ { ldtm++ } / HMMA/ { hmma++ }
END { printf "%-110s %6d %5d %5d %5d %5d %5d\n", name, n, mma, ldg, stg, ldtm, hmma }' | sort -k3 -n -r | awk '$3+$4+$5+$6+$7 > 0' | c++filt
