#!/bin/bash
# SASS census of the shipped library: per kernel, how many tcgen05 MMAs (UTC*MMA), TMA loads/prefetches (UTMALDG/UTMAPF/UBLKCP),
# TMEM loads/stores (LDTM/STTM), tcgen05.commit (UTCBAR) and legacy mma.sync (HMMA) instructions it contains.
# usage: scripts/sass_census.sh [lib] > profiles/r2_sass_census.txt
LIB=${1:-ultravox_b200/libuvx.so}
echo "# cuobjdump -sass $LIB ($(date -u +%F)); columns: UTC*MMA UTMALDG UTMAPF LDTM STTM UTCBAR HMMA  kernel"
cuobjdump -sass "$LIB" | awk '
/Function :/ { if (name != "") printf "%6d %7d %6d %5d %5d %6d %5d  %s\n", mma, tma, pf, ldtm, sttm, bar, hmma, name; name=$3; mma=tma=pf=ldtm=sttm=bar=hmma=0 }
/UTC[A-Z]*MMA/ { mma++ } /UTMALDG/ { tma++ } /UTMAPF|UBLKPF/ { pf++ } /LDTM/ { ldtm++ } /STTM/ { sttm++ } /UTCBAR/ { bar++ } / HMMA/ { hmma++ }
END { printf "%6d %7d %6d %5d %5d %6d %5d  %s\n", mma, tma, pf, ldtm, sttm, bar, hmma, name }' | while read a b c d e f g n; do printf "%6s %7s %6s %5s %5s %6s %5s  %s\n" $a $b $c $d $e $f $g "$(echo $n | c++filt | cut -c1-110)"; done | sort -k8
