#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTCHMMA / UTCQMMA
# (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor load / store), UBLKCP (bulk copy), SYNCS (mbarrier).
LIB=${1:-speinet_b200/libspeinet_b200.so}
echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%dT%H:%MZ), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
printf "%-62s %8s %6s %6s %8s %8s %7s %6s %6s\n" kernel UTCHMMA LDTM STTM UTMALDG UTMASTG UBLKCP SYNCS SHFL
cuobjdump -sass "$LIB" | awk '
  /Function :/ { if (name != "") printf "%-62s %8d %6d %6d %8d %8d %7d %6d %6d\n", name, m, l, s, tl, ts, b, y, h; name=$3; m=l=s=tl=ts=b=y=h=0 }
  /UTCHMMA|UTCQMMA|UTCOMMA/ {m++} /LDTM/ {l++} /STTM/ {s++} /UTMALDG/ {tl++} /UTMASTG/ {ts++} /UBLKCP/ {b++} /SYNCS/ {y++} /SHFL/ {h++}
  END { printf "%-62s %8d %6d %6d %8d %8d %7d %6d %6d\n", name, m, l, s, tl, ts, b, y, h }' | while read -r line; do
    n=$(echo "$line" | awk '{print $1}' | c++filt | sed 's/(.*//; s/^void //; s/spei:://' | cut -c1-60)
    echo "$line" | awk -v n="$n" '{printf "%-62s %8d %6d %6d %8d %8d %7d %6d %6d\n", n, $2, $3, $4, $5, $6, $7, $8, $9}'
  done
