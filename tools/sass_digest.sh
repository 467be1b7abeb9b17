#!/bin/bash
# Mnemonic digest of the shipped library: which kernels use the tensor core (UTCHMMA), TMEM loads / stores (LDTM / STTM),
# bulk copies (UBLKCP), packed fp32 (FFMA2) and register reallocation (USETMAXREG).
cd "$(dirname "$0")/.."
OUT=profiles/r02_sass_digest.txt
{
  echo "cuobjdump -sass cosypose_b200/libcosyb200.so  (sm_100a only)"
  /usr/local/cuda/bin/cuobjdump -lelf cosypose_b200/libcosyb200.so | head -3
  echo
  printf "%-44s %8s %6s %6s %7s %7s %7s %10s\n" kernel UTCHMMA LDTM STTM UBLKCP FFMA2 LDGSTS USETMAXREG
  /usr/local/cuda/bin/cuobjdump -sass cosypose_b200/libcosyb200.so | awk '
    /Function :/ { if (name != "") flush(); name=$3; delete c }
    { for (m in want) if (index($0, m)) c[m]++ }
    function flush() { n=name; gsub(/^_ZN5cosyb/, "", n); printf "%-44s %8d %6d %6d %7d %7d %7d %10d\n", substr(n,1,44), c["UTCHMMA"], c["LDTM"], c["STTM"], c["UBLKCP"], c["FFMA2"], c["LDGSTS"], c["USETMAXREG"] }
    BEGIN { want["UTCHMMA"]; want["LDTM"]; want["STTM"]; want["UBLKCP"]; want["FFMA2"]; want["LDGSTS"]; want["USETMAXREG"] }
    END { flush() }' | sort | awk '$2+$3+$4+$5+$6+$8 > 0'
} > $OUT
wc -l $OUT
