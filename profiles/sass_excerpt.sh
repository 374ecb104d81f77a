#!/bin/bash
# SASS evidence for the FP64 tensor-core / TMA claims (runs here, no GPU needed):
#   profiles/sass_excerpt.sh > profiles/r02_sass_excerpt.txt
SO=${1:-tdvmc_b200/libtdvmc_b200.so}
echo "# cuobjdump -sass $SO  ($(git rev-parse --short HEAD 2>/dev/null))"
echo "# architectures in the fat binary:"
cuobjdump -lelf $SO | sed 's/^/#   /'
cuobjdump -sass $SO > /tmp/_sass.txt
echo
echo "# instruction counts over the whole library"
for m in DMMA UBLKCP 'SYNCS' 'MUFU.RSQ64H' DFMA DADD DMUL 'LDS.128' 'LDS.64' 'ATOMS' 'REDUX' 'SHFL' 'MATCH'; do
  printf "%-14s %8d\n" "$m" "$(grep -c "[[:space:]]$m" /tmp/_sass.txt)"
done
echo
echo "# per kernel: DMMA / UBLKCP / SYNCS / DFMA counts"
awk '/Function :/{k=$3} /DMMA/{d[k]++} /UBLKCP/{u[k]++} /SYNCS/{s[k]++} /DFMA/{f[k]++} /Function :/{seen[k]=1}
     END{for(k in seen) printf "%-90s DMMA=%-4d UBLKCP=%-3d SYNCS=%-3d DFMA=%d\n", k, d[k], u[k], s[k], f[k]}' /tmp/_sass.txt | sort
echo
echo "# syrk_kernel: first lines carrying each mnemonic (address, instruction)"
awk '/Function :/{on=($3 ~ /syrk_kernel/)} on' /tmp/_sass.txt > /tmp/_syrk.txt
for m in UBLKCP 'SYNCS' DMMA; do grep -m 6 "$m" /tmp/_syrk.txt; done
