#!/bin/bash
B=tools/tma_bw
# L2-resident source (64 MB) vs DRAM (2 GB); stages; issuers; box rows; fewer CTAs
for src in 64 2048; do
  for st in 4 6 8 12; do $B 128 $st 1 $src 148; done
  $B 128 6 2 $src 148
  $B 128 6 6 $src 148
  $B 64 12 1 $src 148
  $B 256 3 1 $src 148
  $B 128 6 1 $src 74
  $B 128 6 1 $src 37
done
