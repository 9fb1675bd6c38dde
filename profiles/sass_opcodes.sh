#!/bin/bash
# Per-kernel SASS opcode counts of the in-tree library (no GPU needed): the tcgen05 / TMEM / bulk-copy mnemonics that prove the
# Blackwell path (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier),
# plus spill traffic (LDL / STL) and the global / SFU instruction mix.   bash profiles/sass_opcodes.sh > profiles/r2b_sass_opcodes.txt (UTMALDG = cp.async.bulk.tensor, UCGABAR = barrier.cluster)
cd "$(dirname "$0")/.."
LIB=fmov_pose_b200/libfmov_b200.so
echo "# cuobjdump -sass $LIB, per kernel: count mnemonic ..."
for sym in $(cuobjdump -elf $LIB 2>/dev/null | grep -oE "_ZN4fmov[A-Za-z0-9_]+" | grep -E "kernel" | sort -u); do
  n=$(cuobjdump -sass -fun "$sym" $LIB 2>/dev/null | grep -cE "^\s+/\*[0-9a-f]{4,6}\*/")
  [ "$n" -eq 0 ] && continue
  printf "%-34s %6d instr: " "$(echo $sym | c++filt | sed 's/(.*//; s/fmov:://; s/void //')" "$n"
  cuobjdump -sass -fun "$sym" $LIB 2>/dev/null | grep -oE "UTCHMMA|LDTM|STTM|UBLKCP|UTCBAR|UTMALDG|UTMAPF|UCGABAR|SYNCS|MUFU\.[A-Z0-9]+|HMMA|LDG|STG|LDS|STS|LDL|STL|RED|ATOM|SHFL|BAR" | sort | uniq -c | awk '{printf "%s %s  ", $1, $2}'
  echo
done
echo "# CTA-pair engine: full mnemonics of the tensor / TMA / cluster instructions (2CTA = cta_group::2)"
for sym in $(cuobjdump -elf $LIB 2>/dev/null | grep -oE "_ZN4fmov[A-Za-z0-9_]+" | grep -E "fine_fwd_kernel|fine_bwd_kernel|sdf_query_kernelILi2" | grep -v "_param_" | sort -u); do
  printf "%-34s " "$(echo $sym | c++filt | sed 's/(.*//; s/fmov:://; s/void //')"
  cuobjdump -sass -fun "$sym" $LIB 2>/dev/null | grep -oE "UTCHMMA[.A-Z0-9_]*|UTMALDG[.A-Z0-9_]*|UTCBAR[.A-Z0-9_]*|UCGABAR[.A-Z_]*|UTCATOMSWS[.A-Z0-9_]*" | sort | uniq -c | awk '{printf "%s %s  ", $1, $2}'
  echo
done
