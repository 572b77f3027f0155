#!/bin/sh
# SASS evidence of the built library: mnemonic counts + the TMA subtree stage + k_shade's block reservation.
#   sh tools/sass_excerpt.sh > profiles/r02_sass_excerpt.txt
LIB=cuda-raytracer_b200/libb2rt.so
cuobjdump -sass $LIB > /tmp/b2rt_sass.txt
echo "# SASS evidence, libb2rt.so round 2, final kernels (cuobjdump -sass $LIB | grep -c <mnemonic>)"
for m in UBLKCP SYNCS.ARRIVE.TRANS64 SYNCS.PHASECHK FMNMX3 REDG.E.MIN.64 ATOMS.CAST.SPIN.64 ATOMS.ADD ATOMS.POPC.INC ATOMG.E.INC ATOMG.E.ADD SHFL.UP SHFL.IDX VOTE PRMT LDS.128 LDG.E.128.CONSTANT LDG.E.128 MUFU.RCP REDG UTMALDG "tcgen05\|UTCMMA"; do
  printf "%-26s %s\n" "$m" "$(grep -c "$m" /tmp/b2rt_sass.txt)"
done
echo
echo "# REDG.E.MIN.64 = packed (t, prim) closest-hit merge in global memory (no return value needed); ATOMS.CAST.SPIN.64 = the same"
echo "# merge in shared memory (per-warp leaf queue); ATOMG.E.INC = k_shade's block reservation (atom.inc with a run-time bound,"
echo "# which ptxas leaves un-aggregated: no VOTEU / SHFL around it); UTMALDG / tcgen05 = 0: nothing on this path is a tensor-tile"
echo "# copy or a dense contraction, the subtree blobs are contiguous byte ranges (1-D bulk copies, UBLKCP)."
echo
echo "# registers / spills / shared memory of every kernel: profiles/r02_registers.txt"
echo
echo "# the subtree stage (one TMA bulk copy per chunk) inside k_traverse<4,false,false>:"
cuobjdump -sass -fun "$(cuobjdump -elf $LIB 2>/dev/null | grep -o '_ZN4b2rt[^ ]*k_traverseILi4ELb0ELb0E[^ ]*' | head -1)" $LIB 2>/dev/null | grep -B6 -A6 "UBLKCP" | grep -v "^\s*/\* 0x" | cut -c1-110 | head -30
echo
echo "# k_shade<false>: reservation of the next output block (lane 0), plain ATOMG.E.INC, result consumed a tile later:"
cuobjdump -sass -fun "$(cuobjdump -elf $LIB 2>/dev/null | grep -o '_ZN4b2rt[^ ]*k_shadeILb0E[^ ]*' | head -1)" $LIB 2>/dev/null | grep -B4 -A4 "ATOMG.E.INC" | cut -c1-110 | head -24
