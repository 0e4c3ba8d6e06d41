#!/bin/bash
# Per-kernel SASS mnemonic counts of the product library's cubins (no GPU needed): which byte-SIMD /
# TMA / vector-memory instructions each kernel really contains.  Output: profiles/<R>_sass_excerpts.txt
R=${1:-r02k}
OUT=profiles/${R}_sass_excerpts.txt
{
  echo "SASS mnemonic counts per kernel (cuobjdump -sass build/*.o, sm_100a, static counts; $(nvcc --version | tail -1))"
  echo "columns: IDP.4A (dp4a) | VABSDIFF4 (byte SAD) | I2IP (saturating pack) | PRMT (byte permute) | SHF (funnel shift) | LDG.E.128 | STG.E.128 | LDS.128 | ATOMS/RED shared | UTMALDG (TMA load) | SYNCS (mbarrier) | SHFL | REDUX"
  for o in build/*.o; do
    case $o in build/host_*|build/vabsdiff_probe*) continue;; esac
    cuobjdump -sass $o > /tmp/sass_$$.txt 2>/dev/null || continue
    awk -v obj=$(basename $o) '
      /Function :/ { if (name != "") pr(); name=$3; delete c; n=0 }
      /^[ \t]+\/\*[0-9a-f]+\*\// { n++; m=$2; if (m ~ /^@/) m=$3;
        if (m ~ /^IDP\.4A/) c["idp"]++; if (m ~ /^VABSDIFF4/) c["vabs"]++; if (m ~ /^I2IP/) c["i2ip"]++;
        if (m ~ /^PRMT/) c["prmt"]++; if (m ~ /^SHF/) c["shf"]++; if (m ~ /^LDG.*128/) c["ldg128"]++;
        if (m ~ /^STG.*128/) c["stg128"]++; if (m ~ /^LDS.*128/) c["lds128"]++; if (m ~ /^ATOMS|^RED/) c["atom"]++;
        if (m ~ /^UTMALDG/) c["tma"]++; if (m ~ /^SYNCS/) c["syncs"]++; if (m ~ /^SHFL/) c["shfl"]++; if (m ~ /^REDUX/) c["redux"]++ }
      function pr() { printf "%-16s %-70s instr %5d | %3d | %3d | %3d | %3d | %3d | %3d | %3d | %3d | %3d | %3d | %3d | %3d | %3d\n", obj, substr(name,1,70), n,
        c["idp"], c["vabs"], c["i2ip"], c["prmt"], c["shf"], c["ldg128"], c["stg128"], c["lds128"], c["atom"], c["tma"], c["syncs"], c["shfl"], c["redux"] }
      END { if (name != "") pr() }' /tmp/sass_$$.txt
  done
  rm -f /tmp/sass_$$.txt
} > $OUT
wc -l $OUT
