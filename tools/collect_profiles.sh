#!/bin/bash
# Copies the evidence of one `tools/gpu_round.sh <tag>` call from gpurun_out/<tag>/ into profiles/ under round-2 names and derives the
# per-source-line instruction attribution of the C2 / C3 / C5 kernels from the exported ncu source pages.  Run it in the container,
# with the SAME build that was profiled still in libavif_b200/csrc/ (the attribution matches instruction offsets).
# usage: bash tools/collect_profiles.sh <tag>
set -e
TAG=${1:?tag}
SRC=gpurun_out/$TAG
cp $SRC/bench.json profiles/r02_bench.json
cp $SRC/launches.csv profiles/r02_launches_bench.csv
cp $SRC/gpu.txt profiles/r02_gpu_host.txt
cp $SRC/traffic.json profiles/traffic.json
{ echo "# Round 2 -- ncu --set full of every configuration's BATCHED kernel (gpurun $TAG, B200, clocks not locked)"; echo
  echo "Captured by \`tools/gpu_round.sh\` -> \`tools/profile_cases.py\` under \`ncu --set full --clock-control none --import-source on\`; the second (warm) launch of each case is tabulated by \`tools/ncu_table.py\`, which fails loudly when a case's kernel is missing from the capture.  Times under ncu are serialised and cold-clock: the performance numbers are the CUDA-event times of \`r02_bench.json\`; this table is for DRAM traffic, issue utilisation, occupancy, registers and instruction counts."; echo
  cat $SRC/ncu_cases.md; } > profiles/r02_ncu_cases.md
TMP=$(mktemp -d)
( cd $TMP && cuobjdump -xelf all $OLDPWD/libavif_b200/csrc/y2r_fast_16_16.o > /dev/null && nvdisasm -g y2r_fast_16_16*.cubin > y2r.sass 2>/dev/null
  cuobjdump -xelf all $OLDPWD/libavif_b200/csrc/r2y_fast_16_16.o > /dev/null && nvdisasm -g r2y_fast_16_16*.cubin > r2y.sass 2>/dev/null )
{ echo "# Round 2 -- executed warp instructions per CUDA source line (ncu source pages of gpurun $TAG; tools/sass_lines.py)"; echo
  for spec in "1 y2r 66355200 C2 forward (4K 4:4:4 10-bit + alpha -> RGBA10 premultiplied, 8 frames)" "3 r2y 66355200 C2 back (RGBA10 premultiplied -> YUVA, 8 frames)" \
              "7 y2r 132710400 C3 (8K 4:2:0 12-bit bilinear + unmultiply, 4 frames)" "11 y2r 66355200 C5 (4K 4:2:0 10-bit bilinear -> RGB16, 8 frames)"; do
    set -- $spec; id=$1; dir=$2; px=$3; shift 3
    echo "## $*"; echo; echo '```'
    python tools/sass_lines.py $SRC/src_k$id.csv $TMP/$dir.sass $px 28 2>&1 || true
    echo '```'; echo
  done; } > profiles/r02_sass_lines.md
rm -rf $TMP
if [ -e gpurun_out/sanitize/memcheck.log ]; then cp gpurun_out/sanitize/memcheck.log profiles/r02_memcheck.log; fi
ls -la profiles/
