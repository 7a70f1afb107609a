#!/bin/bash
# Track writers end to end on one GPU: `genmap map -bg -w` with the runs found on the device (default) vs the
# host scan of the full vector (--host-runs).  Prints the CLI's own timing lines; the files must be identical.
set -e
MBP=${1:-250}
W=$(mktemp -d /dev/shm/gmb_wr_XXXX)
python - "$MBP" "$W" <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import genmap_b200 as gm
from genmap_b200 import synth
mbp, w = float(sys.argv[1]), sys.argv[2]
synth.write_fasta(os.path.join(w, "g.fa"), gm.synth_genome(int(mbp * 1e6), 5, 44))
PY
G=genmap_b200/bin/genmap
/usr/bin/time -v true 2>/dev/null || true
t0=$(date +%s.%N); $G index -F $W/g.fa -I $W/index -xn -v | tail -2; t1=$(date +%s.%N)
echo "index: $(echo "$t1 - $t0" | bc 2>/dev/null || python -c "print($t1-$t0)") s"
for mode in "" "--host-runs"; do
  mkdir -p $W/out$mode
  t0=$(date +%s.%N)
  $G map -I $W/index -O $W/out$mode -K 30 -E 0 -fl -bg -w -v $mode | grep -v Progress
  t1=$(date +%s.%N)
  echo "map -bg -w $mode: $(python -c "print(round($t1-$t0,2))") s wall"
done
cmp $W/out/g.genmap.bedgraph $W/out--host-runs/g.genmap.bedgraph && cmp $W/out/g.genmap.wig $W/out--host-runs/g.genmap.wig && echo "files identical: $(du -sh $W/out | cut -f1)"
rm -rf $W
