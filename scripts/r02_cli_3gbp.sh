#!/bin/bash
# One full `genmap index` + `genmap map` at 3 Gbp through the command line, timed end to end:
#   bash scripts/r02_cli_3gbp.sh N_GPUS [ref]     ("ref": also the unmodified reference on the same genome, CPU)
# Everything lives in /dev/shm; the -v lines of both programs give the breakdown.
N=${1:-1}; REF=${2:-}
W=/dev/shm/gmb_cli3g; rm -rf $W; mkdir -p $W/out gpurun_out
G=genmap_b200/bin/genmap
python - <<PY
import sys, time
sys.path.insert(0, ".")
import genmap_b200 as gm
from genmap_b200 import synth
t = time.time(); seqs = gm.synth_genome(3_000_000_000, 24, 45); print("genome generated in %.1f s" % (time.time() - t), flush=True)
t = time.time(); synth.write_fasta("$W/genome.fa", seqs); print("FASTA written in %.1f s" % (time.time() - t), flush=True)
PY
ls -la $W/genome.fa
t0=$(date +%s%N)
if [ -n "$REF" ]; then $G index -F $W/genome.fa -I $W/index -v -xf; else $G index -F $W/genome.fa -I $W/index -v -xn; fi
t1=$(date +%s%N); echo "== genmap index (B200 build): $(( (t1 - t0) / 1000000 )) ms wall"; ls -la $W/index | head -30
run() { # label, args...
  local label=$1; shift
  rm -rf $W/out; mkdir -p $W/out
  local a=$(date +%s%N)
  "$@" | tr '\r' '\n' | grep -v "^Progress\|^File .* Progress" | sed 's/\x1b\[K//g'
  local b=$(date +%s%N)
  echo "== $label: $(( (b - a) / 1000000 )) ms wall"; ls -la $W/out | tail -n +2 | awk '{print "   ", $5, $9}'
}
run "genmap map K=30 E=0 -r -fl -bg, $N GPU(s)" $G map -I $W/index -O $W/out -K 30 -E 0 -r -fl -bg -xg $N -v
run "genmap map K=30 E=0 -bg -w (runs from the device), $N GPU(s)" $G map -I $W/index -O $W/out -K 30 -E 0 -bg -w -xg $N -v
run "genmap map K=30 E=0 -t -fl, $N GPU(s)" $G map -I $W/index -O $W/out -K 30 -E 0 -t -fl -xg $N -v
run "genmap map K=30 E=2 -r -fl -bg, $N GPU(s)" $G map -I $W/index -O $W/out -K 30 -E 2 -r -fl -bg -xg $N -v
if [ -n "$REF" ]; then
  md5sum $W/out/genome.genmap.freq16 > $W/ours_e2.md5 2>/dev/null
  run "genmap map K=30 E=0 -r -fl -bg (ours again, for cmp)" $G map -I $W/index -O $W/out -K 30 -E 0 -r -fl -bg -xg $N -v
  mkdir -p $W/ours; mv $W/out/* $W/ours/
  run "REFERENCE genmap map K=30 E=0 -r -fl -bg -T $(nproc)" oracle/_ref/genmap_ref map -I $W/index -O $W/out -K 30 -E 0 -r -fl -bg -T $(nproc) -v
  for f in genome.genmap.freq16 genome.genmap.bedgraph; do cmp $W/ours/$f $W/out/$f && echo "   $f: identical to the reference's"; done
fi
rm -rf $W
