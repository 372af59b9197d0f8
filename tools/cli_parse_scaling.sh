#!/bin/bash
# Wall clock of `rkmh stream` on one synthetic config-2 FASTQ for several parser thread counts (best of 3 each).
# usage: tools/cli_parse_scaling.sh [reads]
set -e
cd "$(dirname "$0")/.."
N=${1:-8000000}
W=$(mktemp -d)
python - "$N" "$W" <<'PY'
import sys, gzip, shutil, os
import numpy as np
sys.path.insert(0, os.getcwd())
from rkmh_b200.fastx import load_sequences
from rkmh_b200.synth import fast_sample_reads, write_fastq
n, w = int(sys.argv[1]), sys.argv[2]
_, refs, _ = load_sequences(["tests/golden/data/all_pave_ref.fa.gz"])
b, o = fast_sample_reads(2, refs, 0, n, 150)
write_fastq(os.path.join(w, "reads.fq"), b, o)
with gzip.open("tests/golden/data/all_pave_ref.fa.gz", "rb") as fi, open(os.path.join(w, "refs.fa"), "wb") as fo:
    shutil.copyfileobj(fi, fo)
PY
cat "$W/reads.fq" > /dev/null
for t in 1 8 "8 RKMH_NO_DETACH=1" 8 "8 RKMH_NO_DETACH=1" "1 RKMH_NO_DETACH=1"; do
  set -- $t; t=$1; extra=${2:-RKMH_NOP=1}
  best=999; all=""
  for rep in 1 2 3; do
    sleep 1.5                                   # a detached worker of the previous run may still be tearing its context down
    s=$(date +%s.%N)
    env $extra RKMH_PARSE_THREADS=$t rkmh_b200/bin/rkmh stream -r "$W/refs.fa" -f "$W/reads.fq" -k 12 -s 1000 > "$W/out.txt" 2>/dev/null
    e=$(date +%s.%N)
    d=$(python -c "print(round($e-$s,3))")
    all="$all $d"
    best=$(python -c "print(min($best,$d))")
  done
  echo "parse_threads=$t $extra best_of_3_seconds=$best (all:$all) reads=$N md5=$(md5sum < "$W/out.txt" | cut -c1-12)"
done
echo "--- stage marks (RKMH_TIMING=1), default parser threads"
RKMH_TIMING=1 rkmh_b200/bin/rkmh stream -r "$W/refs.fa" -f "$W/reads.fq" -k 12 -s 1000 2>&1 > "$W/out.txt" | grep "rkmh:" || true
echo "--- stage marks, one parser thread"
RKMH_PARSE_THREADS=1 RKMH_TIMING=1 rkmh_b200/bin/rkmh stream -r "$W/refs.fa" -f "$W/reads.fq" -k 12 -s 1000 2>&1 > "$W/out.txt" | grep "rkmh:" || true
echo "--- empty read file (start-up only)"
: > "$W/empty.fq"; s=$(date +%s.%N); rkmh_b200/bin/rkmh stream -r "$W/refs.fa" -f "$W/empty.fq" -k 12 -s 1000 > /dev/null 2>&1; e=$(date +%s.%N); echo "startup_seconds=$(python -c "print(round($e-$s,3))")"
echo "--- parser alone (parse-count)"
for t in 1 4 8; do s=$(date +%s.%N); RKMH_PARSE_THREADS=$t rkmh_b200/bin/rkmh parse-count "$W/reads.fq" > /dev/null; e=$(date +%s.%N); echo "parse-count threads=$t seconds=$(python -c "print(round($e-$s,3))")"; done
rm -rf "$W"
