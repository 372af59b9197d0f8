#!/bin/bash
# tools/cli_variants.sh [reads] -- interleaved wall-clock comparison of CLI builds placed in build_variants/cli/rkmh_{base,detach,both} (see profiles/r02_cli_wallclock.txt)
cd "$(dirname "$0")/.."
N=${1:-8000000}
W=$(mktemp -d)
python - "$N" "$W" <<'PY'
import sys, gzip, shutil, os
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
free -g | head -2; df -h "$W" | tail -1
for round in 1 2 3 4; do
  for v in base detach both "both RKMH_NO_DETACH=1"; do
    set -- $v; bin=$1; extra=${2:-RKMH_NOP=1}
    sleep 1.5
    s=$(date +%s.%N)
    env $extra RKMH_TIMING=1 build_variants/cli/rkmh_$bin stream -r "$W/refs.fa" -f "$W/reads.fq" -k 12 -s 1000 > "$W/out.txt" 2> "$W/err.txt"
    e=$(date +%s.%N)
    echo "round $round $v wall $(python -c "print(round($e-$s,3))") ctx $(grep -o '[0-9.]* s  CUDA contexts' $W/err.txt | cut -d' ' -f1) written $(grep -o '[0-9.]* s  results written' $W/err.txt | cut -d' ' -f1)"
  done
done
rm -rf "$W"
