#!/bin/bash
# tools/cli_gz_inputs.sh [reads] -- wall clock of `rkmh stream` on the same reads as plain FASTQ, ordinary gzip and blocked gzip (BGZF)
cd "$(dirname "$0")/.."
N=${1:-4000000}
W=$(mktemp -d)
python - "$N" "$W" <<'PY'
import sys, gzip, shutil, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from rkmh_b200.fastx import load_sequences
from rkmh_b200.synth import fast_sample_reads, write_fastq
from test_host_cli import _bgzf
n, w = int(sys.argv[1]), sys.argv[2]
_, refs, _ = load_sequences(["tests/golden/data/all_pave_ref.fa.gz"])
b, o = fast_sample_reads(2, refs, 0, n, 150)
write_fastq(os.path.join(w, "reads.fq"), b, o)
payload = open(os.path.join(w, "reads.fq"), "rb").read()
open(os.path.join(w, "reads.bgzf.fq.gz"), "wb").write(_bgzf(payload, 65000, level=1))
open(os.path.join(w, "reads.std.fq.gz"), "wb").write(gzip.compress(payload, 1))
with gzip.open("tests/golden/data/all_pave_ref.fa.gz", "rb") as fi, open(os.path.join(w, "refs.fa"), "wb") as fo:
    shutil.copyfileobj(fi, fo)
PY
cat "$W"/reads.* > /dev/null
for f in reads.fq reads.std.fq.gz reads.bgzf.fq.gz "reads.bgzf.fq.gz RKMH_NO_BGZF=1"; do
  set -- $f; file=$1; extra=${2:-RKMH_NOP=1}
  all=""
  for rep in 1 2 3; do
    sleep 1.5
    s=$(date +%s.%N)
    env $extra rkmh_b200/bin/rkmh stream -r "$W/refs.fa" -f "$W/$file" -k 12 -s 1000 > "$W/out.txt" 2>/dev/null
    e=$(date +%s.%N)
    all="$all $(python -c "print(round($e-$s,2))")"
  done
  echo "$file $extra: wall seconds$all  reads=$N md5=$(md5sum < "$W/out.txt" | cut -c1-12) size=$(du -m "$W/$file" | cut -f1)MB"
done
for f in reads.fq reads.std.fq.gz reads.bgzf.fq.gz; do s=$(date +%s.%N); rkmh_b200/bin/rkmh parse-count "$W/$f" > /dev/null; e=$(date +%s.%N); echo "parse-count $f: $(python -c "print(round($e-$s,2))") s"; done
rm -rf "$W"
