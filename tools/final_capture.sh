#!/bin/bash
# tools/final_capture.sh TAG -- on the GPU box: the evidence bench.py and profiles/ refer to, all from ONE tree:
#   full `pytest -m gpu`, smoke, the default bench line of both arms, the ncu launch list of the bench command, one
#   `ncu --set full` capture of each dominant kernel (summarised on the box), and profiles/ncu_current.json (written to
#   gpurun_out/, stamped with the hash of the kernel sources) from those captures.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $OUT/${TAG}_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --reads 2000000 --no-cpu-baseline --cfg3-reads 20000 --cfg4-reads 2000000 --cfg5-bases 50000000 > $OUT/${TAG}_bench_under_ncu.json 2>> $OUT/${TAG}_bench.err
tools/ncu_capture.sh ${TAG}_sieve classify_memo_sieve_kernel sieve
tools/ncu_capture.sh ${TAG}_warp_k12 classify_warp_kernel murmur
tools/ncu_capture.sh ${TAG}_warp_k20 classify_warp_kernel cfg4
tools/ncu_capture.sh ${TAG}_count count_memo_kernel cfg3
tools/ncu_capture.sh ${TAG}_block classify_memo_block_kernel cfg3
SPECS=""
for t in sieve warp_k12 count block; do SPECS="$SPECS $OUT/${TAG}_${t}_ncu_raw.json:$(cat $OUT/${TAG}_$t.reads)"; done
python profiles/ncu_summary.py --current $OUT/ncu_current.json $SPECS
cp $OUT/ncu_current.json profiles/ncu_current.json      # the bench lines below read it: same sources, so traffic / issue_slots are filled in
python bench.py > $OUT/${TAG}_bench.json 2>> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench.err
tail -2 $OUT/${TAG}_gputest.log; cat $OUT/${TAG}_smoke.log | tail -1; ls -la $OUT | head -50; du -sh $OUT
