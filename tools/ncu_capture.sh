#!/bin/bash
# tools/ncu_capture.sh TAG KERNEL_REGEX MODE [READS]
# One `ncu --set full` capture of the first launch of KERNEL_REGEX in `tools/run_kernel.py MODE`, summarised ON THE BOX
# (the reports are ~18 MB each and gpurun_out/ is capped at 64 MiB): keeps profiles-ready JSON / by-line text, drops the report
# unless KEEP_REP=1.
set -u
TAG=$1; KERNEL=$2; MODE=$3; READS=${4:-}
OUT=gpurun_out
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:$KERNEL -c 1 -f -o $OUT/$TAG python tools/run_kernel.py $MODE $READS > $OUT/$TAG.log 2>&1
python profiles/ncu_summary.py $OUT/$TAG.ncu-rep "ncu --set full --clock-control none -k regex:$KERNEL -c 1 python tools/run_kernel.py $MODE $READS" > $OUT/${TAG}_ncu_raw.json 2>> $OUT/$TAG.log
python profiles/ncu_by_line.py $OUT/$TAG.ncu-rep 60 > $OUT/${TAG}_by_line.txt 2>> $OUT/$TAG.log
ncu -i $OUT/$TAG.ncu-rep --page source --csv --print-source sass 2>/dev/null | python tools/sass_counts.py > $OUT/${TAG}_sass_hist.txt 2>> $OUT/$TAG.log
grep -o '"reads_per_launch": [0-9]*' $OUT/$TAG.log | grep -o '[0-9]*$' > $OUT/$TAG.reads
if [ "${KEEP_REP:-0}" != "1" ]; then mv $OUT/$TAG.ncu-rep /tmp/$TAG.ncu-rep; fi
