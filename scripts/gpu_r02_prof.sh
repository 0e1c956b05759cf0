#!/bin/bash
# round-2 profile artefacts: ncu launch list of the bench command, --set full captures of the attention kernel and the
# quantiser (C2, default mode), compute-sanitizer memcheck / racecheck / synccheck on small ragged calls of every kernel
R=r02
mkdir -p gpurun_out
BENCH="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-comparators --no-other-modes"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file gpurun_out/launches_$R.csv $BENCH > gpurun_out/ncu_list_$R.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/attn_$R $BENCH > gpurun_out/ncu_attn_$R.log 2>&1; echo "ncu attn rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:quant_head -s 6 -c 1 -f -o gpurun_out/quant_$R $BENCH > gpurun_out/ncu_quant_$R.log 2>&1; echo "ncu quant rc=$?"
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitizer_${tool}_$R.txt 2>&1; echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_${tool}_$R.txt
done
ls -la gpurun_out/*_$R.* | head -20
