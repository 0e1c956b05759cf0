#!/bin/bash
# scripts/gpu_cycles.sh VARIANT...: SM cycles of the attention kernel per variant (clock-independent), via a light ncu pass
mkdir -p gpurun_out
for v in "$@"; do
  QA_NATIVE_LIB=$PWD/quantumattention_b200/libqattn_sm100_$v.so ncu --metrics sm__cycles_elapsed.max,gpu__time_duration.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:attn_fwd -s 3 -c 4 --csv --log-file gpurun_out/cyc_$v.csv python bench.py --workload ${WL:-C2_flux} --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
  python - "$v" <<'PY'
import csv,sys,collections
v=sys.argv[1]
rows=[r for r in csv.reader(open(f"gpurun_out/cyc_{v}.csv")) if len(r)>10]
hdr=rows[0]; mi,vi=hdr.index("Metric Name"),hdr.index("Metric Value")
agg=collections.defaultdict(list)
for r in rows[1:]: agg[r[mi]].append(float(r[vi].replace(",","")))
print(v, {k.split(".")[0][:28]: round(min(x),1) for k,x in agg.items()})
PY
done
