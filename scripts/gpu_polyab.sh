#!/bin/bash
# Per-head-dimension A/B of the polynomial-exponential share (variants from scripts/build_variant.sh) + the GPU tests
mkdir -p gpurun_out
out=gpurun_out/polyab.txt; : > $out
L=$PWD/quantumattention_b200/libqattn_sm100
for r in 1 2; do
for v in cur a b c e f; do
  QA_NATIVE_LIB=${L}_$v.so timeout 120 python scripts/time_shape.py 256 8192 0 >> $out 2>&1
done
for v in cur g h; do
  QA_NATIVE_LIB=${L}_$v.so timeout 120 python scripts/time_shape.py 64 8192 0 >> $out 2>&1
done
done
for v in cur a b e; do
  QA_NATIVE_LIB=${L}_$v.so timeout 120 python scripts/time_shape.py 256 8192 1 >> $out 2>&1
done
QA_NATIVE_LIB=${L}_cur.so timeout 120 python scripts/time_shape.py 128 8192 0 >> $out 2>&1
QA_NATIVE_LIB=${L}_cur.so timeout 120 python scripts/time_shape.py 64 8192 1 >> $out 2>&1
cat $out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_polyab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_polyab.log
tail -5 gpurun_out/pytest_gpu_polyab.log
