#!/bin/bash
# sort of sub-batch k+1 overlapped with the accumulation of sub-batch k (two streams)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_golden.py -x -q -m gpu -k "task_groups or batched or amt or pipelined or sub_batch or full_size or montgomery or abort or 2pow24 or 2pow20 or multi_device or sharded" > gpurun_out/r2_run33_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_run33_pytest.log
out=gpurun_out/r2_run33_overlap.log
: > $out
for ov in 1 0; do
  for lg in 24 23 21; do
    echo "== e2e overlap=$ov log_n=$lg" >> $out
    MSM_B200_SORT_OVERLAP=$ov timeout 200 python tools/e2e_timing.py $lg 0 2>&1 | grep e2e_ms >> $out
  done
  echo "== batched e2e overlap=$ov" >> $out
  MSM_B200_SORT_OVERLAP=$ov timeout 300 python bench.py --workload batched --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batched dev %.3f ms'%d['ms_per_step'], 'e2e %.3f ms'%d['e2e']['ms_per_step'], d['e2e'].get('upload_sub_batches'), d.get('paths_agree'), d.get('result_matches_golden'))" >> $out
done
echo "== device-resident row split in two (sort B under accumulate A)" >> $out
for q in 0.5 1 2; do
  for ov in 1 0; do
  echo "ratio=$q overlap=$ov" >> $out
  MSM_B200_SORT_OVERLAP=$ov MSM_B200_PIPELINE_RATIO=$q MSM_B200_PIPELINE_DEVICE=1 MSM_B200_PIPELINE=2 PRECOMPUTE=0 timeout 120 python tools/quick_timing.py 24 2>&1 | grep log_L | cut -c40-170 >> $out
  done
done
cat $out
