#!/bin/bash
# Round-2 pass on one GPU box: bench lines (both arms), verifier / MSM / secret-mode probes, then the ncu passes (launch lists of the
# bench and of the two batched verifiers; counters of the new kernels).  Run under gpurun from the repo root; outputs land in
# gpurun_out/ and are copied to profiles/*_r02.* by hand.  Numbers printed under ncu are never bench values.
set -x
O=gpurun_out
python bench.py > $O/bench_r02.json 2> $O/bench_r02.err || tail -5 $O/bench_r02.err
python bench.py --impl reference > $O/bench_r02_reference_arm.json 2> $O/bench_r02_reference_arm.err || tail -5 $O/bench_r02_reference_arm.err
python tools/gpu_probe_shuffle_verify.py 1 64 512 4096 16384 > $O/shuffle_verify_r02.jsonl 2> $O/shuffle_verify.err
python tools/gpu_probe_range_verify.py 1 64 512 4096 > $O/range_verify_r02.jsonl 2> $O/range_verify.err
python tools/gpu_secret_mode.py > $O/secret_mode_r02.jsonl 2> $O/secret_mode.err
python tools/gpu_small_batch.py > $O/small_batch_r02.jsonl 2> $O/small_batch.err
python tools/gpu_shuffle_once.py 4096 2 > $O/plain_shuffle.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --proofs 0 --fixed-points 0 --msm-sweep-max 0 > $O/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file $O/launches_shuffle_verify_r02.csv python tools/gpu_shuffle_once.py 4096 2 > $O/ncu_shuffle.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_rp_|k_msm|k_scan|k_shuffle" -s 40 -c 120 --csv --log-file $O/launches_range_verify_r02.csv python tools/gpu_probe_range_verify.py 4096 > $O/ncu_range.log 2>&1
ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section InstructionStats --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --clock-control none -k regex:"k_shuffle_pass_a_agg|k_shuffle_pass_b_agg|k_straus|k_msm_accumulate|k_msm_prepare" -s 10 -c 6 -o $O/prof_verifier_r02 python tools/gpu_shuffle_once.py 4096 2 > $O/ncu_prof.log 2>&1
ls -la $O | tail -30
