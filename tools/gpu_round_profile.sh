#!/bin/bash
# One GPU-box pass that refreshes everything under profiles/ for the round: bench lines (both arms), sweeps, verifier probes,
# ncu launch lists and the per-kernel counter table.  Run under gpurun from the repo root; outputs land in gpurun_out/.
set -x
O=gpurun_out
python bench.py > $O/bench.json 2> $O/bench.err || tail -5 $O/bench.err
python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err || tail -5 $O/bench_reference_arm.err
python tools/gpu_msm_sweep.py 10 24 2 > $O/msm_sweep.jsonl 2> $O/msm_sweep.err
python tools/gpu_probe_range_verify.py 1 64 512 4096 > $O/range_verify.jsonl 2> $O/range_verify.err
python tools/gpu_probe_shuffle_verify.py 1 64 4096 > $O/shuffle_verify.jsonl 2> $O/shuffle_verify.err
python tools/gpu_probe_verify_two_ctx.py 4096 > $O/verify_two_ctx.jsonl 2> $O/verify_two_ctx.err
python tools/gpu_small_batch.py > $O/small_batch.jsonl 2> $O/small_batch.err
# ncu passes (numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --proofs 0 --fixed-points 0 > $O/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_msm|k_scan|k_first_bad|k_key_to|k_point_export|k_emit" --launch-skip 56 -c 60 --csv --log-file $O/launches_msm.csv python tools/gpu_msm_sweep.py 20 20 1 > $O/ncu_msm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_rp_fold|k_msm|k_scan" --launch-skip 60 -c 60 --csv --log-file $O/launches_range_verify.csv python tools/gpu_probe_range_verify.py 4096 > $O/ncu_range.log 2>&1
bash tools/ncu_kernel_table.sh
ls -la $O | tail -30
