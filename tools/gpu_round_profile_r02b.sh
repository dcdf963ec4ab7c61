#!/bin/bash
# Round-2 closing pass on one GPU box (after the warp-cooperative Horner chain and the identity-only aggregate check): bench lines
# of both arms, verifier / small-batch probes, MSM sweep, then the ncu launch list of the bench command.  Outputs land in
# gpurun_out/ and are copied to profiles/*_r02b.* by hand.  Numbers printed under ncu are never bench values.
set -x
O=gpurun_out
python bench.py > $O/bench_r02b.json 2> $O/bench_r02b.err || tail -5 $O/bench_r02b.err
python bench.py --impl reference > $O/bench_r02b_reference_arm.json 2> $O/bench_r02b_reference_arm.err || tail -5 $O/bench_r02b_reference_arm.err
python tools/gpu_probe_shuffle_verify.py 1 64 512 4096 16384 > $O/shuffle_verify_r02b.jsonl 2> $O/shuffle_verify.err
python tools/gpu_probe_range_verify.py 1 64 512 4096 > $O/range_verify_r02b.jsonl 2> $O/range_verify.err
python tools/gpu_small_batch.py > $O/small_batch_r02b.jsonl 2> $O/small_batch.err
python tools/gpu_msm_sweep.py > $O/msm_sweep_r02b.jsonl 2> $O/msm_sweep.err
python tools/gpu_probe_prepared.py > $O/msm_prepared_r02b.jsonl 2> $O/msm_prepared.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_r02b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --proofs 0 --fixed-points 0 --msm-sweep-max 0 > $O/ncu_bench.log 2>&1
ls -la $O | tail -20
