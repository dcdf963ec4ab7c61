#!/bin/bash
# ncu passes of the two batched verifiers (round 2): launch lists without the context set-up kernels, and counters of the
# transcript / Straus / Pippenger kernels.  Run under gpurun after tools/gpu_shuffle_once.py has exited 0 in the same call.
set -x
O=gpurun_out
python tools/gpu_shuffle_once.py 4096 2 > $O/plain_shuffle.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 120 --csv --log-file $O/launches_shuffle_verify_r02.csv python tools/gpu_shuffle_once.py 4096 2 > $O/ncu_shuffle.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 120 --csv --log-file $O/launches_range_verify_r02.csv python tools/gpu_range_once.py 4096 16 2 > $O/ncu_range.log 2>&1
# counters of the transcript / Straus / Pippenger kernels (captured once: profiles/ncu_verifier_kernels_r02.json):
# ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section InstructionStats --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --clock-control none --profile-from-start off -k 'regex:k_shuffle_pass_a_agg|k_shuffle_pass_b_agg|k_straus|k_msm_accumulate|k_msm_prepare' -c 5 -o $O/prof_verifier_r02 python tools/gpu_shuffle_once.py 4096 2
ls -la $O | tail -12
