#!/bin/bash
# round 2, GPU call 1h (one B200): FINAL kernel (sector-completing stores + warp-uniform select path, enabled per geometry): parity suite,
# the bench lines of all four workloads, the driver's command, launch list, ncu of the cavity kernel, upload trace
set -x
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02_tests_gpu_n1.log
tail -4 gpurun_out/r02_tests_gpu_n1.log
timeout 500 python bench.py --steps 200 --warmup 20 > gpurun_out/r02_bench_c5_n1.json 2> gpurun_out/r02_bench_c5_n1.err
cut -c1-200 gpurun_out/r02_bench_c5_n1.json; tail -3 gpurun_out/r02_bench_c5_n1.err
for w in c2 c3 c4; do
  timeout 300 python bench.py --workload $w --steps 300 --warmup 20 --no-dropin > gpurun_out/r02_bench_${w}_n1.json 2> gpurun_out/r02_bench_${w}_n1.err
  cut -c1-200 gpurun_out/r02_bench_${w}_n1.json
done
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_c5_n1_driver_cmd.json 2> gpurun_out/r02_bench_c5_n1_driver_cmd.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c5_n1.csv python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-parity --no-configs1 > gpurun_out/r02_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none -k 'regex:^k_step$' -s 14 -c 1 -f -o /tmp/r02_ncu_cavity384 python scripts/r02_probe.py one cavity 384 > gpurun_out/r02_ncu_cavity384.log 2>&1
ncu -i /tmp/r02_ncu_cavity384.ncu-rep --page raw --csv > gpurun_out/r02_ncu_cavity384_raw.csv 2>/dev/null
LUMA_B200_TRACE=1 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-dropin --no-configs1 --no-parity 2>&1 >/dev/null | grep luma_b200_upload > gpurun_out/r02_upload_trace.txt; cat gpurun_out/r02_upload_trace.txt
timeout 120 python scripts/r02_probe.py walls3 > gpurun_out/r02_probe_auto.txt 2>&1; cat gpurun_out/r02_probe_auto.txt
du -sh gpurun_out
