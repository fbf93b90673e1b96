set -x
# launch list of the default bench command (per-launch times are cold-cache and serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s3_launches_c2.csv python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/s3_launches_c2.log 2>&1
# full captures: k_bc (cavity lid), Smagorinsky k_step + k_bc of the cylinder case, time-averaging k_step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bc -s 5 -c 1 -o gpurun_out/s3_kbc -f python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/s3_ncu_kbc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 5 -c 1 -o gpurun_out/s3_kstep_c4 -f python bench.py --workload c4 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/s3_ncu_kstep_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 5 -c 1 -o gpurun_out/s3_kstep_c3 -f python bench.py --workload c3 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/s3_ncu_kstep_c3.log 2>&1
# final bench lines
timeout 300 python bench.py > gpurun_out/s3_bench4_n1.json 2> gpurun_out/s3_bench4_n1.err
timeout 300 python bench.py --workload c4 --steps 300 --warmup 10 --no-cpu > gpurun_out/s3_bench4_c4_n1.json 2> gpurun_out/s3_bench4_c4_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/s3_bench4_ref.json 2> gpurun_out/s3_bench4_ref.err
ls -la gpurun_out | tail -12; cut -c1-250 gpurun_out/s3_bench4_n1.json gpurun_out/s3_bench4_c4_n1.json gpurun_out/s3_bench4_ref.json
