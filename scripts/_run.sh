# round-end style check on one B200:  gpurun --timeout 900 -- 'bash scripts/_run.sh'
set -x
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s6_tests_gpu.log
cat gpurun_out/s6_tests_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s6_smoke.log 2>&1; tail -5 gpurun_out/s6_smoke.log
timeout 300 python bench.py > gpurun_out/s6_bench_n1.json 2> gpurun_out/s6_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/s6_bench_ref.json 2> gpurun_out/s6_bench_ref.err
cut -c1-300 gpurun_out/s6_bench_n1.json gpurun_out/s6_bench_ref.json
