#!/bin/bash
# round 2, GPU call 3 (eight B200):  gpurun --gpus 8 --timeout 1000 -- 'bash scripts/r02_run8.sh'
# (a) slab parity vs the serial oracle on 4 and 8 ranks (2 ranks: scripts/r02_run2.sh), all three transports, and the
#     full-size 1 GPU == 8 GPUs checksums of BASELINE configs[2] / configs[3];
# (b) bench lines: c5 (configs[4], 384^3 per GPU) at 1/2/4/8, c3 (configs[2]) and c4 (configs[3]) at 2/4/8.
# Independent runs share the box on disjoint GPU sets (CUDA_VISIBLE_DEVICES) to keep the lease short.
set -x
mkdir -p gpurun_out
HALO=${HALO:-fused}
W=tests/mgpu_worker.py
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
C8="chan3d,cyl3d,chan2d,cyl2d,slipchan3d,sliptunnel2d,sliptunnel3d,fevel2d,fevel3d,fevel2d_tav,tunnel2d_tav"
C4="$C8,cav3d_32,cav3d_tav,felid3d,kbc2d_cyl,kbc3d_chan"
run() {   # run <gpus csv> <n> <port> <workload> <steps> <tag> [extra...]
  local devs=$1 n=$2 port=$3 w=$4 st=$5 tag=$6; shift 6
  if [ $n -eq 1 ]; then
    CUDA_VISIBLE_DEVICES=$devs timeout 400 python bench.py --workload $w --steps $st --warmup 20 --halo $HALO "$@" > gpurun_out/r02_bench_${tag}.json 2> gpurun_out/r02_bench_${tag}.err
  else
    CUDA_VISIBLE_DEVICES=$devs timeout 400 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --workload $w --steps $st --warmup 20 --halo $HALO "$@" > gpurun_out/r02_bench_${tag}.json 2> gpurun_out/r02_bench_${tag}.err
  fi
}
# ---- phase A: 4-rank parity on GPUs 0-3, the 1-GPU full-size checksums on GPUs 4 and 5, a 2-GPU bench on 6,7 ----
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 $TR --nproc-per-node 4 --master-port 29614 $W $C4 > gpurun_out/r02_mgpu_n4.log 2>&1 ) &
( CUDA_VISIBLE_DEVICES=4 timeout 600 $TR --nproc-per-node 1 --master-port 29641 $W fullsize:c4:20 > gpurun_out/r02_fullsize_c4_n1.log 2>&1 ) &
( CUDA_VISIBLE_DEVICES=5 timeout 600 $TR --nproc-per-node 1 --master-port 29642 $W fullsize:c3:12 > gpurun_out/r02_fullsize_c3_n1.log 2>&1 ) &
run 6,7 2 29621 c5 200 c5_n2 --no-cpu &
wait
grep -c "mgpu ok" gpurun_out/r02_mgpu_n4.log; tail -2 gpurun_out/r02_mgpu_n4.log
# ---- phase B: everything that needs all eight GPUs ----
timeout 900 $TR --nproc-per-node 8 --master-port 29618 $W $C8 > gpurun_out/r02_mgpu_n8.log 2>&1
grep -c "mgpu ok" gpurun_out/r02_mgpu_n8.log; tail -2 gpurun_out/r02_mgpu_n8.log
timeout 300 $TR --nproc-per-node 8 --master-port 29648 $W fullsize:c4:20 > gpurun_out/r02_fullsize_c4_n8.log 2>&1
timeout 300 $TR --nproc-per-node 8 --master-port 29649 $W fullsize:c3:12 > gpurun_out/r02_fullsize_c3_n8.log 2>&1
grep -h "fullsize checksums" gpurun_out/r02_fullsize_*.log
run 0,1,2,3,4,5,6,7 8 29601 c5 200 c5_n8 --no-cpu
run 0,1,2,3,4,5,6,7 8 29602 c3 300 c3_n8 --no-e2e --no-cpu
run 0,1,2,3,4,5,6,7 8 29603 c4 300 c4_n8 --no-e2e --no-cpu
# ---- phase C: 4-, 2- and 1-GPU lines side by side ----
run 0,1,2,3 4 29611 c5 200 c5_n4 --no-cpu & run 4,5,6,7 4 29612 c3 300 c3_n4 --no-e2e --no-cpu & wait
run 0,1,2,3 4 29613 c4 300 c4_n4 --no-e2e --no-cpu & run 4,5 2 29622 c3 300 c3_n2 --no-e2e --no-cpu & run 6,7 2 29623 c4 300 c4_n2 --no-e2e --no-cpu & wait
run 0 1 0 c5 200 c5_n1_box8 --no-cpu --no-dropin --no-configs1
for f in gpurun_out/r02_bench_c*_n*.json; do echo $f; python -c "
import json
d=json.load(open('$f')); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['clocks'], (d.get('parity_check') or {}).get('bitwise'))"; done
