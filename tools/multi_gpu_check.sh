#!/bin/bash
# usage: tools/multi_gpu_check.sh N   -- on a box with N GPUs: (1) sharded bench parity check at a small scale,
# (2) the C drop-in driver with NGSCOV_GPUS=N against its own single-GPU output on a generated BAM (host inflate and
# device inflate), (3) the full-size sharded bench line.  Outputs under gpurun_out/.
N=${1:-2}
set -x
mkdir -p gpurun_out /tmp/mg && cd "$(dirname "$0")/.."
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 --scale 0.02 --no-e2e --no-cpu --check 2>&1 | grep -i "check:\|error" | tail -3
tools/bamgen --config 2 --scale 0.05 --seed 7 --seq none -o /tmp/mg/c2.bam 2>/dev/null
( cd /tmp/mg && rm -rf one many manyg && mkdir one many manyg
  ( cd one && NGSCOV_VERBOSE=1 /root/repo/highperformancengs_b200/host/bam2depth -o o -w 20000 -W ../c2.bam 2>&1 | grep -c "ngscov:" )
  ( cd many && NGSCOV_GPUS=$N NGSCOV_VERBOSE=1 /root/repo/highperformancengs_b200/host/bam2depth -o o -w 20000 -W ../c2.bam 2>&1 | grep "ngscov: shard" | cut -c1-200 )
  ( cd manyg && NGSCOV_GPUS=$N NGSCOV_GPU_INFLATE=1 NGSCOV_VERBOSE=1 /root/repo/highperformancengs_b200/host/bam2depth -o o -w 20000 -W ../c2.bam 2>&1 | grep -c "ngscov: shard" )
  for d in many manyg; do for f in c2.bam.1.bedGraph o.1.depth o.1.wig o.1.chromSize.txt; do cmp one/$f $d/$f && echo "$d/$f identical to the single-GPU run"; done; done ) 2>&1 | tee gpurun_out/multi_gpu_cli_$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1800 gpurun_out/bench_n$N.json
