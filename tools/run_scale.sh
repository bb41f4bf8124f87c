#!/bin/bash
# usage: tools/run_scale.sh N   (inside gpurun --gpus N): one bench.py run at N GPUs, JSON line -> gpurun_out/p/scale_N.json
N=$1
mkdir -p gpurun_out/p
if [ "$N" == "1" ]; then
  GP_MLL_TIMING=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/p/scale_1.json 2> gpurun_out/p/scale_1.err
else
  GP_MLL_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/p/scale_$N.json 2> gpurun_out/p/scale_$N.err
fi
tail -c 1500 gpurun_out/p/scale_$N.json | head -c 400; echo
grep "gp_mll timing" gpurun_out/p/scale_$N.err | sed -n '4,5p;16,17p'
