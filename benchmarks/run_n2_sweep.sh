mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_dense_ext.py tests/test_gpu_large_parity.py -q 2>&1 | tail -12) > gpurun_out/r2_n2_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_n2_default.json 2> gpurun_out/r2_n2_default.err
for p in 0.99 0.9 0.8 0.6 0.2; do
  $TR bench.py --gpus 2 --steps 5 --warmup 3 --p-local $p --no-e2e --no-cpu --no-parity > gpurun_out/r2_n2_p$p.json 2> gpurun_out/r2_n2_p$p.err
done
$TR bench.py --gpus 2 --steps 5 --warmup 3 --dist nccl --no-e2e --no-cpu > gpurun_out/r2_n2_nccl.json 2> gpurun_out/r2_n2_nccl.err
tail -12 gpurun_out/r2_n2_tests.log
head -c 300 gpurun_out/r2_n2_default.json; echo; tail -3 gpurun_out/r2_n2_default.err
for p in 0.99 0.9 0.8 0.6 0.2; do python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2_n2_p$p.json')); print('$p', d['ms_per_step'], d['value'], d['engine']['halo']['remote_edge_fraction'], d['engine']['halo']['nvlink_gbs_per_gpu_in_gather'])
except Exception as e: print('$p ERR', e)
"; done
