# 2 GPUs of one box (gpurun --gpus 2): the world-size-2 GPU tests and the sharded bench lines
O=gpurun_out/final_n2
mkdir -p $O
python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_gpu_multi.log 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
$T bench.py --gpus 2 --no-cpu-baseline > $O/bench_n2_cfgA.json 2> $O/bench_n2_cfgA.err
for c in cfgB cfgC cfgD; do $T bench.py --gpus 2 --config $c > $O/bench_n2_$c.json 2> $O/bench_n2_$c.err; done
tail -3 $O/pytest_gpu_multi.log
