set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
env | grep -i nccl > $O/env_nccl.txt; python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
[ -x profiles/tools/build/fp32_pipe_probe ] || (mkdir -p profiles/tools/build && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/tools/build/fp32_pipe_probe profiles/tools/fp32_pipe_probe.cu)
profiles/tools/build/fp32_pipe_probe > $O/fp32_pipe_probe.jsonl 2>&1
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err
for c in cfgB cfgC cfgD bunny; do python bench.py --config $c > $O/bench_n1_$c.json 2> $O/bench_n1_$c.err; done
python bench.py --batch 296 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_a.csv python bench.py --batch 296 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_launches.log 2>&1
python bench_stages.py --case k4 > $O/k4_stage.json 2>&1 && ncu --set full --clock-control none --import-source on -k regex:score_batch --launch-skip 2 --launch-count 1 -o $O/r2_k4_final python bench_stages.py --case k4 > $O/ncu_k4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_mask --launch-skip 1 --launch-count 1 -o $O/r2_k1_b296_final python bench.py --batch 296 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_k1.log 2>&1
tail -2 $O/pytest_gpu.log
ls -la $O
