set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu_a.json 2> gpurun_out/r02_bench_1gpu_a.err; echo "bench rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu-baseline --sustained-seconds 0"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_ncu_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_stack_tc -s 4 -c 1 -o gpurun_out/r02_conv_fp16c $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
tail -c 600 gpurun_out/r02_bench_1gpu_a.err
