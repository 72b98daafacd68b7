# final check of the round: GPU tests, smoke, the small-batch ncu capture, and (time permitting) a bench line
SECONDS=0
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/s3_pytest_gpu.log 2>&1; echo "pytest rc=$? at ${SECONDS}s"
timeout 50 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s3_smoke.log 2>&1; echo "smoke rc=$? at ${SECONDS}s"
timeout 40 python tools/gpu_small_batch_ncu.py > gpurun_out/s3_small_plain.log 2>&1 && \
timeout 80 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/s3_small_batch \
    python tools/gpu_small_batch_ncu.py > gpurun_out/s3_small_ncu.log 2>&1; echo "ncu rc=$? at ${SECONDS}s"
if [ $SECONDS -lt 270 ]; then
  timeout $((375 - SECONDS)) python bench.py --no-cpu-baseline > gpurun_out/s3_bench_1gpu.json 2> gpurun_out/s3_bench_1gpu.err; echo "bench rc=$? at ${SECONDS}s"
fi
tail -3 gpurun_out/s3_pytest_gpu.log; cat gpurun_out/s3_small_plain.log
