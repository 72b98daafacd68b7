# accuracy of the headline arithmetic (fp16c as created) on 100,000 positions vs the fp32 mode, three weight sets
SECONDS=0
for s in 1 2 3; do
  WSEED=$s CONFIGS=0:64:1 SUSTAIN=0.3 timeout 55 python tools/gpu_partial_split.py > gpurun_out/s5_accuracy_seed$s.log 2>&1
  echo "seed $s rc=$? at ${SECONDS}s"; tail -n 1 gpurun_out/s5_accuracy_seed$s.log | cut -c1-400
done
