# ncu evidence for the byte kernels at 4 Mi positions (beyond L2): launch list with DRAM bytes, then --set full of one launch each
SECONDS=0
timeout 60 python tools/gpu_byte_kernels_ncu.py > gpurun_out/s4_bytes_plain.log 2>&1; echo "plain rc=$? at ${SECONDS}s"
timeout 70 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/s4_ncu_byte_kernels.csv python tools/gpu_byte_kernels_ncu.py > gpurun_out/s4_bytes_ncu1.log 2>&1; echo "ncu1 rc=$? at ${SECONDS}s"
timeout 70 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'encode_kernel|legal_mask_kernel' -c 2 -f \
    -o gpurun_out/s4_byte_kernels python tools/gpu_byte_kernels_ncu.py > gpurun_out/s4_bytes_ncu2.log 2>&1; echo "ncu2 rc=$? at ${SECONDS}s"
tail -2 gpurun_out/s4_bytes_plain.log gpurun_out/s4_bytes_ncu1.log gpurun_out/s4_bytes_ncu2.log
