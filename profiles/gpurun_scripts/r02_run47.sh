# round 2, run 47: ncu --set full of the persistent hop kernel at 256 streams (where do its ~100 phases spend their 10-14 us)
mkdir -p gpurun_out
timeout 300 python profiles/hop_once.py > gpurun_out/r02_run47_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:stream_hop -s 6 -c 1 -o gpurun_out/r02_run47_hop python profiles/hop_once.py > gpurun_out/r02_run47_ncu.log 2>&1; tail -3 gpurun_out/r02_run47_ncu.log
ls -la gpurun_out/r02_run47_hop.ncu-rep
