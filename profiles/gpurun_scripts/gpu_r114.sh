for W in cfg4 cfg3; do timeout 300 python profiles/gpurun_scripts/stream_probe.py $W 2>&1 | tail -6; done
echo done
