mkdir -p gpurun_out
for W in cfg4 cfg3 cfg1; do
timeout 600 python profiles/gpurun_scripts/model_breakdown.py $W > gpurun_out/r108_${W}_breakdown.txt 2>&1; echo "== $W"; tail -24 gpurun_out/r108_${W}_breakdown.txt | cut -c1-80,150-215 | head -16
done
echo done
