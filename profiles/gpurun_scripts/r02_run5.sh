# round 2, run 5: what bounds the wide GEMM when the tensor work is removed (experiments build; results are garbage)
mkdir -p gpurun_out
export PS_B200_LIB=$PWD/puresound_b200/libpuresound_b200_exp.so
for dbg in 0 16 17 18 20 24 19 21 22 23 27 29 30 31 1 2 4 8; do
  PS_WIDE_DBG=$dbg python profiles/gemm_probe.py 2>&1 | tail -1
done | tee gpurun_out/r02_run5_probe.txt
