# round 2, run 6: does the 2000 ns suspend hint of mbarrier.try_wait delay wake-ups?  Same kernel, hint 0 (plain try_wait) / 200 / 2000 ns
mkdir -p gpurun_out
for lib in libps_exp_h0 libps_exp_h200 libpuresound_b200_exp; do
  export PS_B200_LIB=$PWD/puresound_b200/$lib.so
  for dbg in 0 16 31 20; do
    echo -n "$lib "; PS_WIDE_DBG=$dbg python profiles/gemm_probe.py 2>&1 | tail -1
  done
  echo -n "$lib narrow "; PS_TC_WIDE=0 python profiles/gemm_probe.py 2>&1 | tail -1
done | tee gpurun_out/r02_run6_hint.txt
