# A/B of the tcgen05 GEMM knobs (env) on a few shapes + config 2; output in gpurun_out/ab.log
run() { echo "== $1" >> gpurun_out/ab.log; env $1 python profiles/gemm_suite.py 1 30 2>&1 | grep -E "mode 0 \[1,56,56,64->64,k3,s1,dual\]|mode 0 \[1,7,7,512|mode 0 \[8,56|mode 1 \[8,28|mode 2 \[8" >> gpurun_out/ab.log; env $1 python profiles/bench_config.py config2 200 2>&1 | tail -1 >> gpurun_out/ab.log; }
rm -f gpurun_out/ab.log
python -m pytest tests/test_kernels_gpu.py -q -k conv 2>&1 | tail -1 >> gpurun_out/ab.log
run "BRE_TC_PRODUCERS=2 BRE_FUSE_BNACT=0"
run "BRE_TC_PRODUCERS=2 BRE_FUSE_BNACT=1"
run "BRE_TC_PRODUCERS=1 BRE_FUSE_BNACT=0"
run "BRE_TC_PRODUCERS=4 BRE_FUSE_BNACT=0"
for e in "BRE_FUSE_BNACT=0 BRE_TC_TMA=0" "BRE_FUSE_BNACT=0 BRE_TC_TMA=1" "BRE_FUSE_BNACT=1 BRE_TC_TMA=0"; do
  echo "== config3 test $e" >> gpurun_out/ab.log
  env $e python -m pytest tests/test_engine_gpu.py -q --tb=short -k config3 2>&1 | grep -E "AssertionError|passed|failed" >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
