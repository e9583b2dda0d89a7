# usage: tc_sweep.sh  (prints device time per launch for representative ResNet-18 shapes, both back ends)
for shape in "0 @ 1 56 56 64 64 3 1 1 1" "0 @ 1 28 28 128 128 3 1 1 1" "0 @ 1 14 14 256 256 3 1 1 1" "0 @ 1 7 7 512 512 3 1 1 1" "1 @ 1 56 56 64 64 3 1 1 1" "1 @ 1 14 14 256 256 3 1 1 1" "1 @ 1 28 28 128 256 3 2 1 1" "2 @ 1 14 14 256 256 3 1 1 0" "2 @ 1 7 7 512 512 3 1 1 0" "0 @ 8 56 56 64 64 3 1 1 1" "0 @ 8 14 14 256 256 3 1 1 1" "0 @ 1 8 16 64 128 1 1 0 0"; do
  for be in $BACKENDS; do
    python profiles/run_one_gemm.py ${shape/@/$be} 50 2>&1 | tail -1
  done
done
