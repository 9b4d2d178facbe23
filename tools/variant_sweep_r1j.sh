mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r1j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1j_pytest.log
for v in base inp1 cp1 cp2; do
  if [ $v == base ]; then L=""; else L="dps_ttc_b200/build_variants/libdpsttc_$v.so"; fi
  for n in 8 32; do
    DPSTTC_LIB=$L timeout 90 python tools/kernel_bench.py --n $n --iters 50 --graph --only inpaint,gather > gpurun_out/r1j_var_${v}_n$n.jsonl 2> gpurun_out/r1j_var_${v}_n$n.err
  done
  DPSTTC_LIB=$L timeout 90 python tools/kernel_bench.py --n 128 --only inpaint,gather > gpurun_out/r1j_var_${v}_n128.jsonl 2> gpurun_out/r1j_var_${v}_n128.err
done
tail -3 gpurun_out/r1j_pytest.log
