# One `ncu --set full` capture of each top kernel of the final build inside a real 148-walker sample (single lane, full
# batch); summaries go to profiles/ via tools/ncu_summary.py. Usage (on the GPU box): bash tools/ncu_top_kernels.sh
set -x
for spec in "svd_small_kernel 30" "panel_qr_reg_kernel 400" "apply_cols_kernel 300" "gett_large_kernel 120"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/r2_final_$1 \
      python tools/quick_time.py --walkers 148 --reps 1 --profile 0 > gpurun_out/r2_final_$1.log 2>&1
  tail -2 gpurun_out/r2_final_$1.log
done
ls -la gpurun_out/*.ncu-rep
