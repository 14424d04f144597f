# final tree: ncu launch list of the bench command + per-kernel evidence of one profiled step
export OMP_NUM_THREADS=16
mkdir -p gpurun_out
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/r02_final_launches.csv \
    python bench.py --steps 2 --warmup 1 --no_cpu_baseline > gpurun_out/r02_final_launches_run.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r02_final_launches.csv)"
timeout 150 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r02_final_kernels.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    python tools/profile_step.py > gpurun_out/r02_final_kernels_run.log 2>&1
echo "kernel evidence rc=$? lines=$(wc -l < gpurun_out/r02_final_kernels.csv)"
