# round-2 GPU session 2: profiles (ncu launch list + per-kernel evidence + one --set full capture), sanitizer logs,
# the file-backed asynchronous input pipeline
export OMP_NUM_THREADS=16
mkdir -p gpurun_out profiles
NCU=ncu
# (1) launch list of the bench command itself (gpu__time_duration only)
$NCU --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no_cpu_baseline > gpurun_out/r02_launches_run.log 2>&1
# (2) per-kernel evidence of one profiled step
$NCU --profile-from-start off --clock-control none --csv --log-file gpurun_out/r02_kernels.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    python tools/profile_step.py > gpurun_out/r02_kernels_run.log 2>&1
python tools/summarize_ncu_kernels.py gpurun_out/r02_kernels.csv profiles/r02_kernel_evidence > gpurun_out/r02_kernels_sum.log 2>&1
# (3) --set full of the dominant conv kernel (2 launches)
$NCU --profile-from-start off --set full --clock-control none --import-source on -k regex:tapgemm_kmajor_pair_halo -s 20 -c 2 \
    -o gpurun_out/r02_conv_halo_full python tools/profile_step.py > gpurun_out/r02_full_run.log 2>&1
# (4) sanitizer
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file profiles/r02_sanitizer_$tool.log python tools/sanitize_small.py > gpurun_out/r02_san_$tool.out 2>&1
  tail -3 profiles/r02_sanitizer_$tool.log
done
# (5) file-backed input pipeline through run.py --benchmark
python tools/make_wav_dataset.py /tmp/os2s_wavs 128 15.0 > gpurun_out/r02_mkdata.log 2>&1
OS2S_DATASET_CSV=/tmp/os2s_wavs/data.csv timeout 600 python run.py --config_file=configs/jasper10x5_files.py --mode=train --benchmark --bench_steps=40 > gpurun_out/r02_run_benchmark_files.log 2>&1
tail -6 gpurun_out/r02_run_benchmark_files.log
tail -5 gpurun_out/r02_kernels_sum.log; head -30 profiles/r02_kernel_evidence.md
