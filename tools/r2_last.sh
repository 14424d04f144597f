export OMP_NUM_THREADS=16
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "perturbation or featur or augment or spectrogram" > gpurun_out/r02_last_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02_last_tests.log | cut -c1-200
