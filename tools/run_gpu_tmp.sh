set -x
python tools/dct_tf32_study.py > gpurun_out/dct_tf32_study_r02.json 2> gpurun_out/dct_tf32.err; cat gpurun_out/dct_tf32_study_r02.json | head -60; tail -3 gpurun_out/dct_tf32.err
timeout 1500 python -m pytest tests/test_gpu_unit.py -m gpu -q 2>&1 | tail -5
