export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q --tb=short -x 2>&1 | tail -6
timeout 900 python tools/run_continual.py 2>&1 | tail -1 | tee gpurun_out/continual_5task.json | cut -c1-1500
