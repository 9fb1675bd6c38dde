timeout 600 python -m pytest tests/test_gpu_mlp_core.py tests/test_gpu_train_step.py -q -x 2>&1 | tail -15
