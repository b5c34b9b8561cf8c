timeout 1500 python -m pytest tests/test_loss_gpu.py tests/test_infonce_gpu.py -q -m gpu --timeout 600 -x 2>&1 | tail -25
