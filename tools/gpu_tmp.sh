timeout 900 python -m pytest tests/test_loss_gpu.py -q -m gpu -x -k "shared_s or full_size" 2>&1 | tail -15
