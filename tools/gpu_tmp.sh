timeout 900 python -m pytest tests/test_loss_gpu.py -q -m gpu --timeout 600 2>&1 | tail -8
