#!/bin/bash
# round 2: parity after the unused-modality / unequal-weight changes (loss + exchange suites)
timeout 900 python -m pytest tests/test_loss_gpu.py tests/test_loss_exchange_gpu.py -q -m gpu --timeout 600 -x 2>&1 | tail -6
