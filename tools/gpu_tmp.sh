timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -s 2>&1 | grep -E "saturated|trained regime|passed|failed|FAILED|Error|assert" | head -60
