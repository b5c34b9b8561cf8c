#!/bin/bash
# round 2: whole GPU suite + smoke at HEAD (after the last host-side changes)
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | cut -c1-330
