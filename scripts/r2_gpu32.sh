#!/bin/bash
# last check of the committed tree: parity suite + smoke + a short bench
( timeout 1200 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cold 2>&1 | tail -1 | cut -c1-220
