#!/bin/sh
python tools/quick_time.py hivrt 2clr 2>&1 | grep -v "^   counters"
python tools/quick_parity.py
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
