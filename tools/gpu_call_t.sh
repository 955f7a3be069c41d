#!/bin/sh
# timing only (NoCutoff lines of hivrt + 2clr), optional environment through the caller
python tools/quick_time.py hivrt 2clr 2>&1 | grep -A2 "method=0"
