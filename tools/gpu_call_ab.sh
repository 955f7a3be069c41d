#!/bin/sh
echo "== default 320"; python tools/quick_time.py hivrt 2clr 2>&1 | grep -A1 "method=0" | grep "k_tree_gamma" | sed 's/.*k_tree_gamma/k_tree_gamma/'
for n in 192 256 448; do echo "== $n"; AGBNP_B200_LIB=variants/gam$n/libagbnp_b200.so python tools/quick_time.py hivrt 2clr 2>&1 | grep -A1 "method=0" | grep "k_tree_gamma" | sed 's/.*k_tree_gamma/k_tree_gamma/'; done
python tools/quick_parity.py hivrt_standin 2clr
