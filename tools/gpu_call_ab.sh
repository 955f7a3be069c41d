#!/bin/sh
for caps in "512,192,64" "576,224,96" "640,256,96"; do
 echo "== caps $caps default"; AGBNP_B200_INIT_CAPS=$caps python tools/quick_time.py hivrt 2>&1 | grep -A1 "method=0" | grep k_tree | sed 's/k_born=.*//'
 echo "== caps $caps 1-warp CTAs"; AGBNP_B200_INIT_CAPS=$caps AGBNP_B200_LIB=variants/warp1/libagbnp_b200.so python tools/quick_time.py hivrt 2>&1 | grep -A1 "method=0" | grep k_tree | sed 's/k_born=.*//'
done
