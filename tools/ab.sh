#!/bin/bash
# Same-box A/B of two library builds: A = committed HEAD, B = working tree.
#   tools/ab.sh '<command printing a number>'      (run from the repo root, needs gpurun)
set -e
rm -rf /tmp/abuild && mkdir -p /tmp/abuild && git archive HEAD | tar -x -C /tmp/abuild
(cd /tmp/abuild && python -m news_recommendation_project_v2_b200.build > /dev/null 2>&1)
cp /tmp/abuild/news_recommendation_project_v2_b200/libnrb200.so tools/libnrb200_A.so
/usr/local/graft/bin/gpurun --timeout 900 -- "for i in 1 2 3; do for L in tools/libnrb200_A.so news_recommendation_project_v2_b200/libnrb200.so; do echo -n \"\$L \"; NRB200_LIB=\$PWD/\$L $1; done; done" 2>&1 | tail -8
rm -f tools/libnrb200_A.so
