#!/bin/bash
# GPU box: tools/ab_lib2.sh over the given builds, output under gpurun_out/<tag>_ab.txt.   usage: gpu_ab_only.sh <tag> lib...
tag=$1; shift
bash tools/ab_lib2.sh "$@" 2>&1 | tee gpurun_out/${tag}_ab.txt
