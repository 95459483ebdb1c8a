#!/bin/bash
# Developer tool: stage timings of kernel build variants (gpufluidsimulation_b200/lib/variants/*.so) at one size.
# usage: tools/ab_variants.sh [n=512] [combos=p1]
n=${1:-512}; combos=${2:-p1}
for so in gpufluidsimulation_b200/lib/libbimocq_b200.so gpufluidsimulation_b200/lib/variants/*.so; do
  BMQ_LIB=$PWD/$so python tools/stage_ab.py $n $combos 2>&1 | tail -n 2
done
