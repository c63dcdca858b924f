#!/bin/bash
# Launch-geometry sweep of the sorted MH kernel on C2 (needs a library built with --all-cfgs).
# usage: scripts/sweep_sorted.sh "cfg:nc cfg:nc ..."   (nc 0 = the geometry's own choice)
for item in ${1:-"0:0 1:0"}; do
    cfg=${item%%:*}; nc=${item##*:}
    echo "=== cfg $cfg nc $nc"
    if [ "$nc" = "0" ]; then
        PTMCMC_SORT_CFG=$cfg timeout 300 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep -E "kernel|rep|accept"
    else
        PTMCMC_SORT_CFG=$cfg PTMCMC_SORT_NC=$nc timeout 300 python scripts/quick_bench.py 20 8192 32 1000 2 2>&1 | grep -E "kernel|rep|accept"
    fi
done
