#!/bin/bash
# device time of the six contractions of one bulk site of a row absorption (10x10, D=8, chi=64), W walkers
W=${1:-37}
export PEPS_EINSUM_TIME=5
run() { python tools/run_einsum.py --spec "$1" --da "$2" --db "$3" --W $W --reps 1 2>&1 | grep einsum; }
run "apb,kea->kepb" 64,8,64 512,8,64
run "kepb,epfo->kofb" 512,8,8,64 8,8,8,8
run "apb,fbj->apfj" 64,8,64 8,64,64
run "apfj,epfo->eaoj" 64,8,8,64 8,8,8,8
run "kea,eaoj->koj" 512,8,64 8,64,8,64
run "eaoj,toj->eat" 8,64,8,64 64,8,64
