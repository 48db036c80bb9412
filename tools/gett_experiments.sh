#!/bin/bash
# device time of contraction variants (layout experiments), W walkers
W=${1:-37}
export PEPS_EINSUM_TIME=5
run() { python tools/run_einsum.py --spec "$1" --da "$2" --db "$3" --W $W --reps 1 2>&1 | grep einsum; }
echo "# BTen step, current layouts"
run "apx,xyz->apyz" 64,8,64 64,8,64
run "apyz,yfop->zafo" 64,8,8,64 8,8,8,8
run "zafo,zfb->aob" 64,64,8,8 64,8,64
echo "# BTen step, contracted-index-major intermediates"
run "apx,xyz->pyaz" 64,8,64 64,8,64
run "pyaz,yfop->zfao" 8,8,64,64 8,8,8,8
run "zfao,zfb->aob" 64,8,64,8 64,8,64
echo "# absorption forward, current"
run "apb,kea->ekpb" 64,8,64 320,8,64
run "ekpb,epfo->kofb" 8,320,8,64 8,8,8,8
echo "# absorption forward, (e,p)-major tmp1"
run "apb,kea->epkb" 64,8,64 320,8,64
run "epkb,epfo->kofb" 8,8,320,64 8,8,8,8
echo "# absorption backward, current"
run "apb,fbj->apfj" 64,8,64 8,64,64
run "apfj,epfo->eaoj" 64,8,8,64 8,8,8,8
run "kea,eaoj->koj" 320,8,64 8,64,8,64
run "eaoj,toj->eat" 8,64,8,64 64,8,64
echo "# absorption backward, (p,f)-major Y"
run "apb,fbj->pfaj" 64,8,64 8,64,64
run "pfaj,epfo->eaoj" 8,8,64,64 8,8,8,8
