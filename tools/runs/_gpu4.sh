cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P="python tools/profile_solve.py"
$P --batch 37888 --max-iter 6 --repeat 2 2>&1 | tail -1 | cut -c1-110
$P --batch 37888 --max-iter 6 --repeat 2 --prefetch 0 2>&1 | tail -1 | cut -c1-110
for l in 2 4 8; do $P --batch 37888 --max-iter 6 --repeat 2 --lockstep $l 2>&1 | tail -1 | cut -c1-110; done
$P --batch 75776 --max-iter 6 --repeat 2 --tiles 2 2>&1 | tail -1 | cut -c1-110
$P --batch 75776 --max-iter 6 --repeat 2 2>&1 | tail -1 | cut -c1-110
for l in 2 4 8; do $P --batch 65536 --max-iter 0 --repeat 2 --lockstep $l 2>&1 | tail -1 | cut -c1-130; done
$P --batch 65536 --max-iter 0 --repeat 2 --prefetch 0 2>&1 | tail -1 | cut -c1-130
