P="python tools/profile_solve.py"
for v in "--merge 8 --divcall 1" "--divcall 0" "--merge 24" "--merge 24 --divcall 0"; do
$P --max-iter 0 --batch 65536 --repeat 2 $v 2>&1 | tail -1 | cut -c1-140
$P --max-iter 0 --batch 96 --repeat 2 $v 2>&1 | tail -1 | cut -c1-140
$P --max-iter 3 --batch 37888 --repeat 2 $v 2>&1 | tail -1 | cut -c1-140
done
