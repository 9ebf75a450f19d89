python -m pytest tests -m gpu -x -q -s -k "compiled and not l1l2" 2>&1 | grep -v "^init\|^/root/ref" | tail -6
P="python tools/profile_solve.py --max-iter 0 --repeat 2"
for T in 1 2 4; do
$P --batch 65536 --tiles $T 2>&1 | tail -1 | cut -c1-130
$P --batch 262144 --workload mpc_dcmotor --tiles $T 2>&1 | tail -1 | cut -c1-130
done
$P --batch 37888 --max-iter 3 --tiles 1 2>&1 | tail -1 | cut -c1-100
$P --batch 37888 --max-iter 3 --tiles 2 2>&1 | tail -1 | cut -c1-100
$P --batch 75776 --max-iter 3 --tiles 2 2>&1 | tail -1 | cut -c1-100
$P --batch 151552 --max-iter 3 --tiles 4 2>&1 | tail -1 | cut -c1-100
