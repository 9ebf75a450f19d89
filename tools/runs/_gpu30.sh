cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
M="gpu__time_duration.sum,smsp__inst_executed.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction_lookup_hit.sum,gcc__cache_requests_type_instruction_lookup_miss.sum,gcc__cache_requests_type_constant.sum,gcc__cache_requests_type_constant_lookup_miss.sum,lts__t_requests_srcunit_gcc.sum,lts__t_requests_srcunit_gcc_lookup_hit.sum,lts__t_requests_srcunit_gcc_lookup_miss.sum,gcc__xbar2gcc_sectors.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_ltcfabric.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"
prof() { # name batch skip mopts
  n=$1; b=$2; sk=$3; shift; shift; shift
  timeout 600 ncu --metrics $M --clock-control none -k regex:tcm_solve_kernel -s $sk -c 1 --csv --log-file gpurun_out/r03f_ncu_$n.csv python tools/profile_solve.py --batch $b --max-iter 8 "$@" > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r03f_ncu_$n.csv")) if len(r)>10]
h=rows[0]; 
for r in rows[1:]:
    d=dict(zip(h,r)); print("  $n", d["Metric Name"], d["Metric Value"])
PY
}
prof pf0 56832 3 --mopts prefetch=0
prof pf0ls1 56832 3 --mopts prefetch=0,lockstep=1
