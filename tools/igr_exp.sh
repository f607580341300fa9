for m in 0 1 2 4 6 7; do
  P2C_SS_DBG=$m python bench.py --workload igr --steps 3 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dbg', $m, 'linear_act ms', round(d['roofline']['ms'],2), 'step', round(d['ms_per_step'],2))"
done
