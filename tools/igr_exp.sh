# (build first with P2C_NVCC_FLAGS=-DP2C_SS_DEBUG python -m point2cyl_b200.build --force: the toggles are compiled out otherwise)
# the implicit-network bench with parts of the layer kernel switched off (P2C_SS_DBG bits: 1 epilogue body, 2 correction
# read, 4 correction MMAs, 8 softplus' stores, 16 activation math, 32 TMA store of H)
for m in ${IGR_EXP_MODES:-0 1 8 16 32 24 56}; do
  P2C_SS_DBG=$m python bench.py --workload igr --steps 3 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dbg', $m, 'linear_act ms', round(d['roofline']['ms'],2), 'step', round(d['ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['stages'].items() if k in ('p2c_linear_act',)})"
done
