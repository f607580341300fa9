echo "streamed-weight kernel (W from smem)"; P2C_TC_FORCE_SS=1 python tools/bench_layers_dbg.py run 2>&1 | tail -5
echo "W-in-TMEM kernel"; python tools/bench_layers_dbg.py run 2>&1 | tail -5
