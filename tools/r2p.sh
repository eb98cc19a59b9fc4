for dbg in 1 3 5 7; do
ZS_CHAIN_DBG=$dbg timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'chain_pmlp|chain_qkvattn2' --launch-skip 4 --launch-count 2 --csv --log-file gpurun_out/r2p_dbg$dbg.csv python tools/diag_decoder.py 2146689 --once > /dev/null 2>&1
echo "dbg $dbg rc $?"
done
