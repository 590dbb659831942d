for pct in 58 100 -1; do
echo "== DSVCU_CARVEOUT=$pct"
DSVCU_CARVEOUT=$pct GOPN=24 timeout 200 python tools/scale_probe.py 1,16,32 2>&1 | grep -a "^threads"
done
DSVCU_CARVEOUT=58 timeout 100 python tools/scale_probe_dec.py 1,32 2>&1 | tail -2
DSVCU_CARVEOUT=-1 timeout 100 python tools/scale_probe_dec.py 1,32 2>&1 | tail -2
