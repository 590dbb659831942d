timeout 500 python -m pytest tests/test_encode.py tests/test_encops.py tests/test_options.py tests/test_golden.py tests/test_pool.py -m gpu -x -q 2>&1 | tail -3
DSV_PROFILE=2 GOPN=24 timeout 200 python tools/scale_probe.py 1,32 > gpurun_out/probe_s2.log 2>&1
grep -a "^threads" gpurun_out/probe_s2.log
for k in "motion search" "sub+fwd+quant+inv" "reconstruct+filters"; do grep -a -o "$k [0-9.]*" gpurun_out/probe_s2.log | awk -v k="$k" '{v=$NF; s+=v; n++; if(n==1) f=v} END{print k, "solo", f, "loaded avg", (s-f)/(n-1)}'; done
timeout 120 python tools/opbench.py 2>&1 | cut -c1-140 | head -3
