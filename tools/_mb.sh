timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_multi or device_circuit" 2>&1 | tail -4
for f in 3; do echo "== multi_f=$f (5 CTAs)"; PW_MULTI_F=$f timeout 120 python tools/multi_bench.py 12 c128 2>&1 | tail -13; PW_MULTI_F=$f timeout 120 python tools/multi_bench.py 12 c64 2>&1 | tail -13; done
