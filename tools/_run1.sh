set -u
O=gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -4
timeout 200 python tools/bench_gemm.py 2>&1 | grep bnbwd | cut -c1-400
timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_fullscale_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -12
timeout 200 python bench.py --no-cpu-baseline > $O/r2g_bench.json 2>/dev/null; python -c "
import json; d=json.loads(open('$O/r2g_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e'].get('ms_per_step'), json.dumps(d['parity_check'])[:900]); print(d.get('gpu_launches'))"
