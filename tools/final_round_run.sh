#!/bin/bash
# One gpurun call that regenerates the round's 1-GPU artefacts under gpurun_out/final_* (copied to profiles/r2_* afterwards).
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.txt 2>&1
timeout 400 python bench.py > $O/final_bench_1gpu.json 2> $O/final_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_ref.json 2> $O/final_bench_ref.err
timeout 400 python tools/bench_configs.py > $O/final_configs.jsonl 2> $O/final_configs.err
timeout 200 python tools/bench_spmm.py --widths 256 128 40 --variants 0 2 --iters 15 --weighted-only > $O/final_spmm_sweep.jsonl 2>/dev/null
timeout 200 python tools/bench_spmm.py --widths 256 128 --variants 0 2 --iters 15 --weighted-only --p-local 0.8 >> $O/final_spmm_sweep.jsonl 2>/dev/null
timeout 200 python tools/bench_gemm.py > $O/final_gemm_microbench.jsonl 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py $O/final_launches.csv 19 > $O/final_launch_summary.txt 2>/dev/null
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmm_rows_bulk_kernel -s 4 -c 1 -f -o $O/final_spmm_k256 python tools/bench_spmm.py --widths 256 --iters 4 --weighted-only > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmm_rows_gather4_kernel -s 4 -c 1 -f -o $O/final_spmm_k128 python tools/bench_spmm.py --widths 128 --iters 4 --weighted-only > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_kernel -s 30 -c 1 -f -o $O/final_gemm_bnbwd python tools/bench_gemm.py > /dev/null 2>&1
cat $O/final_pytest.txt; tail -1 $O/final_smoke.txt; wc -l $O/final_bench_1gpu.json $O/final_bench_ref.json $O/final_configs.jsonl; ls -la $O/*.ncu-rep
