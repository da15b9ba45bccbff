export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_stack_gpu.py tests/test_model_gpu.py tests/test_training_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-extra-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'mem', round(d['memory']['peak_allocated_gb'],1), {k: round(v['ms_per_step'],2) for k,v in d['kernels'].items() if k in ('papr_stack_bf16','papr_wgrad_bf16')})"; }
run PAPR_BWD_OVERLAP=0 PAPR_BWD_SLICE_ROWS=8388608
run PAPR_BWD_OVERLAP=1
run PAPR_BWD_DGRAD_CTAS=96 PAPR_BWD_WGRAD_CTAS=52
run PAPR_BWD_DGRAD_CTAS=80 PAPR_BWD_WGRAD_CTAS=68
run PAPR_BWD_SLICE_ROWS=1048576
run PAPR_BWD_SLICE_ROWS=4194304
