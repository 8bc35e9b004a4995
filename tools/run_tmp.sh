python -m pytest tests/test_gpu_parity.py -q -x -k "engines_agree or against_oracle" 2>&1 | tail -3
for v in 0 4 5 6; do python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --grid 96 --variant $v | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2/96 variant', $v, '%.4g'%d['value'], '%.3f'%d['ms_per_step'], '%.3f'%d['roofline']['kernel_ms'])"; done
