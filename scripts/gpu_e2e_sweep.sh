#!/bin/bash
for th in 2 3 4; do for ch in 1 2 4 8; do
  echo "== threads $th chunk $ch"
  VP_E2E_THREADS=$th VP_E2E_CHUNK=$ch timeout 300 python bench.py --steps 48 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'lat', round(d['latency_mode']['value']))"
done; done
