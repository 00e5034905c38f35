#!/bin/bash
# Developer tuning: outer panel width of gp_potrf_f64 against the headline bench
for nb in 256 384 512 768 1024; do
  GP_POTRF_NB=$nb python bench.py --steps 6 --warmup 3 --no-secondary --no-cpu 2>/dev/null | NB=$nb python -c "import json,sys,os; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(os.environ['NB'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['executed_tile_tflops'])"
done
