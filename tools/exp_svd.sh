python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in 0 1 0 1; do
  echo "svd_blocked=$v"
  PEPS_SVD_BLOCKED=$v python bench.py --steps 3 --warmup 3 --secondary 0 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],2), d['roofline']['per_class_ms'])"
done
