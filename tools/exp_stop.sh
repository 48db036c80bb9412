for cfg in "0 1" "1 1" "1 2" "1 3" "0 1"; do
  set -- $cfg
  echo "early=$1 stride=$2"
  PEPS_QR_EARLY_STOP=$1 PEPS_QR_STOP_STRIDE=$2 python bench.py --steps 3 --warmup 3 --secondary 0 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],2), d['roofline']['per_class_ms'])"
done
