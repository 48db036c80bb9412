for cfg in "4 148 1" "4 148 0" "2 148 0" "1 148 0" "8 296 0" "4 296 0" "6 222 0"; do
  set -- $cfg
  echo "streams=$1 walkers=$2 early=$3"
  PEPS_QR_EARLY_STOP=$3 python bench.py --steps 2 --warmup 2 --secondary 0 --no-cpu-baseline --e2e-steps 0 --streams $1 --walkers $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],2), d['roofline']['per_class_ms'])"
done
