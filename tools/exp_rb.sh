for cfg in "4 148 256" "4 148 128" "1 148 128" "4 148 192"; do
  set -- $cfg
  echo "streams=$1 walkers=$2 rb=$3"
  PEPS_QR_RB=$3 python bench.py --steps 2 --warmup 3 --secondary 0 --no-cpu-baseline --e2e-steps 0 --streams $1 --walkers $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],2), d['roofline']['per_class_ms'])"
done
