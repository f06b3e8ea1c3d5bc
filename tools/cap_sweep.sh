for cfg in "LPC_CTA_CAP=2 1" "LPC_CTA_CAP=1 1" "LPC_CTA_CAP=1 2" "LPC_CTA_CAP=1 3"; do
  set -- $cfg
  echo "== $1 streams=$2"
  env $1 timeout 200 python tools/stream_sweep.py lpc 64 640 $2 2>&1 | tail -1
done
