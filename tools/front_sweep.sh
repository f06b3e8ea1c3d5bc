for cfg in "" "3,2" "3,4" "3,8" "6,2" "6,4" "6,8" "1,4"; do
  echo "== LPC_FRONT=$cfg"
  LPC_FRONT=$cfg timeout 200 python tools/stream_sweep.py lpc 64 640 1 2>&1 | tail -1
done
