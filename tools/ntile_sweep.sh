for cfg in "LPC_TC_NTILE_AUTO=0" "LPC_TC_NTILE_AUTO=1" "LPC_TC_NTILE_AUTO=0 LPC_TC_NTILE_MAX=128" "LPC_TC_NTILE_AUTO=0 LPC_TC_NTILE_MAX=64"; do
  echo "== $cfg"
  env $cfg timeout 200 python tools/stream_sweep.py lpc 64 640 1 2>&1 | tail -1
  env $cfg timeout 200 python tools/stream_sweep.py lpc 1 640 1 2>&1 | tail -1
  env $cfg timeout 200 python tools/stream_sweep.py yolov10b 1 640 1 2>&1 | tail -1
  env $cfg timeout 200 python tools/stream_sweep.py yolov10s 256 640 1 2>&1 | tail -1
done
