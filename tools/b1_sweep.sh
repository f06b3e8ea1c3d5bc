for cfg in "LPC_TC_DEEP_RING=0" "LPC_TC_DEEP_RING=1"; do
  echo "== $cfg"
  for s in 320 640 960; do env $cfg timeout 200 python tools/stream_sweep.py yolov10b 1 $s 1 2>&1 | tail -1; done
  env $cfg timeout 200 python tools/stream_sweep.py lpc 1 640 1 2>&1 | tail -1
  env $cfg timeout 200 python tools/stream_sweep.py lpc 64 640 1 2>&1 | tail -1
done
