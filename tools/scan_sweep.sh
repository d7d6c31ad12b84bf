for m in 1 4 8 12 16 33; do
  echo "scan_min $m"
  RTU_SCAN_MIN=$m python tools/quickbench.py Teapot/scene2.xml 1920 1080 whitted 128 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms'], d['kernels_ms'])"
  RTU_SCAN_MIN=$m python tools/quickbench.py Project11/scene.xml 800 600 path 64 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms'], d['kernels_ms'])"
done
