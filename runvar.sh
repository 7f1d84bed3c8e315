for f in variants_k8_b1 variants_k8_b2 variants_k4_b1 variants_k4_b2 variants_k4_b3 variants_k4_b4; do
  echo "== $f"
  ENRGY_B200_LIB=$PWD/$f.so python bench.py --steps 5 --warmup 3 --t 600 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['kernel'])"
done
