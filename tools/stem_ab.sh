timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -x 2>&1 | tail -5
for v in "A=1" "YB_STEM_MINB1=1" "YB_STEM_CTAS=3" "YB_STEM_CTAS=2"; do
  echo "== $v"; env $v python tools/layer_table.py n 256 2>/dev/null | sed -n 2,2p
done
