for v in "A=1" "YB_STEM_MINB1=1" "YB_STEM_MINB1=1 YB_STEM_CTAS=2"; do
  echo "== $v"; env $v python tools/layer_table.py n 256 2>/dev/null | sed -n 2,2p
done
