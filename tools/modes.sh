python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 0 1 2; do echo "== cfg $c"; B2S_DWT_CFG=$c python tools/quick_time.py 32 8 db10 1 2>&1 | grep -E "dwt_...@L[12]"; done
