python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/quick_time.py 32 8 db10 1 2>&1 | grep -E "notch|prologue|epilogue|dwt_...  |sum|end"
