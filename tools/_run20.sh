timeout 900 python -m pytest tests -m gpu -x -q -s -k "band_lu" 2>&1 | tail -12
