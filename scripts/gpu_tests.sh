make -s -C oracle
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -30
