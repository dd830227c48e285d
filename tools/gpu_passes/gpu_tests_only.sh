cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -6
