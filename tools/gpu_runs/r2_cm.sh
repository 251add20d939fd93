timeout 900 python -m pytest tests/test_gpu_cm_networks.py -x -q -s 2>&1 | grep -v "Warning\|parse_version\|^$" | tail -30
