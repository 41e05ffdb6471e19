set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc; free -g | head -2
python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -30
