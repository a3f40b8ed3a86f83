timeout 600 python -m pytest tests -m gpu -x -q -k "multi_gpu or torchrun or symmetric or padded" 2>&1 | tail -5
