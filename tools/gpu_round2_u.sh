#!/bin/bash
timeout 900 python -m pytest tests/test_attn_gpu.py -m gpu -q 2>&1 | grep -E "^E|passed|failed" | head -20
