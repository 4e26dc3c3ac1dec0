#!/usr/bin/env bash
# reference arm + our arm back to back, as the driver runs them
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2>&1 | tail -1
python bench.py --gpus 1 --steps 5 --warmup 3 2>&1 | tail -1
