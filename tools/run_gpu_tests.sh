#!/bin/bash
# usage (under gpurun): bash tools/run_gpu_tests.sh [pytest args]
python -m pytest tests -x -q -m gpu "$@" 2>&1 | tail -40
