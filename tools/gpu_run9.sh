#!/bin/bash
export ACX_WATCHDOG_MS=30000
echo "### pytest gpu"
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -n 15
echo "### smoke"
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -n 3
echo "### bench N=1"
timeout 300 python bench.py 2>&1 | grep -E '^\{|rror' | tail -n 2 | cut -c1-1600
