#!/bin/bash
export ACX_WATCHDOG_MS=30000
echo "### pytest gpu (no -x)"
timeout 1200 python -m pytest tests/ -q -m gpu 2>&1 | tail -n 40
