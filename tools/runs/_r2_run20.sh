#!/bin/bash
timeout 300 python -m pytest tests/test_joblight.py -x -q -m gpu 2>&1 | grep -E "^E|Error|assert" | head -20
