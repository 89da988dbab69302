#!/bin/bash
python -m pytest tests -m gpu -x -q -k "solve or smoke" 2>&1 | tail -3
bash scripts/r02_chain.sh 1
