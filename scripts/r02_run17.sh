#!/bin/bash
python -m pytest tests -m gpu -x -q -k "rayleigh or eul_operators or variants" 2>&1 | tail -6
