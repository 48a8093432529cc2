#!/usr/bin/env python
# coding=utf-8
"""
scripts/sptk/compute_circular_srp.py of funcwj/setk on libsetk_b200's CUDA kernels: the same positional
arguments, flags and defaults; implemented in setk_b200/cli_tools.py (circular_srp_main).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from setk_b200.cli_tools import circular_srp_main  # noqa: E402

if __name__ == "__main__":
    circular_srp_main()
