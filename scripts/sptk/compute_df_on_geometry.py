#!/usr/bin/env python
# coding=utf-8
"""
scripts/sptk/compute_df_on_geometry.py of funcwj/setk on libsetk_b200's CUDA kernels: the same positional
arguments, flags and defaults; implemented in setk_b200/cli_tools.py (df_on_geometry_main).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from setk_b200.cli_tools import df_on_geometry_main  # noqa: E402

if __name__ == "__main__":
    df_on_geometry_main()
