"""python -m bathymetric_gnn_b200.dropin <reference script> [script arguments ...]"""
import sys

from . import run_script

if len(sys.argv) < 2:
    sys.exit("usage: python -m bathymetric_gnn_b200.dropin <script.py> [args ...]")
run_script(sys.argv[1], sys.argv[2:])
