"""`python -m dual_threshold_optimization_b200 ...` -- forwards to the C++ CLI (same flags as src/main.rs:20-66)."""
import os
import sys

from ._capi import CLI_PATH

if not os.path.exists(CLI_PATH):
    sys.exit(f"{CLI_PATH} not built: run __graft_entry__.build() first")
os.execv(CLI_PATH, [CLI_PATH] + sys.argv[1:])
