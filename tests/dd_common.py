"""Shared helpers of the double-double LP tests (tests/test_dd_cpu.py, tests/test_gpu_dd.py)."""
import __graft_entry__ as g

random_lp = g.load_package().problems.random_lp
