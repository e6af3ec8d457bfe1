#!/usr/bin/env python3
"""Golden for --expansion 3 (mkl_matrix_power recursion, HapHiC_cluster.py:2017-2023, 2033): the unmodified reference's
mcl() on the planted-block matrix of mcl_block200.npz.  Run in the build container:  python tests/golden/make_expansion_golden.py"""
import os
import sys

if os.environ.get("PYTHONHASHSEED") != "0":
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden as mg

ref = mg.import_reference()
ref.logger.setLevel(20)
link = mg.block_matrix(4, 50, seed=7)
mg.mcl_case(ref, "block200_e3", link, [1.4, 2.0], expansion=3)
link = mg.block_matrix(5, 40, seed=9, noise=0.05)
mg.mcl_case(ref, "block200_e4", link, [2.0], expansion=4, keep_iters=2)
