"""two-front elimination (CS_TWIST=1) against the one-front path: parity on the reference-arm mesh, factor time on 109x32x460"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
print("CS_TWIST", os.environ.get("CS_TWIST"), json.dumps(bench.small_mesh_check(torch, 0)))
