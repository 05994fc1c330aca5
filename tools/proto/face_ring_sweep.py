"""K4' ring kernel: chunk-height sweep (CS_FACE_ZC) at 4.86 M cells -- run one value per process"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
r = bench.stencil_roofline(torch, 0)
print(os.environ.get("CS_FACE_ZC", "16"), json.dumps({k: r["k4_face_f64"][k] for k in ("us_per_launch", "frac")}))
