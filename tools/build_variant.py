"""Builds a variant of the library with extra -D flags into tools/_build/ (A/B experiments):
    python tools/build_variant.py poly4 -DHRIEMO_ATTN_POLY_OF_8=4"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import build as B
out = os.path.join(ROOT, "tools", "_build", f"libhriemo_{sys.argv[1]}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *sys.argv[2:], *[os.path.join(B.CSRC, s) for s in B.SOURCES], "-o", out], check=True)
print(out)
