"""Build a tuning variant of the library: one csrc file recompiled with extra -D flags, linked with the default objects.

    python tools/build_variant.py NAME loudness_tile.cu -DSSB_SERIAL_F=32 -DSSB_SERIAL_MINB=2

-> soundscope_b200/_variants/lib_NAME.so (select it with SSB_LIB=...; tuning experiments only)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soundscope_b200 import build as B

name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
vdir = os.path.join(B.HERE, "_variants")
os.makedirs(vdir, exist_ok=True)
obj = os.path.join(vdir, f"{name}_{src[:-3]}.o")
r = subprocess.run([B.NVCC, *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, src), "-o", obj], capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stdout + r.stderr)
open(obj + ".log", "w").write(r.stdout + r.stderr)
objs = [os.path.join(B.BUILD, os.path.basename(s)[:-3] + ".o") for s in B.sources() if os.path.basename(s) != src] + [obj]
so = os.path.join(vdir, f"lib_{name}.so")
r = subprocess.run([B.NVCC, "-shared", "-cudart", "static", "-ccbin", B.HOST_CXX, "-o", so, *objs], capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stdout + r.stderr)
print(so)
