"""Where does the persistent 3xTF32 kernel spend its time?  Variants (results are WRONG by construction):
   nosplit = splitter does not touch shared memory; mma1 = one MMA per k-step instead of three."""
import os, sys, subprocess
shapes = [("4096^3 bn256", "M=4096,N=4096,K=4096,bn=256"), ("4096^3 bn128", "M=4096,N=4096,K=4096,bn=128"),
          ("8192x768x768 bn192", "M=8192,N=768,K=768,bn=192"), ("8192x768x3072 bn128", "M=8192,N=768,K=3072,bn=128"),
          ("AV b96 bn64", "M=1024,N=64,K=1024,nb=96,bn=64")]
code = '''
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from scripts import tc_probe_lib as L
print("%-8s %-24s %8.3f ms %7.1f TFLOP/s" % (os.environ.get("DLGRAD_TC_EXPERIMENT","full"), sys.argv[1], *L.time_shape(**eval("dict(" + sys.argv[2] + ")"))))
'''
for exp in ("", "nosplit", "mma1", "nosplit,mma1"):
    for name, kw in shapes:
        env = dict(os.environ, DLGRAD_TC_EXPERIMENT=exp, DLGRAD_TC_KERNEL="v2")
        subprocess.run([sys.executable, "-c", code, name, kw], env=env)
