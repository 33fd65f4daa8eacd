"""Dev tool: build tuning variants of the library (extra -D flags) next to the product build.
    python scripts/variants.py name1:-DSB_X=1,-DSB_Y=2 name2:-DSB_X=3 ...
-> solidboolean_b200/lib/variants/libsb_<name>.so ; run with SB_LIB_PATH=<that file>."""
import os, subprocess, sys, concurrent.futures as cf
sys.path.insert(0, ".")
from solidboolean_b200 import build as B
out = os.path.join(B.LIBDIR, "variants"); os.makedirs(out, exist_ok=True)
def one(spec):
    name, _, flags = spec.partition(":")
    flags = [f for f in flags.split(",") if f]
    od = os.path.join(B.OBJDIR, "v_" + name); os.makedirs(od, exist_ok=True)
    objs = []
    for src in B.SOURCES:
        obj = os.path.join(od, src.replace(".cu", ".o"))
        r = subprocess.run([B.NVCC] + B.FLAGS + flags + ["-c", os.path.join(B.CSRC, src), "-o", obj], capture_output=True, text=True)
        if r.returncode: raise RuntimeError(r.stderr)
        if src == "sb_classify.cu":
            for l in (r.stdout + r.stderr).splitlines():
                if "Used" in l or "spill" in l: print(name, l.strip())
        objs.append(obj)
    lib = os.path.join(out, "libsb_%s.so" % name)
    subprocess.run([B.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs, check=True)
    return lib
with cf.ThreadPoolExecutor(4) as ex:
    for lib in ex.map(one, sys.argv[1:]): print(lib)
