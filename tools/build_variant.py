"""Development aid: build libhvs_b200_<name>.so variants of the library with extra -D flags on one source
(default: the fused backward; HVS_VARIANT_BASE=mhc_stream_fwd.cu picks another) for timing experiments.
usage: build_variant.py name -DFOO -DBAR=1 ...   ->  hvs_b200/build/variants/"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hvs_b200
from hvs_b200 import build as b

def main():
    name, defs = sys.argv[1], sys.argv[2:]
    b.build()                                   # the other objects
    obj_dir = os.path.join(b.PKG_DIR, "build")
    vdir = os.path.join(obj_dir, "variants")
    os.makedirs(vdir, exist_ok=True)
    obj = os.path.join(vdir, f"fused_{name}.o")
    cmd = [b._nvcc(), *b.NVCC_FLAGS, *defs, "-Xptxas=-v", "-c", os.path.join(b.CSRC, os.environ.get("HVS_VARIANT_SRC", os.environ.get("HVS_VARIANT_BASE", "mhc_stream_bwd_fused.cu"))), "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        print(r.stdout); sys.exit(1)
    for line in r.stdout.splitlines():
        if "spill" in line or "Used" in line: print(name, line.strip())
    base = os.environ.get("HVS_VARIANT_BASE", "mhc_stream_bwd_fused.cu")      # the library source the variant replaces
    objs = [os.path.join(obj_dir, s.replace(".cu", ".o")) for s in b.SOURCES if s != base] + [obj]
    out = os.path.join(vdir, f"libhvs_b200_{name}.so")
    subprocess.run([b._nvcc(), "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
    print(out)

if __name__ == "__main__":
    main()
