"""Development aid: HVS_VARIANT=<name> makes the tools load hvs_b200/build/variants/libhvs_b200_<name>.so."""
import os

def use_variant():
    name = os.environ.get("HVS_VARIANT")
    if not name:
        return
    from hvs_b200 import build as b
    b.LIB_PATH = os.path.join(b.PKG_DIR, "build", "variants", f"libhvs_b200_{name}.so")
    b.is_fresh = lambda: True
    print(f"[variant] {b.LIB_PATH}")
