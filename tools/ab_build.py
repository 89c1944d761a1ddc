"""Build library variants for on-GPU A/B runs:  python tools/ab_build.py name=DEF1,DEF2 name2= ...   -> libhitsir_<name>.so"""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "single-image-super-resolution-application_b200")
spec = importlib.util.spec_from_file_location("_b", os.path.join(PKG, "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
for arg in sys.argv[1:]:
    name, _, defs = arg.partition("=")
    defines = [d for d in defs.split(",") if d]
    print(b.build(defines=defines, lib=os.path.join(PKG, f"libhitsir_{name}.so"), verbose=False))
