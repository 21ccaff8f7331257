"""Build an experiment variant of the product library: tools/build_variant.py NAME "-DFOO=1 ..."
-> spfft_b200/lib/variants/libspfft_b200_NAME.so (select it with SPFFT_B200_LIB=<path>)."""
import os
import shutil
import sys

name, defs = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
os.environ["SPFFT_B200_DEFS"] = defs
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.OBJDIR = os.path.join(ROOT, "build", "obj_" + name)
vdir = os.path.join(g.LIBDIR, "variants")
os.makedirs(vdir, exist_ok=True)
g.LIB = os.path.join(vdir, f"libspfft_b200_{name}.so")
print(g.build_library(force=True))
