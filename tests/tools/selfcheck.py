"""Guard-column integrity check with the development build (QC_DEBUG_HOOKS=1 csrc/build.sh -> libqcart_dbg.so): runs tests/tools/sanitize.py's
cases and reads back how many guard cells of the shared-memory lines were found non-zero after the last substep (must be 0).

    QC_DEBUG_HOOKS=1 bash deepreinforcementlearningcontrolofquantumcartpoles_b200/csrc/build.sh && python tests/tools/selfcheck.py
"""
import ctypes, os, runpy, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = os.path.join(ROOT, "deepreinforcementlearningcontrolofquantumcartpoles_b200", "libqcart_dbg.so")
assert os.path.exists(lib), "build the development library first"
os.environ["QCART_LIB"] = lib
sys.argv = [sys.argv[0]]
runpy.run_path(os.path.join(ROOT, "tests", "tools", "sanitize.py"), run_name="__main__")
h = ctypes.CDLL(lib)
h.qc_debug_guard_errors.restype = ctypes.c_uint
n = h.qc_debug_guard_errors()
print("guard-column errors over all launches:", n)
sys.exit(1 if n else 0)
