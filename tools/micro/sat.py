import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
if os.environ.get("MCSAS_LIB"):
    from mcsas_b200 import _lib
    _lib.LIB_PATH = os.environ["MCSAS_LIB"]
sys.argv = ["x", "none"]
import tune
tune.saturated(os.environ.get("TAG", "sat"), {})
if os.environ.get("BASE"):
    tune.run("auto", {})
