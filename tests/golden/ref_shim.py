"""The reference-import shim lives in `oracle/ref_shim.py` (shared with bench.py's reference arm); re-exported here for
`make_golden.py` and older imports."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from oracle.ref_shim import *  # noqa: F401,F403,E402
from oracle.ref_shim import install, reference_available, reference_modules  # noqa: F401,E402
