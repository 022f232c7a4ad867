"""Oracle constants (TEST INFRASTRUCTURE ONLY): a single import point so that every oracle module computes with
the product's table (xrt_b200/physconsts.py), which tests/test_abi.py pins bit for bit on the values dumped
from the reference (tests/golden/constants.npz; reference: xrt/backends/raycing/physconsts.py:5-42)."""
import importlib.util as _ilu
import os as _os

_path = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'xrt_b200', 'physconsts.py')
_spec = _ilu.spec_from_file_location('_xrt_b200_physconsts', _path)
_mod = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_mod)  # loaded by path: importing the xrt_b200 package would pull in torch for nothing
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith('_')})
