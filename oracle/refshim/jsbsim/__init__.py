"""Stub of the `jsbsim` PyPI module backed by the CPU oracle. TEST INFRASTRUCTURE ONLY."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.f16_oracle import OracleFDM  # noqa: E402

__version__ = "0.0-oracle-restatement"


class FGFDMExec:
    def __init__(self, root_dir, pm_root=None):
        self._root = root_dir
        self._fdm = OracleFDM()

    def set_debug_level(self, level):
        pass

    def load_model(self, model, add_model_to_path=True):
        if model != "f16":
            raise RuntimeError("the oracle restates aircraft/f16 only")
        return True

    def set_property_value(self, name, value):
        self._fdm.set_property_value(name, value)

    def get_property_value(self, name):
        return self._fdm.get_property_value(name)

    def __getitem__(self, name):
        return self._fdm.get_property_value(name)

    def __setitem__(self, name, value):
        self._fdm.set_property_value(name, value)

    def run_ic(self):
        return self._fdm.run_ic()

    def run(self):
        return self._fdm.run()

    def get_delta_t(self):
        return 1.0 / 120.0
