"""Import shim: the product lives in the directory ``uav-wrf-les-ppo-lstm_b200/`` (the
name the project layout prescribes, which is not a valid Python identifier).  This
package makes it importable as ``uav_wrf_les_ppo_lstm_b200``."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "uav-wrf-les-ppo-lstm_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
del _f
