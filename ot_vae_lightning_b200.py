"""Import shim: the product package lives in the directory `ot-vae-lightning_b200/` (not a valid Python
identifier), so this module points its `__path__` there and runs the package body.  After
`import ot_vae_lightning_b200`, sub-modules import as `ot_vae_lightning_b200.ot.w2_utils` etc."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "ot-vae-lightning_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f, _os
