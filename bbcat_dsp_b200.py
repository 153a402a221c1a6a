"""Import alias: the package directory is named ``bbcat-dsp_b200`` (hyphen), which Python cannot
import by name.  This module turns itself into that package: ``import bbcat_dsp_b200``."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "bbcat-dsp_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
