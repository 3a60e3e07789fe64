"""Import shim: exposes the package directory ``vvc-extension-mm_b200/`` (hyphen in the name, as the
repo layout requires) under the importable name ``vvc_extension_mm_b200``."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "vvc-extension-mm_b200")]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
