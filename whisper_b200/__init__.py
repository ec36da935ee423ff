"""Import shim: the product lives in the directory `whisper.coreml_b200/` (named after the
reference repo, so not an importable identifier); `import whisper_b200` resolves there."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "whisper.coreml_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
