"""Import alias: the product package lives in `gnn-lm_b200/` (a name Python cannot import
directly); `import gnnlm_b200` resolves to it."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gnn-lm_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _os, _f
