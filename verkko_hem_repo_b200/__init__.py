"""Import alias: the package directory is ``verkko-hem-repo_b200/`` (not a valid Python identifier),
so ``import verkko_hem_repo_b200`` resolves to it through this two-line shim."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "verkko-hem-repo_b200")]
exec(open(_os.path.join(__path__[0], "__init__.py")).read())
