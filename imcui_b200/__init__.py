"""Importable alias for the `image-matching-webui_b200/` package directory (its name has a
hyphen, which Python cannot import).  `import imcui_b200.hloc.extractors.superpoint` resolves
inside that directory, mirroring the reference's `imcui.hloc.*` module tree."""
from pathlib import Path as _Path

_PKG = _Path(__file__).resolve().parent.parent / "image-matching-webui_b200"
__path__ = [str(_PKG)]
exec(compile((_PKG / "__init__.py").read_text(), str(_PKG / "__init__.py"), "exec"))
