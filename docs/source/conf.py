"""Sphinx configuration of the gossipy_b200 documentation (MyST markdown + autodoc).

Build:  python -m sphinx -b html docs/source docs/_build/html     (sphinx + myst-parser; neither is needed to READ the
docs -- every page is plain markdown)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

project = "gossipy_b200"
author = "gossipy_b200 developers"
release = "0.2"
extensions = ["sphinx.ext.autodoc", "sphinx.ext.napoleon", "sphinx.ext.viewcode"]
try:
    import myst_parser  # noqa: F401
    extensions.append("myst_parser")
except Exception:
    pass
source_suffix = {".md": "markdown", ".rst": "restructuredtext"}
master_doc = "index"
autodoc_mock_imports = ["gossipy_b200._C"]
autodoc_member_order = "bysource"
html_theme = "alabaster"
exclude_patterns = ["_build"]
