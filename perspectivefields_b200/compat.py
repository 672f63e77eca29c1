"""Optional alias so that callers written against the reference (``from perspective2d import PerspectiveFields``,
``perspective2d.perspectivefields.model_zoo``; demo/demo.py:1-9) run unchanged."""
import sys
import types


def install():
    from . import perspectivefields, variants

    if "perspective2d" in sys.modules and not getattr(sys.modules["perspective2d"], "__pf_b200_alias__", False):
        raise RuntimeError("a real 'perspective2d' package is already imported")
    pkg = types.ModuleType("perspective2d")
    pkg.__pf_b200_alias__ = True
    pkg.PerspectiveFields = perspectivefields.PerspectiveFields
    sub = types.ModuleType("perspective2d.perspectivefields")
    sub.PerspectiveFields = perspectivefields.PerspectiveFields
    sub.model_zoo = variants.model_zoo
    pkg.perspectivefields = sub
    sys.modules["perspective2d"] = pkg
    sys.modules["perspective2d.perspectivefields"] = sub
    return pkg
